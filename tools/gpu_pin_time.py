"""How long does page-locking the file mapping take? (decides whether it can be done per call)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import mbf_bam_b200 as mb
reads = int(os.environ.get("MBF_BENCH_READS", "50000000"))
path, iv, giv, info, _ = bench.ensure_bam(os.environ.get("MBF_BENCH_DIR", "/tmp/mbf_bench"), reads, 0, 1, lambda: None)
rd = mb.Reader(path)
rd.count_reads(1, iv, giv)
for i in range(3):
    t = time.perf_counter(); ok = rd.pin_file(); t1 = time.perf_counter()
    rd.count_reads(1, iv, giv); t2 = time.perf_counter()
    rd.unpin_file(); t3 = time.perf_counter()
    print("pin %s %.1f ms, count %.1f ms, unpin %.1f ms" % (ok, 1e3 * (t1 - t), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
for i in range(2):
    t = time.perf_counter(); rd.count_reads(1, iv, giv); print("staged count %.1f ms" % (1e3 * (time.perf_counter() - t)))
