"""Host-side costs of one call at the bench's scale, on the box that runs the bench."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import mbf_bam_b200 as mb
reads = int(os.environ.get("MBF_BENCH_READS", "50000000"))
path, iv, giv, info, _ = bench.ensure_bam(os.environ.get("MBF_BENCH_DIR", "/tmp/mbf_bench"), reads, 0, 1, lambda: None)


def best(f, n=7):
    b = 1e9
    for _ in range(n):
        t = time.perf_counter(); f(); b = min(b, time.perf_counter() - t)
    return 1e3 * b


rd = mb.Reader(path)
G = sum(len(v) for v in iv.values())
z = np.arange(G, dtype=np.uint64).tobytes(); zf = np.zeros(G, dtype=np.float64).tobytes(); sc = bytes([1] * len(iv))
print("open %.2f ms" % best(lambda: mb.Reader(path)))
print("parse gene_intervals + chunk table + shard plan %.2f ms" % best(lambda: rd.plan(giv, 0, 1)))
print("parse intervals (+ plan without gene table) %.2f ms" % best(lambda: rd.plan(iv, 0, 1)))
print("parse intervals + assemble two result dicts %.2f ms" % best(lambda: rd.assemble(1, iv, z, z, zf, zf, sc, 5)))
