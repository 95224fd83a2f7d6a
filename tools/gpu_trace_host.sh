#!/bin/bash
# host-side timeline of a few e2e calls (MBF_BAM_TRACE) on the default bench file
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
export MBF_BENCH_READS=${MBF_BENCH_READS:-50000000}
python - <<'P' 2>&1 | grep -v "finish:" | tail -${TAIL:-80}
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
import mbf_bam_b200 as mb
reads = int(os.environ["MBF_BENCH_READS"])
path, iv, giv, info, _ = bench.ensure_bam(bench.BENCH_CONFIGS[1], os.environ.get("MBF_BENCH_DIR"), reads, 0, lambda: None)
for i in range(5):
    mb.count_reads_stranded(path, None, iv, giv)
os.environ["MBF_BAM_TRACE"] = "1"
for i in range(2):
    t = time.time()
    f, r = mb.count_reads_stranded(path, None, iv, giv)
    dt = (time.time() - t) * 1e3
    st = mb.last_stats()
    sys.stderr.write("call %d: %.2f ms python wall, C call %.2f ms, gpu_all %.2f ms\n" % (i, dt, st["ms_total"], st["ms_gpu_all"]))
P
