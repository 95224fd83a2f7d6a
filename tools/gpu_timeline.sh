#!/bin/bash
# GPU timeline of one end-to-end call and one resident call (MBF_BAM_TIMELINE=1): completion time of every launch / copy
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
mkdir -p gpurun_out
export MBF_BENCH_READS=${MBF_BENCH_READS:-50000000}
python - <<'P' 2> gpurun_out/timeline.log
import os, sys
sys.path.insert(0, os.getcwd())
import bench
import mbf_bam_b200 as mb
reads = int(os.environ["MBF_BENCH_READS"])
path, iv, giv, info, _ = bench.ensure_bam(bench.BENCH_CONFIGS[1], os.environ.get("MBF_BENCH_DIR"), reads, 0, lambda: None)
for i in range(5):
    mb.count_reads_stranded(path, None, iv, giv)
os.environ["MBF_BAM_TIMELINE"] = "1"
sys.stderr.write("== e2e call\n")
mb.count_reads_stranded(path, None, iv, giv)
sys.stderr.write("e2e gpu_all %.2f\n" % mb.last_stats()["ms_gpu_all"])
os.environ["MBF_BAM_TIMELINE"] = "0"
rd = mb.Reader(path); rd.set_device(0); rd.preload()
rd.count_reads(1, iv, giv)
os.environ["MBF_BAM_TIMELINE"] = "1"
sys.stderr.write("== resident call\n")
rd.count_reads(1, iv, giv)
sys.stderr.write("resident gpu_all %.2f\n" % rd.stats()["ms_gpu_all"])
P
grep -c "mbfbam-tl" gpurun_out/timeline.log; grep "gpu_all" gpurun_out/timeline.log
