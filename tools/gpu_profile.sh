#!/bin/bash
# ncu evidence: launch list (time per launch) and one --set full capture of the top kernels.
# usage: gpu_profile.sh <tag>
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
export MBF_BENCH_READS=${MBF_BENCH_READS:-4000000}
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
CMD="python tools/profile_target.py"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_inflate_batched|k_frame_walk|k_count_reads|k_scan' -c 8 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/plain_$TAG.log gpurun_out/ncu_launch_$TAG.log gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
