#!/bin/bash
# ncu evidence for the bench command itself: launch list (time per launch), then one --set full capture of
# the top kernels.  Each ncu pass runs only after the same command has exited 0 without ncu.
# usage: gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
CMD="python bench.py --steps 2 --warmup 1"
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2> gpurun_out/plain2_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:'k_tokens_bf|k_lz_resolve|k_frame_walk|k_count_reads|k_scan' -c 10 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -c 600 gpurun_out/plain_$TAG.log; echo; tail -2 gpurun_out/ncu_launch_$TAG.log gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out | tail -8
