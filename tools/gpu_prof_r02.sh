#!/bin/bash
# ncu evidence on one resident count_reads_stranded job (tools/profile_target.py, 20M reads):
#   launches_$TAG.csv : every launch with its duration (cold-cache, serialised: compare shares)
#   prof_$TAG.ncu-rep : --set full of one steady-state window's kernels
TAG=${1:-r02}
mkdir -p gpurun_out
export MBF_BENCH_READS=${MBF_BENCH_READS:-20000000}
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
CMD="python tools/profile_target.py"
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2> gpurun_out/plain2_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:'k_tokens_par|k_tokens_bf|k_lz_resolve|k_frame_walk|k_count_reads|k_scan|k_frame_fix' --launch-skip ${SKIP:-18} -c ${COUNT:-9} -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 2 gpurun_out/plain_$TAG.log; tail -n 3 gpurun_out/ncu_full_$TAG.log
