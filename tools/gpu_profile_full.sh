#!/bin/bash
# one --set full capture of a steady-state window of the bench command: the default skip lands on the
# second full-size window of the resident (value) phase (e2e and pinned phases: 2 x 3 jobs x 10 windows x 7
# matched kernels come first)
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
CMD="python bench.py --steps 2 --warmup 1"
$CMD > gpurun_out/plain2_$TAG.log 2> gpurun_out/plain2_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:'k_tokens_bf|k_lz_resolve|k_frame_walk|k_count_reads|k_scan' --launch-skip ${SKIP:-434} -c 7 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_full_$TAG.log
