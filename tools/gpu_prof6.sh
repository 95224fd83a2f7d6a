#!/bin/bash
mkdir -p gpurun_out
export MBF_BENCH_READS=${MBF_BENCH_READS:-20000000}
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
export MBF_BAM_INFLATE_D=6
for W in 512; do echo "== W=$W"; MBF_BAM_WINDOW_MB=$W python tools/profile_target.py 2>&1 | tail -1; MBF_BAM_WINDOW_MB=$W python tools/profile_target.py 2>&1 | tail -1; done
export MBF_BAM_WINDOW_MB=512
python tools/profile_target.py > gpurun_out/plain_inf6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_inflate' -c 1 -f -o gpurun_out/prof_inf6 python tools/profile_target.py > gpurun_out/ncu_inf6.log 2>&1
tail -n 2 gpurun_out/plain_inf6.log
