#!/bin/bash
# source-level ncu capture of the inflate kernels
mkdir -p gpurun_out
export MBF_BENCH_READS=${MBF_BENCH_READS:-2000000}
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
for D in ${SWEEP:-0 1}; do
  export MBF_BAM_INFLATE_D=$D
  python tools/profile_target.py > gpurun_out/plain_inf$D.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'k_inflate' -c 1 -f -o gpurun_out/prof_inf$D python tools/profile_target.py > gpurun_out/ncu_inf$D.log 2>&1
  tail -n 2 gpurun_out/plain_inf$D.log
done
ls -la gpurun_out | head -20
