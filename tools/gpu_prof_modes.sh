#!/bin/bash
# ncu --set full of two steady-state k_tokens_par launches for each pass-1 mode in $MODES (resident 20M-read job)
mkdir -p gpurun_out
export MBF_BENCH_READS=${MBF_BENCH_READS:-20000000}
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
for d in ${MODES:-400 410}; do
  export MBF_BAM_INFLATE_D=$d
  timeout 300 python tools/profile_target.py > gpurun_out/plain_m$d.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tokens_par|k_par_headers|k_lz_resolve' --launch-skip ${SKIP:-6} -c ${COUNT:-3} -f -o gpurun_out/prof_m$d python tools/profile_target.py > gpurun_out/ncu_m$d.log 2>&1
  tail -n 1 gpurun_out/plain_m$d.log; tail -n 2 gpurun_out/ncu_m$d.log
done
