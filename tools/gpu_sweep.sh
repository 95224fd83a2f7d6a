#!/bin/bash
# parity tests, then stage timings of one job for several inflate configurations
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
export MBF_BENCH_READS=${MBF_BENCH_READS:-10000000}
for D in ${SWEEP:-0 4 8 16 32}; do
  echo "== MBF_BAM_INFLATE_D=$D"
  MBF_BAM_INFLATE_D=$D timeout 600 python tools/profile_target.py 2>&1 | tail -2
  MBF_BAM_INFLATE_D=$D timeout 600 python tools/profile_target.py 2>&1 | tail -1
done
