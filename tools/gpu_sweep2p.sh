#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
MBF_BAM_INFLATE_D=${DTEST:-2} timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
export MBF_BENCH_READS=${MBF_BENCH_READS:-20000000}
for D in ${SWEEP:-1 2}; do for W in ${WINS:-256 768}; do
  echo "== D=$D W=$W"
  MBF_BAM_WINDOW_MB=$W MBF_BAM_INFLATE_D=$D timeout 600 python tools/profile_target.py 2>&1 | tail -1
  MBF_BAM_WINDOW_MB=$W MBF_BAM_INFLATE_D=$D timeout 600 python tools/profile_target.py 2>&1 | tail -1
done; done
