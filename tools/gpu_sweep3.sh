#!/bin/bash
# CONFIGS="D:W:RING ..." -> stage timings of one resident job each (twice); optional DTEST runs the gpu tests first
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
if [ -n "$DTEST" ]; then MBF_BAM_INFLATE_D=$DTEST timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4; fi
export MBF_BENCH_READS=${MBF_BENCH_READS:-20000000}
for C in ${CONFIGS:-1:256:2048}; do
  IFS=: read D W R <<< "$C"
  echo "== D=$D W=$W RING=$R"
  for i in 1 2; do
    MBF_BAM_WINDOW_MB=$W MBF_BAM_INFLATE_D=$D MBF_BAM_RESOLVE_RING=$R timeout 600 python tools/profile_target.py 2>&1 | tail -1 | sed 's/records [0-9]* launches//; s/fwd [0-9]* rev [0-9]*//'
  done
done
