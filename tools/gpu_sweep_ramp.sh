#!/bin/bash
# e2e of the default bench under first-window / ramp settings: SWEEP="DIV:RAMP_PINNED:PCT ..."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
for C in ${SWEEP:-8:0:150 8:1:300 16:1:300 8:1:200}; do
  IFS=: read D R P <<< "$C"
  MBF_BAM_FIRST_WINDOW_DIV=$D MBF_BAM_RAMP_PINNED=$R MBF_BAM_RAMP_PCT=$P timeout ${BENCH_TIMEOUT:-200} python bench.py --steps 5 --warmup 3 --no-parity ${BENCH_ARGS:-} > gpurun_out/bench_ramp.json 2> gpurun_out/bench_ramp.err
  echo "div=$D ramp_pinned=$R pct=$P rc=$?"; python - <<P
import json
try:
    j=json.load(open("gpurun_out/bench_ramp.json"))
    print("  e2e %.1f M/s (%.1f ms)  resident %.1f ms" % (j["value"]/1e6, j["ms_per_step"], j["resident"]["ms_per_step"]))
except Exception as e:
    print("  no result", e); print(open("gpurun_out/bench_ramp.err").read()[-1500:])
P
done
