#!/bin/bash
# e2e of the default bench under several window sizes / first-window divisors: SWEEP="W:DIV ..."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
for C in ${SWEEP:-512:8 256:4 192:4 128:2}; do
  IFS=: read W D <<< "$C"
  MBF_BAM_WINDOW_MB=$W MBF_BAM_FIRST_WINDOW_DIV=$D timeout ${BENCH_TIMEOUT:-200} python bench.py --steps 5 --warmup 3 --no-parity ${BENCH_ARGS:-} > gpurun_out/bench_w$W.json 2> gpurun_out/bench_w$W.err
  echo "W=$W div=$D rc=$?"; python - <<P
import json
try:
    j=json.load(open("gpurun_out/bench_w$W.json"))
    print("  e2e %.1f M/s (%.1f ms)  resident %.1f ms  stages %s" % (j["value"]/1e6, j["ms_per_step"], j["resident"]["ms_per_step"], {k:round(v,1) for k,v in j["stage_ms_rank0"].items()}))
except Exception as e:
    print("  no result", e); print(open("gpurun_out/bench_w$W.err").read()[-1500:])
P
done
