#!/bin/bash
# K1 development loop: inflate parity tests (optionally under compute-sanitizer), then A/B timing of the pass-1 variants
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
if [ -n "$SANITIZE" ]; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "inflate_matches_zlib or long_matches or alternate" > gpurun_out/sanitize.log 2>&1
  echo "sanitize rc=$?"; tail -25 gpurun_out/sanitize.log
fi
timeout ${PYTEST_TIMEOUT:-600} python -m pytest tests -x -q -m gpu -k "${K:-inflate or alternate or counts_match or baseline_config}" > gpurun_out/pytest_k1.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_k1.log
for d in ${MODES:-400 401 402 403 308}; do
  MBF_BAM_INFLATE_D=$d timeout ${BENCH_TIMEOUT:-150} python bench.py --steps 3 --warmup 2 --no-parity ${BENCH_ARGS:-} > gpurun_out/bench_d$d.json 2> gpurun_out/bench_d$d.err
  echo "mode $d rc=$?"; python - <<P
import json
try:
    j=json.load(open("gpurun_out/bench_d$d.json"))
    print("  e2e %.1f M/s (%.1f ms)  resident %.1f ms  stages %s" % (j["value"]/1e6, j["ms_per_step"], j["resident"]["ms_per_step"], {k:round(v,1) for k,v in j["stage_ms_rank0"].items()}))
except Exception as e:
    print("  no result", e); print(open("gpurun_out/bench_d$d.err").read()[-1500:])
P
done
