#!/bin/bash
# one GPU call: build, GPU test suite, smoke, bench (both arms) -> gpurun_out/
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
timeout 900 python bench.py --impl reference ${BENCH_ARGS:-} > gpurun_out/bench_full_ref.json 2>> gpurun_out/bench_full.err
cat gpurun_out/bench_full_ref.json
