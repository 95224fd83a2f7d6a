#!/bin/bash
# the driver's bench invocation (default size) for both arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_full_ref.json 2>> gpurun_out/bench_full.err
cat gpurun_out/bench_full_ref.json
