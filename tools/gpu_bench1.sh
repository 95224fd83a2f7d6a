#!/bin/bash
# parity + first bench numbers + launch list
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
export MBF_BENCH_READS=${MBF_BENCH_READS:-10000000}
timeout 1500 python bench.py --steps 2 --warmup 1 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "bench rc=$?"; tail -5 gpurun_out/bench1.err; cat gpurun_out/bench1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench1_ref.json 2>> gpurun_out/bench1.err
cat gpurun_out/bench1_ref.json
