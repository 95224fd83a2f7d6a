#!/bin/bash
# multi-GPU bench as the driver launches it (one rank per GPU under torchrun), then one process fanning out over
# the same GPUs, then the same file on one GPU
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
export MBF_BENCH_READS=${MBF_BENCH_READS:-50000000}
timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "torchrun rc=$?"; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
if [ -n "$SINGLE" ]; then
timeout 900 python bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_n${N}_oneproc.json 2> gpurun_out/bench_n${N}_oneproc.err
echo "one process rc=$?"; tail -3 gpurun_out/bench_n${N}_oneproc.err; cat gpurun_out/bench_n${N}_oneproc.json
fi
if [ -n "$REF" ]; then
timeout 900 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_n${N}_ref.json 2>> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n${N}_ref.json
fi
