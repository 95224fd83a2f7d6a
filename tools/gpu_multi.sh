#!/bin/bash
# multi-GPU bench as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
export MBF_BENCH_READS=${MBF_BENCH_READS:-20000000}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --reads $((MBF_BENCH_READS * N)) > gpurun_out/bench_n1_same.json 2>> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n1_same.json
