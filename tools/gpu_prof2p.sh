#!/bin/bash
mkdir -p gpurun_out
export MBF_BENCH_READS=${MBF_BENCH_READS:-8000000}
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
export MBF_BAM_INFLATE_D=${DPROF:-208} MBF_BAM_WINDOW_MB=768
python tools/profile_target.py > gpurun_out/plain_2p.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_inflate_tokens|k_tokens_bf|k_lz_resolve' -c 2 -f -o gpurun_out/prof_2p python tools/profile_target.py > gpurun_out/ncu_2p.log 2>&1
tail -n 2 gpurun_out/plain_2p.log
