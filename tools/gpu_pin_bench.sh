#!/bin/bash
# host-ingest microbenchmark on the GPU box (tools/pin_bench.cu); results in gpurun_out/pin_bench.log
mkdir -p gpurun_out
{
  nproc; lscpu | head -25; free -g; df -h /tmp /dev/shm; cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/shmem_enabled
  nvidia-smi topo -m 2>/dev/null | head -20
  ./tools/pin_bench /tmp 2048 8
  ./tools/pin_bench /dev/shm 2048 8
} > gpurun_out/pin_bench.log 2>&1
tail -5 gpurun_out/pin_bench.log
