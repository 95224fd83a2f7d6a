#!/bin/bash
# bench.py, both arms, for the BASELINE.json configs in $CFGS on ${GPUS:-1} GPU(s) -> gpurun_out/bench_cfg<k>_n<N>[_ref].json
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
N=${GPUS:-1}
for k in ${CFGS:-0 2 4 3}; do
  if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"; fi
  timeout ${BENCH_TIMEOUT:-900} $RUN bench.py --gpus $N --config $k --steps ${STEPS:-3} --warmup ${WARMUP:-3} ${BENCH_ARGS:-} > gpurun_out/bench_cfg${k}_n$N.json 2> gpurun_out/bench_cfg${k}_n$N.err
  echo "cfg $k n $N rc=$?"; tail -c 600 gpurun_out/bench_cfg${k}_n$N.err
  python - <<P
import json
try:
    j = json.loads(open("gpurun_out/bench_cfg${k}_n$N.json").read().strip().splitlines()[-1])
    print("  value %.1f M reads/s  ms/step %.1f  reads %d  fns %s  parity %s" % (j["value"] / 1e6, j["ms_per_step"], j["config"]["reads"], j["config"]["functions_per_step"], j.get("parity")))
except Exception as e:
    print("  no result", e)
P
  if [ -z "$NO_REF" ]; then
    timeout 600 python bench.py --impl reference --gpus $N --config $k --steps 1 --warmup 1 ${BENCH_ARGS:-} > gpurun_out/bench_cfg${k}_n${N}_ref.json 2>> gpurun_out/bench_cfg${k}_n$N.err
    python -c "
import json
j=json.loads(open('gpurun_out/bench_cfg${k}_n${N}_ref.json').read().strip().splitlines()[-1]); print('  reference arm %.1f M reads/s (%s)' % (j['value']/1e6, j['cpu_baseline']['sample']))"
  fi
  rm -f /tmp/mbf_bench/cfg*_l6.bam /dev/shm/mbf_bench/cfg*_l6.bam
done
