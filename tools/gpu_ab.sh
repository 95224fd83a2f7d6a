#!/bin/bash
# A/B of environment settings on the bench: AB="NAME=VALUE NAME=VALUE ..." (each run = bench.py, our arm)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for KV in ${AB}; do
  echo "== $KV"
  env $KV timeout 900 python bench.py 2> gpurun_out/bench_ab.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    s=d['stage_ms_rank0']
    print('value %.4g e2e %.4g pinned %.4g ms/step %.1f e2e_ms %.1f inflate %.1f tokens %.1f' % (d['value'], d['e2e']['value'], (d.get('e2e_pinned_file') or {}).get('value',0), d['ms_per_step'], d['e2e'].get('ms_per_step',0), s['ms_gpu_inflate'], s['ms_gpu_tokens']))
"
  tail -2 gpurun_out/bench_ab.err
done
