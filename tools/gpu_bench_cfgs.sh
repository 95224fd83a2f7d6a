#!/bin/bash
# bench.py (our arm only) under several K1 / window configurations: CONFIGS="D:W:TOKENS_PER_SM:RESOLVE_RING ..."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
for C in ${CONFIGS:-1:256:0:2048}; do
  IFS=: read D W T R <<< "$C"
  echo "== D=$D W=$W TOKENS_PER_SM=$T RING=$R"
  MBF_BAM_WINDOW_MB=$W MBF_BAM_INFLATE_D=$D MBF_BAM_TOKENS_PER_SM=${T:-0} MBF_BAM_RESOLVE_RING=${R:-2048} timeout 900 python bench.py 2> gpurun_out/bench_cfg.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.4g e2e %.4g pinned %.4g ms/step %.1f e2e_ms %.1f inflate %.1f' % (d['value'], d['e2e']['value'], (d.get('e2e_pinned_file') or {}).get('value',0), d['ms_per_step'], d['e2e'].get('ms_per_step',0), d['stage_ms_rank0']['ms_gpu_inflate']))
"
  tail -2 gpurun_out/bench_cfg.err
done
