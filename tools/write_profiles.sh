#!/bin/bash
# turn gpurun_out/{launches,prof,plain2,bench_full*}_$TAG into the committed summaries under profiles/
TAG=${1:-r01tp}
cd "$(dirname "$0")/.."
cp gpurun_out/launches_$TAG.csv profiles/${TAG}_launches.csv
cp gpurun_out/bench_full.json profiles/${TAG}_bench_n1.json
cp gpurun_out/bench_full_ref.json profiles/${TAG}_bench_n1_reference.json
python - > profiles/${TAG}_launch_summary.txt <<PY
import csv, collections, json
rows=[r for r in csv.reader(open('gpurun_out/launches_$TAG.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
agg=collections.OrderedDict(); n=0
for r in rows[1:]:
    if r[h.index('Metric Name')]!='gpu__time_duration.sum': continue
    us=float(r[iv].replace(',',''))/1000
    k=r[ik].split('(')[0]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=us; n+=1
tot=sum(a[1] for a in agg.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none -c 400  python bench.py --steps 2 --warmup 1")
print("(50M reads, cfg2 shape; two-pass K1; the first 400 launches belong to the streamed (e2e) phase, whose windows")
print(" ramp up from 64 MiB to 512 MiB; cold-cache serialized times: compare SHARES)")
print(f"launches {n}, total {tot/1000:.1f} ms")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:46s} n={a[0]:4d} total_us={a[1]:11.1f} avg_us={a[1]/a[0]:9.1f} share={a[1]/tot:6.3f}")
print()
print("bench.py stage_ms_rank0 of the same command without ncu (CUDA events on the launching streams, per step, resident")
print("phase; K1 of consecutive windows runs on alternating streams, so the K1 spans overlap and their sum exceeds ms_gpu_all):")
d=json.loads(open('gpurun_out/plain2_$TAG.log').read().strip().splitlines()[-1])
s=d['stage_ms_rank0']
print(json.dumps(s))
den=s['ms_gpu_scan']+s['ms_gpu_inflate']+s['ms_gpu_frame']+s['ms_gpu_count']
print("K1 share of the summed kernel spans: %.3f" % (s['ms_gpu_inflate']/den))
PY
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep | grep -v "bank_conf" > /tmp/full_summary.txt
(echo "ncu --set full --clock-control none --import-source on -k regex:'k_tokens_bf|k_lz_resolve|k_frame_walk|k_count_reads|k_scan' --launch-skip 434 -c 7  python bench.py --steps 2 --warmup 1"
 echo "(one steady-state 512 MiB window of the resident phase of the 50M-read cfg2 file; the .ncu-rep stays in gpurun_out/, not committed)"
 echo; cat /tmp/full_summary.txt; echo
 echo "k_tokens_bf: algorithmic bytes of this launch = blocks (8 x grid of k_lz_resolve / 2 = 4 x grid) x 80.5 KB (compressed stream + 4 B per"
 echo "token; bench.py algorithmic_bytes_per_step / BGZF blocks per step); dram read + write above / that = bench.py K1_TRAFFIC_RATIO_NCU"
 echo; echo "== top SASS instructions of k_tokens_bf<8> by stall samples (tools/ncu_sass_hot.py)"
 python tools/ncu_sass_hot.py gpurun_out/prof_$TAG.ncu-rep k_tokens_bf 40 | awk 'NR==1 || NR%2==0'
 echo; echo "== top SASS instructions of k_lz_resolve<2048> by stall samples"
 python tools/ncu_sass_hot.py gpurun_out/prof_$TAG.ncu-rep k_lz_resolve 30 | awk 'NR==1 || NR%2==0') > profiles/${TAG}_ncu_full_summary.txt
