#!/bin/bash
# gpurun_out/*_$TAG (tools/gpu_final.sh) -> committed summaries under profiles/
TAG=${1:-r02f}
cd "$(dirname "$0")/.."
cp gpurun_out/launches_$TAG.csv profiles/${TAG}_launches.csv
cp gpurun_out/bench_$TAG.json profiles/${TAG}_bench_n1.json
cp gpurun_out/bench_${TAG}_ref.json profiles/${TAG}_bench_n1_reference.json
cp gpurun_out/pytest_gpu_$TAG.log profiles/${TAG}_pytest_gpu.log
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
print("ncu --metrics gpu__time_duration.sum --clock-control none -c 600  python bench.py --steps 2 --warmup 1 --no-parity")
print("(50M reads, cfg2 shape; the first 600 launches: the warm-up and timed end-to-end calls; cold-cache serialised times: compare SHARES)")
print(f"launches {n}, total {tot/1000:.1f} ms")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:46s} n={a[0]:4d} total_us={a[1]:11.1f} avg_us={a[1]/a[0]:9.1f} share={a[1]/tot:6.3f}")
print()
print("bench.py of the same command without ncu (CUDA events on the launching streams, per call; K1 of consecutive windows runs on")
print("alternating streams, so the K1 spans overlap each other and their sum exceeds ms_gpu_all):")
d=json.loads(open('gpurun_out/plain_$TAG.json').read().strip().splitlines()[-1])
print(json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "resident_ms": d["resident"]["ms_per_step"], "stage_ms_rank0": d["stage_ms_rank0"]}))
for r in d["roofline_all"]:
    print("  %-14s ms_per_call %7.2f  achieved %8.1f GB/s  frac %.4f" % (r["kernel"], r["ms_per_call"], r["achieved"], r["frac"]))
PY
python tools/ncu_traffic.py gpurun_out/prof_$TAG.ncu-rep gpurun_out/profile_stats.json profiles/${TAG}_ncu_full_summary.txt > profiles/ncu_traffic.json
(echo "ncu --set full --clock-control none --import-source on -k regex:'k_tokens_par|k_par_headers|k_lz_resolve|k_frame_walk|k_count_reads|k_scan|k_frame_fix' --launch-skip 20 -c 10  python tools/profile_target.py"
 echo "(one count_reads_stranded job over a resident 20M-read cfg2 file: 512 MiB windows of ~27 k BGZF members; the .ncu-rep stays in gpurun_out/)"
 echo; python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep | grep -v "bank_conf"
 echo; echo "== where k_tokens_par's instructions go (tools/ncu_sass_profile.py, 64-instruction buckets with > 1 % of the instructions or samples)"
 python tools/ncu_sass_profile.py gpurun_out/prof_$TAG.ncu-rep k_tokens_par 64 | python -c "
import sys
for l in sys.stdin:
    p=l.split()
    try: ex=float(p[2].rstrip('%')); sm=float(p[7].rstrip('%'))
    except Exception: print(l.rstrip()); continue
    if ex>1.0 or sm>2.0: print(l.rstrip())"
 echo; echo "== where k_lz_resolve's instructions go"
 python tools/ncu_sass_profile.py gpurun_out/prof_$TAG.ncu-rep k_lz_resolve 32 | python -c "
import sys
for l in sys.stdin:
    p=l.split()
    try: ex=float(p[2].rstrip('%')); sm=float(p[7].rstrip('%'))
    except Exception: print(l.rstrip()); continue
    if ex>1.5 or sm>3.0: print(l.rstrip())"
 echo; echo "== DRAM traffic against algorithmic bytes per launch (tools/ncu_traffic.py -> profiles/ncu_traffic.json, read by bench.py)"
 cat profiles/ncu_traffic.json) > profiles/${TAG}_ncu_full_summary.txt
