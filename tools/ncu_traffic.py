"""profiles/ncu_traffic.json from an `ncu --set full` report of tools/profile_target.py and the job's statistics:

    python tools/ncu_traffic.py gpurun_out/prof_X.ncu-rep gpurun_out/profile_stats.json profiles/rNN_... > profiles/ncu_traffic.json

Per kernel: dram__bytes_read.sum + dram__bytes_write.sum of one launch, the algorithmic bytes of the same launch
(per-member averages of the job x the members of the launch, which the grid sizes give), and their ratio --
what bench.py's roofline.traffic scales with.  Algorithmic bytes as in DESIGN.md section 3 / bench.py."""
import csv
import json
import subprocess
import sys

rep, stats_path, source = sys.argv[1], sys.argv[2], sys.argv[3]
st = json.load(open(stats_path))
B = st["n_bgzf_blocks"]
per_member = {"C": st["compressed_bytes"] / B, "U": st["uncompressed_bytes"] / B, "T": st["n_tokens"] / B, "R": st["n_records"] / B}
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
col = {n: i for i, n in enumerate(h)}


def val(r, name):
    return float(r[col[name]].replace(",", "")) if r[col[name]] else 0.0


def unit_scale(name):
    u = rows[1][col[name]]
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


launches = {}
for r in rows[2:]:
    k = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    launches.setdefault(k, []).append(r)
# members of the captured window: k_lz_resolve runs one warp per member, four per CTA
lz = next(v for k, v in launches.items() if k.startswith("k_lz_resolve"))[0]
members = val(lz, "launch__grid_size") * 4
alg = {
    "k_tokens": per_member["C"] + 4 * per_member["T"],
    "k_lz_resolve": 4 * per_member["T"] + per_member["U"],
    "k_scan": 2 * per_member["C"],
    "k_frame": 8 * per_member["R"],
    "k_count": 96 * per_member["R"],
}
match = {"k_tokens": "k_tokens_par", "k_lz_resolve": "k_lz_resolve", "k_scan": "k_scan", "k_frame": "k_frame_walk", "k_count": "k_count_reads"}
res = {}
for key, prefix in match.items():
    ls = [r for k, v in launches.items() if k.startswith(prefix) for r in v]
    if not ls:
        continue
    # stages of several launches per window (scan: count + emit, frame walk: count + emit): sum one of each
    names = sorted({r[col["Kernel Name"]] for r in ls})
    dram = 0.0
    ms = 0.0
    for n in names:
        r = next(x for x in ls if x[col["Kernel Name"]] == n)
        dram += val(r, "dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum")
        ms += val(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(rows[1][col["gpu__time_duration.sum"]], 1.0)
    a = alg[key] * members
    if key == "k_scan":  # its own window: one CTA per 4 KiB tile, every byte tested by both passes
        a = 2.0 * 4096.0 * val(next(x for x in ls if x[col["Kernel Name"]] == names[0]), "launch__grid_size")
    res[key] = {"dram_bytes_per_launch": int(dram), "algorithmic_bytes_per_launch": int(a), "dram_bytes_per_algorithmic_byte": dram / a,
                "ms_alone_under_ncu": ms, "achieved_alone_gbs": a / (ms / 1e3) / 1e9 if ms > 0 else None, "members_in_launch": int(members), "kernels": names, "source": source}
json.dump(res, sys.stdout, indent=1)
print()
