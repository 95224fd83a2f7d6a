"""One count_reads_stranded job over a cfg2-shaped BAM (size from MBF_BENCH_READS): the command ncu wraps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mbf_bam_b200 as mb  # noqa: E402

reads = int(os.environ.get("MBF_BENCH_READS", "4000000"))
cfg = int(os.environ.get("MBF_BENCH_CONFIG", "1"))
path, iv, giv, info, _ = bench.ensure_bam(bench.BENCH_CONFIGS[cfg], os.environ.get("MBF_BENCH_DIR"), reads, 0, lambda: None)
rd = mb.Reader(path)
rd.set_device(0)
if os.environ.get("MBF_PROFILE_RESIDENT", "1") != "0":
    rd.preload()
else:
    rd.count_reads(1, iv, giv)  # warm the pipeline buffers
f, r = rd.count_reads(1, iv, giv)
st = rd.stats()
import json  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "profile_stats.json"), "w") as fh:
    json.dump({k: st[k] for k in ("n_records", "n_bgzf_blocks", "compressed_bytes", "uncompressed_bytes", "n_tokens", "n_windows")}, fh)
print("records", st["n_records"], "launches", st["n_kernel_launches"], "fwd", f["_total"], "rev", r["_total"],
      {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")})
