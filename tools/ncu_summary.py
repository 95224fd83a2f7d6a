"""Print the handful of ncu metrics we look at for every kernel of a report: python tools/ncu_summary.py x.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
stall = [i for i, n in enumerate(h) if "issue_stalled" in n and n.endswith("per_issue_active.ratio") and "not_issued" not in n]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")])
    for w in WANT:
        if w in h:
            print(f"   {w:70s} {r[h.index(w)]} {units[h.index(w)]}")
    st = sorted(((float(r[i] or 0), h[i]) for i in stall), reverse=True)[:6]
    for v, n in st:
        print(f"   stall {n.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.2f}")
