"""Top SASS instructions of one kernel by stall samples: python tools/ncu_sass_hot.py rep kernel-regex [n]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                     capture_output=True, text=True).stdout.splitlines()
rows = [r for r in csv.reader(out[1:]) if len(r) > 10]
h = rows[0]
rows = [h] + [r for r in rows[1:] if r[0].startswith('0x')]
ia, isrc, ismp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ismp] or 0) for r in rows[1:])
totex = sum(int(r[iex] or 0) for r in rows[1:])
print("samples", tot, "inst", totex)
base = int(rows[1][ia], 16)
top = sorted(rows[1:], key=lambda r: -int(r[ismp] or 0))[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]
for r in top:
    st = sorted(((int(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:3]
    print(f"{int(r[ia],16)-base:5x} {100*int(r[ismp])/tot:5.2f}% ex {100*int(r[iex])/totex:5.2f}%  {r[isrc].strip()[:58]:58s} " + " ".join(f"{n}:{v}" for v, n in st if v))
