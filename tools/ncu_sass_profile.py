"""Executed-instruction profile of one kernel by address range: python tools/ncu_sass_profile.py rep kernel-regex [bucket=64]
Prints, for consecutive groups of `bucket` SASS instructions, the share of warp-level executed instructions, the
average active threads and the share of stall samples -- shows where a kernel's instructions actually go."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                     capture_output=True, text=True).stdout.splitlines()
rows = [r for r in csv.reader(out[1:]) if len(r) > 10]
h = rows[0]
rows = [r for r in rows[1:] if r[0].startswith('0x')]
ia, isrc, ismp, iex, ith = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
# several launches of the kernel may be in the report: fold by address
agg = {}
order = []
for r in rows:
    a = int(r[ia], 16)
    if a not in agg:
        agg[a] = [r[isrc].strip(), 0, 0, 0]
        order.append(a)
    agg[a][1] += int(r[iex] or 0); agg[a][2] += int(r[ith] or 0); agg[a][3] += int(r[ismp] or 0)
totex = sum(v[1] for v in agg.values()); tots = sum(v[3] for v in agg.values())
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
base = order[0]
print("instructions executed (warp level):", totex, " stall samples:", tots)
for i in range(0, len(order), B):
    grp = order[i:i + B]
    ex = sum(agg[a][1] for a in grp); th = sum(agg[a][2] for a in grp); sm = sum(agg[a][3] for a in grp)
    if ex == 0 and sm == 0:
        continue
    print(f"{grp[0]-base:6x}-{grp[-1]-base:6x}  ex {100*ex/totex:5.1f}%  thr/inst {th/max(1,ex):5.1f}  samples {100*sm/max(1,tots):5.1f}%   {agg[grp[0]][0][:50]}")
