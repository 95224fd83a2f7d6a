#!/usr/bin/env python
"""Extract the CIGAR golden vectors from the reference's own unit tests.

Source: /root/reference/src/bam_ext.rs:50-382 (`spliced_reads`, aligned blocks)
and :384-604 (`test_introns`).  Each record of the (absent) sample BAM is
described by a comment `//<CIGAR> - <index>` followed by asserts on the
expected (start, stop) tuples.  pos = blocks[0].start.

Run once in the build container (the reference tree does not exist on the GPU
box); the output `cigar_golden.json` is committed next to this script.
"""
import json
import re
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/bam_ext.rs"
lines = open(SRC).read().split("\n")

start_blocks = next(i for i, l in enumerate(lines) if "fn spliced_reads" in l)
start_introns = next(i for i, l in enumerate(lines) if "fn test_introns" in l)

com = re.compile(r"^\s*//(\S+) - (\d+)\s*$")
blk = re.compile(r"assert!\(blocks\[(\d+)\] == \((\d+), (\d+)\)\);")
iln = re.compile(r"assert_eq!\(introns\.len\(\), (\d+)\);")
iv = re.compile(r"assert_eq!\(introns\[(\d+)\], \((\d+), (\d+)\)\);")

recs = {}
cur = None
for l in lines[start_blocks:start_introns]:
    m = com.match(l)
    if m:
        cur = int(m.group(2))
        recs.setdefault(cur, {"cigar": m.group(1), "blocks": {}, "introns": {}, "n_introns": None})
        assert recs[cur]["cigar"] == m.group(1)
        continue
    m = blk.search(l)
    if m:
        recs[cur]["blocks"][int(m.group(1))] = (int(m.group(2)), int(m.group(3)))
cur = None
for l in lines[start_introns:]:
    m = com.match(l)
    if m:
        cur = int(m.group(2))
        assert recs[cur]["cigar"] == m.group(1), (cur, recs[cur]["cigar"], m.group(1))
        continue
    m = iln.search(l)
    if m:
        recs[cur]["n_introns"] = int(m.group(1))
        continue
    m = iv.search(l)
    if m:
        recs[cur]["introns"][int(m.group(1))] = (int(m.group(2)), int(m.group(3)))

out = []
for idx in sorted(recs):
    r = recs[idx]
    blocks = [r["blocks"][k] for k in sorted(r["blocks"])]
    assert sorted(r["blocks"]) == list(range(len(blocks)))
    introns = [r["introns"][k] for k in sorted(r["introns"])]
    out.append({
        "index": idx, "cigar": r["cigar"], "pos": blocks[0][0],
        "blocks": blocks, "n_introns": r["n_introns"], "introns": introns,
    })
open(__file__.replace("make_cigar_golden.py", "cigar_golden.json"), "w").write("[\n" + ",\n".join(json.dumps(r) for r in out) + "\n]\n")
print(len(out), "records")
