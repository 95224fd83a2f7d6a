"""CPU suite, part 1: the oracle itself (pinned against the reference's golden vectors) and the
bit-level logic of the DEFLATE core shared with the kernels."""
import json
import os
import subprocess
import zlib

import pytest

from conftest import ROOT, make_case
from mbf_bam_b200.synth import bamio
from oracle import c_oracle, py_oracle

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "cigar_golden.json")))


def test_golden_has_all_58_records():
    assert len(GOLDEN) == 58 and [g["index"] for g in GOLDEN] == list(range(58))


@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: "%d_%s" % (g["index"], g["cigar"]))
def test_cigar_golden_oracles(g):
    """reference src/bam_ext.rs:50-382 (blocks) and :384-604 (introns)."""
    cig = bamio.parse_cigar(g["cigar"])
    words = [(l << 4) | op for op, l in cig]
    blocks = [tuple(b) for b in g["blocks"]]
    introns = [tuple(b) for b in g["introns"]]
    assert py_oracle.cigar_blocks(g["pos"], cig) == blocks
    assert c_oracle.cigar_pairs(words, g["pos"], False) == blocks
    assert len(py_oracle.cigar_introns(g["pos"], cig)) == g["n_introns"]
    got = py_oracle.cigar_introns(g["pos"], cig)
    assert got[:len(introns)] == introns  # the reference asserts a prefix of the introns for some records
    assert c_oracle.cigar_pairs(words, g["pos"], True) == got


@pytest.mark.parametrize("seed,kw", [
    (1, {}), (2, dict(paired=True)), (3, dict(split_records=True, level=9)),
    (4, dict(sync_every=700)), (5, dict(exon_mode=False)), (6, dict(n_contigs=4, contig_len=900_000)),
])
def test_c_oracle_matches_py_oracle(tmp_path, seed, kw):
    path, contigs, recs, iv, giv = make_case(tmp_path, seed, **kw)
    assert c_oracle.count_reads_unstranded(path, None, iv, giv, nthreads=3) == py_oracle.count_reads_unstranded(path, None, iv, giv)
    assert c_oracle.count_reads_stranded(path, None, iv, giv, nthreads=2) == py_oracle.count_reads_stranded(path, None, iv, giv)
    a = c_oracle.count_reads_weighted(path, None, iv, giv, nthreads=2)
    b = py_oracle.count_reads_weighted(path, None, iv, giv)
    for x, y in zip(a, b):
        assert set(x) == set(y)
        for k in x:
            assert abs(x[k] - y[k]) <= 1e-9 * max(1.0, abs(y[k]))
    assert c_oracle.count_introns(path, nthreads=2) == py_oracle.count_introns(path)
    got = c_oracle.chunks(path, None, giv)
    name2tid = {n: i for i, (n, _) in enumerate(contigs)}
    want = py_oracle.make_chunks([(c, name2tid[c], dict(contigs)[c]) for c in giv],
                                 {c: py_oracle.build_intervals(giv[c])[0] for c in giv})
    assert got == [(t, s, e) for _, t, s, e in want]


def test_oracle_semantics_hand_case(tmp_path):
    """Hand-computed expectations for the rules of SURVEY 0.4-0.6 (counters.rs:46-104,151)."""
    R = bamio.Rec
    nh2 = bamio.aux_int(b"NH", 2)
    refs = [("c1", 3_000_000)]
    iv = {"c1": [("gA", 1, [100], [200]), ("gB", -1, [150, 400], [250, 500]), ("gC", 0, [999_990], [1_000_020])]}
    giv = {"c1": [("gA", 1, [100], [200]), ("gB", -1, [150], [500]), ("gC", 0, [999_990], [1_000_020])]}
    recs = [
        R(0, 90, b"a", 0, 255, [(0, 20)], 20),                    # gA fwd (+ strand, forward read)
        R(0, 160, b"b", 16, 255, [(0, 10)], 10),                  # gA (rev for +) and gB (fwd for -)
        R(0, 160, b"mm", 0, 3, [(0, 10)], 10, nh2),               # multi-mapper gA,gB ...
        R(0, 170, b"mm", 0, 3, [(0, 10)], 10, nh2),               # ... same qname, same chunk: counts once
        R(0, 180, b"s", 0, 255, [(0, 10), (3, 215), (0, 10)], 20),  # blocks hit gA+gB and gB again: gB once
        R(0, 300, b"o", 0, 255, [(0, 10)], 10),                   # outside
        R(0, 999_995, b"e", 0, 255, [(0, 10)], 10),               # gC (strand 0 behaves like -1): rev
        R(0, 1_000_010, b"f", 16, 255, [(0, 5)], 5),              # chunk 0 was extended to 1_000_021: still chunk 0
        R(0, 2_500_000, b"mm", 0, 3, [(0, 10)], 10, nh2),         # outside, other chunk
    ]
    path = str(tmp_path / "hand.bam")
    bamio.write_bam(path, refs, recs)
    assert py_oracle.make_chunks([("c1", 0, 3_000_000)], {"c1": py_oracle.build_intervals(giv["c1"])[0]})[0][3] == 1_000_021
    for orc in (py_oracle, c_oracle):
        f, r = orc.count_reads_stranded(path, None, iv, giv)
        assert (f["gA"], r["gA"]) == (1 + 1 + 1, 1)           # a, mm-set, s | b
        u = orc.count_reads_unstranded(path, None, iv, giv)
        assert u["gA"] == 4 and u["gB"] == 3 and u["gC"] == 2
        assert u["_outside"] == 2 and u["_total"] == 9 and u["_c1"] == 9
        assert f["_outside"] == 2 and "_outside" not in r
        assert f["gC"] == 1 and r["gC"] == 1                   # e: forward read on strand 0 -> rev; f: reverse read -> fwd
        assert f["gB"] == 1 and r["gB"] == 2                   # b fwd | mm-set, s rev


def test_inflate_oracle_is_zlib(tmp_path):
    path, *_ = make_case(tmp_path, 11, n_reads=4000)
    data = open(path, "rb").read()
    want = b"".join(p for _, _, p in bamio.iter_bgzf_blocks(data))
    assert c_oracle.inflate_buffer(data, len(want)) == want


@pytest.mark.parametrize("kw", [dict(), dict(level=1), dict(level=9, split_records=True), dict(sync_every=500),
                                dict(strategy=zlib.Z_FIXED), dict(level=0), dict(strategy=zlib.Z_HUFFMAN_ONLY),
                                dict(strategy=zlib.Z_RLE, payload_limit=3000)],
                         ids=["default", "l1", "l9split", "syncflush", "fixed", "stored", "huffonly", "rle_small"])
def test_deflate_core_logic_on_host(tmp_path, kw):
    """The DEFLATE core the kernel runs (mbf_bam_b200/csrc/inflate_core.cuh), compiled for the host by
    g++ and compared with zlib member by member -- logic only; the GPU parity test is in test_gpu_*."""
    exe = str(tmp_path / "host_emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_emu.cpp"), "-lz"])
    path, *_ = make_case(tmp_path, 12, n_reads=6000, **kw)
    out = subprocess.run([exe, path], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
