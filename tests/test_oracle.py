"""CPU suite, part 1: the oracle itself (pinned against the reference's golden vectors) and the
bit-level logic of the DEFLATE core shared with the kernels."""
import json
import os
import subprocess
import zlib

import pytest

from conftest import ROOT, make_case
from synth import bamio
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


# ------------------------------------------------------------------------------------------------ bio's tree
def _check_avl(n):
    """-> (height, max_end, min_start, max_start); asserts the AVL + search-tree + max-augmentation invariants"""
    if n is None:
        return 0, -1, None, None
    lh, lm, lmin, lmax = _check_avl(n.left)
    rh, rm, rmin, rmax = _check_avl(n.right)
    assert abs(lh - rh) <= 1 and n.height == 1 + max(lh, rh)
    assert n.max == max(n.end, lm, rm)
    if lmax is not None:
        assert lmax <= n.start            # left: start <= node.start
    if rmin is not None:
        assert rmin >= n.start            # right: start > node.start when inserted; rotations can move an equal start here
    lo = lmin if lmin is not None else n.start
    hi = rmax if rmax is not None else n.start
    return n.height, n.max, lo, hi


def test_bio_tree_restatement_properties():
    """The restated bio IntervalTree: AVL invariants hold after every insert, find() yields
    exactly the overlapping intervals (compared with the brute-force scan) and yields them in the order
    node / right subtree / left subtree."""
    from hypothesis import given, settings, strategies as st

    ivs_st = st.lists(st.tuples(st.integers(0, 300), st.integers(0, 60)), min_size=0, max_size=60)

    @settings(max_examples=150, deadline=None)
    @given(ivs_st, st.lists(st.tuples(st.integers(0, 320), st.integers(0, 50)), min_size=1, max_size=10))
    def prop(raw, queries):
        ivs = [(s, s + l, i, 1) for i, (s, l) in enumerate(raw)]
        t = py_oracle.BioTree()
        for s, e, g, sd in ivs:
            t.insert(s, e, (g, sd))
            _check_avl(t.root)
        order = []

        def walk(n):
            if n is None:
                return
            order.append(n.value[0])
            walk(n.right)
            walk(n.left)

        walk(t.root)
        rank = {g: k for k, g in enumerate(order)}
        assert sorted(order) == list(range(len(ivs)))
        for q0, ql in queries:
            q1 = q0 + ql
            got = [v[0] for _, _, v in t.find(q0, q1)]
            want = [iv[2] for iv in py_oracle.find(ivs, q0, q1)]
            assert sorted(got) == sorted(want)
            assert got == sorted(got, key=lambda g: rank[g])  # pruning never reorders

    prop()


def test_bio_tree_known_shapes():
    """Hand-checked AVL shapes: ascending inserts rotate left, equal starts go left."""
    t = py_oracle.BioTree()
    for i, s in enumerate([10, 20, 30]):
        t.insert(s, s + 5, (i, 1))
    assert t.root.value[0] == 1 and t.root.left.value[0] == 0 and t.root.right.value[0] == 2
    assert [v[0] for _, _, v in t.find(0, 100)] == [1, 2, 0]     # node, right, left
    t = py_oracle.BioTree()
    for i in range(3):
        t.insert(10, 15 + i, (i, 1))                              # equal starts descend left: 0 <- 1 <- 2, then rotate right
    assert t.root.value[0] == 1 and t.root.left.value[0] == 2 and t.root.right.value[0] == 0
    assert [v[0] for _, _, v in t.find(12, 13)] == [1, 0, 2]


def _stacked_genes(n_contigs=2):
    """Gene layouts where SEVERAL genes cover a 1 Mb boundary, with ends 1 bp apart: the chunk table depends on
    which one bio's tree yields first."""
    import random

    rng = random.Random(77)
    giv = {}
    contigs = []
    for c in range(n_contigs):
        name = "chr%s" % "ABCD"[c]
        contigs.append((name, 4_200_000 + c))
        lst = []
        for k in range(40):
            b = rng.choice([1_000_000, 2_000_000, 3_000_000])
            s = b - rng.randrange(1, 4000)
            e = b + rng.choice([1, 2, 3, 5, 100, 101, 102, 2000, 2001])
            lst.append(("g%d_%d" % (c, k), rng.choice([1, -1]), [s], [e]))
        for k in range(20):
            s = rng.randrange(0, 4_000_000)
            lst.append(("h%d_%d" % (c, k), 1, [s], [s + rng.randrange(10, 30000)]))
        rng.shuffle(lst)
        giv[name] = lst
    return contigs, giv


def test_chunk_edges_follow_tree_order(built, tmp_path):
    """chunked_genome.rs:98: the edge moves past the FIRST covering interval in bio's order.  Product host code,
    C oracle and the pure-Python tree must agree on layouts where that choice matters."""
    contigs, giv = _stacked_genes()
    path = str(tmp_path / "edges.bam")
    bamio.write_bam(path, contigs, [bamio.Rec(0, 10, b"r", 0, 255, [(0, 10)], 10)])
    want = py_oracle.make_chunks([(n, i, l) for i, (n, l) in enumerate(contigs)],
                                 {n: py_oracle.build_intervals(giv[n])[0] for n, _ in contigs})
    want = [(t, s, e) for _, t, s, e in want]
    assert c_oracle.chunks(path, None, giv) == want
    assert [tuple(x) for x in built.Reader(path).chunks(giv)] == want
    # the choice does matter here: "largest end" gives another table
    greedy = []
    for n, tid, length in [(n, i, l) for i, (n, l) in enumerate(contigs)]:
        ivs = py_oracle.build_intervals(giv[n])[0]
        start = 0
        while True:
            stop = start + py_oracle.CHUNK_SIZE
            while True:
                cov = [iv for iv in ivs if iv[0] <= stop < iv[1]]
                if not cov:
                    break
                stop = max(iv[1] for iv in cov) + 1
            greedy.append((tid, start, stop))
            start = stop
            if start >= length:
                break
    assert greedy != want


@pytest.mark.parametrize("seed,kw", [(31, {}), (32, dict(exon_mode=False, genes_per_contig=120)), (33, dict(paired=True, genes_per_contig=200))])
def test_each_read_counts_once_oracles_agree(tmp_path, seed, kw):
    """each_read_counts_once=True (counters.rs:83-92,173-182): the C oracle (array AVL) against the pure-Python
    tree; overlapping genes make the first hit differ from 'any hit'."""
    path, contigs, recs, iv, giv = make_case(tmp_path, seed, n_reads=4000, **kw)
    a = c_oracle.count_reads_unstranded(path, None, iv, giv, True, nthreads=2)
    assert a == py_oracle.count_reads_unstranded(path, None, iv, giv, True)
    f = c_oracle.count_reads_stranded(path, None, iv, giv, True, nthreads=2)
    assert f == py_oracle.count_reads_stranded(path, None, iv, giv, True)
    b = c_oracle.count_reads_unstranded(path, None, iv, giv, False)
    assert a["_outside"] == b["_outside"] and a["_total"] <= b["_total"]
    if kw:
        assert a != b  # with dense overlapping genes the flag changes results


def test_property_c_oracle_vs_py_oracle_and_host_chunk_table(built, tmp_path):
    """Random gene sets and reads (hypothesis): the C restatement (tree walk, BAI seeks, threads) against the brute-force
    Python one, and the product's host-side chunk table against both -- every count, both `each_read_counts_once`
    settings, introns, chunk edges."""
    from hypothesis import HealthCheck, given, settings, strategies as st
    from synth import small

    @settings(max_examples=20, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.integers(0, 10 ** 6), st.integers(1, 3), st.integers(1, 80), st.booleans(), st.booleans())
    def prop(seed, n_contigs, genes, exon_mode, paired):
        contigs, recs, iv, giv = small.simple_case(seed, n_reads=500, n_contigs=n_contigs, contig_len=1_200_000,
                                                   genes_per_contig=genes, exon_mode=exon_mode, paired=paired)
        path = str(tmp_path / ("p%d.bam" % seed))
        bamio.write_bam(path, contigs, recs, level=1)
        for once in (False, True):
            assert c_oracle.count_reads_unstranded(path, None, iv, giv, once, nthreads=2) == \
                py_oracle.count_reads_unstranded(path, None, iv, giv, once)
            assert c_oracle.count_reads_stranded(path, None, iv, giv, once, nthreads=2) == \
                py_oracle.count_reads_stranded(path, None, iv, giv, once)
        assert c_oracle.count_introns(path, nthreads=2) == py_oracle.count_introns(path)
        name2tid = {n: i for i, (n, _) in enumerate(contigs)}
        want = [(t, s, e) for _, t, s, e in py_oracle.make_chunks(
            [(c, name2tid[c], dict(contigs)[c]) for c in giv], {c: py_oracle.build_intervals(giv[c])[0] for c in giv})]
        assert c_oracle.chunks(path, None, giv) == want
        assert built.Reader(path).chunks(giv) == want
        os.unlink(path)
        os.unlink(path + ".bai")

    prop()
