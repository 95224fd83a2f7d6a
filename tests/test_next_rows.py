"""quantify_gene_reads, calculate_duplicate_distribution and
count_reads_primary_only_right_strand_only_by_barcode.  CPU part: the Python restatement
against hand-computed answers (the reference holds no test for any of them: parity unpinned); GPU part: the
device path, through the C ABI, against the restatement."""
import os
import random

import pytest

from conftest import make_case
from synth import bamio
from synth.bamio import Rec, aux_int
from oracle import c_oracle, py_oracle


def pack_seq(bases):
    code = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15}
    out = bytearray((len(bases) + 1) // 2)
    for i, b in enumerate(bases):
        out[i >> 1] |= code[b] << (4 if i % 2 == 0 else 0)
    return bytes(out)


def with_seq(rec, bases, low_nibble_junk=0):
    rec.l_seq = len(bases)
    s = bytearray(pack_seq(bases))
    if len(bases) & 1:
        s[-1] |= low_nibble_junk & 0xF
    rec.seq = bytes(s)
    rec.qual = b"\x25" * len(bases)
    return rec


def add_duplicates(rng, recs, contigs, n_groups=300):
    """groups of records sharing (tid, pos): same/different SEQ, both strands, odd lengths with junk in the spare nibble"""
    out = list(recs)
    for _ in range(n_groups):
        tid = rng.randrange(len(contigs))
        pos = rng.randrange(0, contigs[tid][1] - 200)
        n_seq = rng.randint(1, 3)
        for s in range(n_seq):
            ln = rng.choice([25, 36, 50, 51, 1, 7])
            bases = "".join(rng.choice("ACGT") for _ in range(ln))
            for strand in rng.sample([0, 16], rng.randint(1, 2)):
                for c in range(rng.choice([1, 1, 2, 3, 5, 12])):
                    r = Rec(tid, pos, b"dup%d" % len(out), strand | rng.choice([0, 0, 0x400, 0x100]), 255, [(0, ln)], ln,
                            aux_int(b"NH", 1))
                    out.append(with_seq(r, bases, rng.randrange(16)))
    out.sort(key=lambda r: (r.tid if r.tid >= 0 else 1 << 30, r.pos))
    return out


def dup_case(tmp_path, seed, n_reads=3000, **kw):
    from synth import small

    write_kw = {k: kw.pop(k) for k in list(kw) if k in ("level", "split_records", "payload_limit")}
    rng = random.Random(seed)
    contigs, recs, iv, giv = small.simple_case(seed, n_reads=n_reads, random_seq=True, **kw)
    recs = add_duplicates(rng, recs, contigs)
    path = os.path.join(str(tmp_path), "dup_%d.bam" % seed)
    bamio.write_bam(path, contigs, recs, **write_kw)
    return path, contigs, recs, iv, giv


# ------------------------------------------------------------------------------------------------ CPU
def test_quantify_hand_case(tmp_path):
    """reference src/count_reads/quantify.rs:98-146 by hand."""
    contigs = [("c1", 3_000_000), ("c2", 10_000)]
    iv = {"c1": [("gA", 1, [100, 300], [200, 400]), ("gB", -1, [150], [350]), ("gA", 1, [5000], [5100])],
          "c2": [("gC", 0, [0], [5000])]}
    giv = {"c1": [("gA", 1, [100], [400]), ("gB", -1, [150], [350]), ("gA2", 1, [5000], [5100])], "c2": [("gC", 0, [0], [5000])]}
    recs = [
        Rec(0, 120, b"a", 0, 255, [(0, 50)], 50),          # gA fwd (congruent), gB (strand -1, forward read -> incongruent)
        Rec(0, 120, b"b", 0, 255, [(0, 50)], 50),
        Rec(0, 120, b"c", 16, 255, [(0, 50)], 50),         # reverse: gA incongruent, gB congruent
        Rec(0, 190, b"d", 0, 3, [(0, 5), (3, 110), (0, 20)], 25, aux_int(b"NH", 5)),  # two blocks in gA's two exons: once; NH ignored
        Rec(0, 5050, b"e", 0x100, 255, [(0, 10)], 10),     # secondary: counted; third tuple (duplicate id gA)
        Rec(0, 999_990, b"f", 0, 255, [(0, 50)], 50),      # nothing
        Rec(1, 10, b"g", 16, 255, [(0, 10)], 10),          # strand 0 behaves like -1: reverse read congruent
        Rec(1, 10, b"h", 0, 255, [(0, 10)], 10),
    ]
    p = str(tmp_path / "q.bam")
    bamio.write_bam(p, contigs, recs)
    fwd, rev = py_oracle.quantify_gene_reads(p, None, iv, giv)
    # the duplicated id gA keeps the LAST gene_no of its contig (index 2), per chunk
    assert fwd == {"gA": [(5050, 1)], "gB": [(120, 1)], "gC": [(10, 1)]}
    assert rev == {"gA": [], "gB": [(120, 2), (190, 1)], "gC": [(10, 1)]}
    iv["c1"][2] = ("gA3", 1, [5000], [5100])
    fwd, rev = py_oracle.quantify_gene_reads(p, None, iv, giv)
    assert fwd == {"gA": [(120, 2), (190, 1)], "gA3": [(5050, 1)], "gB": [(120, 1)], "gC": [(10, 1)]}
    assert rev == {"gA": [(120, 1)], "gA3": [], "gB": [(120, 2), (190, 1)], "gC": [(10, 1)]}
    # only contigs of gene_intervals are scanned
    fwd, rev = py_oracle.quantify_gene_reads(p, None, iv, {"c2": giv["c2"]})
    assert fwd == {"gC": [(10, 1)]} and rev == {"gC": [(10, 1)]}


def test_duplicate_distribution_hand_case(tmp_path):
    """reference src/duplicate_distribution.rs:44-86 by hand."""
    contigs = [("c1", 1000), ("c2", 1000)]
    mk = lambda tid, pos, flag, bases, junk=0: with_seq(Rec(tid, pos, b"x", flag, 255, [(0, len(bases))], len(bases)), bases, junk)
    recs = [
        mk(0, 10, 0, "ACGT"), mk(0, 10, 0, "ACGT"), mk(0, 10, 0, "ACGT"),   # one key, 3 copies
        mk(0, 10, 16, "ACGT"),                                              # other strand: own key
        mk(0, 10, 0, "ACGA"),                                               # other sequence
        mk(0, 11, 0, "ACGT"),                                               # other position
        mk(0, 20, 0, "ACG", 0), mk(0, 20, 0, "ACG", 9),                     # odd length: spare nibble is not part of the key
        mk(0, 20, 0, "ACGA"),                                               # ... but a real fourth base is
        mk(0, 999, 0, "AC"), mk(0, 1000, 0, "AC"),                          # pos == target_len is outside fetch(tid, 0, len)
        mk(1, 10, 0, "ACGT"), mk(1, 10, 0, "ACGT"),                         # same (pos, seq) on another contig: own key
        Rec(-1, -1, b"u", 4, 0, [], 0),
    ]
    p = str(tmp_path / "d.bam")
    bamio.write_bam(p, contigs, recs)
    assert py_oracle.calculate_duplicate_distribution(p) == {3: 1, 1: 5, 2: 2}


@pytest.mark.parametrize("seed,kw", [(41, {}), (42, dict(exon_mode=False, genes_per_contig=150)), (43, dict(paired=True, n_contigs=4)),
                                     (44, dict(split_records=True, level=1))])
def test_c_oracle_matches_py_oracle_quantify_and_duplicates(tmp_path, seed, kw):
    """the C restatement (tree walk, BAI seeks, threads) against the brute-force Python one"""
    path, contigs, recs, iv, giv = dup_case(tmp_path, seed, n_reads=4000, **kw)
    assert c_oracle.quantify_gene_reads(path, None, iv, giv, nthreads=3) == py_oracle.quantify_gene_reads(path, None, iv, giv)
    assert c_oracle.calculate_duplicate_distribution(path, nthreads=3) == py_oracle.calculate_duplicate_distribution(path)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture()
def mb(built, has_gpu):
    assert has_gpu, "GPU tests need a CUDA device (no CPU fallback exists)"
    return built


@pytest.mark.gpu
@pytest.mark.parametrize("seed,kw", [(41, {}), (42, dict(exon_mode=False, genes_per_contig=150)), (43, dict(paired=True, n_contigs=4)),
                                     (44, dict(split_records=True, level=1))])
def test_quantify_matches_oracle(mb, tmp_path, seed, kw):
    path, contigs, recs, iv, giv = dup_case(tmp_path, seed, n_reads=6000, **kw)
    assert mb.quantify_gene_reads(path, None, iv, giv) == py_oracle.quantify_gene_reads(path, None, iv, giv)
    st = mb.last_stats()
    assert st["n_records"] >= len(recs) - 1 and st["n_kernel_launches"] > 0


@pytest.mark.gpu
def test_quantify_hand_case_on_device(mb, tmp_path):
    contigs = [("c1", 3_000_000), ("c2", 10_000)]
    iv = {"c1": [("gA", 1, [100, 300], [200, 400]), ("gB", -1, [150], [350]), ("gA", 1, [5000], [5100])],
          "c2": [("gC", 0, [0], [5000])]}
    giv = {"c1": [("gA", 1, [100], [400])], "c2": [("gC", 0, [0], [5000])]}
    recs = [Rec(0, 120, b"a", 0, 255, [(0, 50)], 50), Rec(0, 120, b"c", 16, 255, [(0, 50)], 50),
            Rec(0, 190, b"d", 0, 3, [(0, 5), (3, 110), (0, 20)], 25, aux_int(b"NH", 5)),
            Rec(0, 5050, b"e", 0x100, 255, [(0, 10)], 10), Rec(1, 10, b"g", 16, 255, [(0, 10)], 10)]
    p = str(tmp_path / "q.bam")
    bamio.write_bam(p, contigs, recs)
    want = py_oracle.quantify_gene_reads(p, None, iv, giv)
    assert want[0] == {"gA": [(5050, 1)], "gB": [(120, 1)], "gC": [(10, 1)]}
    assert mb.quantify_gene_reads(p, None, iv, giv) == want
    # equal ids on two contigs: merged; a shared position is the reference's panic
    iv2 = {"c1": [("g", 1, [100], [200])], "c2": [("g", 1, [0], [5000])]}
    giv2 = iv2
    assert mb.quantify_gene_reads(p, None, iv2, giv2) == py_oracle.quantify_gene_reads(p, None, iv2, giv2)
    recs.append(Rec(1, 120, b"z", 0, 255, [(0, 10)], 10))
    bamio.write_bam(p, contigs, recs)
    with pytest.raises(ValueError, match="still not disjoint"):
        py_oracle.quantify_gene_reads(p, None, iv2, giv2)
    with pytest.raises(ValueError, match="still not disjoint"):
        mb.quantify_gene_reads(p, None, iv2, giv2)


@pytest.mark.gpu
def test_quantify_many_windows_and_shards(mb, tmp_path, monkeypatch):
    path, contigs, recs, iv, giv = dup_case(tmp_path, 45, n_reads=60000, n_contigs=3)
    want = py_oracle.quantify_gene_reads(path, None, iv, giv)
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "1")
    mb.release_device(0)
    try:
        assert mb.quantify_gene_reads(path, None, iv, giv) == want
        assert mb.last_stats()["n_windows"] > 2
        # chunk-range shards: positions are owned by one shard, so the per-gene lists concatenate
        merged = ({}, {})
        for s in range(3):
            rd = mb.Reader(path)
            rd.set_device(0)
            rd.set_shard(s, 3)
            part = rd.quantify_gene_reads(iv, giv)
            for b in (0, 1):
                for g, lst in part[b].items():
                    merged[b].setdefault(g, []).extend(lst)
        assert tuple({g: sorted(l) for g, l in m.items()} for m in merged) == want
    finally:
        monkeypatch.delenv("MBF_BAM_WINDOW_MB")
        mb.release_device(0)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,kw", [(51, {}), (52, dict(split_records=True, n_contigs=3)), (53, dict(payload_limit=3000))])
def test_duplicate_distribution_matches_oracle(mb, tmp_path, seed, kw):
    path, contigs, recs, iv, giv = dup_case(tmp_path, seed, n_reads=5000, **kw)
    want = py_oracle.calculate_duplicate_distribution(path)
    assert len(want) >= 4
    assert mb.calculate_duplicate_distribution(path) == want
    assert sum(k * v for k, v in want.items()) == mb.last_stats()["n_records_owned"]


@pytest.mark.gpu
def test_duplicate_distribution_hand_case_windows_and_shards(mb, tmp_path, monkeypatch):
    contigs = [("c1", 1000), ("c2", 1000)]
    mk = lambda tid, pos, flag, bases, junk=0: with_seq(Rec(tid, pos, b"x", flag, 255, [(0, len(bases))], len(bases)), bases, junk)
    recs = [mk(0, 10, 0, "ACGT"), mk(0, 10, 0, "ACGT"), mk(0, 10, 0, "ACGT"), mk(0, 10, 16, "ACGT"), mk(0, 10, 0, "ACGA"),
            mk(0, 11, 0, "ACGT"), mk(0, 20, 0, "ACG", 0), mk(0, 20, 0, "ACG", 9), mk(0, 20, 0, "ACGA"), mk(0, 999, 0, "AC"),
            mk(0, 1000, 0, "AC"), mk(1, 10, 0, "ACGT"), mk(1, 10, 0, "ACGT"), Rec(-1, -1, b"u", 4, 0, [], 0)]
    p = str(tmp_path / "d.bam")
    bamio.write_bam(p, contigs, recs)
    assert mb.calculate_duplicate_distribution(p) == {3: 1, 1: 5, 2: 2}
    path, contigs, recs, iv, giv = dup_case(tmp_path, 54, n_reads=50000, n_contigs=3)
    want = py_oracle.calculate_duplicate_distribution(path)
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "1")
    mb.release_device(0)
    try:
        assert mb.calculate_duplicate_distribution(path) == want
        assert mb.last_stats()["n_windows"] > 2
        total = {}
        for s in range(4):   # groups never straddle a chunk: shard histograms add
            rd = mb.Reader(path)
            rd.set_device(0)
            rd.set_shard(s, 4)
            for k, v in rd.calculate_duplicate_distribution().items():
                total[k] = total.get(k, 0) + v
        assert total == want
    finally:
        monkeypatch.delenv("MBF_BAM_WINDOW_MB")
        mb.release_device(0)


# ------------------------------------------------------------------------------------------------ by barcode
from synth.bamio import aux_str  # noqa: E402

BY_BARCODE = "count_reads_primary_only_right_strand_only_by_barcode"


def barcode_case(tmp_path, seed, n_reads=4000, bad_chunks=True, **kw):
    """single-cell-like tags: XC from a small set of barcodes, XM UMIs from a tiny alphabet so that (pos, UMI) repeat"""
    from synth import small

    write_kw = {k: kw.pop(k) for k in list(kw) if k in ("level", "split_records", "payload_limit")}
    rng = random.Random(seed)
    contigs, recs, iv, giv = small.simple_case(seed, n_reads=n_reads, **kw)
    barcodes = [("".join(rng.choice("ACGT") for _ in range(rng.choice([8, 12, 12, 16])))).encode() for _ in range(25)]
    out = []
    for r in recs:
        tags = aux_str(b"XC", rng.choice(barcodes)) + aux_str(b"XM", "".join(rng.choice("AC") for _ in range(3)).encode())
        if rng.random() < 0.5:
            r.aux = tags + r.aux
        else:
            r.aux = r.aux + tags
        out.append(r)
        if rng.random() < 0.3:   # PCR duplicates and near-duplicates at the same position
            for _ in range(rng.randint(1, 3)):
                d = Rec(r.tid, r.pos, r.qname + b"d", r.flag ^ rng.choice([0, 0, 16]), r.mapq, list(r.cigar), r.l_seq,
                        aux_str(b"XM", "".join(rng.choice("AC") for _ in range(3)).encode()) + aux_str(b"XC", rng.choice(barcodes[:4])))
                out.append(d)
    if bad_chunks:
        # chunks that the reference drops: a right-strand hit without XC / with an integer XC / without XM
        for tid, (name, _) in enumerate(contigs):
            g = iv[name][3]
            if g[1] == 1:
                flag = 0
            else:
                flag = 16
            pos = g[2][0]
            if tid == 0:
                out.append(Rec(tid, pos, b"noxc", flag, 255, [(0, 30)], 30, aux_str(b"XM", b"AAA")))
            elif tid == 1:
                out.append(Rec(tid, pos, b"intxc", flag, 255, [(0, 30)], 30, aux_int(b"XC", 7) + aux_str(b"XM", b"AAA")))
            else:
                out.append(Rec(tid, pos, b"noxm", flag, 255, [(0, 30)], 30, aux_str(b"XC", b"ACGTACGT")))
    # a read without tags that hits nothing on its strand is harmless
    out.append(Rec(0, 7, b"plain", 0, 255, [(0, 3)], 3, b""))
    out.sort(key=lambda r: (r.tid if r.tid >= 0 else 1 << 30, r.pos))
    path = os.path.join(str(tmp_path), "bc_%d.bam" % seed)
    bamio.write_bam(path, contigs, out, **write_kw)
    return path, contigs, out, iv, giv


def test_by_barcode_hand_case(tmp_path):
    """reference src/count_reads/by_barcode.rs:101-182 by hand."""
    contigs = [("c1", 2_500_000)]
    iv = {"c1": [("gP", 1, [100], [200]), ("gM", -1, [150], [400]), ("gFar", 1, [1_200_000], [1_200_100])]}
    giv = iv
    tag = lambda bc, umi: aux_str(b"XC", bc) + aux_str(b"XM", umi)
    recs = [
        Rec(0, 120, b"a", 0, 255, [(0, 50)], 50, tag(b"AAAA", b"u1")),      # forward: gP (+) counts, gM (-) is the wrong strand
        Rec(0, 120, b"b", 0, 255, [(0, 50)], 50, tag(b"AAAA", b"u1")),      # same (gene, barcode, pos, strand, UMI): once
        Rec(0, 120, b"c", 0, 255, [(0, 50)], 50, tag(b"AAAA", b"u2")),      # second UMI
        Rec(0, 121, b"d", 0, 255, [(0, 50)], 50, tag(b"AAAA", b"u1")),      # other position: counts again
        Rec(0, 121, b"e", 0, 255, [(0, 50)], 50, tag(b"CCCC", b"u1")),      # other barcode
        Rec(0, 121, b"f", 0x100, 255, [(0, 50)], 50, tag(b"CCCC", b"u9")),  # secondary: ignored
        Rec(0, 160, b"g", 16, 255, [(0, 50)], 50, tag(b"AAAA", b"u1")),     # reverse: gM only
        Rec(0, 160, b"h", 16, 255, [(0, 10), (3, 100), (0, 10)], 20, tag(b"AAAA", b"u1")),  # two blocks in gM: one set entry
        Rec(0, 500, b"i", 0, 255, [(0, 50)], 50, b""),                      # no tags, no hit: harmless
        Rec(0, 1_200_010, b"j", 0, 255, [(0, 50)], 50, tag(b"GGGG", b"u1")),  # second chunk
        Rec(0, 1_200_020, b"k", 0, 255, [(0, 50)], 50, aux_str(b"XC", b"GGGG")),  # ... which fails: XM missing -> nothing from chunk 2
    ]
    p = str(tmp_path / "b.bam")
    bamio.write_bam(p, contigs, recs)
    genes, barcodes, counts = getattr(py_oracle, BY_BARCODE)(p, None, iv, giv, "straight")
    got = sorted((genes[g], barcodes[b], n) for g, b, n in counts)
    assert got == [("gM", "AAAA", 1), ("gP", "AAAA", 3), ("gP", "CCCC", 1)]
    with pytest.raises(ValueError):
        getattr(py_oracle, BY_BARCODE)(p, None, iv, giv, "hamming")


@pytest.mark.parametrize("seed,kw", [(61, {}), (62, dict(exon_mode=False, genes_per_contig=150, n_contigs=3)),
                                     (63, dict(paired=True, split_records=True)), (64, dict(bad_chunks=False, level=1))])
def test_c_oracle_matches_py_oracle_by_barcode(tmp_path, seed, kw):
    path, contigs, recs, iv, giv = barcode_case(tmp_path, seed, n_reads=4000, **kw)
    assert getattr(c_oracle, BY_BARCODE)(path, None, iv, giv, "straight", nthreads=3) == \
        getattr(py_oracle, BY_BARCODE)(path, None, iv, giv, "straight")


@pytest.mark.gpu
def test_by_barcode_hand_case_on_device(mb, tmp_path):
    contigs = [("c1", 2_500_000)]
    iv = {"c1": [("gP", 1, [100], [200]), ("gM", -1, [150], [400]), ("gFar", 1, [1_200_000], [1_200_100])]}
    tag = lambda bc, umi: aux_str(b"XC", bc) + aux_str(b"XM", umi)
    recs = [Rec(0, 120, b"a", 0, 255, [(0, 50)], 50, tag(b"AAAA", b"u1")), Rec(0, 120, b"b", 0, 255, [(0, 50)], 50, tag(b"AAAA", b"u1")),
            Rec(0, 120, b"c", 0, 255, [(0, 50)], 50, tag(b"AAAA", b"u2")), Rec(0, 121, b"d", 0, 255, [(0, 50)], 50, tag(b"AAAA", b"u1")),
            Rec(0, 121, b"e", 0, 255, [(0, 50)], 50, tag(b"CCCC", b"u1")), Rec(0, 121, b"f", 0x100, 255, [(0, 50)], 50, tag(b"CCCC", b"u9")),
            Rec(0, 160, b"g", 16, 255, [(0, 50)], 50, tag(b"AAAA", b"u1")),
            Rec(0, 160, b"h", 16, 255, [(0, 10), (3, 100), (0, 10)], 20, tag(b"AAAA", b"u1")),
            Rec(0, 500, b"i", 0, 255, [(0, 50)], 50, b""), Rec(0, 1_200_010, b"j", 0, 255, [(0, 50)], 50, tag(b"GGGG", b"u1")),
            Rec(0, 1_200_020, b"k", 0, 255, [(0, 50)], 50, aux_str(b"XC", b"GGGG"))]
    p = str(tmp_path / "b.bam")
    bamio.write_bam(p, contigs, recs)
    fn = getattr(mb, BY_BARCODE)
    genes, barcodes, counts = fn(p, None, iv, iv, "straight")
    assert sorted((genes[g], barcodes[b], n) for g, b, n in counts) == [("gM", "AAAA", 1), ("gP", "AAAA", 3), ("gP", "CCCC", 1)]
    assert (genes, barcodes, counts) == getattr(py_oracle, BY_BARCODE)(p, None, iv, iv, "straight")
    with pytest.raises(ValueError, match="invalid umi_strategy"):
        fn(p, None, iv, iv, "hamming")
    # a barcode longer than the device keeps is an error, not a silent truncation
    recs[0] = Rec(0, 120, b"a", 0, 255, [(0, 50)], 50, tag(b"A" * 60, b"u1"))
    bamio.write_bam(p, contigs, recs)
    with pytest.raises(ValueError, match="barcode"):
        fn(p, None, iv, iv, "straight")


@pytest.mark.gpu
@pytest.mark.parametrize("seed,kw", [(61, {}), (62, dict(exon_mode=False, genes_per_contig=150, n_contigs=3)),
                                     (63, dict(paired=True, split_records=True)), (64, dict(bad_chunks=False, level=1))])
def test_by_barcode_matches_oracle(mb, tmp_path, seed, kw):
    path, contigs, recs, iv, giv = barcode_case(tmp_path, seed, n_reads=6000, **kw)
    want = getattr(py_oracle, BY_BARCODE)(path, None, iv, giv, "straight")
    assert len(want[2]) > 50
    assert getattr(mb, BY_BARCODE)(path, None, iv, giv, "straight") == want


@pytest.mark.gpu
def test_by_barcode_many_windows_and_shards(mb, tmp_path, monkeypatch):
    path, contigs, recs, iv, giv = barcode_case(tmp_path, 65, n_reads=50000, n_contigs=3, random_seq=True)
    want = getattr(py_oracle, BY_BARCODE)(path, None, iv, giv, "straight")
    as_set = lambda res: sorted((res[0][g], res[1][b], n) for g, b, n in res[2])
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "1")
    mb.release_device(0)
    try:
        assert getattr(mb, BY_BARCODE)(path, None, iv, giv, "straight") == want
        assert mb.last_stats()["n_windows"] > 2
        parts = []
        for s in range(3):   # chunk-range shards: entries are per chunk, so the shard lists concatenate
            rd = mb.Reader(path)
            rd.set_device(0)
            rd.set_shard(s, 3)
            parts += as_set(getattr(rd, BY_BARCODE)(iv, giv, "straight"))
        assert sorted(parts) == as_set(want)
    finally:
        monkeypatch.delenv("MBF_BAM_WINDOW_MB")
        mb.release_device(0)


@pytest.mark.gpu
def test_quantify_and_duplicates_on_a_bench_shaped_file(mb, tmp_path):
    """1 M reads of the configs[2] shape (25 human-like contigs, 20 k genes, spliced reads, multi-mappers) against the C
    restatement: every (gene, position) list and the whole duplicate histogram"""
    from synth import model

    path = str(tmp_path / "cfg3.bam")
    iv, giv, info = model.build_config("cfg3", path, reads=1_000_000)
    got = mb.quantify_gene_reads(path, None, iv, giv)
    want = c_oracle.quantify_gene_reads(path, None, iv, giv, nthreads=8)
    assert got == want
    assert sum(len(v) for v in want[0].values()) > 100_000
    assert mb.calculate_duplicate_distribution(path) == c_oracle.calculate_duplicate_distribution(path, nthreads=8)
