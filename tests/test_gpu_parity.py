"""GPU suite: every stage and every entry point against the oracle, through the C ABI (the host
extension only converts arguments).  Integer results bit-exact; weighted within 1e-9 relative."""
import json
import os
import random
import zlib

import pytest

from conftest import ROOT, make_case
from synth import bamio
from oracle import c_oracle, py_oracle

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "cigar_golden.json")))


@pytest.fixture()
def mb(built, has_gpu):
    assert has_gpu, "GPU tests need a CUDA device (no CPU fallback exists)"
    return built


def assert_weighted_close(a, b):
    for x, y in zip(a, b):
        assert set(x) == set(y)
        for k in x:
            assert abs(x[k] - y[k]) <= 1e-9 * max(1.0, abs(y[k])), (k, x[k], y[k])


def check_all(mb, path, iv, giv):
    assert mb.count_reads_unstranded(path, None, iv, giv) == c_oracle.count_reads_unstranded(path, None, iv, giv, nthreads=4)
    st = mb.last_stats()
    assert mb.count_reads_stranded(path, None, iv, giv) == c_oracle.count_reads_stranded(path, None, iv, giv, nthreads=4)
    assert_weighted_close(mb.count_reads_weighted(path, None, iv, giv), c_oracle.count_reads_weighted(path, None, iv, giv, nthreads=4))
    assert mb.count_introns(path) == c_oracle.count_introns(path, nthreads=4)
    return st


# ------------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("kw", [dict(), dict(level=1), dict(level=9, split_records=True), dict(sync_every=500),
                                dict(strategy=zlib.Z_FIXED), dict(level=0), dict(strategy=zlib.Z_HUFFMAN_ONLY),
                                dict(strategy=zlib.Z_RLE, payload_limit=3000)],
                         ids=["default", "l1", "l9split", "syncflush", "fixed", "stored", "huffonly", "rle_small"])
def test_inflate_matches_zlib(mb, tmp_path, kw):
    path, *_ = make_case(tmp_path, 12, n_reads=6000, **kw)
    data = open(path, "rb").read()
    want = b"".join(p for _, _, p in bamio.iter_bgzf_blocks(data))
    got = mb.inflate_buffer(data)
    assert len(got) == len(want)
    assert got == want
    assert mb.last_stats()["n_bgzf_blocks"] == len(list(bamio.iter_bgzf_blocks(data)))


def test_inflate_random_and_text(mb, tmp_path):
    rng = random.Random(5)
    p = tmp_path / "r.bgzf"
    payload = bytes(rng.getrandbits(8) for _ in range(300_000)) + b"abc" * 50_000 + b"\0" * 200_000 + \
        b"".join(b"%d," % rng.randrange(10 ** 6) for _ in range(100_000))
    with open(p, "wb") as fh:
        w = bamio.BgzfWriter(fh)
        w.write(payload, atomic=False)
        w.close()
    assert mb.inflate_buffer(p.read_bytes()) == payload


def test_inflate_long_matches_near_far_and_overlapping(mb, tmp_path, monkeypatch):
    """Long matches (up to 258 bytes) whose source is in the output ring, far behind it (already flushed), or
    overlaps the destination (run-length style): the three routes of the LZ77 pass, for every inflate kernel."""
    rng = random.Random(19)
    a = bytes(rng.getrandbits(8) for _ in range(5000))
    b = bytes(rng.getrandbits(8) for _ in range(700))
    parts = [a, b, a[:3000], bytes(rng.getrandbits(8) for _ in range(40)), b * 3, a[1000:4000],
             b"xy" * 400, b"q" * 700, a[::-1][:2500], a[200:4900]]
    for _ in range(200):
        o = rng.randrange(0, 4500)
        parts.append(a[o:o + rng.randrange(3, 400)])
        parts.append(bytes(rng.getrandbits(8) for _ in range(rng.randrange(0, 6))))
    payload = b"".join(parts)
    assert 40_000 < len(payload) < 65_000  # a single BGZF block: distances up to 32 KiB stay possible
    p = tmp_path / "m.bgzf"
    with open(p, "wb") as fh:
        w = bamio.BgzfWriter(fh, level=9)
        w.write(payload, atomic=False)
        w.close()
    data = p.read_bytes()
    for d in ("400", "401", "402", "403", "308", "304"):
        monkeypatch.setenv("MBF_BAM_INFLATE_D", d)
        assert mb.inflate_buffer(data) == payload, d


def test_inflate_many_windows(mb, tmp_path, monkeypatch):
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "1")
    rng = random.Random(6)
    payload = b"".join(b"%x\t" % rng.getrandbits(40) for _ in range(700_000))
    p = tmp_path / "w.bgzf"
    with open(p, "wb") as fh:
        w = bamio.BgzfWriter(fh, level=1)
        w.write(payload, atomic=False)
        w.close()
    data = p.read_bytes()
    assert len(data) > 3 * (1 << 20)
    assert mb.inflate_buffer(data) == payload
    assert mb.last_stats()["n_windows"] >= 3


def test_window_size_follows_inflate_ratio(mb, tmp_path, monkeypatch):
    """Windows are capped so that their inflated size stays below the 32-bit offset limit; with the limit
    lowered to 4 MiB the file must be cut into more windows and still inflate to the same bytes."""
    rng = random.Random(21)
    payload = b"".join(b"%x\t" % rng.getrandbits(24) for _ in range(2_000_000))
    p = tmp_path / "c.bgzf"
    with open(p, "wb") as fh:
        w = bamio.BgzfWriter(fh, level=6)
        w.write(payload, atomic=False)
        w.close()
    data = p.read_bytes()
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "16")
    assert mb.inflate_buffer(data) == payload
    n_free = mb.last_stats()["n_windows"]
    monkeypatch.setenv("MBF_BAM_U_LIMIT_MB", "4")
    assert mb.inflate_buffer(data) == payload
    assert mb.last_stats()["n_windows"] > n_free


def test_inflate_rejects_corruption(mb, tmp_path):
    path, *_ = make_case(tmp_path, 13, n_reads=3000)
    data = bytearray(open(path, "rb").read())
    blocks = list(bamio.iter_bgzf_blocks(bytes(data)))
    off, csize, _ = blocks[1]
    for k in range(40, 60):
        data[off + k] ^= 0x5A
    with pytest.raises(ValueError, match="Could not read bam"):
        mb.inflate_buffer(bytes(data))


def test_read_error_on_the_staging_thread_is_a_value_error(mb, tmp_path, monkeypatch):
    """The bytes of window k+1 are read on a helper thread: a file that shrinks under the reader must surface as
    ValueError from the call, and the reader/pipeline must stay usable afterwards."""
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "4")  # windows of 1, 1.5, 2.25 MiB: all but the first go through the helper thread
    path, contigs, recs, iv, giv = make_case(tmp_path, 23, n_reads=60_000, random_seq=True)
    size = os.path.getsize(path)
    assert size > 3 * (1 << 20)
    want = c_oracle.count_reads_stranded(path, None, iv, giv)
    rd = mb.Reader(path)
    assert tuple(rd.count_reads(1, iv, giv)) == want
    data = open(path, "rb").read()
    with open(path, "r+b") as fh:
        fh.truncate(size // 2)
    with pytest.raises(ValueError, match="Could not read bam"):
        rd.count_reads(1, iv, giv)
    with open(path, "wb") as fh:
        fh.write(data)
    assert tuple(rd.count_reads(1, iv, giv)) == want


# ------------------------------------------------------------------------------------------------ K3
@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: "%d_%s" % (g["index"], g["cigar"]))
def test_cigar_golden_on_device(mb, g):
    """reference src/bam_ext.rs:50-604 golden records through the device CIGAR expander."""
    words = [(l << 4) | op for op, l in bamio.parse_cigar(g["cigar"])]
    assert [tuple(x) for x in mb.cigar_expand(words, g["pos"], False)] == [tuple(b) for b in g["blocks"]]
    got = [tuple(x) for x in mb.cigar_expand(words, g["pos"], True)]
    assert len(got) == g["n_introns"]
    assert got[:len(g["introns"])] == [tuple(b) for b in g["introns"]]


# ------------------------------------------------------------------------------------------------ end to end
@pytest.mark.parametrize("seed,kw", [
    (1, {}), (2, dict(paired=True)), (3, dict(split_records=True, level=9)),
    (4, dict(sync_every=700)), (5, dict(exon_mode=False)), (6, dict(n_contigs=4, contig_len=900_000)),
    (7, dict(n_reads=60_000)), (8, dict(n_reads=20_000, split_records=True, payload_limit=2000)),
])
def test_counts_match_oracle(mb, tmp_path, seed, kw):
    path, contigs, recs, iv, giv = make_case(tmp_path, seed, **kw)
    st = check_all(mb, path, iv, giv)
    assert st["n_kernel_launches"] > 0


def test_counts_many_windows_and_key_overflow(mb, tmp_path, monkeypatch):
    """Small windows (carry between windows, multi-mapper sets spanning windows) and a multi-mapper key
    buffer that overflows (emit-only re-run)."""
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "1")
    monkeypatch.setenv("MBF_BAM_MM_CAP", "64")
    path, contigs, recs, iv, giv = make_case(tmp_path, 9, n_reads=100_000, level=1, random_seq=True)
    assert os.path.getsize(path) > 3 * (1 << 20)
    st = check_all(mb, path, iv, giv)
    assert st["n_windows"] >= 3


def test_split_records_across_windows(mb, tmp_path, monkeypatch):
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "1")
    path, contigs, recs, iv, giv = make_case(tmp_path, 10, n_reads=80_000, level=1, split_records=True, random_seq=True)
    check_all(mb, path, iv, giv)


def test_read_hitting_many_genes(mb, tmp_path):
    """More distinct genes under one read than the per-thread register set holds (SEEN_K)."""
    R = bamio.Rec
    refs = [("c1", 2_000_000), ("c2", 1_500_000)]
    genes = [("g%02d" % i, 1 if i % 2 else -1, [1000 + i, 1200 + i], [1100 + i, 1300 + i]) for i in range(25)]
    iv = {"c1": genes, "c2": [("h", 1, [10], [20])]}
    giv = {"c1": [(g[0], g[1], [g[2][0]], [g[3][-1]]) for g in genes], "c2": [("h", 1, [10], [20])]}
    nh3 = bamio.aux_int(b"NH", 3, "s")
    recs = [R(0, 1005 + i, b"x%d" % (i % 7), 16 * (i % 2), 255, [(0, 60), (3, 150), (0, 40)], 100,
              nh3 if i % 3 == 0 else b"") for i in range(400)]
    recs += [R(1, 12, b"y", 0, 255, [(0, 5)], 5)]
    path = str(tmp_path / "many.bam")
    bamio.write_bam(path, refs, recs)
    check_all(mb, path, iv, giv)


def test_contig_subsets_and_empty(mb, tmp_path):
    path, contigs, recs, iv, giv = make_case(tmp_path, 14, n_contigs=3)
    sub_g = {k: giv[k] for k in list(giv)[1:]}
    sub_g["not_in_header"] = giv[list(giv)[0]]
    sub_i = dict(iv)
    sub_i["not_in_header"] = iv[list(iv)[0]]
    assert mb.count_reads_unstranded(path, None, sub_i, sub_g) == c_oracle.count_reads_unstranded(path, None, sub_i, sub_g)
    assert mb.count_reads_stranded(path, None, sub_i, sub_g) == c_oracle.count_reads_stranded(path, None, sub_i, sub_g)
    assert mb.count_reads_unstranded(path, None, {}, {}) == {}
    assert mb.count_reads_stranded(path, None, {}, {}) == ({}, {})
    with pytest.raises(ValueError, match="gene_intervals but not in intervals"):
        mb.count_reads_unstranded(path, None, {}, sub_g)
    # a BAM without any record
    p2 = str(tmp_path / "empty.bam")
    bamio.write_bam(p2, contigs, [])
    assert mb.count_reads_unstranded(p2, None, iv, giv) == c_oracle.count_reads_unstranded(p2, None, iv, giv)
    assert mb.count_introns(p2) == c_oracle.count_introns(p2)


def test_nh_type_error(mb, tmp_path):
    R = bamio.Rec
    refs = [("c1", 100_000)]
    iv = {"c1": [("g", 1, [10], [90])]}
    path = str(tmp_path / "nhz.bam")
    bamio.write_bam(path, refs, [R(0, 20, b"a", 0, 255, [(0, 10)], 10, bamio.aux_str(b"NH", b"2"))])
    with pytest.raises(ValueError, match="NH tag is not an integer"):
        mb.count_reads_unstranded(path, None, iv, iv)


def test_shards_sum_to_whole(mb, tmp_path):
    """Chunk-range sharding: per-shard dense results add up to the single-shard result."""
    import numpy as np

    path, contigs, recs, iv, giv = make_case(tmp_path, 15, n_reads=40_000, n_contigs=3)
    for mode in (0, 1, 2):
        rd = mb.Reader(path)
        whole = rd.count_reads_raw(mode, iv, giv)
        for n in (2, 3, 5):
            acc = None
            for i in range(n):
                rd.set_shard(i, n)
                part = rd.count_reads_raw(mode, iv, giv)
                arrs = [np.frombuffer(part[0], dtype=np.uint64), np.frombuffer(part[1], dtype=np.uint64),
                        np.frombuffer(part[2], dtype=np.float64), np.frombuffer(part[3], dtype=np.float64),
                        np.frombuffer(part[4], dtype=np.uint8), np.uint64(part[5])]
                acc = arrs if acc is None else [a + b if k != 4 else np.maximum(a, b) for k, (a, b) in enumerate(zip(acc, arrs))]
            rd.set_shard(0, 1)
            assert acc[0].tobytes() == whole[0] and acc[1].tobytes() == whole[1]
            assert np.allclose(acc[2], np.frombuffer(whole[2], dtype=np.float64), rtol=1e-12, atol=0)
            assert np.allclose(acc[3], np.frombuffer(whole[3], dtype=np.float64), rtol=1e-12, atol=0)
            assert acc[4].tobytes() == whole[4] and int(acc[5]) == whole[5]
    # introns: shard dicts merge to the whole
    rd = mb.Reader(path)
    whole = rd.count_introns()
    merged = {}
    for i in range(3):
        rd.set_shard(i, 3)
        for c, d in rd.count_introns().items():
            tgt = merged.setdefault(c, {})
            for k, v in d.items():
                tgt[k] = tgt.get(k, 0) + v
    assert merged == whole == c_oracle.count_introns(path)


def test_preloaded_file_gives_same_result(mb, tmp_path):
    path, contigs, recs, iv, giv = make_case(tmp_path, 16, n_reads=30_000)
    rd = mb.Reader(path)
    a = rd.count_reads(1, iv, giv)
    rd.preload()
    b = rd.count_reads(1, iv, giv)
    assert rd.stats()["h2d_bytes"] == 0
    rd.drop_preload()
    assert a == b == c_oracle.count_reads_stranded(path, None, iv, giv)


# ------------------------------------------------------------------------------------------------ BASELINE.json configs
@pytest.mark.parametrize("cfg,reads", [("cfg1", 1_000_000), ("cfg2", 1_500_000), ("cfg3", 1_500_000), ("cfg4", 1_000_000)],
                         ids=["cfg1_unstranded_1M", "cfg2_stranded", "cfg3_weighted_introns", "cfg4_paired"])
def test_baseline_config_shapes(mb, tmp_path, cfg, reads):
    """The BASELINE.json config shapes (tools/bamgen, reduced read counts except config 1 which runs at
    its full 1M reads): every entry point against the multi-threaded C oracle."""
    from synth import model

    path = str(tmp_path / (cfg + ".bam"))
    iv, giv, info = model.build_config(cfg, path, reads=reads, threads=os.cpu_count() or 4)
    assert abs(info["records"] - reads) <= 16  # multi-mapper groups are emitted whole
    nt = min(16, os.cpu_count() or 4)
    got_u = mb.count_reads_unstranded(path, None, iv, giv)
    assert mb.last_stats()["n_records"] == info["records"]
    assert got_u == c_oracle.count_reads_unstranded(path, None, iv, giv, nthreads=nt)
    assert mb.count_reads_stranded(path, None, iv, giv) == c_oracle.count_reads_stranded(path, None, iv, giv, nthreads=nt)
    assert_weighted_close(mb.count_reads_weighted(path, None, iv, giv), c_oracle.count_reads_weighted(path, None, iv, giv, nthreads=nt))
    assert mb.count_introns(path) == c_oracle.count_introns(path, nthreads=nt)
    assert got_u["_total"] > 0 and got_u["_outside"] > 0


@pytest.mark.parametrize("d", ["400", "401", "402", "403", "410", "411", "302", "304", "308", "316"])
def test_alternate_inflate_kernels_agree(mb, tmp_path, monkeypatch, d):
    """Every pass-1 variant (MBF_BAM_INFLATE_D: the CTA-per-member kernel in its three-decode and single-decode forms
    and several shapes, the lane-per-member kernel with 2..16 decoders per warp) must give the same bytes as zlib, on
    members with one and several deflate blocks, fixed and dynamic codes, tiny and full-size members."""
    monkeypatch.setenv("MBF_BAM_INFLATE_D", d)
    for k, kw in enumerate([dict(), dict(sync_every=700), dict(strategy=zlib.Z_FIXED), dict(level=1, payload_limit=3000),
                            dict(level=9, split_records=True), dict(strategy=zlib.Z_HUFFMAN_ONLY)]):
        path, *_ = make_case(tmp_path, 170 + k, n_reads=30_000, random_seq=(k % 2 == 0), **kw)
        data = open(path, "rb").read()
        want = b"".join(p for _, _, p in bamio.iter_bgzf_blocks(data))
        assert mb.inflate_buffer(data) == want, (d, kw)
    # highly compressible: long runs, two-bit tokens, members of a few hundred bytes of compressed data
    payload = (b"\0" * 150_000 + b"ab" * 60_000 + bytes(range(256)) * 300 + b"x" * 65_000) * 3
    p = tmp_path / "runs.bgzf"
    with open(p, "wb") as fh:
        w = bamio.BgzfWriter(fh)
        w.write(payload, atomic=False)
        w.close()
    assert mb.inflate_buffer(p.read_bytes()) == payload, d


def test_pinned_file_mapping_gives_same_result(mb, tmp_path):
    path, contigs, recs, iv, giv = make_case(tmp_path, 18, n_reads=30_000, random_seq=True)
    rd = mb.Reader(path)
    a = rd.count_reads(1, iv, giv)
    if not rd.pin_file():
        pytest.skip("platform refuses to page-lock a file mapping")
    b = rd.count_reads(1, iv, giv)
    rd.unpin_file()
    c = rd.count_reads(1, iv, giv)
    assert a == b == c == c_oracle.count_reads_stranded(path, None, iv, giv)


# ------------------------------------------------------------------------------------------------ round 2
@pytest.mark.parametrize("seed,kw", [(31, {}), (32, dict(exon_mode=False, genes_per_contig=120)),
                                     (33, dict(paired=True, genes_per_contig=200)), (34, dict(n_reads=40_000, genes_per_contig=400))])
def test_each_read_counts_once(mb, tmp_path, seed, kw):
    """each_read_counts_once=True (reference counters.rs:83-92,173-182): first hit in bio's tree order."""
    path, contigs, recs, iv, giv = make_case(tmp_path, seed, **kw)
    nt = 4
    assert mb.count_reads_unstranded(path, None, iv, giv, True) == c_oracle.count_reads_unstranded(path, None, iv, giv, True, nthreads=nt)
    assert mb.count_reads_stranded(path, None, iv, giv, True) == c_oracle.count_reads_stranded(path, None, iv, giv, True, nthreads=nt)
    assert mb.count_reads_stranded(path, None, iv, giv, each_read_counts_once=False) == c_oracle.count_reads_stranded(path, None, iv, giv, nthreads=nt)
    if seed != 34:
        assert mb.count_reads_unstranded(path, None, iv, giv, True) == py_oracle.count_reads_unstranded(path, None, iv, giv, True)
    with pytest.raises(ValueError, match="each_read_counts_once"):
        mb.count_reads_weighted(path, None, iv, giv, True)


def test_chunk_edges_with_stacked_genes(mb, tmp_path):
    """Counts on a layout where the chunk table depends on bio's iteration order (chunked_genome.rs:98)."""
    from test_oracle import _stacked_genes
    import random

    contigs, giv = _stacked_genes()
    rng = random.Random(3)
    recs = []
    for i in range(30_000):
        tid = rng.randrange(len(contigs))
        b = rng.choice([1_000_000, 2_000_000, 3_000_000])
        pos = b + rng.randrange(-5000, 5000) if i % 2 else rng.randrange(0, 4_000_000)
        nh = bamio.aux_int(b"NH", rng.choice([1, 1, 2, 3]))
        recs.append(bamio.Rec(tid, pos, b"q%d" % (i // 2), rng.choice([0, 16]), 255, [(0, 30), (3, rng.randrange(1, 3000)), (0, 20)], 50, nh))
    recs.sort(key=lambda r: (r.tid, r.pos))
    path = str(tmp_path / "stacked.bam")
    bamio.write_bam(path, contigs, recs)
    check_all(mb, path, giv, giv)
    assert mb.count_reads_unstranded(path, None, giv, giv) == py_oracle.count_reads_unstranded(path, None, giv, giv)


def test_property_gpu_vs_python_oracle(mb, tmp_path):
    """Random gene sets and reads (SURVEY App. A.6 edge cases come from synth.small) against the brute-force
    pure-Python oracle -- the independent statement of the semantics, not the C restatement."""
    from hypothesis import HealthCheck, given, settings, strategies as st
    from synth import small

    @settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.integers(0, 10 ** 6), st.integers(1, 3), st.integers(1, 60), st.booleans(), st.booleans(), st.sampled_from([0, 1, 6]))
    def prop(seed, n_contigs, genes, exon_mode, paired, level):
        contigs, recs, iv, giv = small.simple_case(seed, n_reads=600, n_contigs=n_contigs, contig_len=1_300_000,
                                                   genes_per_contig=genes, exon_mode=exon_mode, paired=paired)
        path = str(tmp_path / ("p%d.bam" % seed))
        bamio.write_bam(path, contigs, recs, level=level)
        assert mb.count_reads_unstranded(path, None, iv, giv) == py_oracle.count_reads_unstranded(path, None, iv, giv)
        assert mb.count_reads_stranded(path, None, iv, giv) == py_oracle.count_reads_stranded(path, None, iv, giv)
        assert mb.count_reads_stranded(path, None, iv, giv, True) == py_oracle.count_reads_stranded(path, None, iv, giv, True)
        assert_weighted_close(mb.count_reads_weighted(path, None, iv, giv), py_oracle.count_reads_weighted(path, None, iv, giv))
        assert mb.count_introns(path) == py_oracle.count_introns(path)
        assert mb.quantify_gene_reads(path, None, iv, giv) == py_oracle.quantify_gene_reads(path, None, iv, giv)
        assert mb.calculate_duplicate_distribution(path) == py_oracle.calculate_duplicate_distribution(path)
        fn = "count_reads_primary_only_right_strand_only_by_barcode"   # (no XC/XM tags here: chunks with a right-strand hit drop out)
        assert getattr(mb, fn)(path, None, iv, giv, "straight") == getattr(py_oracle, fn)(path, None, iv, giv, "straight")
        os.unlink(path)
        os.unlink(path + ".bai")

    prop()


def test_inflated_buffers_grow_in_the_middle_of_a_job(mb, tmp_path, monkeypatch):
    """A window that inflates to more than the buffers hold makes U (and what is sized from it) grow while the
    previous window's second half (intron pass 2, emit-only re-run) is still outstanding; only the growing
    window's own buffers may be replaced.  Fresh pipeline, 1-2 MiB windows, inflate ratio ~10."""
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "2")
    monkeypatch.setenv("MBF_BAM_MM_CAP", "64")
    path, contigs, recs, iv, giv = make_case(tmp_path, 51, n_reads=400_000)
    assert os.path.getsize(path) > 4 * (1 << 20)
    for fn in ("introns", "stranded"):
        mb.release_device(0)
        if fn == "introns":
            assert mb.count_introns(path) == c_oracle.count_introns(path, nthreads=4)
        else:
            assert mb.count_reads_stranded(path, None, iv, giv) == c_oracle.count_reads_stranded(path, None, iv, giv, nthreads=4)
        st = mb.last_stats()
        assert st["n_windows"] >= 3 and st["uncompressed_bytes"] > 8 * st["compressed_bytes"]
    mb.release_device(0)


def test_job_restarts_with_smaller_windows_after_a_ratio_jump(mb, tmp_path, monkeypatch):
    """Windows are planned from the largest inflate ratio seen so far; a much more compressible stretch can
    push one past the 32-bit offset range.  The job then starts again with smaller windows (lowered limits
    here: 48 MiB hard, 40 MiB planning target)."""
    rng = random.Random(8)
    payload = rng.randbytes(int(38.35 * (1 << 20))) + bytes(40 << 20)
    p = tmp_path / "jump.bgzf"
    with open(p, "wb") as fh:
        w = bamio.BgzfWriter(fh, level=1)
        w.write(payload, atomic=False)
        w.close()
    data = p.read_bytes()
    monkeypatch.setenv("MBF_BAM_WINDOW_MB", "16")
    monkeypatch.setenv("MBF_BAM_U_LIMIT_MB", "40")
    monkeypatch.setenv("MBF_BAM_U_HARD_LIMIT_MB", "48")
    got = mb.inflate_buffer(data)
    assert got == payload
    assert mb.last_stats()["n_retries"] >= 1


def test_fan_out_over_devices(mb, tmp_path):
    """One call, several devices (the rayon pool of counters.rs:217-259 at GPU granularity): every visible GPU
    gets a chunk range; results must not depend on the device list."""
    path, contigs, recs, iv, giv = make_case(tmp_path, 52, n_reads=60_000, n_contigs=4)
    n = mb.device_count()
    want = c_oracle.count_reads_stranded(path, None, iv, giv, nthreads=4)
    want_i = c_oracle.count_introns(path, nthreads=4)
    lists = [[0], list(range(n)), list(range(n))[::-1]]
    for devs in lists:
        rd = mb.Reader(path)
        rd.set_devices(devs)
        assert tuple(rd.count_reads(1, iv, giv)) == want
        assert rd.stats()["n_devices"] == len(devs)
        assert rd.count_introns() == want_i
        assert rd.count_reads(0, iv, giv, True) == c_oracle.count_reads_unstranded(path, None, iv, giv, True)
    # device fan-out composes with explicit shards (what one rank of a multi-node job would do)
    import numpy as np

    rd = mb.Reader(path)
    rd.set_devices(list(range(n)))
    whole = rd.count_reads_raw(1, iv, giv)
    acc = None
    for i in range(3):
        rd.set_shard(i, 3)
        part = rd.count_reads_raw(1, iv, giv)
        arrs = [np.frombuffer(part[0], dtype=np.uint64), np.frombuffer(part[1], dtype=np.uint64), np.uint64(part[5])]
        acc = arrs if acc is None else [a + b for a, b in zip(acc, arrs)]
    assert acc[0].tobytes() == whole[0] and acc[1].tobytes() == whole[1] and int(acc[2]) == whole[5]
    with pytest.raises(ValueError):
        rd.set_devices([0, 0]) or rd.count_reads(1, iv, giv)


def test_concurrent_calls_from_threads(mb, tmp_path):
    """No process-wide lock: calls from several threads are serialised per device and all give the right answer."""
    import threading

    path, contigs, recs, iv, giv = make_case(tmp_path, 53, n_reads=20_000)
    want = c_oracle.count_reads_stranded(path, None, iv, giv)
    out = [None] * 4

    def work(k):
        out[k] = mb.count_reads_stranded(path, None, iv, giv)

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert all(o == want for o in out)


def test_page_locked_file_cache(mb, tmp_path, monkeypatch):
    """hostmap.hpp: after MBF_BAM_PIN_AFTER jobs on a file its mapping is page-locked and later jobs DMA straight
    from the page cache; results are the same on the staged path, the pread path and the page-locked path."""
    path, contigs, recs, iv, giv = make_case(tmp_path, 54, n_reads=50_000, random_seq=True)
    want = c_oracle.count_reads_stranded(path, None, iv, giv)
    mb.host_cache_clear()
    monkeypatch.setenv("MBF_BAM_PIN_AFTER", "2")
    stats = []
    for _ in range(4):
        assert mb.count_reads_stranded(path, None, iv, giv) == want
        stats.append(mb.last_stats())
    assert stats[0]["h2d_pinned_bytes"] == 0 and stats[1]["h2d_pinned_bytes"] == 0
    if stats[2]["h2d_pinned_bytes"] == 0:
        pytest.skip("platform refuses to page-lock a file mapping")
    assert stats[2]["ms_pin"] > 0 and stats[3]["ms_pin"] == 0
    assert stats[3]["h2d_pinned_bytes"] == stats[3]["h2d_bytes"] > 0
    mb.host_cache_clear()
    monkeypatch.setenv("MBF_BAM_PIN_AFTER", "never")
    monkeypatch.setenv("MBF_BAM_STAGE", "pread")
    assert mb.count_reads_stranded(path, None, iv, giv) == want
    assert mb.last_stats()["h2d_pinned_bytes"] == 0
    # a rewritten file (other size/mtime) is never served from a stale mapping
    monkeypatch.delenv("MBF_BAM_STAGE")
    monkeypatch.setenv("MBF_BAM_PIN_AFTER", "0")
    assert mb.count_reads_stranded(path, None, iv, giv) == want
    path2, contigs2, recs2, iv2, giv2 = make_case(tmp_path, 55, n_reads=30_000, random_seq=True)
    os.replace(path2, path)
    os.replace(path2 + ".bai", path + ".bai")
    assert mb.count_reads_stranded(path, None, iv2, giv2) == c_oracle.count_reads_stranded(path, None, iv2, giv2)
    mb.host_cache_clear()


def test_csi_index(mb, tmp_path):
    """the same file under a CSI index (two binning schemes): identical results, also when sharded"""
    from synth import small

    contigs, recs, iv, giv = small.simple_case(78, n_reads=30000, n_contigs=3, random_seq=True)
    pa, pb, pc = (str(tmp_path / n) for n in ("a.bam", "b.bam", "c.bam"))
    bamio.write_bam(pa, contigs, recs)
    bamio.write_bam(pb, contigs, recs, csi=(14, 5))
    bamio.write_bam(pc, contigs, recs, csi=(12, 6))
    want = c_oracle.count_reads_stranded(pa, None, iv, giv, nthreads=4)
    want_i = c_oracle.count_introns(pa, nthreads=4)
    for p in (pb, pc):
        assert mb.count_reads_stranded(p, None, iv, giv) == want
        assert mb.count_introns(p) == want_i
        f, r = {}, {}
        for s in range(3):
            rd = mb.Reader(p)
            rd.set_device(0)
            rd.set_shard(s, 3)
            a, b = rd.count_reads(1, iv, giv)
            for src, dst in ((a, f), (b, r)):
                for k, v in src.items():
                    dst[k] = dst.get(k, 0) + v
        assert (f, r) == want
