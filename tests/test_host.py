"""CPU suite, part 2: the C-ABI library loads and exports every declared symbol, the host-side control
plane (header/BAI parsing, ChunkedGenome table, argument handling, error convention) behaves like the
reference's -- no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, make_case
from oracle import c_oracle


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "mbfbam.h")).read()
    names = set(re.findall(r"\b(mbfbam_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 15
    lib = ctypes.CDLL(os.path.join(ROOT, "mbf_bam_b200", "libmbfbam_cuda.so"))
    for n in sorted(names):
        assert hasattr(lib, n), "missing export " + n
    lib.mbfbam_version.restype = ctypes.c_char_p
    assert lib.mbfbam_version().decode() == built.__version__


def test_module_surface_matches_reference(built):
    """reference src/lib.rs:286-299 (+ count_reads_weighted named by the north star)."""
    import inspect

    import mbf_bam

    for fn in ("count_reads_unstranded", "count_reads_stranded", "count_reads_weighted", "count_introns", "quantify_gene_reads",
               "calculate_duplicate_distribution", "count_reads_primary_only_right_strand_only_by_barcode"):
        assert callable(getattr(mbf_bam, fn))
    for fn in ("subtract_bam", "annotate_barcodes_from_fastq"):   # the writers: present, and honest about it
        with pytest.raises(NotImplementedError):
            getattr(mbf_bam, fn)("a", "b", "c")
    assert isinstance(mbf_bam.__version__, str)
    doc = built.count_reads_unstranded.__doc__
    for arg in ("filename", "index_filename", "intervals", "gene_intervals", "each_read_counts_once"):
        assert arg in doc
    assert "index_filename" in built.count_introns.__doc__
    assert inspect.isbuiltin(built.count_reads_stranded) or callable(built.count_reads_stranded)


def test_open_errors_are_value_errors(built, tmp_path):
    """BamError -> ValueError (reference src/lib.rs:26-32), message of src/bam_ext.rs:16-18."""
    with pytest.raises(ValueError, match="Could not read bam"):
        built.count_introns(str(tmp_path / "missing.bam"))
    path, contigs, recs, iv, giv = make_case(tmp_path, 21, n_reads=200)
    os.rename(path + ".bai", path + ".bai.moved")
    with pytest.raises(ValueError, match="Could not read bam"):
        built.count_reads_unstranded(path, None, iv, giv)
    # explicit index path (IndexedReader::from_path_and_index)
    rd = built.Reader(path, path + ".bai.moved")
    assert [n for n, _ in rd.targets()] == [c for c, _ in contigs]
    junk = tmp_path / "junk.bam"
    junk.write_bytes(b"this is not a bam file at all" * 10)
    with pytest.raises(ValueError, match="Could not read bam"):
        built.Reader(str(junk), path + ".bai.moved")


def test_argument_errors(built, tmp_path):
    path, contigs, recs, iv, giv = make_case(tmp_path, 22, n_reads=200)
    with pytest.raises(TypeError):
        built.count_reads_unstranded(path, None, [1, 2], giv)          # &PyDict argument
    with pytest.raises(ValueError, match="Python error"):
        built.count_reads_unstranded(path, None, {"chrA": "nope"}, giv)
    with pytest.raises(ValueError, match="Python error"):
        built.count_reads_unstranded(path, None, {"chrA": [("g", 1000, [1], [2])]}, giv)  # strand > i8


def test_chunk_table_matches_oracle(built, tmp_path):
    """ChunkedGenome (reference src/count_reads/chunked_genome.rs:72-119): host planner vs oracle."""
    for seed in (31, 32, 33):
        path, contigs, recs, iv, giv = make_case(tmp_path, seed, n_reads=300, n_contigs=3, genes_per_contig=60)
        rd = built.Reader(path)
        assert [tuple(c) for c in rd.chunks(giv)] == c_oracle.chunks(path, None, giv)
        assert [tuple(c) for c in rd.chunks(None)] == c_oracle.chunks(path, None, None)
        sub = {k: giv[k] for k in list(giv)[:1]}
        sub["not_in_header"] = giv[list(giv)[0]]
        assert [tuple(c) for c in rd.chunks(sub)] == c_oracle.chunks(path, None, sub)


def _strip_pseudo_bins(bai: bytes) -> bytes:
    """the same index without the samtools metadata bin 37450"""
    import struct

    out = bytearray(bai[:8])
    p = 8
    (n_ref,) = struct.unpack_from("<i", bai, 4)
    for _ in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", bai, p)
        p += 4
        kept = []
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", bai, p)
            rec = bai[p:p + 8 + 16 * n_chunk]
            p += len(rec)
            if b != 37450:
                kept.append(rec)
        out += struct.pack("<i", len(kept)) + b"".join(kept)
        (n_intv,) = struct.unpack_from("<i", bai, p)
        out += bai[p:p + 4 + 8 * n_intv]
        p += 4 + 8 * n_intv
    return bytes(out + bai[p:])


def test_shard_plan_same_with_and_without_metadata_bin(built, tmp_path):
    """Reference begin/end offsets come from the metadata pseudo-bin when the index has one and from the chunk
    lists otherwise; the byte ranges planned from both must be identical (whole file and mid-contig cuts)."""
    import mbf_bam_b200 as mb

    path, contigs, recs, iv, giv = make_case(tmp_path, 31, n_reads=20_000, n_contigs=5)
    bai = open(path + ".bai", "rb").read()
    stripped = _strip_pseudo_bins(bai)
    assert len(stripped) < len(bai)
    alt = str(tmp_path / "nometa.bai")
    open(alt, "wb").write(stripped)
    a, b = mb.Reader(path), mb.Reader(path, alt)
    for n in (1, 2, 3, 7):
        for i in range(n):
            assert a.plan(giv, i, n) == b.plan(giv, i, n), (i, n)
    assert a.chunks(giv) == b.chunks(giv)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_result_dict_assembly_matches_oracle(built, tmp_path, seed):
    """Dense count vectors -> the reference's dict layout (counters.rs:236-254 `insert` per contig, :262-273 `+=`,
    :309-314 `_outside` only forward): gene ids duplicated within and across contigs, ids that collide with the
    special keys, unscanned contigs.  Host-only, so it runs without a GPU."""
    import random

    import numpy as np

    import mbf_bam_b200 as mb
    from oracle import c_oracle

    rng = random.Random(seed)
    path, contigs, recs, iv0, giv0 = make_case(tmp_path, 40 + seed, n_reads=200, n_contigs=4)
    names = [c for c, _ in contigs]
    pool = ["g%d" % i for i in range(12)] + ["_total", "_outside", "_" + names[0]]
    iv = {c: [(rng.choice(pool), rng.choice([1, -1]), [10 * k], [10 * k + 5]) for k in range(rng.randrange(1, 9))]
          for c in names}
    G = sum(len(v) for v in iv.values())
    fwd = np.array([rng.randrange(0, 1000) for _ in range(G)], dtype=np.uint64)
    rev = np.array([rng.randrange(0, 1000) for _ in range(G)], dtype=np.uint64)
    wf = np.array([rng.random() * 10 for _ in range(G)], dtype=np.float64)
    wr = np.array([rng.random() * 10 for _ in range(G)], dtype=np.float64)
    scanned = np.array([rng.random() < 0.8 for _ in names], dtype=np.uint8)
    outside = rng.randrange(0, 10 ** 6)
    rd = mb.Reader(path)
    args = (iv, fwd.tobytes(), rev.tobytes(), wf.tobytes(), wr.tobytes(), scanned.tobytes(), outside)
    r = dict(names=names, gene_ids=[[t[0] for t in iv[c]] for c in names], scanned=scanned, fwd=fwd, rev=rev,
             wfwd=wf, wrev=wr, outside=outside)
    assert rd.assemble(0, *args) == c_oracle._assemble_unstranded(r)
    assert tuple(rd.assemble(1, *args)) == c_oracle._assemble_dual(r, "fwd", "rev", int)
    got = rd.assemble(2, *args)
    want = c_oracle._assemble_dual(r, "wfwd", "wrev", float)
    for a, b in zip(got, want):
        assert set(a) == set(b)
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-9 * max(1.0, abs(b[k])), k


def test_interval_dict_errors_are_value_errors(built, tmp_path):
    """Malformed interval dicts raise ValueError (lib.rs:118-134 extraction errors), never crash."""
    import mbf_bam_b200 as mb

    path, contigs, recs, iv, giv = make_case(tmp_path, 44, n_reads=200)
    rd = mb.Reader(path)
    c = contigs[0][0]
    for bad in ({c: "nope"}, {c: [("g", 1, [1])]}, {c: [("g", 1, [1], 5)]}, {c: [("g", "x", [1], [2])]},
                {c: [("g", 1, [-1], [2])]}, {c: [("g", 1, [1], [2 ** 33])]}, {c: [(5, 1, [1], [2])]}, {7: []},
                {c: [("g", 300, [1], [2])]}, {c: [["g", 1, [1], [2]]]}):
        with pytest.raises(ValueError):
            rd.plan(bad, 0, 1)


def test_no_cpu_fallback(built, has_gpu, tmp_path):
    """Without a CUDA device every compute entry point must fail loudly."""
    if has_gpu:
        pytest.skip("GPU present")
    path, contigs, recs, iv, giv = make_case(tmp_path, 23, n_reads=200)
    for call in (lambda: built.count_reads_stranded(path, None, iv, giv),
                 lambda: built.count_introns(path),
                 lambda: built.inflate_buffer(open(path, "rb").read()),
                 lambda: built.cigar_expand([16], 5)):
        with pytest.raises(ValueError, match="no CUDA device"):
            call()


def test_csi_index_gives_the_same_plan_as_bai(built, tmp_path):
    """CSI indices (BGZF compressed, per-bin loffset, min_shift / depth in the header; htslib looks for them before
    the .bai): chunk table and shard byte ranges must cover what the BAI-derived ones cover."""
    import shutil

    from synth import bamio, small

    contigs, recs, iv, giv = small.simple_case(77, n_reads=20000, n_contigs=3, random_seq=True)
    pa, pb, pc = (str(tmp_path / n) for n in ("a.bam", "b.bam", "c.bam"))
    bamio.write_bam(pa, contigs, recs)
    bamio.write_bam(pb, contigs, recs, csi=(14, 5))
    bamio.write_bam(pc, contigs, recs, csi=(12, 6))
    assert os.path.exists(pb + ".csi") and not os.path.exists(pb + ".bai")
    ra, rb, rc = built.Reader(pa), built.Reader(pb), built.Reader(pc)
    assert ra.chunks(giv) == rb.chunks(giv) == rc.chunks(giv) == c_oracle.chunks(pa, None, giv)
    for n in (1, 2, 5):
        for i in range(n):
            lo_a, hi_a, ranges_a = ra.plan(giv, i, n)
            for r in (rb, rc):
                lo, hi, ranges = r.plan(giv, i, n)
                assert (lo, hi) == (lo_a, hi_a)
                for b, e in ranges_a:   # every byte the BAI plan reads is read under the CSI plan too
                    assert any(b2 <= b and e <= e2 for b2, e2 in ranges), (n, i, ranges_a, ranges)
    # an explicit index file name decides by content, not by extension; a .csi beside a .bai wins (htslib's order)
    shutil.copy(pb + ".csi", str(tmp_path / "explicit.idx"))
    assert built.Reader(pb, str(tmp_path / "explicit.idx")).chunks(giv) == ra.chunks(giv)
    open(pb + ".bai", "wb").write(b"not an index")
    assert built.Reader(pb).chunks(giv) == ra.chunks(giv)
    with open(pc + ".csi", "r+b") as fh:   # damaged CSI: a ValueError, not a crash
        fh.seek(40)
        fh.write(b"\xff" * 30)
    with pytest.raises(ValueError, match="Could not read bam"):
        built.Reader(pc)
