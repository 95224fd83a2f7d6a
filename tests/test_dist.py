"""CPU suite, part 3: the N > 1 path without GPUs -- chunk-range shard planning (host code of the CUDA
library, no kernel runs) and the cross-rank merge of dense per-shard results over torch.distributed
(gloo, world_size 2), exactly as bench.py merges them over NCCL."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, make_case
from oracle import c_oracle, py_oracle


def test_shard_plan_partitions_the_chunk_table(built, tmp_path):
    path, contigs, recs, iv, giv = make_case(tmp_path, 41, n_reads=20_000, n_contigs=4, contig_len=3_200_000)
    rd = built.Reader(path)
    chunks = rd.chunks(giv)
    size = os.path.getsize(path)
    for n in (1, 2, 3, 4, 7, 8):
        covered = []
        for i in range(n):
            lo, hi, ranges = rd.plan(giv, i, n)
            assert 0 <= lo <= hi <= len(chunks)
            covered.append((lo, hi))
            for b, e in ranges:
                assert 0 <= b < e <= size
            # byte ranges are sorted and disjoint
            assert all(ranges[k][1] < ranges[k + 1][0] for k in range(len(ranges) - 1))
        # contiguous, non-overlapping, complete cover of the chunk table
        assert covered[0][0] == 0 and covered[-1][1] == len(chunks)
        assert all(covered[k][1] == covered[k + 1][0] for k in range(n - 1))
    # every record of an owned chunk lies inside the shard's byte ranges (checked through virtual offsets
    # recomputed from the file)
    from synth import bamio

    data = open(path, "rb").read()
    starts = []
    u = 0
    for coff, csize, payload in bamio.iter_bgzf_blocks(data):
        starts.append((u, coff, csize))
        u += len(payload)
    ub = b"".join(p for _, _, p in bamio.iter_bgzf_blocks(data))
    import bisect
    import struct

    ustarts = [s[0] for s in starts]
    # walk records, remember (tid, pos, coffset of the block holding the record start)
    hdr, _ = bamio.read_bam(path)
    p = 12 + struct.unpack_from("<i", ub, 4)[0]
    for _ in hdr.refs:
        ln = struct.unpack_from("<i", ub, p)[0]
        p += 8 + ln
    recpos = []
    while p < len(ub):
        bs, tid, pos = struct.unpack_from("<iii", ub, p)
        k = bisect.bisect_right(ustarts, p) - 1
        recpos.append((tid, pos, starts[k][1], starts[k][1] + starts[k][2]))
        p += 4 + bs
    for n in (2, 5):
        for i in range(n):
            lo, hi, ranges = rd.plan(giv, i, n)
            owned = chunks[lo:hi]
            for tid, pos, cb, ce in recpos[::7]:
                if pos < 0 or tid < 0:
                    continue
                if any(t == tid and s <= pos < e for t, s, e in owned):
                    assert any(b <= cb and cb < e for b, e in ranges), (tid, pos, cb, ranges)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, path, iv, giv, out_q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, ROOT)
    from mbf_bam_b200 import dist as mbd
    from oracle import c_oracle as orc

    # No GPU here: the oracle restricted to this rank's chunk range stands in for the device part
    # (Reader.count_reads_raw / Reader.count_introns); planning, merge and assembly are the product's.
    def raw_for(mode, once=False):
        def f(rd):
            lo, hi, _ = rd.plan(giv, rank, world)
            part = orc.raw_counts(path, None, iv, giv, mode, nthreads=1, chunk_lo=lo, chunk_hi=hi, each_read_counts_once=once)
            return (part["fwd"].tobytes(), part["rev"].tobytes(), part["wfwd"].tobytes(), part["wrev"].tobytes(),
                    part["scanned"].tobytes(), int(part["outside"]))
        return f

    def introns_part(rd):
        lo, hi, _ = rd.plan(None, rank, world)
        chunks = rd.chunks(None)[lo:hi]
        names = [n for n, _ in rd.targets()]
        whole = orc.count_introns(path)
        # 1 Mb chunks: keep the gaps of reads whose pos falls into an owned chunk -- recomputed from the records
        from synth import bamio
        from oracle import py_oracle

        _, recs = bamio.read_bam(path)
        out = {}
        for tid, s, e in chunks:
            d = out.setdefault(names[tid], {})
            for r in recs:
                if r.tid != tid or not (s <= (r.pos & 0xFFFFFFFF) < e):
                    continue
                k = 0
                for gap in py_oracle.cigar_introns(r.pos, r.cigar):
                    d[gap] = d.get(gap, 0) + 1
                    k += 1
                key = (0xFFFFFFFF, 0xFFFFFFFF - k)
                d[key] = d.get(key, 0) + 1
        return out, whole

    res = {
        "stranded": mbd.count_reads_stranded(path, None, iv, giv, _raw=raw_for(1)),
        "unstranded": mbd.count_reads_unstranded(path, None, iv, giv, _raw=raw_for(0)),
        "weighted": mbd.count_reads_weighted(path, None, iv, giv, _raw=raw_for(2)),
        "once": mbd.count_reads_stranded(path, None, iv, giv, True, _raw=raw_for(1, True)),
    }
    whole_box = []
    res["introns"] = mbd.count_introns(path, _part=lambda rd: (lambda pw: (whole_box.append(pw[1]), pw[0])[1])(introns_part(rd)))
    res["introns_whole"] = whole_box[0]
    # the 8(f) rows: the Python restatement's whole-job result, cut into per-rank parts, stands in for the device parts
    from oracle import py_oracle

    q_whole = py_oracle.quantify_gene_reads(path, None, iv, giv)
    q_part = tuple({g: [t for t in lst if (t[0] // 1000) % world == rank] for g, lst in d.items()} for d in q_whole)
    res["quantify"] = mbd.quantify_gene_reads(path, None, iv, giv, _part=lambda rd: q_part)
    h_whole = py_oracle.calculate_duplicate_distribution(path)
    h_part = {k: v // world + (1 if rank < v % world else 0) for k, v in h_whole.items()}
    res["dups"] = mbd.calculate_duplicate_distribution(path, _part=lambda rd: {k: v for k, v in h_part.items() if v})
    fn = "count_reads_primary_only_right_strand_only_by_barcode"
    b_whole = getattr(py_oracle, fn)(path, None, iv, giv, "straight")
    half = len(b_whole[2]) // 2
    mine = b_whole[2][:half] if rank == 0 else b_whole[2][half:]
    g_loc, b_loc, e_loc = {}, {}, []
    for gi, bi, n in mine:   # rank-local index space, as a shard's own call would return it
        e_loc.append((g_loc.setdefault(b_whole[0][gi], len(g_loc)), b_loc.setdefault(b_whole[1][bi], len(b_loc)), n))
    res["barcode"] = getattr(mbd, fn)(path, None, iv, giv, "straight", _part=lambda rd: (list(g_loc), list(b_loc), e_loc))
    out_q.put((rank, res))
    dist.destroy_process_group()


def test_two_rank_merge_over_gloo(built, tmp_path):
    """world_size 2 through mbf_bam_b200.dist: each rank counts its chunk range, dense results are all-reduced,
    intron maps gathered and merged, EVERY rank gets the whole-job dicts; must equal the single-process oracle."""
    import torch.multiprocessing as mp

    from test_next_rows import barcode_case  # XC/XM tags, duplicates, and chunks the by-barcode counter drops

    path, contigs, recs, iv, giv = barcode_case(tmp_path, 42, n_reads=8_000, n_contigs=3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, path, iv, giv, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        res = got[rank]
        assert res["stranded"] == c_oracle.count_reads_stranded(path, None, iv, giv)
        assert res["unstranded"] == c_oracle.count_reads_unstranded(path, None, iv, giv)
        assert res["once"] == c_oracle.count_reads_stranded(path, None, iv, giv, True)
        want_w = c_oracle.count_reads_weighted(path, None, iv, giv)
        for a, b in zip(res["weighted"], want_w):
            assert set(a) == set(b) and all(abs(a[k] - b[k]) <= 1e-9 * max(1.0, abs(b[k])) for k in a)
        assert res["introns"] == res["introns_whole"] == c_oracle.count_introns(path)
        assert res["quantify"] == py_oracle.quantify_gene_reads(path, None, iv, giv)
        assert res["dups"] == py_oracle.calculate_duplicate_distribution(path)
        assert len(res["barcode"][2]) > 20
        assert res["barcode"] == getattr(py_oracle, "count_reads_primary_only_right_strand_only_by_barcode")(path, None, iv, giv, "straight")
