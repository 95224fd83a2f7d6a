"""One process per GPU (torchrun / torch.distributed): the same four counters, sharded over the ranks.

The reference runs every chunk of the genome on a rayon pool inside one call (reference
src/count_reads/counters.rs:217-259, :289-322; src/count_reads/introns.rs:66-87).  Here the unit of
parallelism above a GPU is the chunk range: rank r of W counts shard r of W of the ChunkedGenome table on its
own GPU (LOCAL_RANK), the dense per-gene vectors are summed with ONE all-reduce (NCCL; ~1 MB), intron maps are
gathered and merged key-wise (rayon's `reduce` with add_hashmaps / combine_intron_results), and every rank
returns the whole-job result, assembled by the same host code as the single-process functions.  Multi-mapper
sets are chunk-local (counters.rs:36,102-104), so shards never share one: there is no other exchange.

    torchrun --nproc-per-node 8 script.py        # script: torch.cuda.set_device(LOCAL_RANK); init_process_group("nccl")
    from mbf_bam_b200 import dist as mbd
    fwd, rev = mbd.count_reads_stranded(bam, None, intervals, gene_intervals)

(For one process driving several GPUs use the plain functions: they fan out over the visible devices.)
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import MODE_STRANDED, MODE_UNSTRANDED, MODE_WEIGHTED, Reader

_last_stats: dict = {}


def last_stats() -> dict:
    """measurements of this rank's part of the most recent call"""
    return dict(_last_stats)


def _dist():
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("mbf_bam_b200.dist needs an initialised torch.distributed process group")
    return dist


def _local_device(dist) -> int:
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    import torch

    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return dist.get_rank() % n if n else 0


def _tensor_device(dist, group, local):
    import torch

    return torch.device("cuda", local) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def _gather_objects(part, group=None):
    """all_gather_object of `part`; under NCCL the pickled bytes travel through this rank's own GPU (the call uses the
    *current* CUDA device, which a script that never called torch.cuda.set_device would leave at 0 on every rank)"""
    dist = _dist()
    parts = [None] * dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        import torch

        with torch.cuda.device(_local_device(dist)):
            dist.all_gather_object(parts, part, group=group)
    else:
        dist.all_gather_object(parts, part, group=group)
    return parts


def merge_dense(raw, group=None):
    """Sum the dense per-shard results of Reader.count_reads_raw over the ranks.

    raw = (fwd u64 bytes, rev u64 bytes, wfwd f64 bytes, wrev f64 bytes, scanned u8 bytes, outside).  Integer
    vectors travel as int64 (counts are far below 2^63), the weighted ones as float64; `scanned` is OR-ed."""
    import torch

    dist = _dist()
    dev = _tensor_device(dist, group, _local_device(dist))
    fwd = torch.from_numpy(np.frombuffer(raw[0], dtype=np.uint64).astype(np.int64)).to(dev)
    rev = torch.from_numpy(np.frombuffer(raw[1], dtype=np.uint64).astype(np.int64)).to(dev)
    sc = torch.from_numpy(np.frombuffer(raw[4], dtype=np.uint8).astype(np.int64)).to(dev)
    ints = torch.cat([fwd, rev, sc, torch.tensor([int(raw[5])], dtype=torch.int64, device=dev)])
    flts = torch.cat([torch.from_numpy(np.frombuffer(raw[2], dtype=np.float64).copy()).to(dev),
                      torch.from_numpy(np.frombuffer(raw[3], dtype=np.float64).copy()).to(dev)])
    dist.all_reduce(ints, group=group)
    if flts.numel():
        dist.all_reduce(flts, group=group)
    ints = ints.cpu().numpy()
    flts = flts.cpu().numpy()
    g, nc = len(fwd), len(sc)
    return (ints[:g].astype(np.uint64).tobytes(), ints[g:2 * g].astype(np.uint64).tobytes(),
            flts[:g].tobytes(), flts[g:].tobytes(), (ints[2 * g:2 * g + nc] > 0).astype(np.uint8).tobytes(),
            int(ints[2 * g + nc]))


def merge_introns(part: dict, group=None) -> dict:
    """reference src/count_reads/introns.rs:12-22 combine_intron_results over the ranks"""
    parts = _gather_objects(part, group)
    out: dict = {}
    for p in parts:
        for contig, d in p.items():
            tgt = out.setdefault(contig, {})
            for k, v in d.items():
                tgt[k] = tgt.get(k, 0) + v
    return out


def _reader(filename, index_filename, group, device):
    dist = _dist()
    rd = Reader(filename, index_filename)
    rd.set_device(_local_device(dist) if device is None else device)
    rd.set_shard(dist.get_rank(group), dist.get_world_size(group))
    return rd


def _count(mode, filename, index_filename, intervals, gene_intervals, each_read_counts_once, group, device, _raw=None):
    global _last_stats
    rd = _reader(filename, index_filename, group, device)
    raw = _raw(rd) if _raw is not None else rd.count_reads_raw(mode, intervals, gene_intervals, each_read_counts_once)
    _last_stats = rd.stats()
    return rd.assemble(mode, intervals, *merge_dense(raw, group))


def count_reads_unstranded(filename, index_filename, intervals, gene_intervals, each_read_counts_once=None, group=None,
                           device: Optional[int] = None, _raw=None):
    """reference src/lib.rs:138-157, sharded over the ranks of `group`"""
    return _count(MODE_UNSTRANDED, filename, index_filename, intervals, gene_intervals, each_read_counts_once, group, device, _raw)


def count_reads_stranded(filename, index_filename, intervals, gene_intervals, each_read_counts_once=None, group=None,
                         device: Optional[int] = None, _raw=None):
    """reference src/lib.rs:161-181, sharded over the ranks of `group`"""
    return tuple(_count(MODE_STRANDED, filename, index_filename, intervals, gene_intervals, each_read_counts_once, group, device, _raw))


def count_reads_weighted(filename, index_filename, intervals, gene_intervals, each_read_counts_once=None, group=None,
                         device: Optional[int] = None, _raw=None):
    """1/NH weighted counts (not in the reference), sharded over the ranks of `group`"""
    return tuple(_count(MODE_WEIGHTED, filename, index_filename, intervals, gene_intervals, each_read_counts_once, group, device, _raw))


def count_introns(filename, index_filename=None, group=None, device: Optional[int] = None, _part=None):
    """reference src/lib.rs:216-225, sharded over the ranks of `group`"""
    global _last_stats
    rd = _reader(filename, index_filename, group, device)
    part = _part(rd) if _part is not None else rd.count_introns()
    _last_stats = rd.stats()
    return merge_introns(part, group)


def merge_quantify(part, group=None):
    """reference src/count_reads/quantify.rs:189-203 add_hashmaps over the ranks: per gene id the position lists of the
    shards are disjoint (a position belongs to one chunk) and are concatenated"""
    parts = _gather_objects(part, group)
    out = ({}, {})
    for p in parts:
        for b in (0, 1):
            for gene, lst in p[b].items():
                out[b].setdefault(gene, []).extend(lst)
    return tuple({g: sorted(l) for g, l in o.items()} for o in out)


def merge_hist(part: dict, group=None) -> dict:
    """reference src/duplicate_distribution.rs:8-14 add_hashmaps over the ranks"""
    parts = _gather_objects(part, group)
    out: dict = {}
    for p in parts:
        for k, v in p.items():
            out[k] = out.get(k, 0) + v
    return out


def merge_by_barcode(part, group=None):
    """reference src/count_reads/by_barcode.rs:74-92 over the ranks' entry lists, in rank (= chunk) order"""
    parts = _gather_objects(part, group)
    genes, barcodes, counts = {}, {}, []
    for g_names, b_names, entries in parts:
        for gi, bi, n in entries:
            g = genes.setdefault(g_names[gi], len(genes))
            b = barcodes.setdefault(b_names[bi], len(barcodes))
            counts.append((g, b, n))
    return list(genes), list(barcodes), counts


def quantify_gene_reads(filename, index_filename, intervals, gene_intervals, group=None, device: Optional[int] = None, _part=None):
    """reference src/lib.rs:247-264, sharded over the ranks of `group`"""
    global _last_stats
    rd = _reader(filename, index_filename, group, device)
    part = _part(rd) if _part is not None else rd.quantify_gene_reads(intervals, gene_intervals)
    _last_stats = rd.stats()
    return merge_quantify(part, group)


def calculate_duplicate_distribution(filename, index_filename=None, group=None, device: Optional[int] = None, _part=None):
    """reference src/lib.rs:108-116, sharded over the ranks of `group`"""
    global _last_stats
    rd = _reader(filename, index_filename, group, device)
    part = _part(rd) if _part is not None else rd.calculate_duplicate_distribution()
    _last_stats = rd.stats()
    return merge_hist(part, group)


def count_reads_primary_only_right_strand_only_by_barcode(filename, index_filename, intervals, gene_intervals, umi_strategy,
                                                          group=None, device: Optional[int] = None, _part=None):
    """reference src/lib.rs:184-211, sharded over the ranks of `group`"""
    global _last_stats
    rd = _reader(filename, index_filename, group, device)
    part = _part(rd) if _part is not None else rd.count_reads_primary_only_right_strand_only_by_barcode(
        intervals, gene_intervals, umi_strategy)
    _last_stats = rd.stats()
    return merge_by_barcode(part, group)
