"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/libmbf_oracle.so.

Presents the C restatement with the reference's Python signatures
(reference src/lib.rs:138-225) so parity tests read like
`oracle.count_reads_stranded(...) == mbf_bam.count_reads_stranded(...)`.
The dict assembly below restates counters.rs:236-254 (unstranded: `insert`,
last duplicate id wins, `_total` sums everything) and counters.rs:262-273,
:309-314 (stranded: `+=`, `_outside` only in the forward dict) independently
of the product's C++ host extension.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Intervals(C.Structure):
    _fields_ = [
        ("n_contigs", C.c_int32),
        ("contig_names", C.POINTER(C.c_char_p)),
        ("n_genes", C.POINTER(C.c_uint32)),
        ("iv_offsets", C.POINTER(C.c_uint64)),
        ("iv_start", C.POINTER(C.c_uint32)),
        ("iv_end", C.POINTER(C.c_uint32)),
        ("iv_gene", C.POINTER(C.c_uint32)),
        ("iv_strand", C.POINTER(C.c_int8)),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "libmbf_oracle.so")
    src = os.path.join(HERE, "mbf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_chunks.restype = C.c_int64
        _LIB.oracle_inflate_buffer.restype = C.c_int64
    return _LIB


def flatten(intervals: dict):
    """{chr: [(gene_id, strand, starts, stops)]} -> (Intervals struct, keepalive, gene_ids per contig)."""
    names = list(intervals.keys())
    n_genes, offs, st, en, ge, sd, gene_ids = [], [0], [], [], [], [], []
    for c in names:
        lst = intervals[c]
        n_genes.append(len(lst))
        gene_ids.append([t[0] for t in lst])
        for gno, (gid, strand, starts, stops) in enumerate(lst):
            for s, e in zip(starts, stops):
                if s > e:
                    raise ValueError("interval start > end")
                st.append(s); en.append(e); ge.append(gno); sd.append(strand)
        offs.append(len(st))
    a_names = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
    a = dict(
        n_genes=np.asarray(n_genes, dtype=np.uint32), offs=np.asarray(offs, dtype=np.uint64),
        st=np.asarray(st, dtype=np.uint32), en=np.asarray(en, dtype=np.uint32),
        ge=np.asarray(ge, dtype=np.uint32), sd=np.asarray(sd, dtype=np.int8), names=a_names)
    iv = Intervals(
        len(names), C.cast(a_names, C.POINTER(C.c_char_p)),
        a["n_genes"].ctypes.data_as(C.POINTER(C.c_uint32)),
        a["offs"].ctypes.data_as(C.POINTER(C.c_uint64)),
        a["st"].ctypes.data_as(C.POINTER(C.c_uint32)),
        a["en"].ctypes.data_as(C.POINTER(C.c_uint32)),
        a["ge"].ctypes.data_as(C.POINTER(C.c_uint32)),
        a["sd"].ctypes.data_as(C.POINTER(C.c_int8)))
    return iv, a, names, gene_ids


def raw_counts(path, index_path, intervals, gene_intervals, mode, nthreads=1, sample_chunks=0, chunk_lo=-1, chunk_hi=-1,
               each_read_counts_once=False):
    iv, keep1, names, gene_ids = flatten(intervals)
    gv, keep2, _, _ = flatten(gene_intervals)
    tot = int(keep1["n_genes"].sum())
    fwd = np.zeros(tot + 1, dtype=np.uint64)
    rev = np.zeros(tot + 1, dtype=np.uint64)
    wf = np.zeros(tot + 1, dtype=np.float64)
    wr = np.zeros(tot + 1, dtype=np.float64)
    scanned = np.zeros(len(names) + 1, dtype=np.uint8)
    outside = C.c_uint64(0)
    nrec = C.c_uint64(0)
    nck = C.c_int64(0)
    err = C.create_string_buffer(512)
    rc = lib().oracle_count_reads(
        path.encode(), index_path.encode() if index_path else None, C.byref(iv), C.byref(gv),
        C.c_int(mode), C.c_int(1 if each_read_counts_once else 0), C.c_int(nthreads), C.c_int64(sample_chunks), C.c_int64(chunk_lo), C.c_int64(chunk_hi),
        fwd.ctypes.data_as(C.c_void_p), rev.ctypes.data_as(C.c_void_p),
        wf.ctypes.data_as(C.c_void_p), wr.ctypes.data_as(C.c_void_p),
        scanned.ctypes.data_as(C.c_void_p), C.byref(outside), C.byref(nrec), C.byref(nck), err,
        C.c_size_t(512))
    if rc:
        raise ValueError(err.value.decode())
    return dict(names=names, gene_ids=gene_ids, fwd=fwd[:tot], rev=rev[:tot], wfwd=wf[:tot],
                wrev=wr[:tot], scanned=scanned[:len(names)], outside=outside.value,
                n_records=nrec.value, n_chunks=nck.value)


def raw_counts_chunks(path, intervals, gene_intervals, mode, chunk_lo, chunk_hi, nthreads=1):
    """dense counts over the chunks [chunk_lo, chunk_hi) only -- what one shard of a sharded job owns"""
    return raw_counts(path, None, intervals, gene_intervals, mode, nthreads=nthreads, chunk_lo=chunk_lo, chunk_hi=chunk_hi)


def _assemble_unstranded(r):
    out = {}
    base = 0
    total_all = 0
    for ci, c in enumerate(r["names"]):
        ids = r["gene_ids"][ci]
        if r["scanned"][ci]:
            cnts = r["fwd"][base:base + len(ids)]
            last = {}
            for g, gid in enumerate(ids):
                last[gid] = int(cnts[g])
            for gid, v in last.items():
                out[gid] = out.get(gid, 0) + v
            tot = int(cnts.sum())
            out["_" + c] = out.get("_" + c, 0) + tot
            total_all += tot
            out.setdefault("_total", 0)
            out.setdefault("_outside", 0)
        base += len(ids)
    if "_total" in out:
        out["_total"] += total_all
        out["_outside"] += int(r["outside"])
    return out


def _assemble_dual(r, key_f, key_r, conv):
    f, v = {}, {}
    base = 0
    for ci, c in enumerate(r["names"]):
        ids = r["gene_ids"][ci]
        if r["scanned"][ci]:
            for d, arr in ((f, r[key_f]), (v, r[key_r])):
                cnts = arr[base:base + len(ids)]
                tot = conv(0)
                for g, gid in enumerate(ids):
                    d[gid] = d.get(gid, conv(0)) + conv(cnts[g])
                    tot += conv(cnts[g])
                d["_" + c] = d.get("_" + c, conv(0)) + tot
                d["_total"] = d.get("_total", conv(0)) + tot
            f["_outside"] = 0
        base += len(ids)
    if "_outside" in f:
        f["_outside"] = int(r["outside"])
    return f, v


def count_reads_unstranded(path, index_path, intervals, gene_intervals, each_read_counts_once=None,
                           nthreads=1):
    return _assemble_unstranded(raw_counts(path, index_path, intervals, gene_intervals, 0, nthreads,
                                           each_read_counts_once=bool(each_read_counts_once)))


def count_reads_stranded(path, index_path, intervals, gene_intervals, each_read_counts_once=None,
                         nthreads=1):
    return _assemble_dual(raw_counts(path, index_path, intervals, gene_intervals, 1, nthreads,
                                     each_read_counts_once=bool(each_read_counts_once)), "fwd", "rev", int)


def count_reads_weighted(path, index_path, intervals, gene_intervals, each_read_counts_once=None,
                         nthreads=1):
    if each_read_counts_once:
        raise ValueError("each_read_counts_once=True is not defined for count_reads_weighted")
    return _assemble_dual(raw_counts(path, index_path, intervals, gene_intervals, 2, nthreads),
                          "wfwd", "wrev", float)


def count_introns(path, index_path=None, nthreads=1, return_stats=False, sample_chunks=0):
    n = C.c_uint64(0)
    tid = C.POINTER(C.c_int32)()
    st = C.POINTER(C.c_uint32)()
    en = C.POINTER(C.c_uint32)()
    cnt = C.POINTER(C.c_uint64)()
    seen = np.zeros(1 << 16, dtype=np.uint8)
    nrec = C.c_uint64(0)
    nck = C.c_int64(0)
    err = C.create_string_buffer(512)
    L = lib()
    rc = L.oracle_count_introns(path.encode(), index_path.encode() if index_path else None,
                                C.c_int(nthreads), C.c_int64(sample_chunks), C.byref(nck), C.byref(n), C.byref(tid), C.byref(st), C.byref(en),
                                C.byref(cnt), seen.ctypes.data_as(C.c_void_p), C.c_int32(len(seen)),
                                C.byref(nrec), err, C.c_size_t(512))
    if rc:
        raise ValueError(err.value.decode())
    from synth.bamio import iter_bgzf_blocks  # noqa: F401  (header names via reader below)
    names = header_names(path)
    out = {names[t]: {} for t in range(len(names)) if seen[t]}
    for i in range(n.value):
        out[names[tid[i]]][(int(st[i]), int(en[i]))] = int(cnt[i])
    for p in (tid, st, en, cnt):
        L.oracle_free(p)
    if return_stats:
        return out, dict(n_records=nrec.value, n_chunks=nck.value)
    return out


def header_names(path):
    import struct
    import zlib
    data = open(path, "rb").read(1 << 22)
    u = b""
    off = 0
    while off + 18 <= len(data):
        csize = struct.unpack_from("<H", data, off + 16)[0] + 1
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        u += zlib.decompress(data[off + 12 + xlen:off + csize - 8], -15)
        off += csize
        if len(u) >= 12:
            l_text = struct.unpack_from("<i", u, 4)[0]
            if len(u) >= 12 + l_text:
                n_ref = struct.unpack_from("<i", u, 8 + l_text)[0]
                p = 12 + l_text
                names = []
                ok = True
                for _ in range(n_ref):
                    if len(u) < p + 4:
                        ok = False
                        break
                    ln = struct.unpack_from("<i", u, p)[0]
                    if len(u) < p + 8 + ln:
                        ok = False
                        break
                    names.append(u[p + 4:p + 4 + ln - 1].decode())
                    p += 8 + ln
                if ok:
                    return names
    raise ValueError("could not parse BAM header")


def chunks(path, index_path, gene_intervals):
    gv = None
    keep = None
    if gene_intervals is not None:
        gv, keep, _, _ = flatten(gene_intervals)
    cap = 1 << 20
    tid = np.zeros(cap, dtype=np.int32)
    st = np.zeros(cap, dtype=np.uint32)
    en = np.zeros(cap, dtype=np.uint32)
    n = lib().oracle_chunks(path.encode(), index_path.encode() if index_path else None,
                            C.byref(gv) if gv is not None else None, tid.ctypes.data_as(C.c_void_p),
                            st.ctypes.data_as(C.c_void_p), en.ctypes.data_as(C.c_void_p), C.c_int64(cap))
    if n < 0:
        raise ValueError("Could not read bam")
    return [(int(tid[i]), int(st[i]), int(en[i])) for i in range(n)]


def cigar_pairs(cigar_words, pos, want_introns):
    arr = np.asarray(cigar_words, dtype=np.uint32)
    out = np.zeros(2 * (len(arr) + 1), dtype=np.uint32)
    L = lib()
    L.oracle_cigar_pairs.restype = C.c_int32
    n = L.oracle_cigar_pairs(arr.ctypes.data_as(C.c_void_p), C.c_int32(len(arr)), C.c_int32(pos),
                             C.c_int32(1 if want_introns else 0), out.ctypes.data_as(C.c_void_p),
                             C.c_int32(len(arr) + 1))
    return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)]


def inflate_buffer(data: bytes, cap: int) -> bytes:
    out = np.zeros(cap + 1, dtype=np.uint8)
    src = np.frombuffer(data, dtype=np.uint8)
    n = lib().oracle_inflate_buffer(src.ctypes.data_as(C.c_void_p), C.c_uint64(len(data)),
                                    out.ctypes.data_as(C.c_void_p), C.c_uint64(cap))
    if n < 0:
        raise ValueError("oracle inflate failed")
    return out[:n].tobytes()


def quantify_gene_reads(path, index_path, intervals, gene_intervals, nthreads=1):
    """reference src/lib.rs:247-264: the C restatement's entries, assembled like quantify.rs:148-203 (per contig the last
    gene_no of a duplicated id wins; equal ids on two contigs are merged, a shared position is the reference's panic)."""
    iv, keep1, names, gene_ids = flatten(intervals)
    gv, keep2, _, _ = flatten(gene_intervals)
    n = C.c_uint64(0)
    bucket = C.POINTER(C.c_int32)()
    gene = C.POINTER(C.c_uint32)()
    pos = C.POINTER(C.c_uint32)()
    cnt = C.POINTER(C.c_uint64)()
    scanned = np.zeros(len(names) + 1, dtype=np.uint8)
    err = C.create_string_buffer(512)
    L = lib()
    rc = L.oracle_quantify(path.encode(), index_path.encode() if index_path else None, C.byref(iv), C.byref(gv), C.c_int(nthreads),
                           C.byref(n), C.byref(bucket), C.byref(gene), C.byref(pos), C.byref(cnt),
                           scanned.ctypes.data_as(C.c_void_p), err, C.c_size_t(512))
    if rc:
        raise ValueError(err.value.decode())
    per = ({}, {})
    for i in range(n.value):
        per[bucket[i]].setdefault(int(gene[i]), []).append((int(pos[i]), int(cnt[i])))
    for p in (bucket, gene, pos, cnt):
        L.oracle_free(p)
    out = ({}, {})
    base = 0
    for ci, c in enumerate(names):
        ids = gene_ids[ci]
        if scanned[ci]:
            last = {gid: g for g, gid in enumerate(ids)}
            for b in (0, 1):
                for gid, g in last.items():
                    lst = per[b].get(base + g, [])
                    tgt = out[b].setdefault(gid, {})
                    for p_, n_ in lst:
                        if p_ in tgt:
                            raise ValueError("still not disjoint")
                        tgt[p_] = n_
        base += len(ids)
    return tuple({g: sorted(d.items()) for g, d in o.items()} for o in out)


def calculate_duplicate_distribution(path, index_path=None, nthreads=1):
    """reference src/lib.rs:108-116 through the C restatement."""
    n = C.c_uint64(0)
    k = C.POINTER(C.c_uint32)()
    cnt = C.POINTER(C.c_uint64)()
    nrec = C.c_uint64(0)
    err = C.create_string_buffer(512)
    L = lib()
    rc = L.oracle_duplicate_distribution(path.encode(), index_path.encode() if index_path else None, C.c_int(nthreads),
                                         C.byref(n), C.byref(k), C.byref(cnt), C.byref(nrec), err, C.c_size_t(512))
    if rc:
        raise ValueError(err.value.decode())
    out = {int(k[i]): int(cnt[i]) for i in range(n.value)}
    for p in (k, cnt):
        L.oracle_free(p)
    return out


def count_reads_primary_only_right_strand_only_by_barcode(path, index_path, intervals, gene_intervals, umi_strategy="straight",
                                                          nthreads=1):
    """reference src/lib.rs:184-211 through the C restatement; entries ordered (chunk, gene, barcode) and indices handed
    out in that order, like py_oracle and the product (the reference's order is HashMap iteration)."""
    if umi_strategy != "straight":
        raise ValueError("invalid umi_strategy")
    iv, keep1, names, gene_ids = flatten(intervals)
    gv, keep2, _, _ = flatten(gene_intervals)
    n = C.c_uint64(0)
    gene = C.POINTER(C.c_uint32)()
    chunk = C.POINTER(C.c_uint32)()
    cnt = C.POINTER(C.c_uint32)()
    off = C.POINTER(C.c_uint64)()
    byt = C.POINTER(C.c_char)()
    err = C.create_string_buffer(512)
    L = lib()
    rc = L.oracle_by_barcode(path.encode(), index_path.encode() if index_path else None, C.byref(iv), C.byref(gv), C.c_int(nthreads),
                             C.byref(n), C.byref(gene), C.byref(chunk), C.byref(cnt), C.byref(off), C.byref(byt), err,
                             C.c_size_t(512))
    if rc:
        raise ValueError(err.value.decode())
    flat_ids = [gid for ids in gene_ids for gid in ids]
    entries = []
    for i in range(n.value):
        entries.append((int(chunk[i]), int(gene[i]), C.string_at(C.addressof(byt.contents) + off[i], off[i + 1] - off[i]), int(cnt[i])))
    for p in (gene, chunk, cnt, off, byt):
        if p:
            L.oracle_free(p)
    entries.sort()
    genes, barcodes, counts = {}, {}, []
    for _ck, g, bc, c in entries:
        gi = genes.setdefault(flat_ids[g], len(genes))
        bi = barcodes.setdefault(bc.decode(), len(barcodes))
        counts.append((gi, bi, c))
    return list(genes), list(barcodes), counts
