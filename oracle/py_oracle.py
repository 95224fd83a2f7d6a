"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of mbf_bam's counters.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this.  It is the slow, obviously-correct statement of the semantics; the C
restatement (oracle/mbf_oracle.c) is checked against it on small inputs and is
what the GPU path is compared with at larger sizes.

Pinning status:
  * cigar_blocks / cigar_introns: PINNED by the reference's 58 golden records
    (reference src/bam_ext.rs:50-382 and :384-604 -> tests/golden/cigar_golden.json).
  * count_reads_unstranded / count_reads_stranded / count_introns end to end:
    PARITY UNPINNED by the reference's own tests (src/lib.rs:300-301 says the
    tests live in the callers); the Rust crate cannot be built here (no cargo).
    They restate src/count_reads/counters.rs, chunked_genome.rs, introns.rs
    line by line (citations on each function).
  * count_reads_weighted: does not exist in the reference (src/lib.rs:288-295);
    defined here -- PARITY UNPINNED.

Third-party behaviour restated from the published APIs (sources absent):
rust-htslib 0.24.0 (fetch/read/aux/aligned_blocks/introns) and bio 0.25.0
IntervalTree::find (half-open overlap).
"""
from __future__ import annotations

import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synth.bamio import aux_get_int, read_bam  # noqa: E402

CHUNK_SIZE = 1_000_000  # reference src/count_reads/chunked_genome.rs:75
U32 = 0xFFFFFFFF


def cigar_blocks(pos, cigar):
    """rust-htslib aligned_blocks via reference src/bam_ext.rs:29-34: M/=/X emit, D/N advance."""
    out = []
    p = pos
    for op, ln in cigar:
        if op in (0, 7, 8):
            out.append((p & U32, (p + ln) & U32))
            p += ln
        elif op in (2, 3):
            p += ln
    return out


def cigar_introns(pos, cigar):
    """rust-htslib introns via reference src/bam_ext.rs:36-41: N emits, M/=/X/D advance."""
    out = []
    p = pos
    for op, ln in cigar:
        if op == 3:
            out.append((p & U32, (p + ln) & U32))
            p += ln
        elif op in (0, 2, 7, 8):
            p += ln
    return out


def build_intervals(lst):
    """reference src/count_reads/mod.rs:27-45: gene_no = list index; zip(starts, stops)."""
    ivs = []
    gene_ids = []
    for gene_no, (gid, strand, starts, stops) in enumerate(lst):
        gene_ids.append(gid)
        for s, e in zip(starts, stops):
            if s > e:
                raise ValueError("interval start > end")
            ivs.append((s, e, gene_no, strand))
    return ivs, gene_ids


def find(ivs, b0, b1):
    """bio IntervalTree::find: overlap iff iv.start < b1 and b0 < iv.end (order-free here)."""
    return [iv for iv in ivs if iv[0] < b1 and b0 < iv[1]]


class BioTree:
    """bio 0.25.0 data_structures::interval_tree::IntervalTree (Cargo.lock:100-101; the crate's source is not
    under /root/reference -- restated from its published algorithm; PARITY UNPINNED).

    An AVL tree keyed on interval.start.  insert(): descend LEFT when new.start <= node.start, else right;
    every node on the way back up is repaired (height/max update, or the usual single/double rotation).
    find(): explicit stack seeded with the root; pop a node, if query.start < node.max push its left child
    and, if query.end > node.start, push its right child and yield the node when it overlaps.  The stack is
    LIFO, so hits come in the order: node, right subtree, left subtree.

    Only two things in the counters depend on that order: each_read_counts_once=True (first hit wins,
    reference counters.rs:83-92,173-182) and which covering gene moves a chunk edge when several do
    (chunked_genome.rs:98)."""

    class Node:
        __slots__ = ("start", "end", "value", "max", "height", "left", "right")

        def __init__(self, start, end, value):
            self.start, self.end, self.value = start, end, value
            self.max = end
            self.height = 1
            self.left = self.right = None

    def __init__(self):
        self.root = None

    @staticmethod
    def _h(n):
        return n.height if n is not None else 0

    @classmethod
    def _update(cls, n):
        n.height = 1 + max(cls._h(n.left), cls._h(n.right))
        m = n.end
        if n.left is not None and n.left.max > m:
            m = n.left.max
        if n.right is not None and n.right.max > m:
            m = n.right.max
        n.max = m

    @classmethod
    def _rot_right(cls, n):
        l = n.left
        n.left = l.right
        l.right = n
        cls._update(n)
        cls._update(l)
        return l

    @classmethod
    def _rot_left(cls, n):
        r = n.right
        n.right = r.left
        r.left = n
        cls._update(n)
        cls._update(r)
        return r

    @classmethod
    def _repair(cls, n):
        lh, rh = cls._h(n.left), cls._h(n.right)
        if abs(lh - rh) <= 1:
            cls._update(n)
            return n
        if rh > lh:
            if cls._h(n.right.left) > cls._h(n.right.right):
                n.right = cls._rot_right(n.right)
            return cls._rot_left(n)
        if cls._h(n.left.right) > cls._h(n.left.left):
            n.left = cls._rot_left(n.left)
        return cls._rot_right(n)

    def insert(self, start, end, value):
        if start > end:
            raise ValueError("interval start > end")  # bio panics
        new = BioTree.Node(start, end, value)
        if self.root is None:
            self.root = new
            return
        path = []
        n = self.root
        while True:
            left = start <= n.start
            path.append((n, left))
            nxt = n.left if left else n.right
            if nxt is None:
                if left:
                    n.left = new
                else:
                    n.right = new
                break
            n = nxt
        for i in range(len(path) - 1, -1, -1):  # the recursion's `self.repair()` on the way back up
            node, _ = path[i]
            rep = BioTree._repair(node)
            if i == 0:
                self.root = rep
            else:
                parent, went_left = path[i - 1]
                if went_left:
                    parent.left = rep
                else:
                    parent.right = rep

    def find(self, q0, q1):
        """yields (start, end, value) of overlapping intervals in bio's iteration order"""
        stack = [self.root] if self.root is not None else []
        while stack:
            n = stack.pop()
            if q0 < n.max:
                if n.left is not None:
                    stack.append(n.left)
                if q1 > n.start:
                    if n.right is not None:
                        stack.append(n.right)
                    if q0 < n.end and n.start < q1:
                        yield (n.start, n.end, n.value)


def build_tree(ivs):
    """reference src/count_reads/mod.rs:27-45: intervals are inserted in list order."""
    t = BioTree()
    for s, e, gene_no, strand in ivs:
        t.insert(s, e, (gene_no, strand))
    return t


def make_chunks(contigs, gene_sets):
    """reference chunked_genome.rs:72-119. contigs: [(name, tid, length)]; gene_sets: name->ivs or None.
    The edge moves past the FIRST interval bio's find(stop..stop+1) yields (chunked_genome.rs:98)."""
    chunks = []
    for name, tid, length in contigs:
        tree = build_tree(gene_sets[name]) if gene_sets is not None else None
        start = 0
        while True:
            stop = start + CHUNK_SIZE
            while tree is not None:
                first = next(tree.find(stop, stop + 1), None)
                if first is None:
                    break
                stop = first[1] + 1
            chunks.append((name, tid, start, stop))
            start = stop
            if start >= length:
                break
    return chunks


def _nh(rec):
    found, val = aux_get_int(rec.aux, b"NH")
    if not found:
        return 1  # counters.rs:66 map_or(1, ...)
    if val is None:
        raise ValueError("NH tag is not an integer")
    return val


def _count(path, intervals, gene_intervals, mode, each_read_counts_once=False):
    hdr, recs = read_bam(path)
    name2tid = {n: i for i, (n, _) in enumerate(hdr.refs)}
    trees = {c: build_intervals(l) for c, l in intervals.items()}
    gene_trees = {c: build_intervals(l) for c, l in gene_intervals.items()}
    contigs = [(c, name2tid[c], hdr.refs[name2tid[c]][1]) for c in gene_trees if c in name2tid]
    for c, _, _ in contigs:
        if c not in trees:
            raise ValueError("contig %r is in gene_intervals but not in intervals" % c)
    chunks = make_chunks(contigs, {c: gene_trees[c][0] for c, _, _ in contigs})
    by_tid = defaultdict(list)
    for r in recs:
        by_tid[r.tid].append(r)
    fwd = defaultdict(int)
    rev = defaultdict(int)
    weighted = mode == "weighted"
    outside_total = 0
    once_trees = {}
    for chr_, tid, start, stop in chunks:
        ivs, gene_ids = trees[chr_]
        if each_read_counts_once and chr_ not in once_trees:
            once_trees[chr_] = build_tree(ivs)
        n = len(gene_ids)
        res = [[0] * n, [0] * n]
        mm = [defaultdict(set), defaultdict(set)]
        outside = 0
        for r in by_tid.get(tid, ()):
            upos = r.pos & U32
            if upos < start or upos >= stop:      # counters.rs:46-48 / :137-139
                continue
            seen = [set(), set()]
            hit = False
            for b0, b1 in cigar_blocks(r.pos, r.cigar):
                if mode == "unstranded" and b0 >= stop:  # counters.rs:52 (live clause)
                    continue
                if each_read_counts_once:   # hits in bio's order; the loop below breaks after the first
                    hits = [(s, e, v[0], v[1]) for s, e, v in once_trees[chr_].find(b0, b1)]
                else:
                    hits = find(ivs, b0, b1)
                for (s, e, gene_no, strand) in hits:
                    hit = True
                    nh = _nh(r)
                    if mode == "unstranded":
                        bucket = 0
                    else:                          # counters.rs:151
                        rv = bool(r.flag & 0x10)
                        bucket = 0 if ((strand == 1 and not rv) or (strand != 1 and rv)) else 1
                    if weighted or nh == 1:
                        seen[bucket].add(gene_no)
                    else:
                        mm[bucket][gene_no].add(r.qname)
                    if each_read_counts_once:
                        break                      # counters.rs:83-85 / :173-175
                if each_read_counts_once and hit:
                    break                          # counters.rs:87-92 / :177-182
            if not hit:
                outside += 1
            if weighted:
                nh = _nh(r) if hit else 1
                w = 1.0 / nh if nh >= 1 else 1.0
            else:
                w = 1
            for b in (0, 1):
                for g in seen[b]:
                    res[b][g] += w
        for b in (0, 1):
            for g, s in mm[b].items():
                res[b][g] += len(s)
        # result assembly: counters.rs:236-254 (unstranded) / :262-273,:309-314 (stranded)
        if mode == "unstranded":
            d = {}
            total = 0
            for g, cnt in enumerate(res[0]):
                d[gene_ids[g]] = cnt              # insert: last duplicate id wins
                total += cnt
            d["_total"] = total
            d["_outside"] = outside
            d["_" + chr_] = total
            for k, v in d.items():
                fwd[k] += v
        else:
            for b, tgt in ((0, fwd), (1, rev)):
                d = defaultdict(int)
                total = 0
                for g, cnt in enumerate(res[b]):
                    d[gene_ids[g]] += cnt
                    total += cnt
                d["_total"] = total
                d["_" + chr_] = total
                for k, v in d.items():
                    tgt[k] += v
            fwd["_outside"] += outside
        outside_total += outside
    if mode == "unstranded":
        return dict(fwd)
    return dict(fwd), dict(rev)


def count_reads_unstranded(path, index_path, intervals, gene_intervals, each_read_counts_once=None):
    """reference src/lib.rs:138-157 -> counters.rs:204-260, :26-106."""
    return _count(path, intervals, gene_intervals, "unstranded", bool(each_read_counts_once))


def count_reads_stranded(path, index_path, intervals, gene_intervals, each_read_counts_once=None):
    """reference src/lib.rs:161-181 -> counters.rs:276-323, :114-202."""
    return _count(path, intervals, gene_intervals, "stranded", bool(each_read_counts_once))


def count_reads_weighted(path, index_path, intervals, gene_intervals, each_read_counts_once=None):
    """NOT IN THE REFERENCE (north_star names it).  Definition: the stranded skeleton
    (no block filter, strand rule of counters.rs:151); every record adds 1/NH (NH
    missing or < 1 -> 1) once per distinct gene it hits, f64; no qname dedup;
    `_total`/`_<chr>` are the float sums, `_outside` stays an integer count."""
    return _count(path, intervals, gene_intervals, "weighted", bool(each_read_counts_once))


def count_introns(path, index_path=None):
    """reference src/lib.rs:216-225 -> introns.rs:56-88, :24-51."""
    hdr, recs = read_bam(path)
    contigs = [(n, i, l) for i, (n, l) in enumerate(hdr.refs)]
    chunks = make_chunks(contigs, None)
    by_tid = defaultdict(list)
    for r in recs:
        by_tid[r.tid].append(r)
    out = {}
    for chr_, tid, start, stop in chunks:
        d = out.setdefault(chr_, {})
        for r in by_tid.get(tid, ()):
            upos = r.pos & U32
            if upos < start or upos >= stop:
                continue
            k = 0
            for s, e in cigar_introns(r.pos, r.cigar):
                d[(s, e)] = d.get((s, e), 0) + 1
                k += 1
            key = (U32, U32 - k)
            d[key] = d.get(key, 0) + 1
    return out


def quantify_gene_reads(path, index_path, intervals, gene_intervals):
    """reference src/lib.rs:247-264 -> src/count_reads/quantify.rs:20-75 (driver), :98-146 (per chunk).

    ({gene_id: [(pos, n_reads_starting_there)]} congruent with the gene's strand, same for incongruent).
    No NH logic, any flag; one alignment counts once per gene (quantify.rs:121-125).  Every gene of a scanned
    contig gets a (possibly empty) list (to_hashmap, quantify.rs:148-157: a duplicated id within one contig keeps
    the LAST gene_no).  The order inside a list is HashMap order in the reference; sorted by position here.
    add_hashmaps (quantify.rs:189-203) panics on a (gene_id, pos) seen twice, which only equal ids on two
    contigs can cause: ValueError here.  PARITY UNPINNED (no test in the reference)."""
    hdr, recs = read_bam(path)
    name2tid = {n: i for i, (n, _) in enumerate(hdr.refs)}
    trees = {c: build_intervals(l) for c, l in intervals.items()}
    gene_trees = {c: build_intervals(l) for c, l in gene_intervals.items()}
    contigs = [(c, name2tid[c], hdr.refs[name2tid[c]][1]) for c in gene_trees if c in name2tid]
    for c, _, _ in contigs:
        if c not in trees:
            raise ValueError("contig %r is in gene_intervals but not in intervals" % c)
    chunks = make_chunks(contigs, {c: gene_trees[c][0] for c, _, _ in contigs})
    by_tid = defaultdict(list)
    for r in recs:
        by_tid[r.tid].append(r)
    out = ({}, {})
    for chr_, tid, start, stop in chunks:
        ivs, gene_ids = trees[chr_]
        res = ([dict() for _ in gene_ids], [dict() for _ in gene_ids])
        for r in by_tid.get(tid, ()):
            upos = r.pos & U32
            if upos < start or upos >= stop:          # quantify.rs:113
                continue
            seen = set()
            for b0, b1 in cigar_blocks(r.pos, r.cigar):
                for (s, e, gene_no, strand) in find(ivs, b0, b1):
                    if gene_no in seen:
                        continue
                    seen.add(gene_no)
                    rv = bool(r.flag & 0x10)
                    congruent = (strand == 1 and not rv) or (strand != 1 and rv)   # quantify.rs:126-127
                    d = res[0 if congruent else 1][gene_no]
                    d[upos] = d.get(upos, 0) + 1
        for b in (0, 1):
            per_chunk = {}
            for g, d in enumerate(res[b]):
                per_chunk[gene_ids[g]] = d            # collect(): the last duplicate id wins
            for gid, d in per_chunk.items():
                tgt = out[b].setdefault(gid, {})
                for pos, n in d.items():
                    if pos in tgt:
                        raise ValueError("still not disjoint")
                    tgt[pos] = n
    return tuple({gid: sorted(d.items()) for gid, d in o.items()} for o in out)


def calculate_duplicate_distribution(path, index_path=None):
    """reference src/lib.rs:108-116 -> src/duplicate_distribution.rs:44-86.

    {k: number of distinct (contig, pos, is_reverse, SEQ) that occur exactly k times}.  The reference walks
    fetch(tid, 0, target_len) of every contig and closes a group whenever `pos` changes (:63-71); in a coordinate
    sorted file that equals grouping by (tid, pos).  seq().as_bytes() decodes l_seq bases, so the unused low
    nibble of an odd-length SEQ is not part of the key.  PARITY UNPINNED."""
    hdr, recs = read_bam(path)
    groups = defaultdict(int)
    for r in recs:
        if r.tid < 0 or r.tid >= len(hdr.refs):
            continue
        if r.pos < 0 or r.pos >= hdr.refs[r.tid][1]:      # not returned by fetch(tid, 0, len)
            continue
        seq = r.seq if r.seq is not None else bytes((r.l_seq + 1) // 2)
        if r.l_seq & 1:
            seq = seq[:-1] + bytes([seq[-1] & 0xF0])
        groups[(r.tid, r.pos, bool(r.flag & 0x10), r.l_seq, seq)] += 1
    out = defaultdict(int)
    for v in groups.values():
        out[v] += 1
    return dict(out)


def aux_get_str(aux, tag):
    """(found, bytes or None if the tag is not of type Z/H) -- rust-htslib Aux::String."""
    import struct
    p = 0
    sizes = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4, b"d": 8}
    while p + 3 <= len(aux):
        t = aux[p:p + 2]
        ty = aux[p + 2:p + 3]
        p += 3
        if ty in (b"Z", b"H"):
            e = aux.index(b"\0", p)
            if t == tag:
                return True, aux[p:e]      # rust-htslib 0.24 aux(): Z and H both give Aux::String
            p = e + 1
        elif ty == b"B":
            sub = aux[p:p + 1]
            cnt = struct.unpack_from("<i", aux, p + 1)[0]
            if t == tag:
                return True, None
            p += 5 + cnt * sizes[sub]
        else:
            if t == tag:
                return True, None
            p += sizes[ty]
    return False, None


def count_reads_primary_only_right_strand_only_by_barcode(path, index_path, intervals, gene_intervals,
                                                          umi_strategy="straight"):
    """reference src/lib.rs:184-211 -> src/count_reads/by_barcode.rs:18-92 (driver), :101-182 (per chunk).

    (genes, barcodes, [(gene_idx, barcode_idx, count)]): per chunk and (gene_no, XC barcode) the number of
    distinct XM UMIs summed over (pos, is_reverse) (:164-176), only reads that are not secondary and lie on the
    gene's strand.  A chunk in which a right-strand hit lacks a string XC/XM tag returns Err and contributes
    NOTHING (:46-68 `_ => {}`).  Index assignment and entry order follow HashMap iteration in the reference, i.e.
    are arbitrary: here chunks in order, then (gene_no, barcode) sorted.  PARITY UNPINNED."""
    if umi_strategy != "straight":
        raise ValueError("invalid umi_strategy")
    hdr, recs = read_bam(path)
    name2tid = {n: i for i, (n, _) in enumerate(hdr.refs)}
    trees = {c: build_intervals(l) for c, l in intervals.items()}
    gene_trees = {c: build_intervals(l) for c, l in gene_intervals.items()}
    contigs = [(c, name2tid[c], hdr.refs[name2tid[c]][1]) for c in gene_trees if c in name2tid]
    for c, _, _ in contigs:
        if c not in trees:
            raise ValueError("contig %r is in gene_intervals but not in intervals" % c)
    chunks = make_chunks(contigs, {c: gene_trees[c][0] for c, _, _ in contigs})
    by_tid = defaultdict(list)
    for r in recs:
        by_tid[r.tid].append(r)
    flat = []
    for chr_, tid, start, stop in chunks:
        ivs, gene_ids = trees[chr_]
        positions = defaultdict(set)
        failed = False
        for r in by_tid.get(tid, ()):
            upos = r.pos & U32
            skipped = upos < start or upos >= stop
            if r.flag & 0x100:                       # by_barcode.rs:127
                continue
            if skipped:
                continue
            rv = bool(r.flag & 0x10)
            for b0, b1 in cigar_blocks(r.pos, r.cigar):
                if b1 < start or b0 >= stop or (b0 < start and b1 >= start):   # by_barcode.rs:131
                    continue
                for (s, e, gene_no, strand) in find(ivs, b0, b1):
                    if (strand == 1 and not rv) or (strand != 1 and rv):
                        okc, bc = aux_get_str(r.aux, b"XC")
                        if not okc or bc is None:
                            failed = True
                            break
                        oku, umi = aux_get_str(r.aux, b"XM")
                        if not oku or umi is None:
                            failed = True
                            break
                        positions[(gene_no, bc, r.pos, rv)].add(umi)
                if failed:
                    break
            if failed:
                break
        if failed:
            continue
        by_gb = defaultdict(int)
        for (gene_no, bc, _p, _r), umis in positions.items():
            by_gb[(gene_no, bc)] += len(umis)
        for (gene_no, bc), n in sorted(by_gb.items()):
            flat.append((gene_ids[gene_no], bc.decode(), n))
    genes, barcodes, counts = {}, {}, []
    for g, b, n in flat:
        gi = genes.setdefault(g, len(genes))
        bi = barcodes.setdefault(b, len(barcodes))
        counts.append((gi, bi, n))
    return list(genes), list(barcodes), counts
