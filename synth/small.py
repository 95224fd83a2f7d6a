"""Small seeded test cases (records + gene models) covering the edge cases of
SURVEY.md App. A.6; used by the parity tests.  Host-side tooling only."""
from __future__ import annotations

import random
from typing import Dict, List, Tuple

from .bamio import Rec, aux_array, aux_int, aux_str, parse_cigar


def gene_model(rng: random.Random, contigs: List[Tuple[str, int]], genes_per_contig: int,
               exon_mode: bool = True, prefix: str = "G"):
    """Return (intervals, gene_intervals) in the reference's argument format
    {chr: [(gene_id, strand, [starts], [stops])]} (reference src/count_reads/mod.rs:27-45)."""
    intervals: Dict[str, list] = {}
    gene_intervals: Dict[str, list] = {}
    gid = 0
    for name, length in contigs:
        ex_list, g_list = [], []
        for _ in range(genes_per_contig):
            span = min(length - 2, int(10 ** rng.uniform(3, 4.6)))
            gs = rng.randrange(0, max(1, length - span))
            ge = gs + span
            strand = rng.choice([1, -1, 1, -1, 0])
            n_ex = rng.randint(1, 5)
            cuts = sorted(rng.sample(range(gs, ge), min(2 * n_ex, ge - gs)))
            starts = cuts[0::2]
            stops = cuts[1::2]
            starts, stops = starts[:len(stops)], stops[:len(starts)]
            if not starts:
                starts, stops = [gs], [ge]
            g_id = "%s%05d" % (prefix, gid)
            gid += 1
            ex_list.append((g_id, strand, starts, stops))
            g_list.append((g_id, strand, [gs], [ge]))
        intervals[name] = ex_list if exon_mode else g_list
        gene_intervals[name] = g_list
    return intervals, gene_intervals


_EDGE_CIGARS = ["10S40M", "20M5I25M", "20M5D30M", "25=25X", "5H45M5H", "10M0N40M", "50M0M",
                "20M1000N30M", "10M500N20M800N20M", "15M3D10M200N25M", "1M2P49M", "0M50M"]


def reads(rng: random.Random, contigs: List[Tuple[str, int]], n: int, read_len: int = 50,
          nh_frac: float = 0.2, spliced_frac: float = 0.2, edge_frac: float = 0.1,
          paired: bool = False, hot: List[Tuple[int, int, int]] = (), random_seq: bool = False) -> List[Rec]:
    """n coordinate-sorted records.  `hot`: (tid, lo, hi) windows that get half of the reads."""
    recs: List[Rec] = []
    qn = 0
    i = 0
    while i < n:
        if hot and rng.random() < 0.5:
            tid, lo, hi = rng.choice(list(hot))
            pos = rng.randrange(lo, hi)
        else:
            tid = rng.randrange(len(contigs))
            pos = rng.randrange(0, max(1, contigs[tid][1] - read_len))
        x = rng.random()
        if x < edge_frac:
            cig = parse_cigar(rng.choice(_EDGE_CIGARS))
        elif x < edge_frac + spliced_frac:
            a = rng.randint(1, read_len - 1)
            gap = int(10 ** rng.uniform(1.8, 4.7))
            cig = [(0, a), (3, gap), (0, read_len - a)]
        else:
            cig = [(0, read_len)]
        flag = rng.choice([0, 16])
        if rng.random() < 0.02:
            flag |= 0x100
        r = rng.random()
        name = b"q%011d" % qn
        qn += 1
        copies = 1
        if r < nh_frac:
            nh = rng.choice([2, 2, 3, 4, 0, 255])
            copies = rng.choice([1, 2, 2, 3])
            ty = rng.choice("cCsSiI")
            aux = aux_int(b"NH", nh if ty != "c" or nh < 128 else 100, ty)
        elif r < nh_frac + 0.3:
            aux = aux_int(b"NH", 1, rng.choice("cCsSiI"))
        elif r < nh_frac + 0.4:
            aux = b""  # NH absent -> 1
        else:
            aux = (aux_str(b"MD", b"50") + aux_array(b"ZB", "S", [1, 2, 3]) + aux_int(b"AS", 98, "C")
                   + aux_int(b"NH", 1, "C") + aux_int(b"HI", 1, "C"))
        for c in range(copies):
            p2 = pos if c == 0 else max(0, pos + rng.randrange(-300, 300))
            t2 = tid
            if c and rng.random() < 0.3:
                t2 = rng.randrange(len(contigs))
                p2 = rng.randrange(0, max(1, contigs[t2][1] - read_len))
            recs.append(Rec(t2, p2, name, flag, 255 if copies == 1 else 3, list(cig), read_len, aux))
            i += 1
        if paired and rng.random() < 0.5:
            mate_pos = pos + rng.randrange(100, 400)
            recs.append(Rec(tid, mate_pos, name, flag ^ 16 | 0x80 | 1, 255, [(0, read_len)], read_len, aux))
            i += 1
    if random_seq:  # incompressible SEQ/QUAL like real data (keeps BGZF members large)
        for r in recs:
            r.seq = rng.randbytes((r.l_seq + 1) // 2)
            r.qual = bytes(rng.choice((37, 37, 37, 25, 11, 2)) for _ in range(r.l_seq))
    # a few structural edge cases
    if n >= 20:
        t0 = 0
        recs.append(Rec(t0, 5, b"unmapped_with_pos", 4, 0, [], read_len, b""))          # flag 4, no cigar
        recs.append(Rec(t0, -1, b"negpos", 4, 0, [], 0, b""))                           # pos -1
        recs.append(Rec(t0, 1_000_000 - 10, b"straddle", 0, 255, [(0, 5), (3, 200), (0, 45)], 50,
                        aux_int(b"NH", 1)))                                              # 2nd block >= stop
        recs.append(Rec(-1, -1, b"nocoord", 4, 0, [], read_len, b""))
    recs.sort(key=lambda r: (r.tid if r.tid >= 0 else 1 << 30, r.pos))
    return recs


def simple_case(seed: int, n_reads: int = 2000, n_contigs: int = 2, contig_len: int = 2_500_000,
                genes_per_contig: int = 30, exon_mode: bool = True, paired: bool = False, random_seq: bool = False):
    rng = random.Random(seed)
    contigs = [("chr%s" % "ABCDEFGH"[i], contig_len + 137 * i) for i in range(n_contigs)]
    intervals, gene_intervals = gene_model(rng, contigs, genes_per_contig, exon_mode)
    hot = []
    for ci, (name, _) in enumerate(contigs):
        for g in gene_intervals[name][:10]:
            hot.append((ci, g[2][0], g[3][0]))
    recs = reads(rng, contigs, n_reads, paired=paired, hot=hot, random_seq=random_seq)
    return contigs, recs, intervals, gene_intervals
