"""Seeded gene models and bench-scale BAM generation (drives tools/bamgen).  Host-side tooling."""
from __future__ import annotations

import json
import os
import random
import subprocess
from typing import Dict, List, Tuple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# GRCh38 primary assembly lengths
HUMAN_CONTIGS = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
    ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
    ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
    ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
    ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415), ("chrM", 16569),
]


def make_gene_model(contigs: List[Tuple[str, int]], n_genes: int, seed: int, mean_exons: float = 3.0,
                    span=(1_000, 200_000), exon_len=(80, 2_000), gene_mode: bool = False):
    """-> (intervals, gene_intervals, genes) ; genes = [(contig_index, strand, starts, stops)].

    intervals = exons (or the gene span when gene_mode), gene_intervals = gene spans, both in the
    reference's argument format.  Gene-span ends are kept >= 2 apart so that the chunk table does not
    depend on bio's tree order."""
    rng = random.Random(seed)
    total = float(sum(l for _, l in contigs))
    intervals: Dict[str, list] = {n: [] for n, _ in contigs}
    gene_intervals: Dict[str, list] = {n: [] for n, _ in contigs}
    genes = []
    used_ends = [set() for _ in contigs]
    # genes per contig proportional to length (at least one on every contig that can hold one)
    alloc = [max(1, int(round(n_genes * l / total))) if l > 2_000 else 0 for _, l in contigs]
    while sum(alloc) > n_genes:
        alloc[alloc.index(max(alloc))] -= 1
    while sum(alloc) < n_genes:
        alloc[alloc.index(max(alloc))] += 1
    gid = 0
    for ci, (name, length) in enumerate(contigs):
        for _ in range(alloc[ci]):
            lo, hi = span
            hi = min(hi, max(lo + 1, length // 4))
            sp = int(lo * (hi / lo) ** rng.random())
            sp = min(sp, length - 2)
            gs = rng.randrange(0, max(1, length - sp))
            ge = gs + sp
            while ge in used_ends[ci] or ge - 1 in used_ends[ci] or ge + 1 in used_ends[ci]:
                ge += 2
            if ge > length:
                ge = length
            used_ends[ci].add(ge)
            strand = 1 if rng.random() < 0.5 else -1
            if gene_mode:
                starts, stops = [gs], [ge]
            else:
                n_ex = 1 + min(7, int(rng.expovariate(1.0 / max(0.01, mean_exons - 1.0))))
                starts, stops = [], []
                p = gs
                for k in range(n_ex):
                    el = rng.randint(*exon_len)
                    if p + el > ge:
                        break
                    starts.append(p)
                    stops.append(p + el)
                    # intron up to what is left, log-uniform-ish
                    left = ge - (p + el)
                    if left < 200:
                        break
                    p = p + el + rng.randint(70, max(71, left // max(1, (n_ex - k))))
                if not starts:
                    starts, stops = [gs], [min(ge, gs + max(1, sp))]
            g_id = "G%06d" % gid
            gid += 1
            intervals[name].append((g_id, strand, starts, stops))
            gene_intervals[name].append((g_id, strand, [gs], [ge]))
            genes.append((ci, strand, starts, stops))
    return intervals, gene_intervals, genes


def write_model(path: str, contigs, genes):
    with open(path, "w") as f:
        f.write("C %d\n" % len(contigs))
        for n, l in contigs:
            f.write("%s %d\n" % (n, l))
        f.write("G %d\n" % len(genes))
        for ci, strand, starts, stops in genes:
            f.write("%d %d %d %s\n" % (ci, strand, len(starts), " ".join("%d %d" % se for se in zip(starts, stops))))


def bamgen_path() -> str:
    exe = os.path.join(ROOT, "tools", "bamgen")
    src = exe + ".c"
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    return exe


def generate_bam(out: str, contigs, genes, reads: int, seed: int, readlen: int = 100, threads: int = 0,
                 level: int = 6, in_gene: float = 0.7, spliced: float = 0.0, mm: float = 0.0,
                 paired: bool = False, edge: float = 0.01) -> dict:
    model = out + ".model.txt"
    write_model(model, contigs, genes)
    cmd = [bamgen_path(), "--genes", model, "--out", out, "--reads", str(reads), "--readlen", str(readlen),
           "--seed", str(seed), "--threads", str(threads or os.cpu_count() or 4), "--level", str(level),
           "--in-gene", str(in_gene), "--spliced", str(spliced), "--mm", str(mm), "--edge", str(edge)]
    if paired:
        cmd.append("--paired")
    res = subprocess.run(cmd, capture_output=True, text=True, check=True)
    return json.loads(res.stdout.strip().splitlines()[-1])


# BASELINE.json `configs`; `reads` may be scaled down by the caller
# `spliced`: fraction of the single-end reads the generator TRIES to move onto an exon junction (tools/bamgen.c
# snap_to_junction); about 53 % of them have one within reach, so 0.66 gives the 35 % spliced reads of SURVEY App. E
# (cfg3) and 0.38 the 20 % of cfg5.  Junctions are shared by many reads, as in RNA-seq data.
CONFIGS = {
    "cfg1": dict(contigs=[("chrA", 5_000_000), ("chrB", 3_000_000)], n_genes=1000, gene_mode=True, span=(1_000, 20_000),
                 reads=1_000_000, readlen=50, seed=1, mm=0.10, spliced=0.0, in_gene=0.5, fn="count_reads_unstranded"),
    "cfg2": dict(contigs=HUMAN_CONTIGS, n_genes=20_000, gene_mode=False, span=(1_000, 200_000),
                 reads=50_000_000, readlen=100, seed=2, mm=0.05, spliced=0.0, in_gene=0.7, fn="count_reads_stranded"),
    "cfg3": dict(contigs=HUMAN_CONTIGS, n_genes=20_000, gene_mode=False, span=(1_000, 200_000),
                 reads=100_000_000, readlen=100, seed=3, mm=0.20, spliced=0.66, in_gene=0.7, fn="count_reads_weighted+count_introns"),
    "cfg4": dict(contigs=HUMAN_CONTIGS, n_genes=20_000, gene_mode=False, span=(1_000, 200_000),
                 reads=200_000_000, readlen=150, seed=4, mm=0.0, spliced=0.10, in_gene=0.7, paired=True, fn="count_reads_stranded"),
    "cfg5": dict(contigs=HUMAN_CONTIGS, n_genes=20_000, gene_mode=False, span=(1_000, 200_000),
                 reads=400_000_000, readlen=100, seed=5, mm=0.10, spliced=0.38, in_gene=0.7, fn="full suite"),
}


def build_config(name: str, out: str, reads: int = 0, threads: int = 0, level: int = 6):
    """Generate the BAM of a named config -> (intervals, gene_intervals, info)."""
    c = CONFIGS[name]
    intervals, gene_intervals, genes = make_gene_model(c["contigs"], c["n_genes"], c["seed"], gene_mode=c["gene_mode"], span=c["span"])
    info = generate_bam(out, c["contigs"], genes, reads or c["reads"], c["seed"], readlen=c["readlen"], threads=threads,
                        level=level, in_gene=c["in_gene"], spliced=c["spliced"], mm=c["mm"], paired=c.get("paired", False))
    return intervals, gene_intervals, info
