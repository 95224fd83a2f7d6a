/* bamgen -- multi-threaded, seeded generator of coordinate-sorted synthetic BAM + BAI files in the
 * shapes named by BASELINE.json `configs`.  zlib + pthreads only.
 *
 * Not on the product path: it produces the inputs of bench.py and of the large parity tests.
 *
 *   bamgen --genes model.txt --out x.bam --reads N [--readlen L] [--seed S] [--threads T]
 *          [--level Z] [--in-gene F] [--spliced F] [--mm F] [--paired] [--edge F]
 *
 * model.txt (written by mbf_bam_b200/synth/model.py):
 *   C <n_contigs>            then n_contigs lines:  <name> <length>
 *   G <n_genes>              then per gene:         <contig_index> <strand> <n_exons> s0 e0 s1 e1 ...
 *
 * Output is deterministic for (model, arguments) and independent of --threads: the genome is cut
 * into 1 Mb bins, every bin draws from its own RNG stream, bins are compressed independently
 * (htslib style: records never straddle BGZF blocks, payload <= 0xff00) and concatenated.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define BIN_SIZE 1000000u
#define MAX_PAYLOAD 0xff00

typedef struct { char name[64]; uint32_t len; } contig_t;
typedef struct { uint32_t contig; int strand; uint32_t n_exons; uint32_t* s; uint32_t* e; } gene_t;
typedef struct { uint32_t s, e, gene, idx; } exon_t; /* sorted by start per contig; idx = exon number in gene */

static contig_t* contigs; static int n_contigs;
static gene_t* genes; static uint32_t n_genes;
static exon_t** cexons; static uint32_t* n_cexons; static uint64_t* cexon_bases; /* per contig */

/* options */
static uint64_t opt_reads = 100000, opt_seed = 1;
static int opt_readlen = 100, opt_threads = 4, opt_level = 6, opt_paired = 0;
static double opt_in_gene = 0.7, opt_spliced = 0.0, opt_mm = 0.0, opt_edge = 0.01;

/* ------------------------------------------------------------------ rng */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t splitmix(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t rnd(rng_t* r) { return splitmix(&r->s); }
static inline double rndf(rng_t* r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rndn(rng_t* r, uint64_t n) { return n ? (uint64_t)(rndf(r) * (double)n) : 0; }
static rng_t rng_for(uint64_t a, uint64_t b, uint64_t c) {
    rng_t r = {opt_seed * 0xD1342543DE82EF95ull + a * 0x9E3779B97F4A7C15ull + b * 0xC2B2AE3D27D4EB4Full + c};
    rnd(&r); rnd(&r);
    return r;
}

/* ------------------------------------------------------------------ growable byte buffer */
typedef struct { uint8_t* p; size_t n, cap; } buf_t;
static void buf_need(buf_t* b, size_t extra) {
    if (b->n + extra > b->cap) {
        b->cap = (b->n + extra) * 2 + 4096;
        b->p = realloc(b->p, b->cap);
        if (!b->p) { fprintf(stderr, "out of memory\n"); exit(1); }
    }
}
static void buf_put(buf_t* b, const void* src, size_t n) { buf_need(b, n); memcpy(b->p + b->n, src, n); b->n += n; }

/* ------------------------------------------------------------------ BAI pieces */
static inline int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (int)(beg >> 14);
    if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (int)(beg >> 17);
    if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (int)(beg >> 20);
    if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (int)(beg >> 23);
    if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (int)(beg >> 26);
    return 0;
}
typedef struct { uint32_t bin; uint64_t v0, v1; } binchunk_t; /* voffsets relative to the task's first byte */
typedef struct { uint32_t win; uint64_t v; } lin_t;
static int cmp_u32(const void* x, const void* y) { uint32_t a1 = *(const uint32_t*)x, b1 = *(const uint32_t*)y; return a1 < b1 ? -1 : a1 > b1; }
static int cmp_binchunk(const void* a, const void* b) {
    const binchunk_t *x = a, *y = b;
    if (x->bin != y->bin) return x->bin < y->bin ? -1 : 1;
    return x->v0 < y->v0 ? -1 : x->v0 > y->v0;
}

/* ------------------------------------------------------------------ a read before encoding */
typedef struct {
    uint32_t pos;
    uint64_t qid;      /* qname number */
    uint16_t flag;
    uint8_t nh, mapq;
    uint8_t kind;      /* 0 plain, 1 spliced (junction from exon), 2 edge cigar, 3 unmapped-with-pos, 4 mate */
    uint32_t aux0, aux1, aux2; /* kind 1: a, gap, (second gap<<?)..; mate: next_pos */
    int32_t tlen;
    uint32_t next_pos;
} read_t;
static int cmp_read(const void* a, const void* b) {
    const read_t *x = a, *y = b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    if (x->qid != y->qid) return x->qid < y->qid ? -1 : 1;
    return (int)x->flag - (int)y->flag;
}

/* multi-mapper copies, bucketed per (contig, bin) by the pre-pass */
typedef struct { uint32_t pos; uint32_t group; uint8_t nh; uint8_t rev; } mmcopy_t;
typedef struct { mmcopy_t* v; uint32_t n, cap; } mmlist_t;

/* ------------------------------------------------------------------ tasks */
typedef struct {
    int contig; uint32_t bin; /* bin index within contig */
    uint64_t n_bg, n_gene;    /* unique reads (single-end) or pairs (paired) starting in this bin */
    uint64_t qid_base;
    mmlist_t mm;
    /* outputs */
    buf_t out;                /* compressed BGZF blocks */
    binchunk_t* chunks; uint32_t n_chunks, cap_chunks;
    lin_t* lin; uint32_t n_lin, cap_lin;
    uint64_t n_mapped, n_unmapped, n_records;
    uint64_t file_off;        /* assigned when written */
} task_t;

static task_t* tasks; static size_t n_tasks;

static const uint8_t QUALS[4] = {37, 25, 11, 2};

typedef struct {
    buf_t payload;    /* current uncompressed block */
    task_t* t;
    z_stream zs;
    uint8_t* cbuf;
} bgzf_w;

static void flush_block(bgzf_w* w) {
    if (!w->payload.n) return;
    deflateReset(&w->zs);
    w->zs.next_in = w->payload.p; w->zs.avail_in = (uInt)w->payload.n;
    w->zs.next_out = w->cbuf + 18; w->zs.avail_out = 65536 - 18 - 8;
    int rc = deflate(&w->zs, Z_FINISH);
    if (rc != Z_STREAM_END) { fprintf(stderr, "deflate failed (incompressible block)\n"); exit(1); }
    uint32_t clen = (uint32_t)w->zs.total_out, bsize = clen + 25;
    uint8_t* h = w->cbuf;
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0};
    memcpy(h, hdr, 16);
    h[16] = bsize & 0xff; h[17] = bsize >> 8;
    uint32_t crc = (uint32_t)crc32(crc32(0, NULL, 0), w->payload.p, (uInt)w->payload.n), isz = (uint32_t)w->payload.n;
    memcpy(h + 18 + clen, &crc, 4);
    memcpy(h + 22 + clen, &isz, 4);
    buf_put(&w->t->out, h, clen + 26);
    w->payload.n = 0;
}

static void add_index(task_t* t, uint32_t pos, uint32_t end, int unmapped, uint64_t v0, uint64_t v1) {
    uint32_t bin = (uint32_t)reg2bin(pos, end);
    if (t->n_chunks && t->chunks[t->n_chunks - 1].bin == bin && t->chunks[t->n_chunks - 1].v1 == v0) {
        t->chunks[t->n_chunks - 1].v1 = v1;
    } else {
        if (t->n_chunks == t->cap_chunks) t->chunks = realloc(t->chunks, (t->cap_chunks = t->cap_chunks * 2 + 64) * sizeof(binchunk_t));
        t->chunks[t->n_chunks++] = (binchunk_t){bin, v0, v1};
    }
    for (uint32_t w = pos >> 14; w <= (end - 1) >> 14; w++) {
        int found = 0;
        for (uint32_t i = t->n_lin; i-- > 0 && t->lin[i].win + 8 >= w;)  /* windows arrive nearly sorted */
            if (t->lin[i].win == w) { found = 1; break; }
        if (!found) {
            if (t->n_lin == t->cap_lin) t->lin = realloc(t->lin, (t->cap_lin = t->cap_lin * 2 + 64) * sizeof(lin_t));
            t->lin[t->n_lin++] = (lin_t){w, v0};
        }
    }
    if (unmapped) t->n_unmapped++; else t->n_mapped++;
}

/* cigar ops: MIDNSHP=X -> 0..8 */
#define CG(op, len) (((uint32_t)(len) << 4) | (op))

static void emit_record(bgzf_w* w, int tid, const read_t* r, const uint32_t* cigar, int n_cigar, uint32_t reflen, rng_t* rg) {
    int L = opt_readlen;
    uint8_t rec[1024];
    char qn[32];
    int lq = snprintf(qn, sizeof qn, "q%011llu", (unsigned long long)r->qid) + 1;
    uint32_t end = r->pos + (reflen && !(r->flag & 4) ? reflen : 1);
    uint32_t body = 32 + (uint32_t)lq + 4u * (uint32_t)n_cigar + (uint32_t)(L + 1) / 2 + (uint32_t)L + 4;
    uint8_t* p = rec;
    int32_t i32; uint16_t u16;
    i32 = (int32_t)body; memcpy(p, &i32, 4); p += 4;
    i32 = tid; memcpy(p, &i32, 4); p += 4;
    i32 = (int32_t)r->pos; memcpy(p, &i32, 4); p += 4;
    *p++ = (uint8_t)lq; *p++ = r->mapq;
    u16 = (uint16_t)reg2bin(r->pos, end); memcpy(p, &u16, 2); p += 2;
    u16 = (uint16_t)n_cigar; memcpy(p, &u16, 2); p += 2;
    u16 = r->flag; memcpy(p, &u16, 2); p += 2;
    i32 = L; memcpy(p, &i32, 4); p += 4;
    i32 = (r->flag & 1) ? tid : -1; memcpy(p, &i32, 4); p += 4;
    i32 = (r->flag & 1) ? (int32_t)r->next_pos : -1; memcpy(p, &i32, 4); p += 4;
    i32 = r->tlen; memcpy(p, &i32, 4); p += 4;
    memcpy(p, qn, (size_t)lq); p += lq;
    memcpy(p, cigar, 4u * (size_t)n_cigar); p += 4 * n_cigar;
    /* SEQ: uniform ACGT nibbles {1,2,4,8}; QUAL: iid {37:.80, 25:.10, 11:.07, 2:.03} */
    for (int i = 0; i < (L + 1) / 2; i += 16) {
        uint64_t x = rnd(rg);
        for (int k = 0; k < 16 && i + k < (L + 1) / 2; k++, x >>= 4) p[i + k] = (uint8_t)((1u << (x & 3)) << 4 | (1u << ((x >> 2) & 3)));
    }
    if (L & 1) p[(L + 1) / 2 - 1] &= 0xf0;
    p += (L + 1) / 2;
    for (int i = 0; i < L; i += 8) {
        uint64_t x = rnd(rg);
        for (int k = 0; k < 8 && i + k < L; k++, x >>= 8) {
            uint32_t u = (uint32_t)(x & 0xff);
            p[i + k] = QUALS[u < 205 ? 0 : u < 230 ? 1 : u < 248 ? 2 : 3];
        }
    }
    p += L;
    *p++ = 'N'; *p++ = 'H'; *p++ = 'C'; *p++ = r->nh;
    size_t n = (size_t)(p - rec);
    if (w->payload.n && w->payload.n + n > MAX_PAYLOAD) flush_block(w);
    uint64_t v0 = ((uint64_t)w->t->out.n << 16) | w->payload.n;
    buf_put(&w->payload, rec, n);
    if (w->payload.n >= MAX_PAYLOAD) flush_block(w);
    uint64_t v1 = ((uint64_t)w->t->out.n << 16) | w->payload.n;
    add_index(w->t, r->pos, end, (r->flag & 4) != 0, v0, v1);
    w->t->n_records++;
}

/* positions of the unique reads / first mates of one bin, sorted; shared by the bin itself and (paired
 * mode) by the next bin, which needs the mates that spill over */
static uint32_t* bin_positions(int c, uint32_t bin, uint64_t n_bg, uint64_t n_gene, uint64_t* n_out) {
    rng_t rg = rng_for(1, (uint64_t)c, bin);
    uint32_t lo = bin * BIN_SIZE, hi = lo + BIN_SIZE;
    uint32_t clen = contigs[c].len;
    if (hi > clen) hi = clen;
    uint32_t maxpos = clen > (uint32_t)opt_readlen + 1 ? clen - (uint32_t)opt_readlen - 1 : 0;
    uint64_t n = n_bg + n_gene;
    uint32_t* pos = malloc((n + 1) * 4);
    uint64_t k = 0;
    for (uint64_t i = 0; i < n_bg; i++) {
        uint32_t p = lo + (uint32_t)rndn(&rg, hi - lo);
        pos[k++] = p > maxpos ? maxpos : p;
    }
    if (n_gene) {
        /* exon pieces inside the bin, weighted by length */
        exon_t* ex = cexons[c]; uint32_t ne = n_cexons[c];
        uint32_t first = 0, cnt = 0;
        uint64_t total = 0;
        /* exons sorted by start: find those overlapping [lo,hi) (max exon length bounded, scan window) */
        uint32_t a = 0, b = ne;
        while (a < b) { uint32_t m = (a + b) / 2; if (ex[m].s + 4000000u < lo) a = m + 1; else b = m; }
        first = a;
        uint32_t i;
        uint64_t* cum = NULL; uint32_t* ids = NULL; uint32_t cap = 0;
        for (i = first; i < ne && ex[i].s < hi; i++) {
            uint32_t s = ex[i].s > lo ? ex[i].s : lo, e = ex[i].e < hi ? ex[i].e : hi;
            if (e <= s) continue;
            if (cnt == cap) { cap = cap * 2 + 64; cum = realloc(cum, cap * 8); ids = realloc(ids, cap * 4); }
            total += e - s; cum[cnt] = total; ids[cnt] = i; cnt++;
        }
        for (uint64_t q = 0; q < n_gene; q++) {
            uint32_t p;
            if (!cnt) p = lo + (uint32_t)rndn(&rg, hi - lo);
            else {
                uint64_t x = rndn(&rg, total);
                uint32_t l2 = 0, h2 = cnt;
                while (l2 < h2) { uint32_t m = (l2 + h2) / 2; if (cum[m] <= x) l2 = m + 1; else h2 = m; }
                exon_t* e1 = &ex[ids[l2]];
                uint32_t s = e1->s > lo ? e1->s : lo;
                uint64_t prev = l2 ? cum[l2 - 1] : 0;
                p = s + (uint32_t)(x - prev);
                /* start the read a little upstream so that it covers the exon base */
                uint32_t back = (uint32_t)rndn(&rg, (uint64_t)opt_readlen);
                p = p > lo + back ? p - back : lo;
            }
            pos[k++] = p > maxpos ? maxpos : p;
        }
        free(cum); free(ids);
    }
    /* sort positions (counting-free: simple qsort, bins hold ~1e4-1e5 reads) */
    qsort(pos, n, 4, cmp_u32);
    *n_out = n;
    return pos;
}

/* the exon (of contig c) that contains position p and has a successor exon in its gene; returns NULL if none */
static const exon_t* junction_exon(int c, uint32_t p) {
    exon_t* ex = cexons[c]; uint32_t ne = n_cexons[c];
    uint32_t a = 0, b = ne;
    while (a < b) { uint32_t m = (a + b) / 2; if (ex[m].s <= p) a = m + 1; else b = m; }
    for (uint32_t i = a; i-- > 0 && ex[i].s + 4000000u >= p;) {
        if (ex[i].e > p && ex[i].idx + 1 < genes[ex[i].gene].n_exons) return &ex[i];
    }
    return NULL;
}

/* Spliced reads sit on junctions, and junctions are shared by many reads (RNA-seq: ~10^5 distinct junctions for 10^7-10^8
 * spliced reads): move a read that is to be spliced onto the nearest exon end (of an exon with a successor) that keeps it
 * inside its bin [lo, hi) and within 50 kb of where it was; `off` in [1, L-1] is how much of the read lies before the
 * junction.  Returns 0 when there is no such exon: the read then stays unspliced. */
static int snap_to_junction(int c, uint32_t pos, uint32_t lo, uint32_t hi, uint32_t off, uint32_t* out) {
    exon_t* ex = cexons[c]; uint32_t ne = n_cexons[c];
    uint32_t a = 0, b = ne;
    while (a < b) { uint32_t m = (a + b) / 2; if (ex[m].s <= pos) a = m + 1; else b = m; }
    uint32_t from = a > 12 ? a - 12 : 0, to = a + 24 < ne ? a + 24 : ne;
    uint32_t best = 0xffffffffu, best_d = 50000u;
    for (uint32_t i = from; i < to; i++) {
        if (ex[i].idx + 1 >= genes[ex[i].gene].n_exons || ex[i].e <= off) continue;
        uint32_t np = ex[i].e - off;
        if (np < lo || np >= hi || np + off <= ex[i].s) continue;  /* the part before the junction stays inside the exon */
        uint32_t d = np > pos ? np - pos : pos - np;
        if (d < best_d) { best_d = d; best = np; }
    }
    if (best == 0xffffffffu) return 0;
    *out = best;
    return 1;
}

static int build_cigar(int c, read_t* r, uint32_t* cg, uint32_t* reflen, rng_t* rg) {
    int L = opt_readlen;
    if (r->kind == 3) { *reflen = 0; return 0; }
    if (r->kind == 2) {
        int n = 0;
        switch (rnd(rg) % 6) {
            case 0: cg[n++] = CG(4, 5); cg[n++] = CG(0, L - 5); *reflen = (uint32_t)(L - 5); break;
            case 1: cg[n++] = CG(0, L / 2); cg[n++] = CG(1, 3); cg[n++] = CG(0, L - L / 2 - 3); *reflen = (uint32_t)(L - 3); break;
            case 2: cg[n++] = CG(0, L / 2); cg[n++] = CG(2, 4); cg[n++] = CG(0, L - L / 2); *reflen = (uint32_t)(L + 4); break;
            case 3: cg[n++] = CG(7, L / 2); cg[n++] = CG(8, L - L / 2); *reflen = (uint32_t)L; break;
            case 4: cg[n++] = CG(0, L - 7); cg[n++] = CG(4, 7); *reflen = (uint32_t)(L - 7); break;
            default: cg[n++] = CG(5, 10); cg[n++] = CG(0, L); *reflen = (uint32_t)L; break;
        }
        return n;
    }
    if (r->kind == 1) {
        /* spliced: snap to an exon junction when the read starts inside an exon that has a successor */
        const exon_t* e = junction_exon(c, r->pos);
        uint32_t a, gap;
        if (e && e->e - r->pos < (uint32_t)L && e->e > r->pos) {
            const gene_t* g = &genes[e->gene];
            a = e->e - r->pos;
            gap = g->s[e->idx + 1] - e->e;
            uint32_t rest = (uint32_t)L - a;
            uint32_t nlen = g->e[e->idx + 1] - g->s[e->idx + 1];
            if (gap > 0 && rest > nlen && e->idx + 2 < g->n_exons && g->s[e->idx + 2] > g->e[e->idx + 1]) {
                uint32_t gap2 = g->s[e->idx + 2] - g->e[e->idx + 1];
                cg[0] = CG(0, a); cg[1] = CG(3, gap); cg[2] = CG(0, nlen); cg[3] = CG(3, gap2); cg[4] = CG(0, rest - nlen);
                *reflen = (uint32_t)L + gap + gap2;
                return 5;
            }
            if (gap == 0) gap = 70;
        } else {
            a = 1 + (uint32_t)rndn(rg, (uint64_t)(L - 1));
            gap = (uint32_t)(70.0 * pow(50000.0 / 70.0, rndf(rg)));  /* log-uniform 70 bp .. 50 kb */
        }
        uint32_t clen = contigs[c].len;
        if ((uint64_t)r->pos + L + gap >= clen) gap = 70;
        cg[0] = CG(0, a); cg[1] = CG(3, gap); cg[2] = CG(0, (uint32_t)L - a);
        *reflen = (uint32_t)L + gap;
        return 3;
    }
    cg[0] = CG(0, L);
    *reflen = (uint32_t)L;
    return 1;
}

static void run_task(task_t* t) {
    int c = t->contig;
    uint64_t n = 0;
    uint32_t* pos = bin_positions(c, t->bin, t->n_bg, t->n_gene, &n);
    uint64_t n_prev = 0; uint32_t* prev_pos = NULL; task_t* pt = NULL;
    if (opt_paired && t->bin > 0) {
        pt = t - 1;  /* tasks of one contig are contiguous */
        prev_pos = bin_positions(c, pt->bin, pt->n_bg, pt->n_gene, &n_prev);
    }
    size_t cap = (size_t)(n * (opt_paired ? 2 : 1) + t->mm.n + (opt_paired ? 4096 : 0) + 16);
    read_t* rs = calloc(cap, sizeof(read_t));
    size_t k = 0;
    rng_t rg = rng_for(2, (uint64_t)c, t->bin);
    uint32_t lo = t->bin * BIN_SIZE, hi = lo + BIN_SIZE;
    uint32_t clen = contigs[c].len;
    uint32_t maxpos = clen > (uint32_t)opt_readlen + 1 ? clen - (uint32_t)opt_readlen - 1 : 0;
    for (uint64_t i = 0; i < n; i++) {
        read_t r; memset(&r, 0, sizeof r);
        r.pos = pos[i]; r.qid = t->qid_base + i; r.nh = 1; r.mapq = 255;
        uint64_t x = rnd(&rg);
        int rev = (int)(x & 1);
        double u = (double)((x >> 8) & 0xffffff) / 16777216.0;
        if (opt_paired) {
            uint32_t d = 200 + (uint32_t)((x >> 32) % 401);
            uint32_t mp = r.pos + d; if (mp > maxpos) mp = maxpos;
            r.flag = rev ? 83 : 99; r.next_pos = mp; r.tlen = (int32_t)(mp + (uint32_t)opt_readlen - r.pos);
            read_t m = r; m.pos = mp; m.flag = rev ? 163 : 147; m.next_pos = r.pos; m.tlen = -r.tlen; m.kind = 4;
            if (u < opt_spliced) r.kind = 1;
            rs[k++] = r;
            if (mp < hi) rs[k++] = m;  /* else it belongs to the next bin, which regenerates it */
        } else {
            r.flag = rev ? 16 : 0;
            if (u < opt_edge * 0.1) { r.kind = 3; r.flag |= 4; r.mapq = 0; }
            else if (u < opt_edge) r.kind = 2;
            else if (u < opt_edge + opt_spliced) {
                uint32_t np;
                if (snap_to_junction(c, r.pos, lo, hi < maxpos ? hi : maxpos, 1u + (uint32_t)((x >> 36) % (uint32_t)(opt_readlen - 1)), &np)) {
                    r.pos = np;
                    r.kind = 1;
                }
            }
            rs[k++] = r;
        }
    }
    if (prev_pos) {
        rng_t pg = rng_for(2, (uint64_t)c, pt->bin);
        for (uint64_t i = 0; i < n_prev; i++) {
            uint64_t x = rnd(&pg);
            int rev = (int)(x & 1);
            uint32_t d = 200 + (uint32_t)((x >> 32) % 401);
            uint32_t mp = prev_pos[i] + d; if (mp > maxpos) mp = maxpos;
            if (mp >= lo && mp < hi + BIN_SIZE) {
                if (mp >= hi) continue;
                read_t m; memset(&m, 0, sizeof m);
                m.pos = mp; m.qid = pt->qid_base + i; m.nh = 1; m.mapq = 255; m.kind = 4;
                m.flag = rev ? 163 : 147; m.next_pos = prev_pos[i]; m.tlen = -(int32_t)(mp + (uint32_t)opt_readlen - prev_pos[i]);
                if (k == cap) { cap *= 2; rs = realloc(rs, cap * sizeof(read_t)); }
                rs[k++] = m;
            }
        }
        free(prev_pos);
    }
    for (uint32_t i = 0; i < t->mm.n; i++) {
        read_t r; memset(&r, 0, sizeof r);
        r.pos = t->mm.v[i].pos; r.qid = 90000000000ull + t->mm.v[i].group; r.nh = t->mm.v[i].nh;
        r.mapq = r.nh == 2 ? 3 : r.nh <= 4 ? 1 : 0;
        r.flag = t->mm.v[i].rev ? 16 : 0;
        if (k == cap) { cap *= 2; rs = realloc(rs, cap * sizeof(read_t)); }
        rs[k++] = r;
    }
    qsort(rs, k, sizeof(read_t), cmp_read);
    bgzf_w w; memset(&w, 0, sizeof w);
    w.t = t; w.cbuf = malloc(65536 + 64);
    deflateInit2(&w.zs, opt_level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    rng_t sg = rng_for(3, (uint64_t)c, t->bin);
    for (size_t i = 0; i < k; i++) {
        uint32_t cg[8], reflen = 0;
        int nc = build_cigar(c, &rs[i], cg, &reflen, &sg);
        emit_record(&w, c, &rs[i], cg, nc, reflen, &sg);
    }
    flush_block(&w);
    deflateEnd(&w.zs);
    free(w.cbuf); free(w.payload.p); free(rs); free(pos);
    free(t->mm.v); t->mm.v = NULL;
}

/* ------------------------------------------------------------------ pool over a batch */
static size_t batch_lo, batch_hi, batch_next;
static void* worker(void* arg) {
    (void)arg;
    for (;;) {
        size_t i = __atomic_fetch_add(&batch_next, 1, __ATOMIC_RELAXED);
        if (i >= batch_hi) break;
        run_task(&tasks[i]);
    }
    return NULL;
}

/* ------------------------------------------------------------------ model + main */
static int cmp_exon(const void* a, const void* b) {
    const exon_t *x = a, *y = b;
    if (x->s != y->s) return x->s < y->s ? -1 : 1;
    return x->e < y->e ? -1 : x->e > y->e;
}
static void load_model(const char* path) {
    FILE* f = fopen(path, "r");
    if (!f) { perror(path); exit(1); }
    char tag[8];
    if (fscanf(f, "%7s %d", tag, &n_contigs) != 2 || tag[0] != 'C') { fprintf(stderr, "bad model\n"); exit(1); }
    contigs = calloc((size_t)n_contigs, sizeof(contig_t));
    for (int i = 0; i < n_contigs; i++) if (fscanf(f, "%63s %u", contigs[i].name, &contigs[i].len) != 2) exit(1);
    if (fscanf(f, "%7s %u", tag, &n_genes) != 2 || tag[0] != 'G') { fprintf(stderr, "bad model\n"); exit(1); }
    genes = calloc((size_t)n_genes + 1, sizeof(gene_t));
    n_cexons = calloc((size_t)n_contigs, 4);
    for (uint32_t g = 0; g < n_genes; g++) {
        gene_t* G = &genes[g];
        if (fscanf(f, "%u %d %u", &G->contig, &G->strand, &G->n_exons) != 3) exit(1);
        G->s = malloc(4 * (size_t)G->n_exons); G->e = malloc(4 * (size_t)G->n_exons);
        for (uint32_t i = 0; i < G->n_exons; i++) if (fscanf(f, "%u %u", &G->s[i], &G->e[i]) != 2) exit(1);
        n_cexons[G->contig] += G->n_exons;
    }
    fclose(f);
    cexons = calloc((size_t)n_contigs, sizeof(exon_t*));
    cexon_bases = calloc((size_t)n_contigs, 8);
    uint32_t* fill = calloc((size_t)n_contigs, 4);
    for (int c = 0; c < n_contigs; c++) cexons[c] = malloc(((size_t)n_cexons[c] + 1) * sizeof(exon_t));
    for (uint32_t g = 0; g < n_genes; g++)
        for (uint32_t i = 0; i < genes[g].n_exons; i++) {
            uint32_t c = genes[g].contig;
            cexons[c][fill[c]++] = (exon_t){genes[g].s[i], genes[g].e[i], g, i};
            cexon_bases[c] += genes[g].e[i] - genes[g].s[i];
        }
    for (int c = 0; c < n_contigs; c++) qsort(cexons[c], n_cexons[c], sizeof(exon_t), cmp_exon);
    free(fill);
}

static void mm_push(mmlist_t* l, mmcopy_t x) {
    if (l->n == l->cap) l->v = realloc(l->v, (l->cap = l->cap * 2 + 16) * sizeof(mmcopy_t));
    l->v[l->n++] = x;
}

int main(int argc, char** argv) {
    const char *model = NULL, *out = NULL;
    for (int i = 1; i < argc; i++) {
        const char* a = argv[i];
        const char* v = i + 1 < argc ? argv[i + 1] : "";
        if (!strcmp(a, "--genes")) { model = v; i++; }
        else if (!strcmp(a, "--out")) { out = v; i++; }
        else if (!strcmp(a, "--reads")) { opt_reads = strtoull(v, NULL, 10); i++; }
        else if (!strcmp(a, "--readlen")) { opt_readlen = atoi(v); i++; }
        else if (!strcmp(a, "--seed")) { opt_seed = strtoull(v, NULL, 10); i++; }
        else if (!strcmp(a, "--threads")) { opt_threads = atoi(v); i++; }
        else if (!strcmp(a, "--level")) { opt_level = atoi(v); i++; }
        else if (!strcmp(a, "--in-gene")) { opt_in_gene = atof(v); i++; }
        else if (!strcmp(a, "--spliced")) { opt_spliced = atof(v); i++; }
        else if (!strcmp(a, "--mm")) { opt_mm = atof(v); i++; }
        else if (!strcmp(a, "--edge")) { opt_edge = atof(v); i++; }
        else if (!strcmp(a, "--paired")) opt_paired = 1;
        else { fprintf(stderr, "unknown option %s\n", a); return 2; }
    }
    if (!model || !out || opt_readlen < 20 || opt_readlen > 400) { fprintf(stderr, "usage: bamgen --genes model.txt --out x.bam --reads N ...\n"); return 2; }
    if (opt_threads < 1) opt_threads = 1;
    if (opt_paired) { opt_mm = 0; opt_edge = 0; }
    load_model(model);

    /* tasks = (contig, 1 Mb bin) */
    double total_len = 0;
    for (int c = 0; c < n_contigs; c++) total_len += contigs[c].len;
    size_t* first_task = calloc((size_t)n_contigs + 1, sizeof(size_t));
    for (int c = 0; c < n_contigs; c++) first_task[c + 1] = first_task[c] + (contigs[c].len + BIN_SIZE - 1) / BIN_SIZE + (contigs[c].len == 0);
    n_tasks = first_task[n_contigs];
    tasks = calloc(n_tasks + 1, sizeof(task_t));
    uint64_t units = opt_paired ? opt_reads / 2 : opt_reads;                 /* reads or pairs */
    uint64_t mm_copies_target = (uint64_t)((double)units * opt_mm);
    uint64_t unique = units - mm_copies_target;
    uint64_t qid = 0, assigned = 0;
    for (int c = 0; c < n_contigs; c++) {
        uint64_t n_c = c == n_contigs - 1 ? unique - assigned : (uint64_t)((double)unique * contigs[c].len / total_len);
        assigned += n_c;
        uint64_t n_gene_c = cexon_bases[c] ? (uint64_t)((double)n_c * opt_in_gene) : 0, n_bg_c = n_c - n_gene_c;
        size_t nb = first_task[c + 1] - first_task[c];
        /* cumulative shares per bin: background by length, gene reads by exon bases */
        double* ebin = calloc(nb + 1, sizeof(double));
        for (uint32_t i = 0; i < n_cexons[c]; i++) {
            exon_t* e = &cexons[c][i];
            for (uint32_t b = e->s / BIN_SIZE; b <= (e->e ? (e->e - 1) / BIN_SIZE : 0) && b < nb; b++) {
                uint32_t lo = b * BIN_SIZE, hi = lo + BIN_SIZE;
                uint32_t s = e->s > lo ? e->s : lo, en = e->e < hi ? e->e : hi;
                if (en > s) ebin[b] += en - s;
            }
        }
        double etot = 0; for (size_t b = 0; b < nb; b++) etot += ebin[b];
        double cl = 0, ce = 0;
        uint64_t done_bg = 0, done_g = 0;
        for (size_t b = 0; b < nb; b++) {
            task_t* t = &tasks[first_task[c] + b];
            t->contig = c; t->bin = (uint32_t)b;
            uint32_t lo = (uint32_t)b * BIN_SIZE, hi = lo + BIN_SIZE; if (hi > contigs[c].len) hi = contigs[c].len;
            cl += contigs[c].len ? (double)(hi - lo) / contigs[c].len : 1.0;
            ce += etot > 0 ? ebin[b] / etot : 0;
            uint64_t tb = b + 1 == nb ? n_bg_c : (uint64_t)((double)n_bg_c * cl);
            uint64_t tg = b + 1 == nb ? n_gene_c : (uint64_t)((double)n_gene_c * ce);
            if (tb < done_bg) tb = done_bg;
            if (tg < done_g) tg = done_g;
            t->n_bg = tb - done_bg; t->n_gene = tg - done_g;
            done_bg = tb; done_g = tg;
            t->qid_base = qid; qid += t->n_bg + t->n_gene;
        }
        free(ebin);
    }
    /* multi-mapper pre-pass: groups of NH copies sharing a qname */
    if (mm_copies_target) {
        rng_t rg = rng_for(4, 0, 0);
        uint64_t made = 0; uint32_t group = 0;
        while (made < mm_copies_target) {
            double u = rndf(&rg);
            int nh = u < 0.5 ? 2 : u < 0.75 ? 3 : u < 0.85 ? 4 : u < 0.92 ? 5 : 6 + (int)rndn(&rg, 5);
            int c = 0; uint32_t p = 0;
            for (int k = 0; k < nh; k++) {
                if (k > 0 && rndf(&rg) < 0.35) {
                    int64_t q = (int64_t)p + (int64_t)rndn(&rg, 1000) - 500;
                    p = q < 0 ? 0 : (uint32_t)q;
                } else {
                    double x = rndf(&rg) * total_len;
                    for (c = 0; c < n_contigs - 1 && x >= contigs[c].len; c++) x -= contigs[c].len;
                    if (n_cexons[c] && rndf(&rg) < opt_in_gene) {
                        exon_t* e = &cexons[c][rndn(&rg, n_cexons[c])];
                        p = e->s + (uint32_t)rndn(&rg, e->e - e->s + 1);
                        uint32_t back = (uint32_t)rndn(&rg, (uint64_t)opt_readlen);
                        p = p > back ? p - back : 0;
                    } else p = (uint32_t)rndn(&rg, contigs[c].len);
                }
                uint32_t maxpos = contigs[c].len > (uint32_t)opt_readlen + 1 ? contigs[c].len - (uint32_t)opt_readlen - 1 : 0;
                if (p > maxpos) p = maxpos;
                mm_push(&tasks[first_task[c] + p / BIN_SIZE].mm, (mmcopy_t){p, group, (uint8_t)nh, (uint8_t)(rnd(&rg) & 1)});
                made++;
            }
            group++;
        }
    }

    FILE* fo = fopen(out, "wb");
    if (!fo) { perror(out); return 1; }
    /* header */
    uint64_t file_off = 0;
    {
        buf_t h = {0, 0, 0};
        char line[256];
        buf_t text = {0, 0, 0};
        buf_put(&text, "@HD\tVN:1.6\tSO:coordinate\n", 26);
        for (int c = 0; c < n_contigs; c++) { int n = snprintf(line, sizeof line, "@SQ\tSN:%s\tLN:%u\n", contigs[c].name, contigs[c].len); buf_put(&text, line, (size_t)n); }
        int32_t i32;
        buf_put(&h, "BAM\1", 4);
        i32 = (int32_t)text.n; buf_put(&h, &i32, 4); buf_put(&h, text.p, text.n);
        i32 = n_contigs; buf_put(&h, &i32, 4);
        for (int c = 0; c < n_contigs; c++) {
            i32 = (int32_t)strlen(contigs[c].name) + 1; buf_put(&h, &i32, 4); buf_put(&h, contigs[c].name, (size_t)i32);
            i32 = (int32_t)contigs[c].len; buf_put(&h, &i32, 4);
        }
        task_t ht; memset(&ht, 0, sizeof ht);
        bgzf_w w; memset(&w, 0, sizeof w);
        w.t = &ht; w.cbuf = malloc(65536 + 64);
        deflateInit2(&w.zs, opt_level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        for (size_t o = 0; o < h.n;) {
            size_t take = h.n - o < MAX_PAYLOAD ? h.n - o : MAX_PAYLOAD;
            buf_put(&w.payload, h.p + o, take); flush_block(&w); o += take;
        }
        deflateEnd(&w.zs);
        fwrite(ht.out.p, 1, ht.out.n, fo); file_off = ht.out.n;
        free(ht.out.p); free(w.cbuf); free(w.payload.p); free(h.p); free(text.p);
    }
    /* batches */
    size_t batch = (size_t)opt_threads * 4;
    uint64_t total_records = 0;
    pthread_t* th = malloc((size_t)opt_threads * sizeof(pthread_t));
    for (size_t lo = 0; lo < n_tasks; lo += batch) {
        batch_lo = lo; batch_hi = lo + batch < n_tasks ? lo + batch : n_tasks; batch_next = lo;
        for (int t = 0; t < opt_threads; t++) pthread_create(&th[t], NULL, worker, NULL);
        for (int t = 0; t < opt_threads; t++) pthread_join(th[t], NULL);
        for (size_t i = batch_lo; i < batch_hi; i++) {
            tasks[i].file_off = file_off;
            if (tasks[i].out.n && fwrite(tasks[i].out.p, 1, tasks[i].out.n, fo) != tasks[i].out.n) { perror("write"); return 1; }
            file_off += tasks[i].out.n;
            total_records += tasks[i].n_records;
            free(tasks[i].out.p); tasks[i].out.p = NULL;
        }
    }
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, fo);
    fclose(fo);

    /* BAI: merge the per-task pieces (tasks are in file order, contig by contig) */
    char* bai_path; if (asprintf(&bai_path, "%s.bai", out) < 0) return 1;
    FILE* fb = fopen(bai_path, "wb");
    if (!fb) { perror(bai_path); return 1; }
    fwrite("BAI\1", 1, 4, fb);
    int32_t i32 = n_contigs; fwrite(&i32, 4, 1, fb);
    for (int c = 0; c < n_contigs; c++) {
        /* collect chunks per bin */
        size_t total = 0;
        for (size_t i = first_task[c]; i < first_task[c + 1]; i++) total += tasks[i].n_chunks;
        binchunk_t* all = malloc((total + 1) * sizeof(binchunk_t));
        size_t k = 0;
        uint64_t n_mapped = 0, n_unmapped = 0, off_beg = 0, off_end = 0; int have = 0;
        for (size_t i = first_task[c]; i < first_task[c + 1]; i++) {
            task_t* t = &tasks[i];
            for (uint32_t j = 0; j < t->n_chunks; j++) {
                binchunk_t x = t->chunks[j];
                x.v0 += t->file_off << 16; x.v1 += t->file_off << 16;
                /* a record ending exactly at a block end gets v1 = (next block, 0): fine */
                all[k++] = x;
                if (!have) { off_beg = x.v0; have = 1; }
                off_end = x.v1;
            }
            n_mapped += t->n_mapped; n_unmapped += t->n_unmapped;
        }
        /* stable sort by bin keeping file order */
        qsort(all, k, sizeof(binchunk_t), cmp_binchunk);
        int32_t n_bin = 0;
        for (size_t i = 0; i < k; i++) if (i == 0 || all[i].bin != all[i - 1].bin) n_bin++;
        if (have) n_bin++;
        fwrite(&n_bin, 4, 1, fb);
        for (size_t i = 0; i < k;) {
            size_t j = i;
            while (j < k && all[j].bin == all[i].bin) j++;
            uint32_t bin = all[i].bin; int32_t nchunk = (int32_t)(j - i);
            fwrite(&bin, 4, 1, fb); fwrite(&nchunk, 4, 1, fb);
            for (size_t q = i; q < j; q++) { fwrite(&all[q].v0, 8, 1, fb); fwrite(&all[q].v1, 8, 1, fb); }
            i = j;
        }
        if (have) {
            uint32_t bin = 37450; int32_t two = 2;
            fwrite(&bin, 4, 1, fb); fwrite(&two, 4, 1, fb);
            fwrite(&off_beg, 8, 1, fb); fwrite(&off_end, 8, 1, fb); fwrite(&n_mapped, 8, 1, fb); fwrite(&n_unmapped, 8, 1, fb);
        }
        free(all);
        /* linear index */
        uint32_t maxw = 0; int anyw = 0;
        for (size_t i = first_task[c]; i < first_task[c + 1]; i++)
            for (uint32_t j = 0; j < tasks[i].n_lin; j++) { if (tasks[i].lin[j].win >= maxw) maxw = tasks[i].lin[j].win; anyw = 1; }
        int32_t n_intv = anyw ? (int32_t)maxw + 1 : 0;
        uint64_t* lin = calloc((size_t)n_intv + 1, 8);
        for (size_t i = first_task[c]; i < first_task[c + 1]; i++)
            for (uint32_t j = 0; j < tasks[i].n_lin; j++) {
                uint64_t v = tasks[i].lin[j].v + (tasks[i].file_off << 16);
                uint32_t w = tasks[i].lin[j].win;
                if (!lin[w] || v < lin[w]) lin[w] = v;
            }
        uint64_t nxt = 0;
        for (int32_t w = n_intv - 1; w >= 0; w--) { if (lin[w]) nxt = lin[w]; else lin[w] = nxt; }
        fwrite(&n_intv, 4, 1, fb);
        fwrite(lin, 8, (size_t)n_intv, fb);
        free(lin);
    }
    uint64_t n_no_coor = 0; fwrite(&n_no_coor, 8, 1, fb);
    fclose(fb);
    printf("{\"records\": %llu, \"bytes\": %llu, \"tasks\": %zu}\n", (unsigned long long)total_records,
           (unsigned long long)(file_off + 28), n_tasks);
    return 0;
}
