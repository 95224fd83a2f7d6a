/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C + zlib + pthreads) of the
 * read-counting path of MarcoMernberger/mbf_bam.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.  The product
 * (mbf_bam_b200/csrc) never links or calls it.
 *
 * Pinning status (same as oracle/py_oracle.py, which this file is checked against):
 *   cigar blocks/introns ... pinned by the reference's 58 golden records
 *                            (src/bam_ext.rs:50-604 -> tests/golden/cigar_golden.json)
 *   count_reads_* / count_introns end to end ... PARITY UNPINNED by the reference's own
 *                            tests (src/lib.rs:300-301); the Rust crate cannot be built here.
 *   count_reads_weighted ... not in the reference; defined in DESIGN.md -- PARITY UNPINNED.
 *
 * Structure mirrors the reference on purpose, so that it doubles as the timed CPU
 * baseline: identical chunk list, a thread pool over chunks, every task opens its own
 * file handle, seeks through the BAI linear index and inflates with zlib
 * (src/count_reads/counters.rs:217-259).
 *
 * Third-party behaviour restated from published APIs (sources not in the reference tree):
 * rust-htslib 0.24.0 (+ bundled htslib): fetch/read/aux/aligned_blocks/introns;
 * bio 0.25.0 IntervalTree::find: half-open overlap a.start < b.end && b.start < a.end.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "../include/mbfbam.h"

#define CHUNK_SIZE 1000000u /* src/count_reads/chunked_genome.rs:75 */

/* ------------------------------------------------------------------ BGZF reader */
typedef struct {
    FILE* f;
    uint8_t* cbuf;
    uint8_t* ubuf;
    uint32_t ulen, upos;
    uint64_t coff;      /* file offset of the block in ubuf */
    uint64_t next_coff; /* file offset of the next block */
    int eof;
    int err;
    z_stream zs;
    int zs_init;
} bgzf_t;

static int bgzf_open(bgzf_t* b, const char* path) {
    memset(b, 0, sizeof *b);
    b->f = fopen(path, "rb");
    if (!b->f) return -1;
    b->cbuf = malloc(65536 + 64);
    b->ubuf = malloc(65536 + 64);
    return 0;
}
static void bgzf_close(bgzf_t* b) {
    if (b->f) fclose(b->f);
    if (b->zs_init) inflateEnd(&b->zs);
    free(b->cbuf);
    free(b->ubuf);
}
static int bgzf_load(bgzf_t* b, uint64_t coff) {
    b->ulen = b->upos = 0;
    if (fseeko(b->f, (off_t)coff, SEEK_SET)) return b->err = -1;
    size_t n = fread(b->cbuf, 1, 18, b->f);
    if (n == 0) { b->eof = 1; return 1; }
    if (n < 18 || b->cbuf[0] != 0x1f || b->cbuf[1] != 0x8b || b->cbuf[2] != 8 || !(b->cbuf[3] & 4))
        return b->err = -2;
    uint32_t xlen = b->cbuf[10] | (b->cbuf[11] << 8);
    /* general extra-field walk (htslib writes XLEN=6 with BC first) */
    uint8_t* extra = b->cbuf + 12;
    uint32_t have = 6;
    if (xlen > 6) {
        if (fread(b->cbuf + 18, 1, xlen - 6, b->f) != xlen - 6) return b->err = -2;
        have = xlen;
    }
    int32_t bsize = -1;
    for (uint32_t p = 0; p + 4 <= have;) {
        uint32_t slen = extra[p + 2] | (extra[p + 3] << 8);
        if (extra[p] == 66 && extra[p + 1] == 67 && slen == 2) bsize = extra[p + 4] | (extra[p + 5] << 8);
        p += 4 + slen;
    }
    if (bsize < 0) return b->err = -2;
    uint32_t csize = (uint32_t)bsize + 1;
    uint32_t hdr = 12 + xlen;
    if (csize < hdr + 8) return b->err = -2;
    uint32_t rest = csize - hdr;
    if (fread(b->cbuf + hdr, 1, rest, b->f) != rest) return b->err = -2;
    uint32_t isize = b->cbuf[csize - 4] | (b->cbuf[csize - 3] << 8) | (b->cbuf[csize - 2] << 16) |
                     ((uint32_t)b->cbuf[csize - 1] << 24);
    if (isize > 65536) return b->err = -2;
    if (!b->zs_init) {
        memset(&b->zs, 0, sizeof b->zs);
        if (inflateInit2(&b->zs, -15) != Z_OK) return b->err = -3;
        b->zs_init = 1;
    } else {
        inflateReset(&b->zs);
    }
    b->zs.next_in = b->cbuf + hdr;
    b->zs.avail_in = rest - 8;
    b->zs.next_out = b->ubuf;
    b->zs.avail_out = 65536;
    int rc = inflate(&b->zs, Z_FINISH);
    if (rc != Z_STREAM_END || b->zs.total_out != isize) return b->err = -2;
    uint32_t crc = b->cbuf[csize - 8] | (b->cbuf[csize - 7] << 8) | (b->cbuf[csize - 6] << 16) |
                   ((uint32_t)b->cbuf[csize - 5] << 24);
    if ((uint32_t)crc32(crc32(0, NULL, 0), b->ubuf, isize) != crc) return b->err = -2;
    b->coff = coff;
    b->next_coff = coff + csize;
    b->ulen = isize;
    return 0;
}
static int bgzf_seek(bgzf_t* b, uint64_t voff) {
    b->eof = 0;
    int rc = bgzf_load(b, voff >> 16);
    if (rc) return rc;
    b->upos = (uint32_t)(voff & 0xffff);
    return b->upos <= b->ulen ? 0 : -2;
}
static size_t bgzf_read(bgzf_t* b, void* dst_, size_t n) {
    uint8_t* dst = dst_;
    size_t got = 0;
    while (got < n) {
        if (b->upos >= b->ulen) {
            if (b->eof || b->err) break;
            if (bgzf_load(b, b->next_coff)) break;
            continue;
        }
        size_t take = b->ulen - b->upos;
        if (take > n - got) take = n - got;
        memcpy(dst + got, b->ubuf + b->upos, take);
        b->upos += (uint32_t)take;
        got += take;
    }
    return got;
}

/* ------------------------------------------------------------------ header + BAI */
typedef struct {
    int32_t n_ref;
    char** names;
    uint32_t* lens;
    /* BAI */
    int32_t bai_n_ref;
    int32_t* n_intv;
    uint64_t** ioffset;
    uint64_t* ref_beg; /* min chunk begin over real bins, 0 if none */
    uint8_t* has_bins;
} bamidx_t;

static int read_header(const char* path, bamidx_t* h) {
    bgzf_t b;
    if (bgzf_open(&b, path)) return -1;
    if (bgzf_load(&b, 0)) { bgzf_close(&b); return -2; }
    char magic[4];
    int32_t l_text;
    if (bgzf_read(&b, magic, 4) != 4 || memcmp(magic, "BAM\1", 4) || bgzf_read(&b, &l_text, 4) != 4) {
        bgzf_close(&b);
        return -2;
    }
    char* text = malloc((size_t)l_text + 1);
    bgzf_read(&b, text, (size_t)l_text);
    free(text);
    bgzf_read(&b, &h->n_ref, 4);
    h->names = calloc((size_t)h->n_ref + 1, sizeof(char*));
    h->lens = calloc((size_t)h->n_ref + 1, 4);
    for (int i = 0; i < h->n_ref; i++) {
        int32_t ln;
        bgzf_read(&b, &ln, 4);
        h->names[i] = malloc((size_t)ln + 1);
        bgzf_read(&b, h->names[i], (size_t)ln);
        h->names[i][ln] = 0;
        int32_t l;
        bgzf_read(&b, &l, 4);
        h->lens[i] = (uint32_t)l;
    }
    int bad = b.err;
    bgzf_close(&b);
    return bad ? -2 : 0;
}
static int read_bai(const char* path, bamidx_t* h) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    char magic[4];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "BAI\1", 4)) { fclose(f); return -2; }
    if (fread(&h->bai_n_ref, 4, 1, f) != 1) { fclose(f); return -2; }
    int n = h->bai_n_ref;
    h->n_intv = calloc((size_t)n + 1, 4);
    h->ioffset = calloc((size_t)n + 1, sizeof(uint64_t*));
    h->ref_beg = calloc((size_t)n + 1, 8);
    h->has_bins = calloc((size_t)n + 1, 1);
    for (int r = 0; r < n; r++) {
        int32_t n_bin;
        if (fread(&n_bin, 4, 1, f) != 1) { fclose(f); return -2; }
        uint64_t mn = UINT64_MAX;
        for (int i = 0; i < n_bin; i++) {
            uint32_t bin;
            int32_t n_chunk;
            if (fread(&bin, 4, 1, f) != 1 || fread(&n_chunk, 4, 1, f) != 1) { fclose(f); return -2; }
            for (int c = 0; c < n_chunk; c++) {
                uint64_t be[2];
                if (fread(be, 8, 2, f) != 2) { fclose(f); return -2; }
                if (bin != 37450 && be[0] < mn) mn = be[0];
            }
        }
        h->has_bins[r] = mn != UINT64_MAX;
        h->ref_beg[r] = mn == UINT64_MAX ? 0 : mn;
        if (fread(&h->n_intv[r], 4, 1, f) != 1) { fclose(f); return -2; }
        h->ioffset[r] = malloc(((size_t)h->n_intv[r] + 1) * 8);
        if (h->n_intv[r] && fread(h->ioffset[r], 8, (size_t)h->n_intv[r], f) != (size_t)h->n_intv[r]) {
            fclose(f);
            return -2;
        }
    }
    fclose(f);
    return 0;
}
static void free_idx(bamidx_t* h) {
    for (int i = 0; i < h->n_ref; i++) free(h->names[i]);
    free(h->names);
    free(h->lens);
    for (int i = 0; i < h->bai_n_ref; i++) free(h->ioffset[i]);
    free(h->ioffset);
    free(h->n_intv);
    free(h->ref_beg);
    free(h->has_bins);
}
static int find_tid(const bamidx_t* h, const char* name) {
    for (int i = 0; i < h->n_ref; i++)
        if (!strcmp(h->names[i], name)) return i;
    return -1;
}
static char* default_bai(const char* bam, const char* bai) {
    char* p;
    if (bai) return strdup(bai);
    if (asprintf(&p, "%s.bai", bam) < 0) return NULL;
    FILE* f = fopen(p, "rb");
    if (f) { fclose(f); return p; }
    free(p);
    size_t n = strlen(bam);
    if (n > 4 && !strcmp(bam + n - 4, ".bam")) {
        p = strdup(bam);
        memcpy(p + n - 4, ".bai", 4);
        return p;
    }
    return NULL;
}

/* ------------------------------------------------------------------ interval trees */
/* bio 0.25.0 data_structures::interval_tree::IntervalTree<u32,(u32,i8)> (reference Cargo.lock:100-101,
 * src/count_reads/mod.rs:19; the crate's source is not under /root/reference: restated from its published
 * algorithm, SURVEY.md App. B.2 -- PARITY UNPINNED).  AVL tree keyed on interval.start; insert descends
 * LEFT when new.start <= node.start, else right, and repairs every node on the way back up; find() pops a
 * node from an explicit stack, pushes its left child when query.start < node.max and, when
 * query.end > node.start, its right child, and yields the node if it overlaps -- so hits come in the order
 * node, right subtree, left subtree.  The counters depend on that order only with each_read_counts_once
 * (counters.rs:83-92,173-182) and at chunk edges covered by several genes (chunked_genome.rs:98). */
typedef struct {
    uint32_t start, end, max, gene;
    int32_t l, r;
    int32_t h;
    int8_t strand;
} tnode_t;
typedef struct {
    uint32_t n;
    int32_t root;
    tnode_t* nd;
} ivset_t;

static inline int t_h(const ivset_t* t, int32_t x) { return x < 0 ? 0 : t->nd[x].h; }
static void t_update(ivset_t* t, int32_t x) {
    tnode_t* n = &t->nd[x];
    int lh = t_h(t, n->l), rh = t_h(t, n->r);
    n->h = 1 + (lh > rh ? lh : rh);
    uint32_t m = n->end;
    if (n->l >= 0 && t->nd[n->l].max > m) m = t->nd[n->l].max;
    if (n->r >= 0 && t->nd[n->r].max > m) m = t->nd[n->r].max;
    n->max = m;
}
static int32_t t_rot_right(ivset_t* t, int32_t x) {
    int32_t y = t->nd[x].l;
    t->nd[x].l = t->nd[y].r;
    t->nd[y].r = x;
    t_update(t, x);
    t_update(t, y);
    return y;
}
static int32_t t_rot_left(ivset_t* t, int32_t x) {
    int32_t y = t->nd[x].r;
    t->nd[x].r = t->nd[y].l;
    t->nd[y].l = x;
    t_update(t, x);
    t_update(t, y);
    return y;
}
static int32_t t_repair(ivset_t* t, int32_t x) {
    int lh = t_h(t, t->nd[x].l), rh = t_h(t, t->nd[x].r);
    if (lh - rh <= 1 && rh - lh <= 1) { t_update(t, x); return x; }
    if (rh > lh) {
        int32_t r = t->nd[x].r;
        if (t_h(t, t->nd[r].l) > t_h(t, t->nd[r].r)) t->nd[x].r = t_rot_right(t, r);
        return t_rot_left(t, x);
    }
    int32_t l = t->nd[x].l;
    if (t_h(t, t->nd[l].r) > t_h(t, t->nd[l].l)) t->nd[x].l = t_rot_left(t, l);
    return t_rot_right(t, x);
}
static int32_t t_insert(ivset_t* t, int32_t x, int32_t nw) {
    if (x < 0) return nw;
    if (t->nd[nw].start <= t->nd[x].start) t->nd[x].l = t_insert(t, t->nd[x].l, nw);
    else t->nd[x].r = t_insert(t, t->nd[x].r, nw);
    return t_repair(t, x);
}
/* src/count_reads/mod.rs:27-45 build_tree: intervals are inserted in list order */
static void build_ivset(const mbfbam_intervals* iv, int c, ivset_t* s) {
    uint64_t a = iv->iv_offsets[c], b = iv->iv_offsets[c + 1];
    s->n = (uint32_t)(b - a);
    s->root = -1;
    s->nd = malloc(((size_t)s->n + 1) * sizeof *s->nd);
    for (uint32_t i = 0; i < s->n; i++) {
        tnode_t* n = &s->nd[i];
        n->start = iv->iv_start[a + i];
        n->end = iv->iv_end[a + i];
        n->max = n->end;
        n->gene = iv->iv_gene[a + i];
        n->strand = iv->iv_strand[a + i];
        n->l = n->r = -1;
        n->h = 1;
        s->root = t_insert(s, s->root, (int32_t)i);
    }
}
static void free_ivset(ivset_t* s) { free(s->nd); }

/* find(q0..q1) as an explicit-stack iterator (bio's IntervalTreeIterator::next) */
typedef struct {
    int32_t stack[160]; /* <= 2 * height entries; an AVL tree of 2^32 nodes is < 47 high */
    int sp;
    uint32_t q0, q1;
} tfind_t;
static void t_find_begin(const ivset_t* t, tfind_t* f, uint32_t q0, uint32_t q1) {
    f->sp = 0;
    f->q0 = q0;
    f->q1 = q1;
    if (t->root >= 0) f->stack[f->sp++] = t->root;
}
static int32_t t_find_next(const ivset_t* t, tfind_t* f) {
    while (f->sp) {
        int32_t x = f->stack[--f->sp];
        const tnode_t* n = &t->nd[x];
        if (f->q0 < n->max) {
            if (n->l >= 0) f->stack[f->sp++] = n->l;
            if (f->q1 > n->start) {
                if (n->r >= 0) f->stack[f->sp++] = n->r;
                if (f->q0 < n->end && n->start < f->q1) return x;
            }
        }
    }
    return -1;
}

/* ------------------------------------------------------------------ chunk table */
typedef struct { int32_t contig; int32_t tid; uint32_t start, stop; } chunk_t;

/* src/count_reads/chunked_genome.rs:93-109: while tree.find(stop..stop+1) yields anything, the edge moves
 * past the FIRST interval it yields. */
static uint32_t extend_stop(const ivset_t* g, uint32_t stop) {
    for (;;) {
        tfind_t f;
        t_find_begin(g, &f, stop, stop + 1);
        int32_t x = t_find_next(g, &f);
        if (x < 0) return stop;
        stop = g->nd[x].end + 1;
    }
}
static chunk_t* make_chunks(const bamidx_t* h, const mbfbam_intervals* genes, const ivset_t* gsets,
                            size_t* n_out) {
    size_t cap = 1024, n = 0;
    chunk_t* out = malloc(cap * sizeof *out);
    int n_contigs = genes ? genes->n_contigs : h->n_ref;
    for (int c = 0; c < n_contigs; c++) {
        int tid = genes ? find_tid(h, genes->contig_names[c]) : c;
        if (tid < 0) continue; /* chunked_genome.rs:21-25 */
        uint32_t len = h->lens[tid], start = 0;
        do {
            uint32_t stop = start + CHUNK_SIZE;
            if (genes) stop = extend_stop(&gsets[c], stop);
            if (n == cap) out = realloc(out, (cap *= 2) * sizeof *out);
            out[n++] = (chunk_t){c, tid, start, stop};
            start = stop;
        } while (start < len);
    }
    *n_out = n;
    return out;
}

/* ------------------------------------------------------------------ exact qname sets */
typedef struct mmnode {
    struct mmnode* next;
    uint32_t gene;
    uint16_t len;
    uint8_t bucket;
    uint8_t name[];
} mmnode;
typedef struct {
    mmnode** tab;
    size_t cap, n;
    char** arenas;
    size_t n_arenas, arena_used, arena_cap;
} mmset_t;
static void mm_init(mmset_t* s) {
    memset(s, 0, sizeof *s);
    s->cap = 1024;
    s->tab = calloc(s->cap, sizeof(mmnode*));
}
static void* mm_alloc(mmset_t* s, size_t n) {
    n = (n + 7) & ~(size_t)7;
    if (!s->n_arenas || s->arena_used + n > s->arena_cap) {
        s->arena_cap = n > (1 << 20) ? n : (1 << 20);
        s->arenas = realloc(s->arenas, (s->n_arenas + 1) * sizeof(char*));
        s->arenas[s->n_arenas++] = malloc(s->arena_cap);
        s->arena_used = 0;
    }
    void* p = s->arenas[s->n_arenas - 1] + s->arena_used;
    s->arena_used += n;
    return p;
}
static void mm_free(mmset_t* s) {
    for (size_t i = 0; i < s->n_arenas; i++) free(s->arenas[i]);
    free(s->arenas);
    free(s->tab);
}
static uint64_t fnv(const uint8_t* p, size_t n, uint64_t h) {
    for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}
/* returns 1 if newly inserted */
static int mm_insert(mmset_t* s, uint8_t bucket, uint32_t gene, const uint8_t* name, size_t len) {
    uint64_t h = fnv(name, len, 1469598103934665603ull ^ ((uint64_t)gene << 1 | bucket));
    if (s->n * 2 > s->cap) {
        size_t ncap = s->cap * 4;
        mmnode** nt = calloc(ncap, sizeof(mmnode*));
        for (size_t i = 0; i < s->cap; i++)
            for (mmnode* p = s->tab[i]; p;) {
                mmnode* nx = p->next;
                uint64_t hh = fnv(p->name, p->len, 1469598103934665603ull ^ ((uint64_t)p->gene << 1 | p->bucket));
                p->next = nt[hh & (ncap - 1)];
                nt[hh & (ncap - 1)] = p;
                p = nx;
            }
        free(s->tab);
        s->tab = nt;
        s->cap = ncap;
    }
    mmnode** slot = &s->tab[h & (s->cap - 1)];
    for (mmnode* p = *slot; p; p = p->next)
        if (p->gene == gene && p->bucket == bucket && p->len == len && !memcmp(p->name, name, len)) return 0;
    mmnode* nn = mm_alloc(s, sizeof(mmnode) + len);
    nn->gene = gene;
    nn->bucket = bucket;
    nn->len = (uint16_t)len;
    memcpy(nn->name, name, len);
    nn->next = *slot;
    *slot = nn;
    s->n++;
    return 1;
}

/* ------------------------------------------------------------------ record helpers */
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

/* htslib bam_aux_get + rust-htslib Aux::Integer.  returns 0 absent, 1 integer (*val), -1 present but
 * not an integer (rust-htslib `.integer()` panics there), -2 malformed aux area. */
static int aux_nh(const uint8_t* p, const uint8_t* e, int64_t* val) {
    while (p + 3 <= e) {
        int match = p[0] == 'N' && p[1] == 'H';
        uint8_t t = p[2];
        p += 3;
        size_t sz;
        switch (t) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'd': sz = 8; break;
            case 'Z': case 'H': {
                if (match) return -1;
                while (p < e && *p) p++;
                if (p >= e) return -2;
                p++;
                continue;
            }
            case 'B': {
                if (match) return -1;
                if (p + 5 > e) return -2;
                uint8_t st = p[0];
                uint32_t cnt = rd32(p + 1);
                size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                p += 5 + (size_t)cnt * es;
                continue;
            }
            default: return -2;
        }
        if (p + sz > e) return -2;
        if (match) {
            switch (t) {
                case 'c': *val = (int8_t)p[0]; return 1;
                case 'C': *val = p[0]; return 1;
                case 's': *val = (int16_t)rd16(p); return 1;
                case 'S': *val = rd16(p); return 1;
                case 'i': *val = (int32_t)rd32(p); return 1;
                case 'I': *val = rd32(p); return 1;
                default: return -1;
            }
        }
        p += sz;
    }
    return 0;
}

/* rust-htslib Record::aux(tag) for string values (types Z and H -> Aux::String): 1 found, 0 absent, -1 another type,
 * -2 malformed */
static int aux_str(const uint8_t* p, const uint8_t* e, char t0, char t1, const uint8_t** sp, size_t* slen) {
    while (p + 3 <= e) {
        int match = p[0] == (uint8_t)t0 && p[1] == (uint8_t)t1;
        uint8_t t = p[2];
        p += 3;
        size_t sz;
        switch (t) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'd': sz = 8; break;
            case 'Z': case 'H': {
                const uint8_t* s0 = p;
                while (p < e && *p) p++;
                if (p >= e) return -2;
                if (match) { *sp = s0; *slen = (size_t)(p - s0); return 1; }
                p++;
                continue;
            }
            case 'B': {
                if (match) return -1;
                if (p + 5 > e) return -2;
                uint8_t st = p[0];
                uint32_t cnt = rd32(p + 1);
                size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                p += 5 + (size_t)cnt * es;
                continue;
            }
            default: return -2;
        }
        if (p + sz > e) return -2;
        if (match) return -1;
        p += sz;
    }
    return 0;
}

/* rust-htslib aligned_blocks / introns through src/bam_ext.rs:28-42. */
int32_t oracle_cigar_pairs(const uint32_t* cigar, int32_t n_ops, int32_t pos, int32_t want_introns,
                           uint32_t* out, int32_t cap) {
    uint32_t p = (uint32_t)pos;
    int32_t n = 0;
    for (int i = 0; i < n_ops; i++) {
        uint32_t op = cigar[i] & 15, len = cigar[i] >> 4;
        int is_m = op == 0 || op == 7 || op == 8;
        if ((is_m && !want_introns) || (op == 3 && want_introns)) {
            if (n < cap) { out[2 * n] = p; out[2 * n + 1] = p + len; }
            n++;
        }
        if (is_m || op == 2 || op == 3) p += len;
    }
    return n;
}

/* ------------------------------------------------------------------ the counting job */
typedef struct {
    const char* bam_path;
    bamidx_t* idx;
    const mbfbam_intervals* iv;
    ivset_t* sets;      /* per intervals-contig */
    int32_t* tid2c;     /* header tid -> contig index in `iv` (-1) */
    uint64_t* gene_base;
    chunk_t* chunks;
    size_t n_chunks;
    size_t next_chunk;  /* atomic */
    int mode;
    int once;  /* each_read_counts_once */
    int introns;
    int quantify; /* quantify_gene_reads: imaps hold (bucket, global gene, pos) -> reads */
    int barcode;  /* by_barcode: entries (chunk, gene, barcode, distinct UMIs) appended to bc_* under `mu` */
    int dups;     /* duplicate distribution: one chunk per reference, imaps hold (0, k, 0) -> groups of k equal reads */
    int nthreads;
    int failed;
    char err[256];
    pthread_mutex_t mu;
    /* outputs */
    uint64_t *fwd, *rev;
    double *wfwd, *wrev; /* nthreads private copies reduced at the end */
    uint64_t outside;
    uint64_t n_records;
    /* introns: per-thread maps merged at the end */
    struct imap* imaps;
    /* by_barcode */
    uint32_t *bc_gene, *bc_chunk, *bc_count;
    uint64_t* bc_off;
    char* bc_bytes;
    size_t bc_n, bc_cap, bc_blen, bc_bcap;
} job_t;

/* (tid,start,stop) -> count, open addressing */
typedef struct imap {
    uint64_t* key; /* start | stop<<32 */
    int32_t* tid;  /* -1 empty */
    uint64_t* cnt;
    size_t cap, n;
} imap;
static void imap_init(imap* m, size_t cap) {
    m->cap = cap; m->n = 0;
    m->key = malloc(cap * 8);
    m->tid = malloc(cap * 4);
    m->cnt = calloc(cap, 8);
    for (size_t i = 0; i < cap; i++) m->tid[i] = -1;
}
static void imap_free(imap* m) { free(m->key); free(m->tid); free(m->cnt); }
static void imap_add(imap* m, int32_t tid, uint32_t s, uint32_t e, uint64_t c);
static void imap_grow(imap* m) {
    imap o = *m;
    imap_init(m, o.cap * 2);
    for (size_t i = 0; i < o.cap; i++)
        if (o.tid[i] >= 0) imap_add(m, o.tid[i], (uint32_t)o.key[i], (uint32_t)(o.key[i] >> 32), o.cnt[i]);
    imap_free(&o);
}
static void imap_add(imap* m, int32_t tid, uint32_t s, uint32_t e, uint64_t c) {
    if ((m->n + 1) * 2 > m->cap) imap_grow(m);
    uint64_t k = (uint64_t)s | ((uint64_t)e << 32);
    uint64_t h = (k ^ ((uint64_t)tid * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    size_t i = h & (m->cap - 1);
    for (;; i = (i + 1) & (m->cap - 1)) {
        if (m->tid[i] < 0) { m->tid[i] = tid; m->key[i] = k; m->cnt[i] = c; m->n++; return; }
        if (m->tid[i] == tid && m->key[i] == k) { m->cnt[i] += c; return; }
    }
}

typedef struct { uint32_t* v; size_t n, cap; } u32vec;
static inline int vec_add_unique(u32vec* s, uint32_t x) {
    for (size_t i = 0; i < s->n; i++) if (s->v[i] == x) return 0;
    if (s->n == s->cap) s->v = realloc(s->v, (s->cap = s->cap ? s->cap * 2 : 16) * 4);
    s->v[s->n++] = x;
    return 1;
}

static void fail(job_t* j, const char* msg) {
    pthread_mutex_lock(&j->mu);
    if (!j->failed) { j->failed = 1; snprintf(j->err, sizeof j->err, "%s", msg); }
    pthread_mutex_unlock(&j->mu);
}

/* duplicate_distribution.rs:65-69: every distinct (strand, sequence) among the reads of one position adds 1 to
 * counts[number of reads that have it] */
static const uint8_t* dup_base;
static const uint32_t* dup_offs;
static __thread const uint8_t* t_dup_base;
static __thread const uint32_t* t_dup_offs;
static int dup_cmp(const void* a_, const void* b_) {
    uint32_t a = *(const uint32_t*)a_, b = *(const uint32_t*)b_;
    uint32_t la = t_dup_offs[a + 1] - t_dup_offs[a], lb = t_dup_offs[b + 1] - t_dup_offs[b];
    if (la != lb) return la < lb ? -1 : 1;
    return memcmp(t_dup_base + t_dup_offs[a], t_dup_base + t_dup_offs[b], la);
}
static void dup_flush(imap* im, const uint8_t* grp, const uint32_t* off, size_t n) {
    if (!n) return;
    (void)dup_base; (void)dup_offs;
    uint32_t* order = malloc(n * 4);
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    t_dup_base = grp; t_dup_offs = off;
    qsort(order, n, 4, dup_cmp);
    size_t run = 1;
    for (size_t i = 1; i <= n; i++) {
        if (i < n && dup_cmp(&order[i - 1], &order[i]) == 0) { run++; continue; }
        imap_add(im, 0, (uint32_t)run, 0, 1);
        run = 1;
    }
    free(order);
}

/* one chunk == count_reads_in_region_{unstranded,stranded} (counters.rs:26-106,114-202)
 * or count_introns (introns.rs:24-51) */
static void run_chunk(job_t* j, int thread, const chunk_t* ck, uint8_t** rbuf, size_t* rcap) {
    bamidx_t* h = j->idx;
    int tid = ck->tid;
    if (tid >= h->bai_n_ref || !h->has_bins[tid]) return;
    uint32_t w = ck->start >> 14;
    if ((int32_t)w >= h->n_intv[tid]) return; /* nothing at or right of start */
    uint64_t voff = h->ioffset[tid][w];
    if (voff == 0) voff = h->ref_beg[tid];
    bgzf_t b;
    if (bgzf_open(&b, j->bam_path)) { fail(j, "Could not read bam: open failed"); return; } /* counters.rs:223 */
    if (bgzf_seek(&b, voff)) { bgzf_close(&b); fail(j, "Fetch error"); return; }
    const int no_sets = j->introns || j->dups;
    const ivset_t* set = no_sets ? NULL : &j->sets[ck->contig];
    uint64_t gbase = no_sets ? 0 : j->gene_base[ck->contig];
    uint32_t ngenes = no_sets ? 0 : j->iv->n_genes[ck->contig];
    uint64_t* lf = NULL; uint64_t* lr = NULL;
    double *wf = NULL, *wr = NULL;
    if (!j->introns && !j->dups && !j->quantify && !j->barcode) {
        if (j->mode == MBFBAM_MODE_WEIGHTED) {
            size_t tot = (size_t)j->gene_base[j->iv->n_contigs];
            wf = j->wfwd + (size_t)thread * tot + gbase;
            wr = j->wrev + (size_t)thread * tot + gbase;
        } else {
            lf = calloc((size_t)ngenes + 1, 8);
            lr = calloc((size_t)ngenes + 1, 8);
        }
    }
    mmset_t mm;
    mm_init(&mm);
    u32vec seen[2] = {{0, 0, 0}, {0, 0, 0}};
    uint64_t outside = 0, nrec = 0;
    imap* im = (j->introns || j->quantify || j->dups) ? &j->imaps[thread] : NULL;
    /* duplicate distribution: the reads of the current position (duplicate_distribution.rs:57-75) */
    uint8_t* grp = NULL; size_t grp_len = 0, grp_cap = 0; uint32_t* grp_off = NULL; size_t grp_n = 0, grp_ncap = 0;
    int64_t grp_pos = -1;
    /* by_barcode: scratch key, and one (gene, barcode) entry per distinct (gene, barcode, pos, strand, UMI) */
    int bc_failed = 0;
    uint8_t* bkey = NULL; size_t bkey_cap = 0;
    uint8_t* bent = NULL; size_t bent_len = 0, bent_cap = 0; uint32_t* bent_off = NULL; size_t bent_n = 0, bent_ncap = 0;
    for (;;) {
        int32_t bs;
        if (bgzf_read(&b, &bs, 4) != 4) break; /* Err(_) ends the loop, counters.rs:41 */
        if (bs < 32) { fail(j, "corrupt BAM record"); break; }
        if ((size_t)bs > *rcap) { *rcap = (size_t)bs * 2; *rbuf = realloc(*rbuf, *rcap); }
        uint8_t* r = *rbuf;
        if (bgzf_read(&b, r, (size_t)bs) != (size_t)bs) break;
        int32_t rtid = (int32_t)rd32(r), pos = (int32_t)rd32(r + 4);
        if (rtid != tid) {
            if (rtid < tid && rtid >= 0) continue;
            break; /* iterator ends at the next reference */
        }
        if (pos >= 0 && (uint32_t)pos >= ck->stop) break;
        uint32_t upos = (uint32_t)pos;
        if (upos < ck->start || upos >= ck->stop) continue; /* counters.rs:46-48 */
        nrec++;
        uint32_t l_qn = r[8], n_cig = rd16(r + 12), flag = rd16(r + 14), l_seq = rd32(r + 16);
        const uint8_t* qn = r + 32;
        const uint8_t* cig = qn + l_qn;
        const uint8_t* aux = cig + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq;
        const uint8_t* end = r + bs;
        if (aux > end) { fail(j, "corrupt BAM record"); break; }
        if (j->introns) {
            /* introns.rs:38-48 */
            uint32_t p = upos, k = 0;
            for (uint32_t i = 0; i < n_cig; i++) {
                uint32_t c = rd32(cig + 4 * i), op = c & 15, len = c >> 4;
                if (op == 3) { imap_add(im, tid, p, p + len, 1); k++; }
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) p += len;
            }
            imap_add(im, tid, UINT32_MAX, UINT32_MAX - k, 1);
            continue;
        }
        if (j->dups) {
            /* key = (is_reverse, decoded sequence): l_seq + the packed bases, spare nibble of an odd length cleared */
            if ((int64_t)pos != grp_pos) {
                dup_flush(im, grp, grp_off, grp_n);
                grp_n = 0; grp_len = 0; grp_pos = pos;
            }
            const uint8_t* seq = cig + 4 * (size_t)n_cig;
            size_t nb = ((size_t)l_seq + 1) / 2, klen = 5 + nb;
            if (grp_len + klen > grp_cap) { grp_cap = (grp_len + klen) * 2 + 64; grp = realloc(grp, grp_cap); }
            if (grp_n + 2 > grp_ncap) { grp_ncap = grp_ncap ? grp_ncap * 2 : 64; grp_off = realloc(grp_off, grp_ncap * 4); }
            grp_off[grp_n++] = (uint32_t)grp_len;
            grp[grp_len] = (flag & 0x10) ? 1 : 0;
            memcpy(grp + grp_len + 1, &l_seq, 4);
            memcpy(grp + grp_len + 5, seq, nb);
            if ((l_seq & 1) && nb) grp[grp_len + 5 + nb - 1] &= 0xf0;
            grp_len += klen;
            grp_off[grp_n] = (uint32_t)grp_len;
            continue;
        }
        if (j->barcode) {
            /* by_barcode.rs:125-162 */
            if (bc_failed || (flag & 0x100)) continue;
            int rv = (flag & 0x10) != 0;
            uint32_t p = upos;
            for (uint32_t i = 0; i < n_cig && !bc_failed; i++) {
                uint32_t c = rd32(cig + 4 * i), op = c & 15, len = c >> 4;
                if (op == 0 || op == 7 || op == 8) {
                    uint32_t b0 = p, b1 = p + len;
                    p += len;
                    if (b1 < ck->start || b0 >= ck->stop || (b0 < ck->start && b1 >= ck->start)) continue; /* :131 */
                    tfind_t f;
                    t_find_begin(set, &f, b0, b1);
                    for (int32_t k; (k = t_find_next(set, &f)) >= 0;) {
                        int st = set->nd[k].strand;
                        if (!((st == 1 && !rv) || (st != 1 && rv))) continue;
                        const uint8_t *xc, *xm;
                        size_t lc, lm;
                        if (aux_str(aux, end, 'X', 'C', &xc, &lc) != 1 || aux_str(aux, end, 'X', 'M', &xm, &lm) != 1) {
                            bc_failed = 1; /* `?`: the chunk's function returns Err, the driver drops the chunk (:46-68) */
                            break;
                        }
                        /* positions[(gene_no, barcode, pos, is_reverse)].insert(umi) */
                        size_t kl = 4 + 4 + 1 + 2 + lc + lm;
                        if (kl > bkey_cap) { bkey_cap = kl * 2; bkey = realloc(bkey, bkey_cap); }
                        uint32_t g = set->nd[k].gene;
                        uint16_t lc16 = (uint16_t)lc;
                        memcpy(bkey, &g, 4); memcpy(bkey + 4, &pos, 4); bkey[8] = (uint8_t)rv; memcpy(bkey + 9, &lc16, 2);
                        memcpy(bkey + 11, xc, lc); memcpy(bkey + 11 + lc, xm, lm);
                        if (mm_insert(&mm, 0, 0, bkey, kl)) {
                            /* a new UMI for this (gene, barcode, pos, strand): one more for (gene, barcode) */
                            size_t el = 4 + lc;
                            if (bent_len + el + 8 > bent_cap) { bent_cap = (bent_len + el + 8) * 2; bent = realloc(bent, bent_cap); }
                            if (bent_n + 2 > bent_ncap) { bent_ncap = bent_ncap ? bent_ncap * 2 : 256; bent_off = realloc(bent_off, bent_ncap * 4); }
                            bent_off[bent_n++] = (uint32_t)bent_len;
                            memcpy(bent + bent_len, &g, 4);
                            memcpy(bent + bent_len + 4, xc, lc);
                            bent_len += el;
                            bent_off[bent_n] = (uint32_t)bent_len;
                        }
                    }
                } else if (op == 2 || op == 3) {
                    p += len;
                }
            }
            continue;
        }
        if (j->quantify) {
            /* quantify.rs:113-142: one alignment counts once per gene, no NH logic, any flag */
            seen[0].n = 0;
            uint32_t p = upos;
            for (uint32_t i = 0; i < n_cig; i++) {
                uint32_t c = rd32(cig + 4 * i), op = c & 15, len = c >> 4;
                if (op == 0 || op == 7 || op == 8) {
                    uint32_t b0 = p, b1 = p + len;
                    p += len;
                    tfind_t f;
                    t_find_begin(set, &f, b0, b1);
                    for (int32_t k; (k = t_find_next(set, &f)) >= 0;) {
                        uint32_t g = set->nd[k].gene;
                        if (!vec_add_unique(&seen[0], g)) continue;
                        int rv = (flag & 0x10) != 0, st = set->nd[k].strand;
                        int congruent = (st == 1 && !rv) || (st != 1 && rv);
                        imap_add(im, congruent ? 0 : 1, (uint32_t)(gbase + g), upos, 1);
                    }
                } else if (op == 2 || op == 3) {
                    p += len;
                }
            }
            continue;
        }
        seen[0].n = seen[1].n = 0;
        int hit = 0, have_nh = 0;
        int64_t nh = 1;
        uint32_t p = upos;
        for (uint32_t i = 0; i < n_cig; i++) {
            uint32_t c = rd32(cig + 4 * i), op = c & 15, len = c >> 4;
            if (op == 0 || op == 7 || op == 8) {
                uint32_t b0 = p, b1 = p + len;
                p += len;
                if (j->mode == MBFBAM_MODE_UNSTRANDED && b0 >= ck->stop) continue; /* counters.rs:52 */
                tfind_t f;
                t_find_begin(set, &f, b0, b1);
                for (int32_t k; (k = t_find_next(set, &f)) >= 0;) { /* counters.rs:61 / :144 tree.find(iv.0..iv.1) */
                    hit = 1;
                    if (!have_nh) {
                        int rc = aux_nh(aux, end, &nh); /* counters.rs:65-66 */
                        if (rc == 0) nh = 1;
                        else if (rc < 0) { fail(j, rc == -1 ? "NH tag is not an integer" : "corrupt aux data"); nh = 1; }
                        have_nh = 1;
                    }
                    int bucket = 0;
                    if (j->mode != MBFBAM_MODE_UNSTRANDED) { /* counters.rs:151 */
                        int rv = (flag & 0x10) != 0, st = set->nd[k].strand;
                        bucket = ((st == 1 && !rv) || (st != 1 && rv)) ? 0 : 1;
                    }
                    uint32_t g = set->nd[k].gene;
                    if (j->mode == MBFBAM_MODE_WEIGHTED || nh == 1) vec_add_unique(&seen[bucket], g);
                    else mm_insert(&mm, (uint8_t)bucket, g, qn, l_qn ? l_qn - 1 : 0);
                    if (j->once) break; /* counters.rs:83-85 / :173-175 */
                }
                if (j->once && hit) break; /* counters.rs:87-92 / :177-182 */
            } else if (op == 2 || op == 3) {
                p += len;
            }
        }
        if (!hit) outside++;
        if (j->mode == MBFBAM_MODE_WEIGHTED) {
            double wgt = nh >= 1 ? 1.0 / (double)nh : 1.0;
            for (size_t i = 0; i < seen[0].n; i++) wf[seen[0].v[i]] += wgt;
            for (size_t i = 0; i < seen[1].n; i++) wr[seen[1].v[i]] += wgt;
        } else {
            for (size_t i = 0; i < seen[0].n; i++) lf[seen[0].v[i]]++;
            for (size_t i = 0; i < seen[1].n; i++) lr[seen[1].v[i]]++;
        }
    }
    if (j->dups) dup_flush(im, grp, grp_off, grp_n);
    free(grp); free(grp_off);
    if (j->barcode && !bc_failed && bent_n) {
        /* by_barcode.rs:164-180: sum of the set sizes per (gene_no, barcode) */
        uint32_t* order = malloc(bent_n * 4);
        for (size_t i = 0; i < bent_n; i++) order[i] = (uint32_t)i;
        t_dup_base = bent; t_dup_offs = bent_off;
        qsort(order, bent_n, 4, dup_cmp);
        pthread_mutex_lock(&j->mu);
        size_t run = 1;
        for (size_t i = 1; i <= bent_n; i++) {
            if (i < bent_n && dup_cmp(&order[i - 1], &order[i]) == 0) { run++; continue; }
            const uint8_t* e0 = bent + bent_off[order[i - 1]];
            size_t el = bent_off[order[i - 1] + 1] - bent_off[order[i - 1]], bl = el - 4;
            uint32_t g;
            memcpy(&g, e0, 4);
            if (j->bc_n + 2 > j->bc_cap) {
                j->bc_cap = j->bc_cap ? j->bc_cap * 2 : 1024;
                j->bc_gene = realloc(j->bc_gene, j->bc_cap * 4); j->bc_chunk = realloc(j->bc_chunk, j->bc_cap * 4);
                j->bc_count = realloc(j->bc_count, j->bc_cap * 4); j->bc_off = realloc(j->bc_off, (j->bc_cap + 1) * 8);
            }
            if (j->bc_blen + bl + 1 > j->bc_bcap) { j->bc_bcap = (j->bc_blen + bl + 1) * 2; j->bc_bytes = realloc(j->bc_bytes, j->bc_bcap); }
            j->bc_gene[j->bc_n] = (uint32_t)(gbase + g);
            j->bc_chunk[j->bc_n] = (uint32_t)(ck - j->chunks);
            j->bc_count[j->bc_n] = (uint32_t)run;
            j->bc_off[j->bc_n] = j->bc_blen;
            memcpy(j->bc_bytes + j->bc_blen, e0 + 4, bl);
            j->bc_blen += bl;
            j->bc_n++;
            j->bc_off[j->bc_n] = j->bc_blen;
            run = 1;
        }
        pthread_mutex_unlock(&j->mu);
        free(order);
    }
    free(bkey); free(bent); free(bent_off);
    if (b.err) fail(j, "corrupt BGZF block");
    if (!j->introns && !j->dups && !j->quantify && !j->barcode && j->mode != MBFBAM_MODE_WEIGHTED) {
        /* counters.rs:102-104 / :195-200: add the sizes of the multi-mapper sets */
        for (size_t i = 0; i < mm.cap; i++)
            for (mmnode* q = mm.tab[i]; q; q = q->next) (q->bucket ? lr : lf)[q->gene]++;
        for (uint32_t g = 0; g < ngenes; g++) {
            if (lf[g]) __atomic_fetch_add(&j->fwd[gbase + g], lf[g], __ATOMIC_RELAXED);
            if (lr[g]) __atomic_fetch_add(&j->rev[gbase + g], lr[g], __ATOMIC_RELAXED);
        }
    }
    __atomic_fetch_add(&j->outside, outside, __ATOMIC_RELAXED);
    __atomic_fetch_add(&j->n_records, nrec, __ATOMIC_RELAXED);
    free(lf); free(lr);
    free(seen[0].v); free(seen[1].v);
    mm_free(&mm);
    bgzf_close(&b);
}

typedef struct { job_t* j; int thread; } targ_t;
static void* worker(void* a_) {
    targ_t* a = a_;
    job_t* j = a->j;
    uint8_t* rbuf = NULL;
    size_t rcap = 0;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next_chunk, 1, __ATOMIC_RELAXED);
        if (i >= j->n_chunks) break;
        run_chunk(j, a->thread, &j->chunks[i], &rbuf, &rcap);
    }
    free(rbuf);
    return NULL;
}
static void run_pool(job_t* j) {
    int nt = j->nthreads < 1 ? 1 : j->nthreads;
    pthread_t* th = malloc((size_t)nt * sizeof *th);
    targ_t* ta = malloc((size_t)nt * sizeof *ta);
    for (int t = 0; t < nt; t++) { ta[t] = (targ_t){j, t}; pthread_create(&th[t], NULL, worker, &ta[t]); }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    free(th); free(ta);
}

static int open_idx(const char* bam, const char* bai, bamidx_t* idx, char* err, size_t errlen) {
    memset(idx, 0, sizeof *idx);
    if (read_header(bam, idx)) { snprintf(err, errlen, "Could not read bam: %s", bam); return MBFBAM_ERR_IO; }
    char* bp = default_bai(bam, bai);
    int rc = bp ? read_bai(bp, idx) : -1;
    free(bp);
    if (rc) { snprintf(err, errlen, "Could not read bam: index of %s", bam); return MBFBAM_ERR_IO; }
    return 0;
}

/* counters.rs:204-260 / :276-323 without the String maps (assembled by the test helper).
 * sample_chunks > 0: run only every k-th chunk so that about `sample_chunks` chunks run
 * (bench.py's bounded CPU sample); n_records_out reports the records they owned.
 * chunk_lo >= 0: only chunks [chunk_lo, chunk_hi) of the chunk table (what one shard owns). */
int oracle_count_reads(const char* bam, const char* bai, const mbfbam_intervals* iv,
                       const mbfbam_intervals* genes, int mode, int each_read_counts_once, int nthreads, int64_t sample_chunks,
                       int64_t chunk_lo, int64_t chunk_hi,
                       uint64_t* fwd, uint64_t* rev, double* wfwd, double* wrev, uint8_t* scanned,
                       uint64_t* outside, uint64_t* n_records_out, int64_t* n_chunks_out, char* err,
                       size_t errlen) {
    bamidx_t idx;
    int rc = open_idx(bam, bai, &idx, err, errlen);
    if (rc) return rc;
    job_t j;
    memset(&j, 0, sizeof j);
    pthread_mutex_init(&j.mu, NULL);
    j.bam_path = bam; j.idx = &idx; j.iv = iv; j.mode = mode; j.once = each_read_counts_once != 0; j.nthreads = nthreads < 1 ? 1 : nthreads;
    j.sets = calloc((size_t)iv->n_contigs + 1, sizeof(ivset_t));
    j.gene_base = calloc((size_t)iv->n_contigs + 1, 8);
    for (int c = 0; c < iv->n_contigs; c++) {
        build_ivset(iv, c, &j.sets[c]);
        j.gene_base[c + 1] = j.gene_base[c] + iv->n_genes[c];
    }
    size_t tot = (size_t)j.gene_base[iv->n_contigs];
    ivset_t* gsets = calloc((size_t)genes->n_contigs + 1, sizeof(ivset_t));
    for (int c = 0; c < genes->n_contigs; c++) build_ivset(genes, c, &gsets[c]);
    size_t nck;
    chunk_t* cks = make_chunks(&idx, genes, gsets, &nck);
    /* chunk.contig currently indexes `genes`; remap to `iv` (counters.rs:224 trees.get(&chunk.chr).unwrap()) */
    rc = 0;
    for (size_t i = 0; i < nck; i++) {
        const char* nm = genes->contig_names[cks[i].contig];
        int c2 = -1;
        for (int c = 0; c < iv->n_contigs; c++) if (!strcmp(iv->contig_names[c], nm)) { c2 = c; break; }
        if (c2 < 0) { snprintf(err, errlen, "contig %s is in gene_intervals but not in intervals", nm); rc = MBFBAM_ERR_ARG; break; }
        cks[i].contig = c2;
        if (scanned) scanned[c2] = 1;
    }
    if (!rc && chunk_lo >= 0) { /* only the chunks [chunk_lo, chunk_hi) of the table (one shard) */
        size_t lo = (size_t)chunk_lo < nck ? (size_t)chunk_lo : nck, hi = (size_t)chunk_hi < nck ? (size_t)chunk_hi : nck;
        if (hi < lo) hi = lo;
        memmove(cks, cks + lo, (hi - lo) * sizeof *cks);
        nck = hi - lo;
    }
    if (!rc) {
        if (sample_chunks > 0 && (size_t)sample_chunks < nck) {
            size_t step = nck / (size_t)sample_chunks, m = 0;
            for (size_t i = 0; i < nck; i += step) cks[m++] = cks[i];
            nck = m;
        }
        j.chunks = cks; j.n_chunks = nck;
        j.fwd = fwd; j.rev = rev;
        if (mode == MBFBAM_MODE_WEIGHTED) {
            j.wfwd = calloc(tot * (size_t)j.nthreads + 1, 8);
            j.wrev = calloc(tot * (size_t)j.nthreads + 1, 8);
        }
        run_pool(&j);
        if (mode == MBFBAM_MODE_WEIGHTED) {
            for (size_t g = 0; g < tot; g++) {
                double a = 0, b = 0;
                for (int t = 0; t < j.nthreads; t++) { a += j.wfwd[(size_t)t * tot + g]; b += j.wrev[(size_t)t * tot + g]; }
                wfwd[g] = a; wrev[g] = b;
            }
            free(j.wfwd); free(j.wrev);
        }
        *outside = j.outside;
        if (n_records_out) *n_records_out = j.n_records;
        if (n_chunks_out) *n_chunks_out = (int64_t)nck;
        if (j.failed) { snprintf(err, errlen, "%s", j.err); rc = MBFBAM_ERR_FORMAT; }
    }
    for (int c = 0; c < iv->n_contigs; c++) free_ivset(&j.sets[c]);
    for (int c = 0; c < genes->n_contigs; c++) free_ivset(&gsets[c]);
    free(j.sets); free(gsets); free(j.gene_base); free(cks);
    free_idx(&idx);
    return rc;
}

/* introns.rs:56-88.  Result arrays are malloc'ed; free with oracle_free. */
int oracle_count_introns(const char* bam, const char* bai, int nthreads, int64_t sample_chunks, int64_t* n_chunks_out,
                         uint64_t* n_out, int32_t** tid_out,
                         uint32_t** start_out, uint32_t** stop_out, uint64_t** count_out,
                         uint8_t* contig_seen, int32_t n_targets_cap, uint64_t* n_records_out, char* err,
                         size_t errlen) {
    bamidx_t idx;
    int rc = open_idx(bam, bai, &idx, err, errlen);
    if (rc) return rc;
    job_t j;
    memset(&j, 0, sizeof j);
    pthread_mutex_init(&j.mu, NULL);
    j.bam_path = bam; j.idx = &idx; j.introns = 1; j.nthreads = nthreads < 1 ? 1 : nthreads;
    size_t nck;
    chunk_t* cks = make_chunks(&idx, NULL, NULL, &nck);
    if (sample_chunks > 0 && (size_t)sample_chunks < nck) { /* bounded sample for the timed CPU baseline */
        size_t step = nck / (size_t)sample_chunks, m = 0;
        for (size_t i = 0; i < nck; i += step) cks[m++] = cks[i];
        nck = m;
    }
    if (n_chunks_out) *n_chunks_out = (int64_t)nck;
    j.chunks = cks; j.n_chunks = nck;
    j.imaps = calloc((size_t)j.nthreads, sizeof(imap));
    for (int t = 0; t < j.nthreads; t++) imap_init(&j.imaps[t], 1 << 12);
    run_pool(&j);
    imap all;
    imap_init(&all, 1 << 12);
    for (int t = 0; t < j.nthreads; t++) {
        imap* m = &j.imaps[t];
        for (size_t i = 0; i < m->cap; i++)
            if (m->tid[i] >= 0) imap_add(&all, m->tid[i], (uint32_t)m->key[i], (uint32_t)(m->key[i] >> 32), m->cnt[i]);
        imap_free(m);
    }
    free(j.imaps);
    *n_out = all.n;
    *tid_out = malloc((all.n + 1) * 4);
    *start_out = malloc((all.n + 1) * 4);
    *stop_out = malloc((all.n + 1) * 4);
    *count_out = malloc((all.n + 1) * 8);
    size_t k = 0;
    for (size_t i = 0; i < all.cap; i++)
        if (all.tid[i] >= 0) {
            (*tid_out)[k] = all.tid[i];
            (*start_out)[k] = (uint32_t)all.key[i];
            (*stop_out)[k] = (uint32_t)(all.key[i] >> 32);
            (*count_out)[k] = all.cnt[i];
            k++;
        }
    imap_free(&all);
    /* introns.rs:75-79: every chunk whose scan did not fail inserts its contig key */
    for (int t = 0; t < idx.n_ref && t < n_targets_cap; t++) contig_seen[t] = 1;
    if (n_records_out) *n_records_out = j.n_records;
    if (j.failed) { snprintf(err, errlen, "%s", j.err); rc = MBFBAM_ERR_FORMAT; }
    free(cks);
    free_idx(&idx);
    return rc;
}
/* quantify.rs:20-75 without the String maps.  Entries (bucket 0 congruent / 1 incongruent, global gene slot, pos, reads);
 * result arrays are malloc'ed; free with oracle_free. */
int oracle_quantify(const char* bam, const char* bai, const mbfbam_intervals* iv, const mbfbam_intervals* genes, int nthreads,
                    uint64_t* n_out, int32_t** bucket_out, uint32_t** gene_out, uint32_t** pos_out, uint64_t** count_out,
                    uint8_t* scanned, char* err, size_t errlen) {
    bamidx_t idx;
    int rc = open_idx(bam, bai, &idx, err, errlen);
    if (rc) return rc;
    job_t j;
    memset(&j, 0, sizeof j);
    pthread_mutex_init(&j.mu, NULL);
    j.bam_path = bam; j.idx = &idx; j.iv = iv; j.mode = MBFBAM_MODE_STRANDED; j.quantify = 1; j.nthreads = nthreads < 1 ? 1 : nthreads;
    j.sets = calloc((size_t)iv->n_contigs + 1, sizeof(ivset_t));
    j.gene_base = calloc((size_t)iv->n_contigs + 1, 8);
    for (int c = 0; c < iv->n_contigs; c++) {
        build_ivset(iv, c, &j.sets[c]);
        j.gene_base[c + 1] = j.gene_base[c] + iv->n_genes[c];
    }
    ivset_t* gsets = calloc((size_t)genes->n_contigs + 1, sizeof(ivset_t));
    for (int c = 0; c < genes->n_contigs; c++) build_ivset(genes, c, &gsets[c]);
    size_t nck;
    chunk_t* cks = make_chunks(&idx, genes, gsets, &nck);
    for (size_t i = 0; i < nck; i++) {
        const char* nm = genes->contig_names[cks[i].contig];
        int c2 = -1;
        for (int c = 0; c < iv->n_contigs; c++) if (!strcmp(iv->contig_names[c], nm)) { c2 = c; break; }
        if (c2 < 0) { snprintf(err, errlen, "contig %s is in gene_intervals but not in intervals", nm); rc = MBFBAM_ERR_ARG; break; }
        cks[i].contig = c2;
        if (scanned) scanned[c2] = 1;
    }
    *n_out = 0;
    if (!rc) {
        j.chunks = cks; j.n_chunks = nck;
        j.imaps = calloc((size_t)j.nthreads, sizeof(imap));
        for (int t = 0; t < j.nthreads; t++) imap_init(&j.imaps[t], 1 << 12);
        run_pool(&j);
        imap all;
        imap_init(&all, 1 << 12);
        for (int t = 0; t < j.nthreads; t++) {
            imap* m = &j.imaps[t];
            for (size_t i = 0; i < m->cap; i++)
                if (m->tid[i] >= 0) imap_add(&all, m->tid[i], (uint32_t)m->key[i], (uint32_t)(m->key[i] >> 32), m->cnt[i]);
            imap_free(m);
        }
        free(j.imaps);
        *n_out = all.n;
        *bucket_out = malloc((all.n + 1) * 4);
        *gene_out = malloc((all.n + 1) * 4);
        *pos_out = malloc((all.n + 1) * 4);
        *count_out = malloc((all.n + 1) * 8);
        size_t k = 0;
        for (size_t i = 0; i < all.cap; i++)
            if (all.tid[i] >= 0) {
                (*bucket_out)[k] = all.tid[i];
                (*gene_out)[k] = (uint32_t)all.key[i];
                (*pos_out)[k] = (uint32_t)(all.key[i] >> 32);
                (*count_out)[k] = all.cnt[i];
                k++;
            }
        imap_free(&all);
        if (j.failed) { snprintf(err, errlen, "%s", j.err); rc = MBFBAM_ERR_FORMAT; }
    }
    for (int c = 0; c < iv->n_contigs; c++) free_ivset(&j.sets[c]);
    for (int c = 0; c < genes->n_contigs; c++) free_ivset(&gsets[c]);
    free(j.sets); free(gsets); free(j.gene_base); free(cks);
    free_idx(&idx);
    return rc;
}

/* duplicate_distribution.rs:44-86: a pool task per reference walks fetch(tid, 0, target_len).  Entries (k, number of
 * distinct (position, strand, sequence) seen exactly k times); free with oracle_free. */
int oracle_duplicate_distribution(const char* bam, const char* bai, int nthreads, uint64_t* n_out, uint32_t** k_out,
                                  uint64_t** count_out, uint64_t* n_records_out, char* err, size_t errlen) {
    bamidx_t idx;
    int rc = open_idx(bam, bai, &idx, err, errlen);
    if (rc) return rc;
    job_t j;
    memset(&j, 0, sizeof j);
    pthread_mutex_init(&j.mu, NULL);
    j.bam_path = bam; j.idx = &idx; j.dups = 1; j.nthreads = nthreads < 1 ? 1 : nthreads;
    size_t nck = (size_t)idx.n_ref;
    chunk_t* cks = calloc(nck + 1, sizeof *cks);
    for (int t = 0; t < idx.n_ref; t++) cks[t] = (chunk_t){t, t, 0, idx.lens[t]};
    j.chunks = cks; j.n_chunks = nck;
    j.imaps = calloc((size_t)j.nthreads, sizeof(imap));
    for (int t = 0; t < j.nthreads; t++) imap_init(&j.imaps[t], 1 << 8);
    run_pool(&j);
    imap all;
    imap_init(&all, 1 << 8);
    for (int t = 0; t < j.nthreads; t++) {
        imap* m = &j.imaps[t];
        for (size_t i = 0; i < m->cap; i++)
            if (m->tid[i] >= 0) imap_add(&all, m->tid[i], (uint32_t)m->key[i], (uint32_t)(m->key[i] >> 32), m->cnt[i]);
        imap_free(m);
    }
    free(j.imaps);
    *n_out = all.n;
    *k_out = malloc((all.n + 1) * 4);
    *count_out = malloc((all.n + 1) * 8);
    size_t k = 0;
    for (size_t i = 0; i < all.cap; i++)
        if (all.tid[i] >= 0) { (*k_out)[k] = (uint32_t)all.key[i]; (*count_out)[k] = all.cnt[i]; k++; }
    imap_free(&all);
    if (n_records_out) *n_records_out = j.n_records;
    if (j.failed) { snprintf(err, errlen, "%s", j.err); rc = MBFBAM_ERR_FORMAT; }
    free(cks);
    free_idx(&idx);
    return rc;
}

/* by_barcode.rs:18-92 without the String maps: entries (global gene slot, chunk index, distinct UMIs, barcode bytes) in
 * the order the pool produced them; free the arrays with oracle_free. */
int oracle_by_barcode(const char* bam, const char* bai, const mbfbam_intervals* iv, const mbfbam_intervals* genes, int nthreads,
                      uint64_t* n_out, uint32_t** gene_out, uint32_t** chunk_out, uint32_t** count_out, uint64_t** off_out,
                      char** bytes_out, char* err, size_t errlen) {
    bamidx_t idx;
    int rc = open_idx(bam, bai, &idx, err, errlen);
    if (rc) return rc;
    job_t j;
    memset(&j, 0, sizeof j);
    pthread_mutex_init(&j.mu, NULL);
    j.bam_path = bam; j.idx = &idx; j.iv = iv; j.mode = MBFBAM_MODE_STRANDED; j.barcode = 1; j.nthreads = nthreads < 1 ? 1 : nthreads;
    j.sets = calloc((size_t)iv->n_contigs + 1, sizeof(ivset_t));
    j.gene_base = calloc((size_t)iv->n_contigs + 1, 8);
    for (int c = 0; c < iv->n_contigs; c++) {
        build_ivset(iv, c, &j.sets[c]);
        j.gene_base[c + 1] = j.gene_base[c] + iv->n_genes[c];
    }
    ivset_t* gsets = calloc((size_t)genes->n_contigs + 1, sizeof(ivset_t));
    for (int c = 0; c < genes->n_contigs; c++) build_ivset(genes, c, &gsets[c]);
    size_t nck;
    chunk_t* cks = make_chunks(&idx, genes, gsets, &nck);
    for (size_t i = 0; i < nck; i++) {
        const char* nm = genes->contig_names[cks[i].contig];
        int c2 = -1;
        for (int c = 0; c < iv->n_contigs; c++) if (!strcmp(iv->contig_names[c], nm)) { c2 = c; break; }
        if (c2 < 0) { snprintf(err, errlen, "contig %s is in gene_intervals but not in intervals", nm); rc = MBFBAM_ERR_ARG; break; }
        cks[i].contig = c2;
    }
    *n_out = 0;
    if (!rc) {
        j.chunks = cks; j.n_chunks = nck;
        run_pool(&j);
        *n_out = j.bc_n;
        *gene_out = j.bc_gene; *chunk_out = j.bc_chunk; *count_out = j.bc_count; *off_out = j.bc_off; *bytes_out = j.bc_bytes;
        if (j.failed) { snprintf(err, errlen, "%s", j.err); rc = MBFBAM_ERR_FORMAT; }
    }
    for (int c = 0; c < iv->n_contigs; c++) free_ivset(&j.sets[c]);
    for (int c = 0; c < genes->n_contigs; c++) free_ivset(&gsets[c]);
    free(j.sets); free(gsets); free(j.gene_base); free(cks);
    free_idx(&idx);
    return rc;
}

void oracle_free(void* p) { free(p); }

int64_t oracle_chunks(const char* bam, const char* bai, const mbfbam_intervals* genes, int32_t* tid,
                      uint32_t* start, uint32_t* stop, int64_t cap) {
    bamidx_t idx;
    char err[64];
    if (open_idx(bam, bai, &idx, err, sizeof err)) return -1;
    ivset_t* gsets = NULL;
    if (genes) {
        gsets = calloc((size_t)genes->n_contigs + 1, sizeof(ivset_t));
        for (int c = 0; c < genes->n_contigs; c++) build_ivset(genes, c, &gsets[c]);
    }
    size_t n;
    chunk_t* cks = make_chunks(&idx, genes, gsets, &n);
    for (size_t i = 0; i < n && (int64_t)i < cap; i++) { tid[i] = cks[i].tid; start[i] = cks[i].start; stop[i] = cks[i].stop; }
    if (genes) { for (int c = 0; c < genes->n_contigs; c++) free_ivset(&gsets[c]); free(gsets); }
    free(cks);
    free_idx(&idx);
    return (int64_t)n;
}

/* Inflate all BGZF members of a buffer with zlib: the known-answer side of the inflate parity test
 * and the CPU inflate timing. returns total bytes or -1. */
int64_t oracle_inflate_buffer(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap) {
    uint64_t off = 0, o = 0;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return -1;
    while (off + 18 <= n) {
        if (in[off] != 0x1f || in[off + 1] != 0x8b) { inflateEnd(&zs); return -1; }
        uint32_t xlen = in[off + 10] | (in[off + 11] << 8);
        uint32_t csize = (in[off + 16] | (in[off + 17] << 8)) + 1u;
        if (off + csize > n) break;
        uint32_t isize = rd32(in + off + csize - 4);
        if (o + isize > cap) { inflateEnd(&zs); return -1; }
        inflateReset(&zs);
        zs.next_in = (Bytef*)(in + off + 12 + xlen);
        zs.avail_in = csize - 12 - xlen - 8;
        zs.next_out = out + o;
        zs.avail_out = isize;
        if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != isize) { inflateEnd(&zs); return -1; }
        o += isize;
        off += csize;
    }
    inflateEnd(&zs);
    return (int64_t)o;
}
