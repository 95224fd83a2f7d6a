// DEFLATE (RFC 1951) decoder core shared by the sm_100a inflate kernel and the host-side
// logic test (tests/host_emu.cpp compiles this header with g++ to exercise the bit-level
// logic without a GPU; the product only ever runs the __global__ kernels in kernels.cu).
//
// Replaces zlib's inflate as reached from htslib's bgzf_read_block under
// `bam.read(&mut read)` (reference src/count_reads/counters.rs:41,132; introns.rs:33).
//
// Layout of one decoder's tables (shared memory on the device, ~3.4 KB):
//   ll_lut[1<<LL_ROOT]  u16  (symbol<<4 | code length); 0xFFFF = invalid, 0x8000 = "longer than root"
//   d_lut [1<<D_ROOT]   u16
//   ll_sym[288], d_sym[32]   canonical (length, symbol) ordered symbol lists for long codes
//   HuffDec              first code / offset / count per code length
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MBF_HD __host__ __device__ __forceinline__
#define MBF_D __device__ __forceinline__
#else
#define MBF_HD inline
#define MBF_D inline
#endif

namespace mbf {

constexpr int LL_ROOT = 10;
constexpr int D_ROOT = 8;
constexpr int CL_ROOT = 7;
constexpr uint16_t LUT_SLOW = 0x8000;     // code longer than the LUT root: canonical search
constexpr uint16_t LUT_INVALID = 0xFFFF;  // no code has this prefix
// valid entries are (symbol << 4 | code length) <= 0x11FF, so `entry < 0x1000` <=> literal

enum InflateError : int {
    INF_OK = 0,
    INF_BAD_BTYPE = 1,
    INF_BAD_STORED = 2,
    INF_BAD_HEADER = 3,
    INF_BAD_CODES = 4,
    INF_BAD_SYMBOL = 5,
    INF_BAD_DISTANCE = 6,
    INF_OUTPUT_OVERRUN = 7,
    INF_INPUT_OVERRUN = 8,
    INF_SIZE_MISMATCH = 9,
};

struct HuffDec {
    uint16_t first[16];
    uint16_t offs[16];
    uint16_t cnt[16];
};

struct DecoderTables {
    uint16_t ll_lut[1 << LL_ROOT];
    uint16_t d_lut[1 << D_ROOT];
    uint16_t ll_sym[288];
    uint16_t d_sym[32];
    HuffDec ll_hd, d_hd;
    uint8_t lens[320];  // code lengths while a dynamic header is parsed
};

// RFC 1951 3.2.5 tables: a host copy (header parsing on the CPU, logic tests) and a
// __constant__ copy for the kernels.
#define MBF_TABLES(Q, P)                                                                          \
    Q uint16_t P##LEN_BASE[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27, \
                                  31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};  \
    Q uint8_t P##LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,                     \
                                  2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};                       \
    Q uint16_t P##DIST_BASE[30] = {1,    2,    3,    4,    5,    7,    9,    13,    17,    25,     \
                                   33,   49,   65,   97,   129,  193,  257,  385,   513,   769,    \
                                   1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577}; \
    Q uint8_t P##DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,             \
                                   6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};           \
    Q uint8_t P##CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
MBF_TABLES(static const, H_)
#ifdef __CUDACC__
MBF_TABLES(__device__ __constant__, D_)
#define MBF_TBL(name, i) D_##name[i]
#else
#define MBF_TBL(name, i) H_##name[i]
#endif
#if defined(__CUDACC__) && !defined(__CUDA_ARCH__)
#undef MBF_TBL
#define MBF_TBL(name, i) H_##name[i]
#endif

MBF_HD uint32_t bitrev32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}

// LSB-first bit reader over 32-bit aligned words (global memory on the device).  One word is
// always prefetched (`nextw`) so that the load latency is off the decode chain.
struct BitReader {
    const uint32_t* words;
    uint32_t wpos;    // next word to prefetch
    uint32_t wlimit;  // words [0, wlimit) may be loaded
    uint64_t buf;
    uint32_t cnt;
    uint32_t nextw;

    MBF_HD uint32_t load(uint32_t i) const {
#ifdef __CUDA_ARCH__
        return i < wlimit ? __ldg(words + i) : 0u;
#else
        return i < wlimit ? words[i] : 0u;
#endif
    }
    // byte pointer `p` (any alignment), `nbytes` of payload
    MBF_HD void init(const uint8_t* p, uint32_t nbytes) {
        uintptr_t a = (uintptr_t)p;
        uint32_t mis = (uint32_t)(a & 3);
        words = (const uint32_t*)(a - mis);
        wlimit = (mis + nbytes + 3) >> 2;
        buf = (uint64_t)load(0) >> (8 * mis);
        cnt = 32 - 8 * mis;
        nextw = load(1);
        wpos = 2;
    }
    // guarantees cnt >= 32 afterwards
    MBF_HD void refill() {
        if (cnt < 32) {
            buf |= (uint64_t)nextw << cnt;
            cnt += 32;
            nextw = load(wpos);
            wpos++;
        }
    }
    MBF_HD uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1); }
    MBF_HD void consume(int n) { buf >>= n; cnt -= n; }
    MBF_HD uint32_t take(int n) { uint32_t v = peek(n); consume(n); return v; }
    // position of the next unread bit, counted from bit 0 of words[0]
    MBF_HD uint32_t bit_position() const { return (wpos - 1) * 32 - cnt; }
    MBF_HD void seek(uint32_t bitpos) {
        const uint32_t i = bitpos >> 5, sh = bitpos & 31;
        buf = (uint64_t)load(i) >> sh;
        cnt = 32 - sh;
        nextw = load(i + 1);
        wpos = i + 2;
    }
    // bytes of payload consumed so far (rounded up to whole bytes), given the initial misalignment
    MBF_HD uint32_t bytes_consumed(uint32_t mis) const {
        uint64_t bits = (uint64_t)(wpos - 1) * 32 - cnt - 8ull * mis;
        return (uint32_t)((bits + 7) >> 3);
    }
};

// Build LUT + canonical lists for one alphabet.  `strict`: the code-length code (zlib type CODES)
// may not be incomplete; litlen/dist may be only if their longest code is 1 bit (zlib inftrees.c).
MBF_HD int build_table(const uint8_t* lens, int n, int root, uint16_t* lut, uint16_t* sym, HuffDec* hd,
                       bool strict) {
    for (int i = 0; i < 16; i++) hd->cnt[i] = 0;
    for (int i = 0; i < n; i++) hd->cnt[lens[i]]++;
    int max = 15;
    while (max > 0 && hd->cnt[max] == 0) max--;
    for (int i = 0; i < (1 << root); i++) lut[i] = LUT_INVALID;
    if (max == 0) return 0;  // no codes at all: every lookup is invalid (zlib allows this)
    int left = 1;
    for (int len = 1; len <= 15; len++) {
        left <<= 1;
        left -= hd->cnt[len];
        if (left < 0) return INF_BAD_CODES;  // over-subscribed
    }
    if (left > 0 && (strict || max != 1)) return INF_BAD_CODES;  // incomplete
    hd->offs[0] = 0;
    hd->offs[1] = 0;
    for (int len = 1; len < 15; len++) hd->offs[len + 1] = hd->offs[len] + hd->cnt[len];
    uint32_t code = 0;
    hd->first[0] = 0;
    for (int len = 1; len <= 15; len++) {
        hd->first[len] = (uint16_t)code;
        code = (code + hd->cnt[len]) << 1;
    }
    uint16_t next[16];
    for (int len = 0; len < 16; len++) next[len] = hd->offs[len];
    for (int s = 0; s < n; s++)
        if (lens[s]) sym[next[lens[s]]++] = (uint16_t)s;
    int total = hd->offs[15] + hd->cnt[15];
    for (int idx = 0; idx < total; idx++) {
        int s = sym[idx];
        int l = lens[s];
        uint32_t c = hd->first[l] + (idx - hd->offs[l]);
        uint32_t r = bitrev32(c) >> (32 - l);
        if (l <= root) {
            uint16_t e = (uint16_t)((s << 4) | l);
            for (uint32_t k = r; k < (1u << root); k += (1u << l)) lut[k] = e;
        } else {
            lut[r & ((1u << root) - 1)] = LUT_SLOW;
        }
    }
    return 0;
}

// Codes longer than the LUT root: canonical search. `bits15` = next 15 stream bits (LSB first).
// Returns (symbol << 4 | len) or 0 if no code matches.
MBF_HD uint32_t slow_decode(const HuffDec* hd, const uint16_t* sym, uint32_t bits15, int root) {
    uint32_t code = bitrev32(bits15) >> 17;  // 15-bit MSB-first
    for (int len = root + 1; len <= 15; len++) {
        int d = (int)(code >> (15 - len)) - (int)hd->first[len];
        if (d >= 0 && d < (int)hd->cnt[len]) return ((uint32_t)sym[hd->offs[len] + d] << 4) | (uint32_t)len;
    }
    return 0;
}

MBF_HD int build_fixed(DecoderTables* T) {
    for (int i = 0; i < 144; i++) T->lens[i] = 8;
    for (int i = 144; i < 256; i++) T->lens[i] = 9;
    for (int i = 256; i < 280; i++) T->lens[i] = 7;
    for (int i = 280; i < 288; i++) T->lens[i] = 8;
    for (int i = 0; i < 30; i++) T->lens[288 + i] = 5;
    int rc = build_table(T->lens, 288, LL_ROOT, T->ll_lut, T->ll_sym, &T->ll_hd, false);
    if (rc) return rc;
    // 30 five-bit codes leave the code incomplete by design; zlib hard-codes this table.
    T->lens[288 + 30] = 5;
    T->lens[288 + 31] = 5;
    return build_table(T->lens + 288, 32, D_ROOT, T->d_lut, T->d_sym, &T->d_hd, false);
}

// Parse a dynamic-Huffman block header (RFC 1951 3.2.7) and build both tables.  The pieces may live
// in different memories (LUTs in shared, the rest in global scratch in the multi-decoder kernel).
MBF_HD int build_dynamic_parts(BitReader& br, uint8_t* lens, uint16_t* ll_lut, uint16_t* ll_sym, HuffDec* ll_hd,
                               uint16_t* d_lut, uint16_t* d_sym, HuffDec* d_hd) {
    br.refill();
    uint32_t hlit = br.take(5) + 257;
    uint32_t hdist = br.take(5) + 1;
    uint32_t hclen = br.take(4) + 4;
    if (hlit > 286 || hdist > 30) return INF_BAD_HEADER;
    uint8_t cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (uint32_t i = 0; i < hclen; i++) {
        br.refill();
        cl[MBF_TBL(CL_ORDER, i)] = (uint8_t)br.take(3);
    }
    // the code-length code's LUT (128 entries) and lists live in d_lut / d_sym until the real
    // distance table is built
    int rc = build_table(cl, 19, CL_ROOT, d_lut, d_sym, d_hd, true);
    if (rc) return rc;
    uint32_t n = hlit + hdist, i = 0;
    while (i < n) {
        br.refill();
        uint32_t e = d_lut[br.peek(CL_ROOT)];
        if (e >= LUT_SLOW) return INF_BAD_HEADER;  // code-length codes are <= 7 bits = CL_ROOT
        uint32_t l = e & 15;
        br.consume(l);
        uint32_t s = e >> 4;
        if (s < 16) {
            lens[i++] = (uint8_t)s;
            continue;
        }
        uint32_t rep, val = 0;
        if (s == 16) {
            if (i == 0) return INF_BAD_HEADER;
            val = lens[i - 1];
            rep = 3 + br.take(2);
        } else if (s == 17) {
            rep = 3 + br.take(3);
        } else {
            rep = 11 + br.take(7);
        }
        if (i + rep > n) return INF_BAD_HEADER;
        for (uint32_t k = 0; k < rep; k++) lens[i++] = (uint8_t)val;
    }
    if (lens[256] == 0) return INF_BAD_HEADER;  // no end-of-block code
    rc = build_table(lens, (int)hlit, LL_ROOT, ll_lut, ll_sym, ll_hd, false);
    if (rc) return rc;
    return build_table(lens + hlit, (int)hdist, D_ROOT, d_lut, d_sym, d_hd, false);
}
MBF_HD int build_dynamic(DecoderTables* T, BitReader& br) {
    return build_dynamic_parts(br, T->lens, T->ll_lut, T->ll_sym, &T->ll_hd, T->d_lut, T->d_sym, &T->d_hd);
}

// Status of decode_symbols
enum { DEC_END_BLOCK = 0, DEC_NEED_FLUSH = 1 };  // errors are returned negated: -(InflateError)

// Decode literal/length symbols of the current Huffman block into `out` until the end-of-block
// symbol (DEC_END_BLOCK) or until out.want_flush() (DEC_NEED_FLUSH).
//
// Out must provide:  bool want_flush();  void put(uint8_t);  bool copy(uint32_t dist, uint32_t len)
// (false if dist reaches before the start of this member's output).
template <class Out>
MBF_HD int decode_symbols(const DecoderTables* T, BitReader& br, Out& out) {
    for (;;) {
        if (out.want_flush()) return DEC_NEED_FLUSH;
        br.refill();
        uint32_t e = T->ll_lut[br.peek(LL_ROOT)];
        if (e >= LUT_SLOW) {
            if (e != LUT_SLOW) return -INF_BAD_SYMBOL;
            e = slow_decode(&T->ll_hd, T->ll_sym, br.peek(15), LL_ROOT);
            if (!e) return -INF_BAD_SYMBOL;
        }
        br.consume(e & 15);
        uint32_t sym = e >> 4;
        if (sym < 256) {
            out.put((uint8_t)sym);
            continue;
        }
        if (sym == 256) return DEC_END_BLOCK;
        sym -= 257;
        if (sym >= 29) return -INF_BAD_SYMBOL;
        uint32_t len = MBF_TBL(LEN_BASE, sym) + br.take(MBF_TBL(LEN_EXTRA, sym));
        br.refill();
        e = T->d_lut[br.peek(D_ROOT)];
        if (e >= LUT_SLOW) {
            if (e != LUT_SLOW) return -INF_BAD_DISTANCE;
            e = slow_decode(&T->d_hd, T->d_sym, br.peek(15), D_ROOT);
            if (!e) return -INF_BAD_DISTANCE;
        }
        br.consume(e & 15);
        uint32_t ds = e >> 4;
        if (ds >= 30) return -INF_BAD_DISTANCE;
        uint32_t dist = MBF_TBL(DIST_BASE, ds) + br.take(MBF_TBL(DIST_EXTRA, ds));
        if (!out.copy(dist, len)) return -INF_BAD_DISTANCE;
    }
}

}  // namespace mbf
