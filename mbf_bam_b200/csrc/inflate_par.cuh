// K1 pass 1, production form: k_tokens_par -- one CTA per BGZF member, every lane decodes its own stretch
// of the SAME deflate block.
//
// The SIMT decoder it replaces (k_tokens_bf, inflate_twopass.cuh; still the fallback for irregular members)
// gave every decoding lane its own member and therefore its own 2.4 KB of tables: 88 decoders per SM, 6 useful
// lanes per instruction, 0.53 us per member.  DEFLATE token streams self-synchronise: a decoder started at an
// arbitrary bit falls into step with the true token chain after a few dozen tokens (measured on the bench
// files with tools/selfsync_sim.cpp: with sub-sequences of 1024 bits 97 % of the blocks need exactly one
// corrected pass).  So the lanes of a CTA share ONE table set and cut the block's bits into sub-sequences:
//
//   stage   the member's compressed words into shared memory (coalesced 16-byte loads), at most NT*SW_MAX words
//           per wave; a typical 20 KB member is one wave
//   header  thread 0 parses the block header out of shared memory; warps 0 and 1 build the literal/length and
//           the distance table (two-level: 9 / 8 root bits, second-level tables for longer codes -- a long code
//           costs one more shared load instead of the canonical search's ~100 instructions)
//   guess   lane i decodes from the first bit of sub-sequence i (lane 0 from the true start) until it has crossed
//           into sub-sequence i+1 and publishes where it stopped
//   sync    every lane whose predecessor stopped somewhere else restarts there; repeated until nobody moves.
//           Converged <=> every lane started where its predecessor stopped <=> the chain is the true one
//           (induction from lane 0), whatever the data: speculation is verified, never trusted
//   emit    token counts -> exclusive scan -> every lane decodes its sub-sequence once more and writes its tokens
//
// Stored blocks, members whose tables overflow the second-level pool, and anything that looks corrupt are
// appended to a fallback list and decoded by k_tokens_bf, which also produces the precise error codes.
//
// Replaces zlib's inflate (Huffman half) under htslib's bgzf_read_block (reference
// src/count_reads/counters.rs:41,132; src/count_reads/introns.rs:33).
#pragma once
#include "inflate_twopass.cuh"

namespace mbf {

constexpr int PT_ROOT_LL = 9;       // root bits of the literal/length table
constexpr int PT_ROOT_D = 8;        // ... of the distance table
constexpr int PT_SUB_CAP = 768;     // shared pool of second-level entries (zlib's bound for 9 root bits is 340)
constexpr uint32_t E2_LINK = 1u << 29;          // with K_SPECIAL: [8:5] second-level bits, [19:10] pool offset
constexpr uint32_t E3_BAD = K_SPECIAL | E2_LINK;  // "no such code": a link to pool slot 0, which holds E3_TERM
constexpr uint32_t E3_TERM = K_SPECIAL;           // terminal: kind 3, consumes nothing
constexpr uint32_t E2_LINKTMP = 0xF0000000u;    // during the build: | second-level bits so far (atomicMax; > E3_BAD)
constexpr uint32_t PF_EOB = 1u << 31, PF_INV = 1u << 30, PF_DEAD = 1u << 29;  // flags beside a bit position
constexpr uint32_t PF_ANY = PF_EOB | PF_INV | PF_DEAD;
constexpr int PT_SW_MIN = 5;        // shortest sub-sequence in words

template <int NT, int SW_MAX>
struct ParShared {
    uint32_t ll[1 << PT_ROOT_LL];
    uint32_t d[1 << PT_ROOT_D];
    uint32_t sub[PT_SUB_CAP];        // during header parsing: scratch of the code-length code
    uint32_t zero;                   // the "distance entry" of a literal: consumes nothing
    uint32_t cb[NT * SW_MAX + 16];   // staged compressed words
    uint32_t exitb[NT + 1];          // exitb[i+1]: where lane i stopped | flags; exitb[0]: the true start
    uint32_t wsum[NT / 32];
    uint32_t run[2][16], first[2][16];
    uint16_t rev[320];               // bit-reversed code of every symbol (table build)
    uint8_t lens[320];
    uint32_t sub_alloc, first_flag;
    int bad;
    uint32_t hdr_bfinal, hdr_hlit, hdr_hdist, hdr_data_bp;
    int hdr_err;                     // 0 ok, else: hand the member to the fallback kernel
    uint32_t member;
};

struct ParArgs {
    TokenArgs t;
    uint32_t* fb_list;   // members for the fallback kernel
    uint32_t* fb_count;
};

// ------------------------------------------------------------------------------------------------
// two-level table of one alphabet, built by one warp.  lens[0..n) in shared memory.  Returns 0, an
// InflateError, or -1 when the second-level pool is exhausted.
template <bool IS_DIST>
__device__ __forceinline__ int build_two_level(const uint8_t* lens, int n, int root, uint32_t* lut, uint32_t* sub,
                                               uint32_t* sub_alloc, uint32_t* run, uint32_t* first, uint16_t* rev, int lane) {
    const uint32_t rmask = (1u << root) - 1u;
    if (lane < 16) run[lane] = 0;
    for (int i = lane; i < (1 << root); i += 32) lut[i] = E3_BAD;
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const uint32_t l = lens[s];
        if (l) atomicAdd(&run[l], 1u);
    }
    __syncwarp();
    if (lane == 0) {
        int left = 1, max = 0, bad = 0;
        uint32_t code = 0;
        for (int len = 1; len <= 15; len++) {
            const uint32_t c = run[len];
            if (c) max = len;
            left = (left << 1) - (int)c;
            if (left < 0) bad = 1;
            first[len] = code;
            code = (code + c) << 1;
        }
        if (max == 0) bad = 0;                   // no codes at all: every lookup stays invalid (zlib allows it)
        else if (left > 0 && max != 1) bad = 1;  // incomplete (zlib allows a single 1-bit code)
        first[0] = (uint32_t)bad;
    }
    __syncwarp();
    if (first[0]) return INF_BAD_CODES;
    if (lane < 16) run[lane] = 0;  // now: symbols of each length seen so far (canonical rank)
    __syncwarp();
    for (int base = 0; base < n; base += 32) {
        const int s = base + lane;
        const uint32_t l = s < n ? lens[s] : 0u;
        const unsigned peers = __match_any_sync(0xffffffffu, l);
        uint32_t rank = 0;
        if (l) rank = run[l] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (l && lane == __ffs(peers) - 1) run[l] += __popc(peers);
        __syncwarp();
        if (l) {
            const uint32_t code = first[l] + rank;
            const uint32_t r = __brev(code) >> (32 - l);
            rev[s] = (uint16_t)r;
            if (l <= (uint32_t)root) {
                uint32_t e = make_entry<IS_DIST, 2>((uint32_t)s, l);
                if (e >= K_SPECIAL) e = E3_BAD;  // a symbol that may not occur (286, 287; distance 30, 31)
                for (uint32_t k = r; k <= rmask; k += (1u << l)) lut[k] = e;
            } else {
                atomicMax(&lut[r & rmask], E2_LINKTMP | (l - (uint32_t)root));  // longest code below this prefix
            }
        }
    }
    __syncwarp();
    bool over = false;
    for (int i = lane; i < (1 << root); i += 32) {
        const uint32_t e = lut[i];
        if ((e & 0xF0000000u) == E2_LINKTMP) {
            const uint32_t sl = e & 15u;
            const uint32_t off = atomicAdd(sub_alloc, 1u << sl);
            if (off + (1u << sl) > (uint32_t)PT_SUB_CAP) over = true;
            else lut[i] = K_SPECIAL | E2_LINK | (sl << 5) | (off << 10);
        }
    }
    if (__any_sync(0xffffffffu, over)) return -1;
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const uint32_t l = lens[s];
        if (l > (uint32_t)root) {
            const uint32_t r = rev[s];
            const uint32_t e = lut[r & rmask];
            const uint32_t sl = (e >> 5) & 15u, off = (e >> 10) & 0x3ffu;
            uint32_t ent = make_entry<IS_DIST, 2>((uint32_t)s, l);
            if (ent >= K_SPECIAL) ent = E3_TERM;
            for (uint32_t k = r >> root; k < (1u << sl); k += (1u << (l - (uint32_t)root))) sub[off + k] = ent;
        }
    }
    __syncwarp();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// bit reader over the staged words (thread 0, block headers only)
struct SmemBits {
    uint32_t s_cb;    // shared address of the staged words
    uint32_t rel;     // bit position relative to the first staged word
    uint32_t limit;   // bits that may be consumed
    uint32_t lo, hi, cur;  // the two words at word index `cur`
    __device__ __forceinline__ void init(uint32_t s_cb_, uint32_t rel_, uint32_t limit_) {
        s_cb = s_cb_; rel = rel_; limit = limit_; cur = 0xffffffffu; lo = hi = 0;
    }
    __device__ __forceinline__ uint32_t peek() {
        const uint32_t wi = rel >> 5;
        if (wi != cur) {
            const uint32_t a = s_cb + (wi << 2);
            lo = lds_u32(a);
            hi = lds_u32(a + 4);
            cur = wi;
        }
        return __funnelshift_r(lo, hi, rel & 31u);
    }
    __device__ __forceinline__ uint32_t take(uint32_t n) {
        const uint32_t v = peek() & ((1u << n) - 1u);
        rel += n;
        return v;
    }
    __device__ __forceinline__ bool overrun() const { return rel > limit; }
};

// One deflate block header at bit position `bp` (member-relative; staged from word cw0).  Thread 0.
// hdr_err != 0: stored block, bad header, or a header that leaves the staged words -> fallback kernel.
template <int NT, int SW_MAX>
__device__ void par_parse_header(ParShared<NT, SW_MAX>& S, uint32_t s_cb, uint32_t bp, uint32_t cw0, uint32_t staged_words,
                                 uint32_t member_bits) {
    SmemBits br;
    br.init(s_cb, bp - cw0 * 32u, min((staged_words - 2u) * 32u, member_bits - cw0 * 32u));
    S.hdr_err = 0;
    S.sub_alloc = 1;  // pool slot 0 is the terminal entry every "no such code" links to
    S.bad = 0;
    S.zero = 0;
    S.hdr_bfinal = br.take(1);
    const uint32_t btype = br.take(2);
    if (btype == 1) {
        for (int i = 0; i < 144; i++) S.lens[i] = 8;
        for (int i = 144; i < 256; i++) S.lens[i] = 9;
        for (int i = 256; i < 280; i++) S.lens[i] = 7;
        for (int i = 280; i < 288; i++) S.lens[i] = 8;
        for (int i = 0; i < 32; i++) S.lens[288 + i] = 5;
        S.hdr_hlit = 288;
        S.hdr_hdist = 32;
    } else if (btype == 2) {
        const uint32_t hlit = br.take(5) + 257u, hdist = br.take(5) + 1u, hclen = br.take(4) + 4u;
        if (hlit > 286u || hdist > 30u) { S.hdr_err = 1; return; }
        uint8_t cl[19];
#pragma unroll
        for (int i = 0; i < 19; i++) cl[i] = 0;
        for (uint32_t i = 0; i < hclen; i++) cl[MBF_TBL(CL_ORDER, i)] = (uint8_t)br.take(3);
        // the code-length code: LUT (128 x u16), symbol list and HuffDec live in the second-level pool for now
        uint16_t* cl_lut = reinterpret_cast<uint16_t*>(S.sub);
        uint16_t* cl_sym = cl_lut + (1 << CL_ROOT);
        HuffDec* cl_hd = reinterpret_cast<HuffDec*>(cl_sym + 32);
        if (build_table(cl, 19, CL_ROOT, cl_lut, cl_sym, cl_hd, true)) { S.hdr_err = 1; return; }
        const uint32_t nn = hlit + hdist;
        const uint32_t s_cl = smem_addr(cl_lut);
        uint32_t i = 0, prev = 0;
        while (i < nn) {
            const uint32_t w = br.peek();
            const uint32_t e = lds_u16(s_cl + ((w & ((1u << CL_ROOT) - 1u)) << 1));
            if (e >= LUT_SLOW) { S.hdr_err = 1; return; }
            const uint32_t cl_bits = e & 15u, s2 = e >> 4;
            if (s2 < 16u) {
                br.rel += cl_bits;
                S.lens[i++] = (uint8_t)s2;
                prev = s2;
            } else {
                uint32_t rep, val = 0;
                if (s2 == 16u) {
                    if (i == 0) { S.hdr_err = 1; return; }
                    val = prev;
                    rep = 3u + ((w >> cl_bits) & 3u);
                    br.rel += cl_bits + 2u;
                } else if (s2 == 17u) {
                    rep = 3u + ((w >> cl_bits) & 7u);
                    br.rel += cl_bits + 3u;
                } else {
                    rep = 11u + ((w >> cl_bits) & 127u);
                    br.rel += cl_bits + 7u;
                }
                if (i + rep > nn) { S.hdr_err = 1; return; }
                for (uint32_t k = 0; k < rep; k++) S.lens[i++] = (uint8_t)val;
                prev = val;
            }
            if (br.overrun()) { S.hdr_err = 1; return; }
        }
        if (S.lens[256] == 0) { S.hdr_err = 1; return; }
        S.hdr_hlit = hlit;
        S.hdr_hdist = hdist;
    } else {
        S.hdr_err = 1;  // stored block (or BTYPE 3): the fallback kernel handles / reports it
        return;
    }
    if (br.overrun()) { S.hdr_err = 1; return; }
    S.hdr_data_bp = cw0 * 32u + br.rel;
}

// ------------------------------------------------------------------------------------------------
// Decode tokens from bit `start_bp` until the position reaches `end_bp` (a token that starts before end_bp is
// decoded whole), the end-of-block code, or an invalid code.  Returns where it stopped | PF_EOB | PF_INV.
// All positions are member-relative; cw0_bits = first staged word * 32.
// second-level lookup without divergence: e = special ? shared[addr] : e
__device__ __forceinline__ uint32_t lds_if_special(uint32_t e, uint32_t addr) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %0, 0xC0000000;\n\t@p ld.shared.u32 %0, [%1];\n\t}" : "+r"(e) : "r"(addr));
    return e;
}

template <bool EMIT>
__device__ __forceinline__ uint32_t par_run(uint32_t s_cb, uint32_t cw0_bits, uint32_t s_ll, uint32_t s_d, uint32_t s_sub,
                                            uint32_t s_zero, uint32_t start_bp, uint32_t end_bp, uint32_t& n_out, uint32_t* tk) {
    uint32_t pos = start_bp - cw0_bits;  // bits from the first staged word
    uint32_t lim = end_bp - cw0_bits;    // set to 0 to leave the loop
    uint32_t a = s_cb + ((pos >> 5) << 2);
    uint32_t lo = lds_u32(a), hi = lds_u32(a + 4), nx = lds_u32(a + 8);
    a += 12;
    uint32_t n = 0, flags = 0;
    // One trip = one token, one straight-line instruction sequence whatever the token is: literal/length
    // lookup, second-level lookup (predicated load), distance lookup (a literal reads a zero word), second-level
    // lookup, predicated token store, window advance.  End of block and invalid codes are table entries of their
    // own (kind bit 31 set): they end the loop through `lim`, not through a branch.
    while (pos < lim) {
        const uint32_t bo = pos & 31u;
        const uint32_t w = __funnelshift_r(lo, hi, bo);
        uint32_t e = lds_u32(s_ll + ((w & ((1u << PT_ROOT_LL) - 1u)) << 2));
        e = lds_if_special(e, s_sub + ((((e >> 10) & 0x3ffu) + ((w >> PT_ROOT_LL) & ~(0xffffffffu << ((e >> 5) & 15u)))) << 2));
        const uint32_t c1 = e & 31u;
        const bool m = (e >> 30) == 1u;
        const uint32_t bo1 = bo + c1;  // <= 31 + 20
        const bool up = bo1 >= 32u;
        const uint32_t w2 = __funnelshift_r(up ? hi : lo, up ? nx : hi, bo1);  // shift amount is taken mod 32
        uint32_t d = lds_u32(m ? s_d + ((w2 & ((1u << PT_ROOT_D) - 1u)) << 2) : s_zero);
        d = lds_if_special(d, s_sub + ((((d >> 10) & 0x3ffu) + ((w2 >> PT_ROOT_D) & ~(0xffffffffu << ((d >> 5) & 15u)))) << 2));
        const uint32_t c2 = d & 31u;
        const bool stop = (int32_t)(e | d) < 0;  // end of block (kind 2), no such code (kind 3: consumes nothing)
        if (EMIT) {
            const uint32_t val = ((e >> 10) & 0xffffu) + ((w & ~(0xffffffffu << c1)) >> ((e >> 5) & 15u));
            const uint32_t dist = ((d >> 10) & 0xffffu) + ((w2 & ~(0xffffffffu << c2)) >> ((d >> 5) & 15u));
            if (!stop) __stcg(tk, m ? (0x80000000u | (val << 16) | ((dist - 1u) & 0x7fffu)) : val);
            tk += stop ? 0 : 1;
        }
        n += stop ? 0u : 1u;
        flags = stop ? (((e | d) & (1u << 30)) ? PF_INV : PF_EOB) : flags;
        lim = stop ? 0u : lim;
        const uint32_t adv = bo1 + c2;  // <= 79
        pos += c1 + c2;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ge.u32 p, %4, 32;\n\t"
            "@p mov.b32 %0, %1;\n\t"
            "@p mov.b32 %1, %2;\n\t"
            "@p ld.shared.u32 %2, [%3];\n\t"
            "@p add.u32 %3, %3, 4;\n\t}"
            : "+r"(lo), "+r"(hi), "+r"(nx), "+r"(a)
            : "r"(adv));
        if (adv >= 64u) { lo = hi; hi = nx; nx = lds_u32(a); a += 4; }  // a token of more than 32 bits: rare
    }
    n_out = n;
    return (cw0_bits + pos) | flags;
}

// ------------------------------------------------------------------------------------------------
template <int NT, int SW_MAX>
__global__ void __launch_bounds__(NT) k_tokens_par(ParArgs pa) {
    extern __shared__ __align__(16) uint8_t par_smem[];
    ParShared<NT, SW_MAX>& S = *reinterpret_cast<ParShared<NT, SW_MAX>*>(par_smem);
    constexpr uint32_t CB_WORDS = NT * SW_MAX + 16;
    const TokenArgs& a = pa.t;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t s_cb = smem_addr(S.cb), s_ll = smem_addr(S.ll), s_d = smem_addr(S.d), s_sub = smem_addr(S.sub),
                   s_zero = smem_addr(&S.zero);
    const uint32_t n_blocks = *a.n_blocks;
    unsigned long long my_tokens = 0;

    for (;;) {
        __syncthreads();  // everybody is done with the previous member (S.member, S.exitb, ...)
        if (t == 0) S.member = atomicAdd(a.queue_head, 1u);
        __syncthreads();
        const uint32_t b = S.member;
        if (b >= n_blocks) break;
        const uint8_t* p = a.cbuf + a.cpos[b];
        const uint32_t clen = a.clen[b], isize = a.isize[b];
        const uint32_t mis = (uint32_t)((uintptr_t)p & 3u);
        const uint32_t* words = reinterpret_cast<const uint32_t*>(p - mis);
        const uint32_t nwords = (mis + clen + 3u) >> 2;       // words[nwords] is still inside the BGZF member (trailer)
        const uint32_t member_bits = (mis + clen) * 8u;
        uint32_t* const tk = a.tok + (a.uoff[b] - a.u_base);
        uint32_t ntok = 0;
        uint32_t bp = mis * 8u;  // true position of the decoder (uniform over the CTA)
        bool fallback = false, done = false;

        while (!done && !fallback) {
            // ---------------------------------------------------------------- stage, header, tables
            __syncthreads();  // the previous block's emit pass has read the staged words
            uint32_t cw0 = bp >> 5;
            uint32_t staged = min(CB_WORDS, nwords + 2u - min(cw0, nwords + 2u));
            for (uint32_t i = t; i < CB_WORDS; i += NT) S.cb[i] = i < staged ? __ldg(words + cw0 + i) : 0u;
            __syncthreads();
            if (t == 0) par_parse_header<NT, SW_MAX>(S, s_cb, bp, cw0, max(staged, 3u), member_bits);
            __syncthreads();
            if (S.hdr_err) { fallback = true; break; }
            const uint32_t bfinal = S.hdr_bfinal, hlit = S.hdr_hlit, hdist = S.hdr_hdist;
            if (t == 64 % NT) S.sub[0] = E3_TERM;  // (the header's scratch lived in the pool)
            if (wid == 0) {
                const int rc = build_two_level<false>(S.lens, (int)hlit, PT_ROOT_LL, S.ll, S.sub, &S.sub_alloc, S.run[0], S.first[0], S.rev, lane);
                if (rc && lane == 0) atomicOr(&S.bad, 1);
            } else if (wid == 1) {
                const int rc = build_two_level<true>(S.lens + hlit, (int)hdist, PT_ROOT_D, S.d, S.sub, &S.sub_alloc, S.run[1], S.first[1],
                                                     S.rev + 288, lane);
                if (rc && lane == 0) atomicOr(&S.bad, 1);
            }
            __syncthreads();
            if (S.bad) { fallback = true; break; }
            bp = S.hdr_data_bp;

            // ---------------------------------------------------------------- waves of NT sub-sequences
            bool eob = false;
            bool first_wave = true;
            while (!eob) {
                if (!first_wave) {  // the first wave uses what was staged with the header
                    __syncthreads();
                    cw0 = bp >> 5;
                    staged = min(CB_WORDS, nwords + 2u - min(cw0, nwords + 2u));
                    for (uint32_t i = t; i < CB_WORDS; i += NT) S.cb[i] = i < staged ? __ldg(words + cw0 + i) : 0u;
                    __syncthreads();
                }
                first_wave = false;
                const uint32_t w0 = bp >> 5;                      // first word of the wave
                if (w0 > nwords) { fallback = true; break; }       // ran off the member without an end-of-block code
                const uint32_t in_buf = CB_WORDS - 8u - (w0 - cw0);
                const uint32_t avail = min(in_buf, nwords + 1u - w0);
                uint32_t sw = (avail + NT - 1) / NT;
                sw |= 1u;                                          // odd: the lanes' words fall into different banks
                sw = min((uint32_t)SW_MAX, max((uint32_t)PT_SW_MIN, sw));
                const uint32_t n_active = min((uint32_t)NT, (avail + sw - 1u) / sw);
                const uint32_t wave_base = w0 * 32u, cw0_bits = cw0 * 32u;
                const bool active = (uint32_t)t < n_active;
                uint32_t my_start = t == 0 ? bp : wave_base + (uint32_t)t * sw * 32u;
                // (the last lane stops at the end of what is staged, not at the end of its nominal stretch)
                const uint32_t my_end = wave_base + min((uint32_t)(t + 1) * sw, avail) * 32u;
                uint32_t n = 0, res = PF_DEAD;
                if (active) res = par_run<false>(s_cb, cw0_bits, s_ll, s_d, s_sub, s_zero, my_start, my_end, n, nullptr);
                if (t == 0) { S.exitb[0] = bp; S.first_flag = NT; }
                S.exitb[t + 1] = res;
                __syncthreads();
                // ---- synchronise: restart wherever the predecessor really stopped
                for (;;) {
                    const uint32_t prev = S.exitb[t];
                    const bool chg = active && !(prev & PF_ANY) && prev != my_start;
                    if (!__syncthreads_or(chg)) break;
                    if (chg) {
                        my_start = prev;
                        res = par_run<false>(s_cb, cw0_bits, s_ll, s_d, s_sub, s_zero, my_start, my_end, n, nullptr);
                        S.exitb[t + 1] = res;
                    }
                    __syncthreads();
                }
                // ---- the chain is true up to the first flagged lane
                if (active && (res & (PF_EOB | PF_INV))) atomicMin(&S.first_flag, (uint32_t)t);
                __syncthreads();
                const uint32_t ff = S.first_flag;                  // NT: no end of block in this wave
                const uint32_t last = ff < (uint32_t)NT ? ff : n_active - 1u;
                const uint32_t last_res = S.exitb[last + 1];
                if (last_res & PF_INV) { fallback = true; break; }
                // ---- token offsets: exclusive scan of the counts of lanes 0..last
                const uint32_t mine = (uint32_t)t <= last ? n : 0u;
                uint32_t incl = mine;
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, dlt);
                    if (lane >= dlt) incl += y;
                }
                if (lane == 31) S.wsum[wid] = incl;
                __syncthreads();
                uint32_t before = 0, total = 0;
#pragma unroll
                for (int k = 0; k < NT / 32; k++) {
                    const uint32_t v = S.wsum[k];
                    if (k < wid) before += v;
                    total += v;
                }
                if (ntok + total > isize) { fallback = true; break; }  // more tokens than output bytes: corrupt
                if ((uint32_t)t <= last && n) {
                    uint32_t n2;
                    par_run<true>(s_cb, cw0_bits, s_ll, s_d, s_sub, s_zero, my_start, my_end, n2, tk + ntok + before + incl - mine);
                }
                ntok += total;
                bp = last_res & ~PF_ANY;
                eob = (last_res & PF_EOB) != 0;
                if (bp > member_bits || (!eob && bp >= member_bits)) { fallback = true; break; }  // ran off the member
            }
            if (fallback) break;
            if (bfinal) done = true;
        }
        if (t == 0) {
            if (fallback) {
                a.ntok[b] = 0;
                pa.fb_list[atomicAdd(pa.fb_count, 1u)] = b;
            } else {
                a.ntok[b] = ntok;
                my_tokens += ntok;
            }
        }
    }
    if (t == 0 && a.n_tokens && my_tokens) atomicAdd(a.n_tokens, my_tokens);
}

}  // namespace mbf
