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
//           into sub-sequence i+1, counts its tokens and publishes where it stopped.  On the way it leaves a
//           CHECKPOINT behind every 256-bit boundary of its stretch: the first token start after the boundary and the
//           number of tokens before it
//   sync    every lane whose predecessor stopped somewhere else restarts there -- but only until it steps exactly onto
//           a checkpoint: from there on its path IS the guess path (a decoder's state is its bit position), so the
//           guess run's count and exit are reused.  Measured on the bench files (tools/merge_sim.cpp; sub-sequences
//           of 1312 bits = 122 tokens): the true chain meets the guess path after 10 tokens on average, 97 % within
//           12 words, and 0.09 % of the lanes never do (they decode to the end of their stretch).
//           Repeated until nobody moves.  Converged <=> every lane started where its predecessor stopped <=> the
//           chain is the true one (induction from lane 0), whatever the data: speculation is verified, never trusted
//   emit    token counts -> exclusive scan -> every lane decodes its stretch once more and writes its tokens
//           (three-decode scheme, the default).  Single-decode scheme (ONEPASS): the guess and sync runs keep their
//           tokens in a transposed scratch area in HBM (token k of lane j at [32 k + j]: a warp's lanes write whole
//           lines) and emit is a copy, [the sync run's tokens][the guess run's tokens from the checkpoint on], through
//           shared-memory tiles.  It issues 6 % fewer instructions and is no faster (measured): every trip
//           then carries the token arithmetic and a store
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
constexpr uint32_t PF_MERGED = 1u << 28;  // (sync run only, never published) stopped on a marked position
constexpr uint32_t PF_ANY = PF_EOB | PF_INV | PF_DEAD;
constexpr int PT_SW_MIN = 5;        // shortest sub-sequence in words
// Three-decode scheme: the guess run remembers, for every 256 bits of its stretch, the first token start behind that
// boundary and how many tokens came before it (one word per checkpoint); a sync run that steps exactly onto a
// checkpoint has met the guess path and stops there.
__host__ __device__ constexpr int par_ncp(int sw_max) { return (sw_max * 32 + 255) >> 8; }
// tokens one run may keep: 2 bits per token over the longest sub-sequence (a match costs at least 2); a run that
// would need more is reported as invalid and the member goes to the fallback kernel
__host__ __device__ constexpr uint32_t par_scratch_cap(int sw_max) { return (uint32_t)sw_max * 16u; }

template <int NT, int SW_MAX, bool ONEPASS = false>
struct ParShared {
    uint32_t ll[1 << PT_ROOT_LL];
    uint32_t d[1 << PT_ROOT_D];
    uint32_t sub[PT_SUB_CAP];        // during header parsing: scratch of the code-length code
    uint32_t zero;                   // the "distance entry" of a literal: consumes nothing
    uint32_t cb[NT * SW_MAX + 16];   // staged compressed words
    uint32_t mk[NT * par_ncp(SW_MAX)];  // par_ncp(SW_MAX) checkpoints per lane
    uint32_t exitb[NT + 1];          // exitb[i+1]: where lane i stopped | flags; exitb[0]: the true start
    uint32_t wsum[NT / 32];
    uint32_t run[2][16], first[2][16];
    uint16_t rev[320];               // bit-reversed code of every symbol (table build)
    alignas(4) uint8_t lens[320];
    uint32_t sub_alloc, first_flag;
    int bad;
    uint32_t hdr_bfinal, hdr_hlit, hdr_hdist, hdr_data_bp;
    int hdr_err;                     // 0 ok, else: hand the member to the fallback kernel
    uint32_t member;
};

// What k_par_headers found at the start of a member.
struct ParHdr {
    uint32_t data_bp;   // first bit of the block's data, relative to the member's first byte
    uint16_t hlit;
    uint8_t hdist;
    uint8_t flags;      // bit 0: BFINAL; bit 1: the code lengths are in hdr_lens (else: parse in k_tokens_par)
};
constexpr uint32_t PAR_HDR_STRIDE = 320;

struct ParArgs {
    TokenArgs t;
    uint32_t* fb_list;   // members for the fallback kernel
    uint32_t* fb_count;
    uint32_t* scratch;   // gridDim.x * NT * 2 * par_scratch_cap(SW_MAX) token slots: per lane [guess run][sync run]
    const ParHdr* hdr;          // per member: the first block header, parsed by k_par_headers
    const uint8_t* hdr_lens;    // per member PAR_HDR_STRIDE bytes: literal/length code lengths, then the distance ones
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
template <class PS>
__device__ void par_parse_header(PS& S, uint32_t s_cb, uint32_t bp, uint32_t cw0, uint32_t staged_words,
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

// MODE PR_COUNT    decode and count
//      PR_EMIT     decode and write the tokens to tk[0], tk[1], ... (three-decode scheme: emit run)
//      PR_COUNT_CP guess run: count, and leave a checkpoint (position | tokens before it << 12, in the lane's words at
//                  s_mk) at the first token start behind every 256-bit boundary of the stretch
//      PR_SYNC_CP  sync run: count until the position equals a checkpoint (result | PF_MERGED, n_out = tokens before it,
//                  aux = the guess run's tokens before it) or the stretch ends
//      PR_GUESS    single-decode scheme: PR_COUNT_CP that also keeps token k in tk[32 k] (a warp's lanes all write row k
//                  in trip k: one 128-byte line)
//      PR_SYNC     single-decode scheme: PR_SYNC_CP that also keeps its tokens
// A keeping run that fills `cap` token slots ends as invalid.
enum { PR_COUNT = 0, PR_EMIT = 1, PR_GUESS = 2, PR_SYNC = 3, PR_COUNT_CP = 4, PR_SYNC_CP = 5 };
template <int MODE>
__device__ __forceinline__ uint32_t par_run(uint32_t s_cb, uint32_t s_mk, uint32_t cw0_bits, uint32_t s_ll, uint32_t s_d,
                                            uint32_t s_sub, uint32_t s_zero, uint32_t start_bp, uint32_t end_bp, uint32_t cap,
                                            uint32_t& n_out, uint32_t* tk, uint32_t base_rel = 0, uint32_t* aux = nullptr) {
    constexpr bool KEEP = MODE == PR_GUESS || MODE == PR_SYNC;
    constexpr bool CP = MODE != PR_COUNT && MODE != PR_EMIT;
    constexpr uint32_t TK_STEP = KEEP ? 32u : 1u;
    uint32_t pos = start_bp - cw0_bits;  // bits from the first staged word
    uint32_t lim = end_bp - cw0_bits;    // set to 0 to leave the loop
    uint32_t a = s_cb + ((pos >> 5) << 2);
    uint32_t lo = lds_u32(a), hi = lds_u32(a + 4), nx = lds_u32(a + 8);
    a += 12;
    uint32_t n = 0, flags = 0;
    uint32_t cp_j = 0, cp_v = 0;                 // PR_COUNT_CP / PR_SYNC_CP: 256-bit region of the stretch, its checkpoint
    // One trip = one token, one straight-line instruction sequence whatever the token is: literal/length
    // lookup, second-level lookup (predicated load), distance lookup (a literal reads a zero word), second-level
    // lookup, predicated token store, window advance.  End of block and invalid codes are table entries of their
    // own (kind bit 31 set): they end the loop through `lim`, not through a branch; so does a marked position.
    while (pos < lim) {
        const uint32_t bo = pos & 31u;
        bool hit = false;
        if (CP) {
            const uint32_t rel = pos - base_rel, j = rel >> 8;  // (tokens are shorter than 256 bits: no region is skipped)
            if (j != cp_j) {
                cp_j = j;
                if (MODE == PR_COUNT_CP || MODE == PR_GUESS) sts_u32(s_mk + (j << 2) - 4u, rel | (n << 12));
                else cp_v = lds_u32(s_mk + (j << 2) - 4u);
            }
            if (MODE == PR_SYNC_CP || MODE == PR_SYNC) hit = (cp_v & 0xfffu) == rel;
        }
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
        const bool end = (int32_t)(e | d) < 0;  // end of block (kind 2), no such code (kind 3: consumes nothing)
        const bool stop = end || hit;
        if (MODE == PR_EMIT || KEEP) {
            const uint32_t val = ((e >> 10) & 0xffffu) + ((w & ~(0xffffffffu << c1)) >> ((e >> 5) & 15u));
            const uint32_t dist = ((d >> 10) & 0xffffu) + ((w2 & ~(0xffffffffu << c2)) >> ((d >> 5) & 15u));
            if (!stop) __stcg(tk, m ? (0x80000000u | (val << 16) | ((dist - 1u) & 0x7fffu)) : val);
            tk += stop ? 0u : TK_STEP;
        }
        n += stop ? 0u : 1u;
        const bool full = KEEP && n >= cap;
        flags = hit ? PF_MERGED : (end ? (((e | d) & (1u << 30)) ? PF_INV : PF_EOB) : (full ? PF_INV : flags));
        lim = (stop || full) ? 0u : lim;
        const uint32_t adv = bo1 + c2;  // <= 79
        pos += hit ? 0u : c1 + c2;
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
    if (MODE == PR_SYNC_CP || MODE == PR_SYNC) *aux = cp_v >> 12;
    return (cw0_bits + pos) | flags;
}

// Single-decode scheme, emit: the warp moves its 32 lanes' tokens from the transposed scratch (token k of lane j at
// [32 k + j]) to the lanes' places in the member's token list, 32 x 32 tokens at a time through a shared tile, so
// that both the loads and the stores are whole lines.  Lane j contributes rows [from, to) of `src`, written to
// dst[0 .. to - from).  `tile`: 32 x 33 words of shared memory owned by the warp.
__device__ __forceinline__ void par_copy_warp(uint32_t* tile, const uint32_t* src, uint32_t from, uint32_t to, uint32_t* dst,
                                              int lane) {
    const uint32_t lo_row = __reduce_min_sync(0xffffffffu, from < to ? from : 0xffffffffu);
    const uint32_t hi_row = __reduce_max_sync(0xffffffffu, from < to ? to : 0u);
    for (uint32_t r0 = lo_row; r0 < hi_row; r0 += 32) {
        __syncwarp();
        const uint32_t rows = min(32u, hi_row - r0);
        for (uint32_t r = 0; r < rows; r++) tile[r * 33 + lane] = __ldcg(src + (size_t)(r0 + r) * 32 + lane);
        __syncwarp();
#pragma unroll 4
        for (int j = 0; j < 32; j++) {
            const uint32_t fj = __shfl_sync(0xffffffffu, from, j), tj = __shfl_sync(0xffffffffu, to, j);
            const unsigned long long dj = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)dst, j);
            const uint32_t row = r0 + lane;  // lane k handles row r0 + k of column j
            if (row >= fj && row < tj) __stcg(reinterpret_cast<uint32_t*>((uintptr_t)dj) + (row - fj), tile[lane * 33 + j]);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Header pre-pass: ONE THREAD per member reads the first block header (BTYPE 1 or 2) and leaves the ~300 code
// lengths in HBM.  The code-length stream is a serial chain of ~100 Huffman symbols; inside k_tokens_par it kept
// one thread busy and 127 waiting for a quarter of the kernel's time.  Here 32 members advance per warp
// instruction and nobody waits.  Anything irregular (stored block, over-subscribed or incomplete code-length
// code, lengths running past their count, a header longer than the member) is left to k_tokens_par's own parser,
// which decides between decoding, the fallback kernel and an error -- this kernel never reports one.
constexpr int PAR_HDR_THREADS = 64;
__global__ void __launch_bounds__(PAR_HDR_THREADS) k_par_headers(TokenArgs a, ParHdr* hdr, uint8_t* hdr_lens) {
    __shared__ uint8_t cl_lut_s[PAR_HDR_THREADS][128];  // (symbol << 3 | length) of the code-length code; 0 = no code
    const uint32_t b = blockIdx.x * PAR_HDR_THREADS + threadIdx.x;
    if (b >= *a.n_blocks) return;
    uint8_t* const lut = cl_lut_s[threadIdx.x];
    const uint8_t* p = a.cbuf + a.cpos[b];
    const uint32_t clen = a.clen[b];
    const uint32_t mis = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(p - mis);
    const uint32_t limit = (mis + clen) * 8u;  // bits (word relative) that belong to the member
    ParHdr h;
    h.data_bp = 0; h.hlit = 0; h.hdist = 0; h.flags = 0;
    uint32_t rel = mis * 8u, cur = 0xffffffffu, lo = 0, hi = 0;
    auto peek = [&]() {
        const uint32_t wi = rel >> 5;
        if (wi != cur) { lo = __ldg(words + wi); hi = __ldg(words + wi + 1); cur = wi; }  // (the trailer follows the data)
        return __funnelshift_r(lo, hi, rel & 31u);
    };
    auto take = [&](uint32_t n) { const uint32_t v = peek() & ((1u << n) - 1u); rel += n; return v; };
    uint32_t* const out = reinterpret_cast<uint32_t*>(hdr_lens + (size_t)b * PAR_HDR_STRIDE);
    bool ok = false;
    const uint32_t bfinal = take(1), btype = take(2);
    if (btype == 1u) {
        for (int i = 0; i < 36; i++) out[i] = 0x08080808u;
        for (int i = 36; i < 64; i++) out[i] = 0x09090909u;
        for (int i = 64; i < 70; i++) out[i] = 0x07070707u;
        for (int i = 70; i < 72; i++) out[i] = 0x08080808u;
        for (int i = 72; i < 80; i++) out[i] = 0x05050505u;
        h.hlit = 288; h.hdist = 32;
        ok = true;
    } else if (btype == 2u) {
        const uint32_t hlit = take(5) + 257u, hdist = take(5) + 1u, hclen = take(4) + 4u;
        uint32_t cl_lo = 0, cl_hi = 0;  // 19 lengths of 3 bits: symbol s at bits 3s (lo: 0..9, hi: 10..18)
        for (uint32_t i = 0; i < hclen; i++) {
            const uint32_t sy = MBF_TBL(CL_ORDER, i), v = take(3);
            if (sy < 10u) cl_lo |= v << (3u * sy); else cl_hi |= v << (3u * (sy - 10u));
        }
        auto cl_of = [&](uint32_t sy) { return sy < 10u ? (cl_lo >> (3u * sy)) & 7u : (cl_hi >> (3u * (sy - 10u))) & 7u; };
        uint32_t kraft = 0;
        for (uint32_t sy = 0; sy < 19u; sy++) { const uint32_t l = cl_of(sy); if (l) kraft += 128u >> l; }
        ok = hlit <= 286u && hdist <= 30u && kraft == 128u;  // a complete code only
        if (ok) {
            uint32_t code = 0;
            for (uint32_t l = 1; l <= 7u; l++) {  // canonical codes in order of (length, symbol)
                for (uint32_t sy = 0; sy < 19u; sy++) {
                    if (cl_of(sy) != l) continue;
                    const uint32_t r = __brev(code) >> (32u - l);
                    for (uint32_t k = r; k < 128u; k += 1u << l) lut[k] = (uint8_t)((sy << 3) | l);
                    code++;
                }
                code <<= 1;
            }
            const uint32_t nn = hlit + hdist;
            uint32_t i = 0, prev = 0, acc = 0;
            auto put = [&](uint32_t v) {
                acc |= v << (8u * (i & 3u));
                if ((i & 3u) == 3u) { out[i >> 2] = acc; acc = 0; }
                i++;
            };
            uint32_t len256 = 0;
            while (i < nn && ok) {
                const uint32_t w = peek();
                const uint32_t e = lut[w & 127u];
                const uint32_t cb = e & 7u, s2 = e >> 3;
                if (s2 < 16u) {
                    rel += cb;
                    if (i == 256u) len256 = s2;
                    put(s2);
                    prev = s2;
                } else {
                    uint32_t rep, val = 0;
                    if (s2 == 16u) {
                        if (i == 0) { ok = false; break; }
                        val = prev;
                        rep = 3u + ((w >> cb) & 3u);
                        rel += cb + 2u;
                    } else if (s2 == 17u) {
                        rep = 3u + ((w >> cb) & 7u);
                        rel += cb + 3u;
                    } else {
                        rep = 11u + ((w >> cb) & 127u);
                        rel += cb + 7u;
                    }
                    if (i + rep > nn) { ok = false; break; }
                    if (i <= 256u && 256u < i + rep) len256 = val;
                    for (uint32_t k = 0; k < rep; k++) put(val);
                    prev = val;
                }
                if (rel > limit) ok = false;
            }
            if (ok && (i & 3u)) out[i >> 2] = acc;
            if (len256 == 0) ok = false;  // no end-of-block code
            h.hlit = (uint16_t)hlit; h.hdist = (uint8_t)hdist;
        }
    }
    if (rel > limit) ok = false;
    h.data_bp = rel - mis * 8u;
    h.flags = (uint8_t)((bfinal ? 1u : 0u) | (ok ? 2u : 0u));
    hdr[b] = h;
}

// ------------------------------------------------------------------------------------------------
template <int NT, int SW_MAX, bool ONEPASS>
__global__ void __launch_bounds__(NT, 896 / NT) k_tokens_par(ParArgs pa) {
    extern __shared__ __align__(16) uint8_t par_smem[];
    typedef ParShared<NT, SW_MAX, ONEPASS> PS;
    PS& S = *reinterpret_cast<PS*>(par_smem);
    static_assert(!ONEPASS || NT * SW_MAX + 16 >= (NT / 32) * 32 * 33, "the staged-word area doubles as the copy tiles");
    constexpr uint32_t CB_WORDS = NT * SW_MAX + 16;
    const TokenArgs& a = pa.t;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t s_cb = smem_addr(S.cb), s_mk = smem_addr(S.mk), s_ll = smem_addr(S.ll), s_d = smem_addr(S.d), s_sub = smem_addr(S.sub),
                   s_zero = smem_addr(&S.zero);
    constexpr uint32_t CAP = par_scratch_cap(SW_MAX);
    constexpr int NCP = par_ncp(SW_MAX);
    const uint32_t s_cp = s_mk + (uint32_t)t * NCP * 4u;  // (three decodes) this lane's checkpoints
    // transposed token scratch of the warp: [guess run | latest sync run][CAP rows][32 lanes]
    uint32_t* const g_tok = pa.scratch + ((size_t)blockIdx.x * (NT / 32) + wid) * 2 * CAP * 32;
    uint32_t* const s_tok = g_tok + (size_t)CAP * 32;
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
            const ParHdr ph = pa.hdr[b];
            if (bp == mis * 8u && (ph.flags & 2u)) {  // the member's first header was parsed by k_par_headers
                const uint32_t nn4 = ((uint32_t)ph.hlit + ph.hdist + 3u) >> 2;
                const uint32_t* src = reinterpret_cast<const uint32_t*>(pa.hdr_lens + (size_t)b * PAR_HDR_STRIDE);
                for (uint32_t i = t; i < nn4; i += NT) reinterpret_cast<uint32_t*>(S.lens)[i] = __ldg(src + i);
                if (t == 0) {
                    S.hdr_err = 0; S.sub_alloc = 1; S.bad = 0; S.zero = 0;
                    S.hdr_bfinal = ph.flags & 1u; S.hdr_hlit = ph.hlit; S.hdr_hdist = ph.hdist;
                    S.hdr_data_bp = mis * 8u + ph.data_bp;
                }
            } else if (t == 0) {
                par_parse_header(S, s_cb, bp, cw0, max(staged, 3u), member_bits);
            }
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
            uint32_t waves = 0;
            while (!eob) {
                if (++waves > 4096u) { fallback = true; break; }  // (every wave consumes at least one word: a guard, as below)
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
                uint32_t cur_start = my_start;      // where the path this lane currently stands for begins
                const uint32_t base_rel = ((w0 - cw0) + (uint32_t)t * sw) * 32u;  // the lane's stretch in staged-word bit positions
                uint32_t n_g = 0, n_s = 0, g_from = 0;  // (single decode) tokens of the guess run; of the sync run; first guess token in use
                uint32_t res_g = PF_DEAD, res = PF_DEAD;
                uint32_t n = 0;
                // ---- guess: every lane decodes its stretch from the first bit
                // ... counts (single decode: and keeps its tokens), and leaves checkpoints for the sync runs
#pragma unroll
                for (int k = 0; k < NCP; k++) S.mk[t * NCP + k] = 0u;
                if (active) {
                    if (ONEPASS) res = res_g = par_run<PR_GUESS>(s_cb, s_cp, cw0_bits, s_ll, s_d, s_sub, s_zero, my_start, my_end, CAP, n_g, g_tok + lane, base_rel);
                    else res = res_g = par_run<PR_COUNT_CP>(s_cb, s_cp, cw0_bits, s_ll, s_d, s_sub, s_zero, my_start, my_end, 0, n_g, nullptr, base_rel);
                }
                n = n_g;
                if (t == 0) { S.exitb[0] = bp; S.first_flag = NT; }
                S.exitb[t + 1] = res;
                __syncthreads();
                // ---- synchronise: restart wherever the predecessor really stopped (single decode: only until the
                // path meets the guess path).  Exits are functions of starts and lane 0 never moves, so this ends
                // after at most NT rounds; the counter is a guard against a hang should that reasoning ever fail.
                uint32_t rounds = 0;
                for (;;) {
                    const uint32_t prev = S.exitb[t];
                    const bool chg = active && !(prev & PF_ANY) && prev != cur_start;
                    if (!__syncthreads_or(chg)) break;
                    if (++rounds > (uint32_t)NT + 2u) { fallback = true; break; }
                    if (chg) {
                        cur_start = prev;
                        uint32_t g_before = 0;
                        const uint32_t r = ONEPASS ? par_run<PR_SYNC>(s_cb, s_cp, cw0_bits, s_ll, s_d, s_sub, s_zero, prev, my_end, CAP, n_s, s_tok + lane, base_rel, &g_before)
                                                   : par_run<PR_SYNC_CP>(s_cb, s_cp, cw0_bits, s_ll, s_d, s_sub, s_zero, prev, my_end, 0, n_s, nullptr, base_rel, &g_before);
                        if (r & PF_MERGED) {  // from the checkpoint on the path is the guess path
                            g_from = g_before;
                            res = res_g;
                        } else {
                            g_from = n_g;  // never met: the sync run's tokens are the lane's tokens
                            res = r;
                        }
                        n = n_s + (n_g - g_from);
                        S.exitb[t + 1] = res;
                    }
                    __syncthreads();
                }
                if (fallback) break;
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
                if (ONEPASS) {
                    // the staged words are spent: their area becomes the warps' copy tiles
                    uint32_t* const tile = reinterpret_cast<uint32_t*>(par_smem + offsetof(PS, cb)) + wid * (32 * 33);
                    uint32_t* const dst = tk + ntok + before + incl - mine;
                    const bool mine_counts = (uint32_t)t <= last;
                    par_copy_warp(tile, s_tok, 0u, mine_counts ? n_s : 0u, dst, lane);
                    par_copy_warp(tile, g_tok, g_from, mine_counts ? n_g : g_from, dst + n_s, lane);
                } else if ((uint32_t)t <= last && n) {
                    uint32_t n2;
                    par_run<PR_EMIT>(s_cb, s_mk, cw0_bits, s_ll, s_d, s_sub, s_zero, cur_start, my_end, 0, n2, tk + ntok + before + incl - mine);
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
