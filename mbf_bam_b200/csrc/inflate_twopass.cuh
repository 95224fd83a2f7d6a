// K1, two-pass form (production).
//
// Profiles of the one-warp-per-block kernel (k_inflate_batched) showed where its time goes: the serial
// Huffman decode runs on one lane per warp, so 85 % of the issued instructions serve a single lane.
// Here the two halves of DEFLATE get the shape each one wants:
//
//   pass 1  k_tokens_bf   SIMT entropy decode.  DL lanes of a warp are decoders, each with its own BGZF
//           block (pulled from a work queue), its own bit window in registers and its own tables
//           (2.4 KB) in shared memory.  The decoders advance in lock step, one symbol per trip, and a trip
//           is ONE straight-line instruction sequence whatever the symbols are; it emits a 32-bit token
//           (literal byte | length + distance) to HBM.  Everything rare -- end of block, headers, stored
//           blocks, table builds (by all 32 lanes), codes longer than the LUT root -- happens outside it.
//   pass 2  k_lz_resolve  one warp per block turns tokens into bytes: batches of up to 32 matches / 512
//           bytes (tokens read 32 at a time, warp scan for the output offsets, literals stored at once),
//           independent matches copied one per lane (16 bytes per step; sources within 1.5 KB from the
//           shared-memory ring, older ones from L2/HBM), dependent matches in order with 32 lanes each.
//
// Tokens of block b live at tok[uoff[b] - U_PREFIX ...] (a token yields >= 1 byte, so ISIZE bounds the
// count); ntok[b] is written by pass 1.
//
// Replaces zlib's inflate under htslib's bgzf_read_block (reference src/count_reads/counters.rs:41).
#pragma once
#include "inflate_batched.cuh"

namespace mbf {

struct TokScratch {  // global memory, one per decoder lane slot: code lengths while a header is parsed
    uint8_t lens[320];
    uint32_t hlit, hdist;
};

struct TokenArgs {
    const uint8_t* cbuf;
    const uint32_t* cpos;
    const uint32_t* clen;
    const uint32_t* isize;
    const uint32_t* uoff;
    uint32_t* tok;           // token storage, indexed by (uoff[b] - u_base + k)
    uint32_t* ntok;          // per block
    uint32_t u_base;         // U_PREFIX
    const uint32_t* n_blocks;
    uint32_t* queue_head;
    uint32_t* err;
    uint32_t* err_info;
    TokScratch* scratch;     // gridDim.x * 32
    unsigned long long* n_tokens;  // statistics (may be null)
    const uint32_t* list;    // null: members 0..n_blocks-1; else the members to decode are list[0..*n_blocks)
};

// first error wins; the failed member gets an empty token list, so pass 2 never walks stale tokens for it
__device__ __forceinline__ void tok_err(const TokenArgs& a, uint32_t code, uint32_t info) {
    if (atomicCAS(a.err, 0u, code) == 0u) *a.err_info = info;
    a.ntok[info] = 0u;
}

// ------------------------------------------------------------------------------------------------
// Pass 1.  A first version with an ordinary branchy symbol loop was no faster than the one-pass kernel:
// with D decoders in lock step every trip issued the union of all paths (literal, match, long code, one
// step of some lane's header).  Here a trip is ONE straight-line sequence for all decoding lanes -- literal/length lookup, predicated window advance, distance lookup (always done,
// applied only to matches), predicated advance, predicated token store -- and everything rare (end of
// block, headers, stored blocks, table builds, long codes) happens outside it.
//
// The bit window is three 32-bit registers (lo, hi, nx) and a bit offset < 32 at the top of a trip:
// peek = one funnel shift, consume = one add, advance = four predicated moves and one prefetching load.
struct BitWin {
    const uint32_t* words;
    uint32_t wpos, wlimit, lo, hi, nx, nx2, bo;
    // words[wlimit] is still inside the BGZF block (the 8-byte gzip trailer follows the payload)
    __device__ __forceinline__ uint32_t load(uint32_t i) const { return __ldg(words + min(i, wlimit)); }
    __device__ __forceinline__ uint32_t init(const uint8_t* p, uint32_t nbytes) {
        const uint32_t mis = (uint32_t)((uintptr_t)p & 3);
        words = reinterpret_cast<const uint32_t*>(p - mis);
        wlimit = (mis + nbytes + 3) >> 2;
        lo = load(0); hi = load(1); nx = load(2); nx2 = load(3);
        wpos = 4;
        bo = mis * 8;
        return mis;
    }
    // One step: bo -= 32.  The word loaded here (nx2) is first READ by the next step (nx = nx2); a trip
    // makes one step at its very end, so a load has a whole trip to arrive (in-order issue: any read
    // of the register, even predicated off, waits for the load).
    __device__ __forceinline__ void refill() {
        // in PTX so that the predicated load writes nx2's own register (the compiler's version loads into
        // a temporary and copies it at the end of the same trip, which waits for the load after all)
        const uint32_t* src = words + min(wpos, wlimit);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ge.u32 p, %4, 32;\n\t"
            "@p mov.b32 %0, %1;\n\t"
            "@p mov.b32 %1, %2;\n\t"
            "@p mov.b32 %2, %3;\n\t"
            "@p ld.global.nc.L2::256B.u32 %3, [%6];\n\t"
            "@p add.u32 %5, %5, 1;\n\t"
            "@p add.u32 %4, %4, -32;\n\t}"
            : "+r"(lo), "+r"(hi), "+r"(nx), "+r"(nx2), "+r"(bo), "+r"(wpos)
            : "l"(src));
    }
    // 32 bits at bit offset o < 64 (o may point into hi:nx)
    __device__ __forceinline__ uint32_t window_at(uint32_t o) const {
        const bool up = o >= 32u;
        return __funnelshift_r(up ? hi : lo, up ? nx : hi, o);  // shift amount is taken mod 32
    }
    __device__ __forceinline__ uint32_t window() const { return __funnelshift_r(lo, hi, bo); }  // bo < 32
    __device__ __forceinline__ uint32_t peek(int n) const { return window() & ((1u << n) - 1); }
    __device__ __forceinline__ void consume(uint32_t n) { bo += n; }
    __device__ __forceinline__ uint32_t take(int n) { refill(); const uint32_t v = peek(n); bo += n; return v; }
    __device__ __forceinline__ void align_byte() { bo = (bo + 7u) & ~7u; }
    __device__ __forceinline__ uint32_t bytes_consumed(uint32_t mis) const { return ((wpos - 4) * 32 + bo - mis * 8 + 7) >> 3; }
    __device__ __forceinline__ bool overrun() const { return wpos > wlimit + 5; }
};

// codes longer than the LUT root: canonical search on shared-memory addresses, format-2 entry
template <bool IS_DIST>
__device__ __forceinline__ uint32_t slow_entry_s(uint32_t s_hd, uint32_t s_sym, uint32_t bits15, int root) {
    const uint32_t code = __brev(bits15) >> 17;
    for (int len = root + 1; len <= 15; len++) {
        const int d = (int)(code >> (15 - len)) - (int)lds_u16(s_hd + (uint32_t)offsetof(HuffDec, first) + 2u * len);
        if (d >= 0 && d < (int)lds_u16(s_hd + (uint32_t)offsetof(HuffDec, cnt) + 2u * len)) {
            const uint32_t o = lds_u16(s_hd + (uint32_t)offsetof(HuffDec, offs) + 2u * len);
            const uint32_t sy = lds_u16(s_sym + 2u * (o + (uint32_t)d));
            return make_entry<IS_DIST, 2>(sy, (uint32_t)len);
        }
    }
    return E2_INVALID;
}

constexpr int LL_ROOT3 = 8;  // shared memory per decoder decides how many decoders an SM holds
constexpr int D_ROOT3 = 7;
struct LaneTables32 {
    uint32_t ll_lut[1 << LL_ROOT3];  // format-2 entries (inflate_batched.cuh)
    uint32_t d_lut[1 << D_ROOT3];    // its first 256 bytes double as the code-length code's 16-bit LUT
    uint32_t zero;                   // the "distance entry" of a literal: consumes nothing
    uint16_t ll_sym[288];
    uint16_t d_sym[32];
    HuffDec ll_hd, d_hd;
};
template <int DL>
struct TokensShared32 {
    LaneTables32 lane[DL];
    uint32_t run[16];
    uint32_t base_slot[32];
};

enum { B_HUFF = 0, B_DONE = 1, B_IDLE = 2, B_HEADER = 3, B_BUILD = 4, B_FIN = 5 };

template <int DL, bool PIPE>
__global__ void __launch_bounds__(32, 16) k_tokens_bf(TokenArgs a) {
    extern __shared__ __align__(16) uint8_t tk_smem[];
    TokensShared32<DL>& S = *reinterpret_cast<TokensShared32<DL>*>(tk_smem);
    const int lane = threadIdx.x;
    LaneTables32& T = S.lane[lane < DL ? lane : 0];
    // one opaque shared-memory base register; the trip addresses everything relative to it (generic
    // pointers into T inside the loop make the compiler rebuild the address from %tid every trip)
    uint32_t s_T = smem_addr(&T);
    {   // round trip through a volatile shared load: ptxas cannot rematerialise the result
        const uint32_t slot = smem_addr(&S.base_slot[lane]);
        asm volatile("st.volatile.shared.u32 [%1], %0;\n\tld.volatile.shared.u32 %0, [%1];" : "+r"(s_T) : "r"(slot) : "memory");
    }
    const uint32_t s_ll = s_T + (uint32_t)offsetof(LaneTables32, ll_lut);
    const uint32_t s_d = s_T + (uint32_t)offsetof(LaneTables32, d_lut);
    const uint32_t s_llsym = s_T + (uint32_t)offsetof(LaneTables32, ll_sym);
    const uint32_t s_dsym = s_T + (uint32_t)offsetof(LaneTables32, d_sym);
    const uint32_t s_llhd = s_T + (uint32_t)offsetof(LaneTables32, ll_hd);
    const uint32_t s_dhd = s_T + (uint32_t)offsetof(LaneTables32, d_hd);
    const uint32_t s_zero = s_T + (uint32_t)offsetof(LaneTables32, zero);
    T.zero = 0u;
    __syncwarp();
    TokScratch* const sc = a.scratch + (size_t)blockIdx.x * DL + (lane < DL ? lane : 0);
    const uint32_t n_blocks = *a.n_blocks;

    BitWin br;
    br.words = nullptr; br.wpos = br.wlimit = 0; br.lo = br.hi = br.nx = br.nx2 = 0; br.bo = 0;
    int st = lane < DL ? B_IDLE : B_DONE;
    uint32_t b = 0, mis = 0, clen_b = 0, isize_b = 0, bfinal = 0;
    uint32_t out_bytes = 0, ntok = 0;
    uint32_t* tk = nullptr;

    for (;;) {
        int err = 0;
        // The decoding lanes stay in this inner loop until one of them needs anything else: no path out
        // of a trip leads to code that reads the read-ahead register, so nothing waits for its load.
        const unsigned hm = __ballot_sync(0xffffffffu, st == B_HUFF);
        if (st == B_HUFF) {
            // values, token, store of one symbol: nothing here feeds the bit position
            auto emit_token = [&](uint32_t e, uint32_t w, uint32_t d, uint32_t w2, bool emit) {
                const bool is_match = (e >> 30) == 1u;
                const uint32_t val = ((e >> 10) & 0xffffu) + ((w & ~(0xffffffffu << (e & 31u))) >> ((e >> 5) & 15u));
                const uint32_t dist = ((d >> 10) & 0xffffu) + ((w2 & ~(0xffffffffu << (d & 31u))) >> ((d >> 5) & 15u));
                const uint32_t token = is_match ? (0x80000000u | (val << 16) | ((dist - 1u) & 0x7fffu)) : val;
                const uint32_t nout = out_bytes + (is_match ? val : 1u);
                const bool ok = emit && nout <= isize_b;  // all predicated: no branch region in the common path
                if (ok) __stcg(tk + ntok, token);         // L2 only: the small L1 beside 220 KB of tables is for the input
                ntok += ok ? 1u : 0u;
                out_bytes = ok ? nout : out_bytes;
                err = (emit && !ok) ? (int)INF_OUTPUT_OVERRUN : err;
            };
            // Software pipelined by one trip: the token of trip i-1 is computed and stored in the same basic
            // block as the two table lookups of trip i, so the (in-order) warp has independent instructions
            // to issue while the lookups and the bit-position chain wait for their operands.
            uint32_t pe = 0, pw = 0, pd = 0, pw2 = 0;
            bool pemit = false, leave;
            do {
                // ------------------------------------------------------------ one symbol, straight line
                const uint32_t w = br.window();
                uint32_t e = lds_u32(s_ll + ((w & ((1u << LL_ROOT3) - 1)) << 2));
                uint32_t bo1 = br.bo + (e & 31u);  // the special entries consume nothing; bo1 <= 51
                uint32_t w2 = br.window_at(bo1);
                // distance lookup in every trip; a literal reads its zero word instead (consumes nothing)
                const bool m0 = (e >> 30) == 1u;
                uint32_t d = lds_u32((m0 ? s_d : s_zero) + ((w2 << 2) & (m0 ? ((1u << D_ROOT3) - 1) << 2 : 0u)));
                if (PIPE) emit_token(pe, pw, pd, pw2, pemit);
                bool emit = true;
                if ((e | d) >> 31) {
                    // ---- rare: end of block, code longer than the LUT root, invalid code
                    if (e >= K_SPECIAL) {
                        e = e == E2_SLOW ? slow_entry_s<false>(s_llhd, s_llsym, w & 0x7fffu, LL_ROOT3) : E2_INVALID;
                        if (e < K_SPECIAL) {  // redo what the trip skipped
                            bo1 = br.bo + (e & 31u);
                            w2 = br.window_at(bo1);
                            d = (e >> 30) == 1u ? lds_u32(s_d + ((w2 & ((1u << D_ROOT3) - 1)) << 2)) : 0u;
                        }
                    }
                    const uint32_t k2 = e >> 30;
                    if (k2 == 3u) {
                        err = err ? err : (int)INF_BAD_SYMBOL;
                        emit = false;
                    } else if (k2 == 2u) {
                        st = bfinal ? B_FIN : B_HEADER;
                        emit = false;
                    } else if (k2 == 1u && d >= K_SPECIAL) {
                        d = d == E2_SLOW ? slow_entry_s<true>(s_dhd, s_dsym, w2 & 0x7fffu, D_ROOT3) : E2_INVALID;
                        if (d >= K_SPECIAL) {
                            err = err ? err : (int)INF_BAD_DISTANCE;
                            emit = false;
                        }
                    }
                }
                br.bo = bo1 + (d & 31u);  // <= 79
                br.refill();              // a second step, if ever needed, is made in the rare section below
                if (PIPE) {
                    pe = e; pw = w; pd = d; pw2 = w2; pemit = emit;
                } else {
                    emit_token(e, w, d, w2, emit);
                }
                leave = st != B_HUFF || err != 0 || br.bo >= 32u;
            } while (!__any_sync(hm, leave));
            if (PIPE) emit_token(pe, pw, pd, pw2, pemit);  // the last trip's token
        }
        __syncwarp();
        if (err) {
            tok_err(a, WERR_INFLATE + (uint32_t)err, b);
            st = B_IDLE;  // drop this member; the window is reported as failed
            err = 0;
        }
        // ---------------------------------------------------------------- rare work, lane by lane
        while (br.bo >= 32u) br.refill();  // a trip that consumed more than 32 bits
        if (st == B_FIN) {
            a.ntok[b] = ntok;
            if (out_bytes != isize_b) tok_err(a, WERR_INFLATE + (uint32_t)INF_SIZE_MISMATCH, b);
            else if (br.bytes_consumed(mis) > clen_b) tok_err(a, WERR_INFLATE + (uint32_t)INF_INPUT_OVERRUN, b);
            if (a.n_tokens) atomicAdd(a.n_tokens, (unsigned long long)ntok);
            st = B_IDLE;
        }
        if (st == B_IDLE) {
            b = atomicAdd(a.queue_head, 1u);
            if (b >= n_blocks) {
                st = B_DONE;
            } else {
                if (a.list) b = a.list[b];
                clen_b = a.clen[b];
                isize_b = a.isize[b];
                mis = br.init(a.cbuf + a.cpos[b], clen_b);
                tk = a.tok + (a.uoff[b] - a.u_base);
                ntok = 0;
                out_bytes = 0;
                bfinal = 0;
                st = B_HEADER;
            }
        }
        while (st == B_HEADER) {
            // whole block header in one go (the other lanes of the warp wait: ~3 headers per 64 KB)
            if (br.overrun()) { err = INF_INPUT_OVERRUN; break; }
            br.refill();
            bfinal = br.take(1);
            const uint32_t btype = br.take(2);
            if (btype == 0) {
                br.align_byte();
                br.refill();
                const uint32_t len = br.take(16);
                br.refill();
                const uint32_t nlen = br.take(16);
                br.refill();
                if ((len ^ 0xffffu) != nlen) { err = INF_BAD_STORED; break; }
                if (out_bytes + len > isize_b) { err = INF_OUTPUT_OVERRUN; break; }
                for (uint32_t i = 0; i < len; i++) {
                    tk[ntok++] = br.take(8);
                    br.refill();
                }
                out_bytes += len;
                if (bfinal) st = B_FIN;
            } else if (btype == 1) {
                for (int i = 0; i < 144; i++) sc->lens[i] = 8;
                for (int i = 144; i < 256; i++) sc->lens[i] = 9;
                for (int i = 256; i < 280; i++) sc->lens[i] = 7;
                for (int i = 280; i < 288; i++) sc->lens[i] = 8;
                for (int i = 0; i < 32; i++) sc->lens[288 + i] = 5;
                sc->hlit = 288;
                sc->hdist = 32;
                br.refill();
                st = B_BUILD;
            } else if (btype == 2) {
                const uint32_t hlit = br.take(5) + 257, hdist = br.take(5) + 1, hclen = br.take(4) + 4;
                br.refill();
                uint8_t cl[19];
                for (int i = 0; i < 19; i++) cl[i] = 0;
                for (uint32_t i = 0; i < hclen; i++) {
                    cl[MBF_TBL(CL_ORDER, i)] = (uint8_t)br.take(3);
                    br.refill();
                }
                uint16_t* cl_lut = reinterpret_cast<uint16_t*>(T.d_lut);
                int rc = (hlit > 286 || hdist > 30) ? (int)INF_BAD_HEADER
                                                    : build_table(cl, 19, CL_ROOT, cl_lut, T.d_sym, &T.d_hd, true);
                const uint32_t nn = hlit + hdist;
                uint32_t i = 0, prev = 0;
                while (!rc && i < nn) {
                    br.refill();
                    const uint32_t e = cl_lut[br.peek(CL_ROOT)];
                    if (e >= LUT_SLOW) { rc = INF_BAD_HEADER; break; }
                    br.consume(e & 15);
                    const uint32_t s2 = e >> 4;
                    if (s2 < 16) {
                        sc->lens[i++] = (uint8_t)s2;
                        prev = s2;
                    } else {
                        uint32_t rep, val = 0;
                        if (s2 == 16) {
                            if (i == 0) { rc = INF_BAD_HEADER; break; }
                            val = prev;
                            rep = 3 + br.take(2);
                        } else if (s2 == 17) {
                            rep = 3 + br.take(3);
                        } else {
                            rep = 11 + br.take(7);
                        }
                        if (i + rep > nn) { rc = INF_BAD_HEADER; break; }
                        for (uint32_t k = 0; k < rep; k++) sc->lens[i++] = (uint8_t)val;
                        prev = val;
                    }
                    br.refill();
                    if (br.overrun()) rc = INF_INPUT_OVERRUN;
                }
                if (rc) { err = rc; break; }
                sc->hlit = hlit;
                sc->hdist = hdist;
                st = B_BUILD;
            } else {
                err = INF_BAD_BTYPE;
                break;
            }
        }
        if (err) {
            tok_err(a, WERR_INFLATE + (uint32_t)err, b);
            st = B_IDLE;
        }
        // ---------------------------------------------------------------- table builds, 32 lanes each
        __syncwarp();
        unsigned need = __ballot_sync(0xffffffffu, st == B_BUILD);
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            TokScratch* t = a.scratch + (size_t)blockIdx.x * DL + src;
            LaneTables32& L = S.lane[src];
            const uint32_t hlit = t->hlit, hdist = t->hdist;
            int rc = t->lens[256] == 0 ? (int)INF_BAD_HEADER
                                       : build_lut_warp<false, 2>(t->lens, (int)hlit, LL_ROOT3, L.ll_lut, L.ll_sym, &L.ll_hd, S.run, lane);
            if (!rc) rc = build_lut_warp<true, 2>(t->lens + hlit, (int)hdist, D_ROOT3, L.d_lut, L.d_sym, &L.d_hd, S.run, lane);
            if (lane == src) {
                if (rc) {
                    tok_err(a, WERR_INFLATE + (uint32_t)rc, b);
                    st = B_IDLE;
                } else {
                    st = B_HUFF;
                }
            }
        }
        if (__all_sync(0xffffffffu, st == B_DONE)) break;
    }
}

// ------------------------------------------------------------------------------------------------
struct ResolveArgs {
    const uint32_t* tok;
    const uint32_t* ntok;
    const uint32_t* isize;
    const uint32_t* uoff;
    uint32_t u_base;
    uint8_t* U;
    const uint32_t* n_blocks;
    uint32_t* err;
    uint32_t* err_info;
};

template <uint32_t MASK>
__device__ __forceinline__ void ring_flush_w(const uint8_t* ring, uint8_t* U, uint32_t from, uint32_t to, int lane) {
    if (from >= to) return;
    uint32_t head_end = min(to, (from + 15u) & ~15u);
    for (uint32_t i = from + lane; i < head_end; i += 32) U[i] = ring[i & MASK];
    uint32_t body_end = to & ~15u;
    if (body_end > head_end) {
        for (uint32_t i = head_end + 16u * lane; i < body_end; i += 512u)
            *(uint4*)(U + i) = *(const uint4*)(ring + (i & MASK));
    } else {
        body_end = head_end;
    }
    for (uint32_t i = body_end + lane; i < to; i += 32) U[i] = ring[i & MASK];
}

// RESOLVE_WARPS warps per CTA, each with its own BGZF block: an SM holds at most 32 CTAs, and this kernel
// needs more than 32 warps per SM to cover the latency of its far-match loads.
constexpr int RESOLVE_WARPS = 4;
template <int RING>
__global__ void __launch_bounds__(32 * RESOLVE_WARPS) k_lz_resolve(ResolveArgs a) {
    constexpr uint32_t RINGB_MASK = RING - 1;            // shadows the constants of the one-pass kernel
    constexpr uint32_t BATCH_NEAR = RING - BATCH_MAX_OUT;
    constexpr uint32_t RINGB = RING;
    __shared__ __align__(16) uint8_t rings[RESOLVE_WARPS][RING];
    __shared__ uint2 queues[RESOLVE_WARPS][32];  // (len << 16 | dist - 1, output offset in the batch) per match
    uint8_t* const ring = rings[threadIdx.x >> 5];
    const uint32_t b = blockIdx.x * RESOLVE_WARPS + (threadIdx.x >> 5);
    if (b >= *a.n_blocks) return;
    const int lane = threadIdx.x & 31;
    const uint32_t ostart = a.uoff[b], oend = ostart + a.isize[b];
    const uint32_t* tk = a.tok + (ostart - a.u_base);
    const uint32_t nt = a.ntok[b];
    uint8_t* const U = a.U;
    const uint32_t s_ring = smem_addr(ring);
    const uint32_t s_q = smem_addr(queues[threadIdx.x >> 5]);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t opos = ostart, flushed = ostart;
    bool bad = false;
    for (uint32_t t0 = 0; t0 < nt;) {
        // ---- older output to HBM first (far matches read it back through L2)
        {
            const uint32_t upto = min(opos & ~15u, oend);
            ring_flush_w<RINGB_MASK>(ring, U, flushed, upto, lane);
            if (upto > flushed) flushed = upto;
        }
        // ---- build a batch: tokens in rounds of 32 until 32 matches or BATCH_MAX_OUT bytes are collected.
        // Literals go straight into the ring, matches into the queue (one per lane for the copy passes:
        // a batch of 32 TOKENS would leave 60 % of the lanes without a match).
        uint32_t nm = 0, acc = 0;
        for (bool more = true; more;) {
            const uint32_t ti = t0 + lane;
            const bool valid = ti < nt;
            const uint32_t e = valid ? __ldg(tk + ti) : 0u;
            if (ti + 32 < nt) asm volatile("prefetch.global.L1 [%0];" ::"l"(tk + ti + 32));
            const bool is_match = (e >> 31) != 0;
            const uint32_t mylen = valid ? (is_match ? ((e >> 16) & 0x1ffu) : 1u) : 0u;
            uint32_t incl = mylen;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += y;
            }
            const unsigned mb = __ballot_sync(0xffffffffu, valid && is_match);
            const uint32_t midx = nm + (uint32_t)__popc(mb & lt_mask);
            const bool fit = valid && acc + incl <= BATCH_MAX_OUT && (!is_match || midx < 32u);
            const unsigned fb = __ballot_sync(0xffffffffu, fit);
            const uint32_t k = fb == 0xffffffffu ? 32u : (uint32_t)__ffs(~fb) - 1u;  // leading tokens that fit
            const uint32_t off = acc + incl - mylen;
            if ((uint32_t)lane < k) {
                if (is_match) {
                    sts_u32(s_q + 8u * midx, (mylen << 16) | (e & 0x7fffu));
                    sts_u32(s_q + 8u * midx + 4u, off);
                } else {
                    sts_u8(s_ring + ((opos + off) & RINGB_MASK), e & 0xffu);
                }
            }
            if (k) {
                acc += __shfl_sync(0xffffffffu, incl, (int)k - 1);
                nm += (uint32_t)__popc(k == 32u ? mb : mb & ((1u << k) - 1u));
                t0 += k;
            }
            more = k == 32u && t0 < nt && nm < 32u && acc < BATCH_MAX_OUT;
        }
        if (acc == 0 || opos + acc > oend) { bad = true; break; }
        __syncwarp();
        // ---- pass 1: one whole match per lane when it reads nothing of this batch
        uint32_t len = 0, dist = 1, off = 0;
        if ((uint32_t)lane < nm) {
            const uint32_t m = lds_u32(s_q + 8u * lane);
            off = lds_u32(s_q + 8u * lane + 4u);
            len = m >> 16;
            dist = (m & 0x7fffu) + 1u;
        }
        // long matches would keep one lane busy for many steps while 31 wait: they take the 32-lane route
        const bool dep = len != 0 && (dist < off + len || len > 32u);
        if (len != 0 && dist > opos + off - ostart) bad = true;
        if (__any_sync(0xffffffffu, bad)) { bad = true; break; }
        const uint32_t dst = opos + off, src = dst - dist;
        const uint32_t len1 = dep ? 0u : len;
        const bool near = dist <= BATCH_NEAR;
        for (uint32_t base = 0; __any_sync(0xffffffffu, base < len1); base += 16) {
            if (base < len1) {
                const uint32_t rem = min(16u, len1 - base);
                uint32_t w[4];
                if (near) {
                    const uint32_t s0 = src + base;
                    const uint32_t sh = (s0 & 3u) * 8u;
                    uint32_t x[5];
#pragma unroll
                    for (int q = 0; q < 5; q++) x[q] = lds_u32(s_ring + (((s0 & ~3u) + 4u * q) & RINGB_MASK));
#pragma unroll
                    for (int q = 0; q < 4; q++) w[q] = __funnelshift_r(x[q], x[q + 1], sh);
                } else {
                    const uint8_t* sp = U + src + base;
#pragma unroll
                    for (int q = 0; q < 4; q++) w[q] = (uint32_t)(4 * q) < rem ? ldu32_cg(sp + 4 * q) : 0u;
                }
                const uint32_t d0 = (dst + base) & RINGB_MASK;
                if (d0 + 16 <= (uint32_t)RINGB) {
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        if ((uint32_t)q < rem) sts_u8(s_ring + d0 + q, (w[q >> 2] >> (8 * (q & 3))) & 0xffu);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        if ((uint32_t)q < rem) sts_u8(s_ring + ((d0 + q) & RINGB_MASK), (w[q >> 2] >> (8 * (q & 3))) & 0xffu);
                }
            }
        }
        __syncwarp();
        // ---- pass 2: matches that read this batch's own output (or overlap themselves) and long ones, in
        // order, 32 lanes each
        unsigned mm = __ballot_sync(0xffffffffu, dep);
        while (mm) {
            const int j = __ffs(mm) - 1;
            mm &= mm - 1;
            const uint32_t lj = __shfl_sync(0xffffffffu, len, j);
            const uint32_t dj = __shfl_sync(0xffffffffu, dist, j);
            const uint32_t dstj = opos + __shfl_sync(0xffffffffu, off, j);
            const uint32_t sj = dstj - dj;
            if (dj > BATCH_NEAR) {  // source already flushed and possibly gone from the ring (dj > lj here)
                for (uint32_t i = lane; i < lj; i += 32) {
                    uint32_t v;
                    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(U + sj + i));
                    sts_u8(s_ring + ((dstj + i) & RINGB_MASK), v);
                }
            } else if (dj >= lj) {
                for (uint32_t i = lane; i < lj; i += 32)
                    sts_u8(s_ring + ((dstj + i) & RINGB_MASK), lds_u8(s_ring + ((sj + i) & RINGB_MASK)));
            } else {
                for (uint32_t i = lane; i < lj; i += 32)
                    sts_u8(s_ring + ((dstj + i) & RINGB_MASK), lds_u8(s_ring + ((sj + i % dj) & RINGB_MASK)));
            }
            __syncwarp();
        }
        opos += acc;
    }
    __syncwarp();
    if (!bad) {
        ring_flush_w<RINGB_MASK>(ring, U, flushed, min(opos, oend), lane);
        if (opos != oend) bad = true;
    }
    if (bad && lane == 0) {
        if (atomicCAS(a.err, 0u, (uint32_t)(WERR_INFLATE + INF_SIZE_MISMATCH)) == 0u) *a.err_info = b;
    }
}

}  // namespace mbf
