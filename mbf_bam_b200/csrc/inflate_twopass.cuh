// K1, two-pass form (production).
//
// Profiles of the one-warp-per-block kernels showed where the time goes: the serial Huffman decode on
// one lane per warp (85 % of the issued instructions serve a single lane), while the LZ77 copies, once
// batched, are cheap.  So the two halves of DEFLATE get the shape each one wants:
//
//   pass 1  k_inflate_tokens   SIMT entropy decode.  Every LANE is a decoder with its own BGZF block
//           (pulled from a work queue), its own bit reader in registers and its own two LUTs
//           (1.5 KB) in shared memory; 32 blocks advance in lock step, one symbol per trip, and emit
//           32-bit tokens (literal byte | length+distance) to HBM.  Nothing of variable length happens
//           inside a trip, so the lanes re-converge every trip.  Rare work is made uniform too:
//           dynamic-block headers are decoded one code length per trip, and Huffman tables are built by
//           all 32 lanes for one decoder at a time.
//   pass 2  k_lz_resolve       one warp per block turns tokens into bytes: 32 tokens per batch, warp
//           scan for the output offsets, literals stored at once, independent matches copied one per
//           lane (16 bytes per step; sources within 1.5 KB from the shared-memory ring, older ones
//           from L2/HBM), dependent matches in order with 32 lanes each.
//
// Tokens of block b live at tok[uoff[b] - U_PREFIX ...] (a token yields >= 1 byte, so ISIZE bounds the
// count); ntok[b] is written by pass 1.
#pragma once
#include "inflate_batched.cuh"

namespace mbf {

constexpr int LL_ROOT2 = 8;
constexpr int D_ROOT2 = 7;

// Everything a lane needs while decoding lives in shared memory (1.6 KB per lane): a trip must never
// wait for HBM/L2 -- with 32 decoders in lock step some lane takes the long-code path in most trips.
struct LaneTables {
    uint16_t ll_lut[1 << LL_ROOT2];  // (symbol << 4 | length), LUT_SLOW, LUT_INVALID
    uint16_t d_lut[1 << D_ROOT2];    // doubles as the code-length code's LUT while a header is parsed
    uint16_t ll_sym[288];
    uint16_t d_sym[32];
    HuffDec ll_hd, d_hd;
};
template <int DL>
struct TokensShared {
    LaneTables lane[DL];
    uint32_t lenx[32];   // LEN_BASE << 8 | LEN_EXTRA
    uint32_t distx[32];  // DIST_BASE << 8 | DIST_EXTRA
    uint32_t run[16];
};

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
};

enum { T_IDLE = 0, T_HEADER = 1, T_CLEN = 2, T_BUILD = 3, T_HUFF = 4, T_STORED = 5, T_FIN = 6, T_DONE = 7 };

__device__ __forceinline__ void tok_err(const TokenArgs& a, uint32_t code, uint32_t info) {
    if (atomicCAS(a.err, 0u, code) == 0u) *a.err_info = info;
}

// DL = decoders per warp (lanes 0..DL-1 decode; all 32 lanes help with table builds).  Fewer decoders
// per warp means more warps per SM for the same shared memory: better latency hiding, fewer lanes
// served per issued instruction.
template <int DL>
__global__ void __launch_bounds__(32) k_inflate_tokens(TokenArgs a) {
    extern __shared__ __align__(16) uint8_t tk_smem[];
    TokensShared<DL>& S = *reinterpret_cast<TokensShared<DL>*>(tk_smem);
    const int lane = threadIdx.x;
    LaneTables& T = S.lane[lane < DL ? lane : 0];
    const uint32_t s_ll = smem_addr(T.ll_lut);
    const uint32_t s_d = smem_addr(T.d_lut);
    const uint32_t s_lenx = smem_addr(S.lenx);
    const uint32_t s_distx = smem_addr(S.distx);
    TokScratch* const sc = a.scratch + (size_t)blockIdx.x * DL + (lane < DL ? lane : 0);
    const uint32_t n_blocks = *a.n_blocks;
    S.lenx[lane] = lane < 29 ? ((uint32_t)MBF_TBL(LEN_BASE, lane) << 8) | MBF_TBL(LEN_EXTRA, lane) : 0u;
    S.distx[lane] = lane < 30 ? ((uint32_t)MBF_TBL(DIST_BASE, lane) << 8) | MBF_TBL(DIST_EXTRA, lane) : 0u;
    __syncwarp();

    BitReader br;
    br.words = nullptr; br.wpos = br.wlimit = 0; br.buf = 0; br.cnt = 0; br.nextw = 0;
    int st = lane < DL ? T_IDLE : T_DONE;
    uint32_t b = 0, mis = 0, clen_b = 0, isize_b = 0, bfinal = 0, stored_left = 0;
    uint32_t out_bytes = 0, ntok = 0, prev_len = 0;
    uint32_t* tk = nullptr;
    uint32_t cl_i = 0, cl_n = 0;  // code-length decoding progress

    for (;;) {
        int err = 0;
        if (st == T_HUFF) {
            // ------------------------------------------------------------ one symbol
            br.refill();
            uint32_t e = lds_u16(s_ll + (((uint32_t)br.buf & ((1u << LL_ROOT2) - 1)) << 1));
            if (e >= LUT_SLOW) {
                e = e == LUT_SLOW ? slow_decode(&T.ll_hd, T.ll_sym, br.peek(15), LL_ROOT2) : 0u;
                if (!e) err = INF_BAD_SYMBOL;
            }
            if (!err) {
                br.consume(e & 15);
                uint32_t sym = e >> 4;
                if (sym < 256) {
                    tk[ntok++] = sym;
                    out_bytes++;
                } else if (sym == 256) {
                    st = bfinal ? T_FIN : T_HEADER;
                } else if (sym >= 257 + 29) {
                    err = INF_BAD_SYMBOL;
                } else {
                    const uint32_t lx = lds_u32(s_lenx + ((sym - 257) << 2));
                    const uint32_t len = (lx >> 8) + br.take(lx & 0xff);
                    br.refill();
                    uint32_t d = lds_u16(s_d + (((uint32_t)br.buf & ((1u << D_ROOT2) - 1)) << 1));
                    if (d >= LUT_SLOW) {
                        d = d == LUT_SLOW ? slow_decode(&T.d_hd, T.d_sym, br.peek(15), D_ROOT2) : 0u;
                        if (!d) err = INF_BAD_DISTANCE;
                    }
                    if (!err && (d >> 4) >= 30) err = INF_BAD_DISTANCE;
                    if (!err) {
                        br.consume(d & 15);
                        const uint32_t dx = lds_u32(s_distx + ((d >> 4) << 2));
                        const uint32_t dist = (dx >> 8) + br.take(dx & 0xff);
                        if (dist > out_bytes) err = INF_BAD_DISTANCE;
                        else {
                            tk[ntok++] = 0x80000000u | (len << 16) | (dist - 1);
                            out_bytes += len;
                        }
                    }
                }
            }
            if (!err && out_bytes > isize_b) err = INF_OUTPUT_OVERRUN;
        } else if (st == T_CLEN) {
            // ------------------------------------------------------------ one code length (header)
            br.refill();
            const uint32_t e = T.d_lut[br.peek(CL_ROOT)];  // LUT of the code-length code
            if (e >= LUT_SLOW) {
                err = INF_BAD_HEADER;
            } else {
                br.consume(e & 15);
                const uint32_t s2 = e >> 4;
                if (s2 < 16) {
                    sc->lens[cl_i++] = (uint8_t)s2;
                    prev_len = s2;
                } else {
                    uint32_t rep, val = 0;
                    if (s2 == 16) {
                        if (cl_i == 0) err = INF_BAD_HEADER;
                        else val = prev_len;
                        rep = 3 + br.take(2);
                    } else if (s2 == 17) {
                        rep = 3 + br.take(3);
                    } else {
                        rep = 11 + br.take(7);
                    }
                    if (cl_i + rep > cl_n) err = INF_BAD_HEADER;
                    if (!err) {
                        for (uint32_t k = 0; k < rep; k++) sc->lens[cl_i++] = (uint8_t)val;
                        prev_len = val;
                    }
                }
                if (!err && cl_i >= cl_n) st = T_BUILD;  // (a missing end-of-block code is caught by the build)
            }
        } else if (st == T_STORED) {
            br.refill();
            tk[ntok++] = br.take(8);
            out_bytes++;
            if (out_bytes > isize_b) err = INF_OUTPUT_OVERRUN;
            if (--stored_left == 0) st = bfinal ? T_FIN : T_HEADER;
        } else if (st == T_HEADER || st == T_IDLE) {
            if (st == T_IDLE) {
                b = atomicAdd(a.queue_head, 1u);
                if (b >= n_blocks) {
                    st = T_DONE;
                } else {
                    const uint8_t* p = a.cbuf + a.cpos[b];
                    mis = (uint32_t)((uintptr_t)p & 3);
                    clen_b = a.clen[b];
                    isize_b = a.isize[b];
                    br.init(p, clen_b);
                    tk = a.tok + (a.uoff[b] - a.u_base);
                    ntok = 0;
                    out_bytes = 0;
                    bfinal = 0;
                    st = T_HEADER;
                }
            }
            if (st == T_HEADER) {
                br.refill();
                bfinal = br.take(1);
                const uint32_t btype = br.take(2);
                if (btype == 0) {
                    br.consume(br.cnt & 7);
                    br.refill();
                    const uint32_t len = br.take(16);
                    br.refill();
                    const uint32_t nlen = br.take(16);
                    if ((len ^ 0xffffu) != nlen) err = INF_BAD_STORED;
                    stored_left = len;
                    st = len ? T_STORED : (bfinal ? T_FIN : T_HEADER);
                } else if (btype == 1) {
                    for (int i = 0; i < 144; i++) sc->lens[i] = 8;
                    for (int i = 144; i < 256; i++) sc->lens[i] = 9;
                    for (int i = 256; i < 280; i++) sc->lens[i] = 7;
                    for (int i = 280; i < 288; i++) sc->lens[i] = 8;
                    for (int i = 0; i < 32; i++) sc->lens[288 + i] = 5;
                    sc->hlit = 288;
                    sc->hdist = 32;
                    st = T_BUILD;
                } else if (btype == 2) {
                    br.refill();
                    const uint32_t hlit = br.take(5) + 257, hdist = br.take(5) + 1, hclen = br.take(4) + 4;
                    uint8_t cl[19];
                    for (int i = 0; i < 19; i++) cl[i] = 0;
                    for (uint32_t i = 0; i < hclen; i++) {
                        br.refill();
                        cl[MBF_TBL(CL_ORDER, i)] = (uint8_t)br.take(3);
                    }
                    // the code-length code's LUT (128 entries) lives in this lane's d_lut
                    int rc = (hlit > 286 || hdist > 30) ? (int)INF_BAD_HEADER
                                                        : build_table(cl, 19, CL_ROOT, T.d_lut, T.d_sym, &T.d_hd, true);
                    if (rc) err = rc;
                    sc->hlit = hlit;
                    sc->hdist = hdist;
                    cl_i = 0;
                    cl_n = hlit + hdist;
                    st = T_CLEN;
                } else {
                    err = INF_BAD_BTYPE;
                }
            }
        }
        if (err) {
            tok_err(a, WERR_INFLATE + (uint32_t)err, b);
            st = T_IDLE;  // drop this member; the window is reported as failed
        }
        // ---------------------------------------------------------------- warp services
        unsigned need = __ballot_sync(0xffffffffu, st == T_BUILD);
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            TokScratch* t = a.scratch + (size_t)blockIdx.x * DL + src;
            LaneTables& L = S.lane[src];
            const uint32_t hlit = t->hlit, hdist = t->hdist;
            int rc = t->lens[256] == 0 ? (int)INF_BAD_HEADER
                                       : build_lut16_warp(t->lens, (int)hlit, LL_ROOT2, L.ll_lut, L.ll_sym, &L.ll_hd, S.run, lane);
            if (!rc) rc = build_lut16_warp(t->lens + hlit, (int)hdist, D_ROOT2, L.d_lut, L.d_sym, &L.d_hd, S.run, lane);
            if (lane == src) {
                if (rc) {
                    tok_err(a, WERR_INFLATE + (uint32_t)rc, b);
                    st = T_IDLE;
                } else {
                    st = T_HUFF;
                }
            }
        }
        if (st == T_FIN) {
            if (out_bytes != isize_b) tok_err(a, WERR_INFLATE + (uint32_t)INF_SIZE_MISMATCH, b);
            else if (br.bytes_consumed(mis) > clen_b) tok_err(a, WERR_INFLATE + (uint32_t)INF_INPUT_OVERRUN, b);
            a.ntok[b] = ntok;
            st = T_IDLE;
        }
        if (__ballot_sync(0xffffffffu, st != T_DONE) == 0) break;
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

__global__ void __launch_bounds__(32) k_lz_resolve(ResolveArgs a) {
    __shared__ __align__(16) uint8_t ring[RINGB];
    const uint32_t b = blockIdx.x;
    if (b >= *a.n_blocks) return;
    const int lane = threadIdx.x;
    const uint32_t ostart = a.uoff[b], oend = ostart + a.isize[b];
    const uint32_t* tk = a.tok + (ostart - a.u_base);
    const uint32_t nt = a.ntok[b];
    uint8_t* const U = a.U;
    const uint32_t s_ring = smem_addr(ring);
    uint32_t opos = ostart, flushed = ostart;
    bool bad = false;
    for (uint32_t t0 = 0; t0 < nt;) {
        // ---- next batch: as many of the next 32 tokens as fit BATCH_MAX_OUT bytes
        const uint32_t ti = t0 + lane;
        const uint32_t e = ti < nt ? __ldg(tk + ti) : 0u;
        const bool is_match = (e >> 31) != 0;
        const uint32_t mylen = ti < nt ? (is_match ? ((e >> 16) & 0x1ffu) : 1u) : 0u;
        uint32_t incl = mylen;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        const unsigned fits = __ballot_sync(0xffffffffu, mylen != 0 && incl <= BATCH_MAX_OUT);
        const uint32_t k = (uint32_t)__popc(fits);  // tokens taken (prefix property: incl is monotone)
        const uint32_t acc = __shfl_sync(0xffffffffu, incl, k ? k - 1 : 0);
        const bool mine = (uint32_t)lane < k;
        const uint32_t off = incl - mylen;
        if (k == 0 || opos + acc > oend) { bad = true; break; }
        // ---- older output to HBM first (far matches read it back through L2)
        {
            const uint32_t upto = min(opos & ~15u, oend);
            ring_flush_w<RINGB_MASK>(ring, U, flushed, upto, lane);
            if (upto > flushed) flushed = upto;
        }
        if (mine && !is_match) sts_u8(s_ring + ((opos + off) & RINGB_MASK), e & 0xffu);
        __syncwarp();
        // ---- pass 1: one whole match per lane when it reads nothing of this batch
        const uint32_t len = mine && is_match ? mylen : 0u;
        const uint32_t dist = (e & 0x7fffu) + 1u;
        const bool dep = len != 0 && dist < off + len;
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
        // ---- pass 2: matches that read this batch's own output (or overlap themselves), in order
        unsigned mm = __ballot_sync(0xffffffffu, dep);
        while (mm) {
            const int j = __ffs(mm) - 1;
            mm &= mm - 1;
            const uint32_t lj = __shfl_sync(0xffffffffu, len, j);
            const uint32_t dj = __shfl_sync(0xffffffffu, dist, j);
            const uint32_t dstj = opos + __shfl_sync(0xffffffffu, off, j);
            const uint32_t sj = dstj - dj;
            if (dj >= lj) {
                for (uint32_t i = lane; i < lj; i += 32)
                    sts_u8(s_ring + ((dstj + i) & RINGB_MASK), lds_u8(s_ring + ((sj + i) & RINGB_MASK)));
            } else {
                for (uint32_t i = lane; i < lj; i += 32)
                    sts_u8(s_ring + ((dstj + i) & RINGB_MASK), lds_u8(s_ring + ((sj + i % dj) & RINGB_MASK)));
            }
            __syncwarp();
        }
        opos += acc;
        t0 += k;
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
