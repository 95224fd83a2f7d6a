// K1, grouped form: every warp hosts G6 = 8 DEFLATE decoders (one BGZF block each), each owned by a
// group of 4 lanes.  The work is cut into warp-uniform phases so that the SIMT lanes stay on short,
// similar paths and every issued instruction serves up to 8 decoders (decode) or 32 lanes (copies):
//
//   A  decode     the 8 group leaders each decode up to Q6 symbols of their block.  Literals go
//                 straight into the group's output ring (shared memory); matches are only queued
//                 as (length, distance, offset).  Block headers are parsed here too (code lengths
//                 -> lens[]); nothing of variable length is copied in this phase.
//   B  copy       pass 1: matches whose source lies entirely before this batch (~90 %) are
//                 independent -- every lane copies whole matches (ring -> ring, or L2 -> ring for
//                 sources older than NEAR6 bytes, which were flushed before the batch);
//                 pass 2: the few matches that read this batch's own output, in order, 4 lanes each.
//   C  build      Huffman tables requested in phase A are built by all 32 lanes, one decoder at a time.
//   D  finish     completed members: flush the ring tail, verify ISIZE and input consumption, and
//                 pull the next block from the window's work queue.
//
// Per decoder 3.9 KB of shared memory (32-bit LUTs with root 8 / 7 bits, 1 KB ring, match queue,
// canonical tables for codes longer than the root) -> 7 warps = 56 decoders per SM.
#pragma once
#include "inflate_batched.cuh"

namespace mbf {

constexpr int G6 = 8;                         // decoders per warp
constexpr int RING6 = 1024;
constexpr uint32_t RING6_MASK = RING6 - 1;
constexpr uint32_t MAXOUT6 = 320;             // output bytes per batch and decoder
constexpr uint32_t NEAR6 = RING6 - MAXOUT6;   // matches with dist <= NEAR6 read the ring
constexpr uint32_t Q6 = 32;                   // symbols per batch
constexpr uint32_t MQ6 = 16;                  // matches per batch
constexpr int LL_ROOT6 = 8;
constexpr int D_ROOT6 = 7;

struct Dec6 {
    uint32_t ll_lut[1 << LL_ROOT6];
    uint32_t d_lut[1 << D_ROOT6];
    __align__(16) uint8_t ring[RING6];
    uint2 mq[MQ6];
    uint16_t ll_sym[288];
    uint16_t d_sym[32];
    HuffDec ll_hd, d_hd;
    uint8_t lens[320];
    uint32_t hlit, hdist;
};
struct Group6Shared {
    Dec6 dec[G6];
    uint32_t run[16];
};

enum { ST6_IDLE = 0, ST6_HEADER = 1, ST6_HUFF = 2, ST6_STORED = 3, ST6_DONE = 4 };
enum { F6_BUILD = 1, F6_FIN = 2, F6_NEW = 4 };

struct InflateArgs6 {
    const uint8_t* cbuf;
    const uint32_t* cpos;
    const uint32_t* clen;
    const uint32_t* isize;
    const uint32_t* uoff;
    uint8_t* U;
    uint32_t* n_blocks;      // device
    uint32_t* queue_head;    // device, zeroed per window
    uint32_t* err;           // WinMeta::err
    uint32_t* err_info;
};

__device__ __forceinline__ void set_err6(const InflateArgs6& a, uint32_t code, uint32_t info) {
    if (atomicCAS(a.err, 0u, code) == 0u) *a.err_info = info;
}

// flush ring bytes [from, to) of one decoder with the 4 lanes of its group
__device__ __forceinline__ void group_flush(const uint8_t* ring, uint8_t* U, uint32_t from, uint32_t to, int sub) {
    if (from >= to) return;
    uint32_t head_end = min(to, (from + 15u) & ~15u);
    for (uint32_t i = from + sub; i < head_end; i += 4) U[i] = ring[i & RING6_MASK];
    uint32_t body_end = to & ~15u;
    if (body_end > head_end) {
        for (uint32_t i = head_end + 16u * sub; i < body_end; i += 64u)
            *(uint4*)(U + i) = *(const uint4*)(ring + (i & RING6_MASK));
    } else {
        body_end = head_end;
    }
    for (uint32_t i = body_end + sub; i < to; i += 4) U[i] = ring[i & RING6_MASK];
}

__global__ void __launch_bounds__(32) k_inflate_group(InflateArgs6 a) {
    extern __shared__ __align__(16) uint8_t g6_smem[];
    Group6Shared& S = *reinterpret_cast<Group6Shared*>(g6_smem);
    const int lane = threadIdx.x;
    const int grp = lane >> 2, sub = lane & 3;
    const int leader_lane = grp << 2;
    const unsigned grp_mask = 0xFu << leader_lane;
    Dec6& D = S.dec[grp];
    const uint32_t s_ll = smem_addr(D.ll_lut);
    const uint32_t s_d = smem_addr(D.d_lut);
    const uint32_t s_ring = smem_addr(D.ring);
    const uint32_t s_mq = smem_addr(D.mq);
    const uint32_t n_blocks = *a.n_blocks;
    uint8_t* const U = a.U;

    // leader-only state
    BitReader br;
    br.words = nullptr; br.wpos = br.wlimit = 0; br.buf = 0; br.cnt = 0; br.nextw = 0;
    uint32_t mis = 0, bfinal = 0, stored_left = 0, b = 0, held = 0, clen_b = 0;
    int st = sub == 0 ? ST6_IDLE : ST6_DONE;
    // group-uniform state
    uint32_t opos = 0, ostart = 0, oend = 0, flushed = 0;

    for (;;) {
        uint32_t nm = 0, acc = 0, flags = 0;
        int err = 0;
        uint32_t new_ostart = 0, new_oend = 0;
        // ================================================================ phase A0: rare leader states
        if (sub == 0 && st != ST6_HUFF && st != ST6_DONE) {
            if (st == ST6_IDLE) {
                b = atomicAdd(a.queue_head, 1u);
                if (b >= n_blocks) {
                    st = ST6_DONE;
                } else {
                    const uint8_t* p = a.cbuf + a.cpos[b];
                    mis = (uint32_t)((uintptr_t)p & 3);
                    clen_b = a.clen[b];
                    br.init(p, clen_b);
                    new_ostart = a.uoff[b];
                    new_oend = new_ostart + a.isize[b];
                    opos = ostart = flushed = new_ostart;
                    oend = new_oend;
                    bfinal = 0;
                    held = 0;
                    flags |= F6_NEW;
                    st = ST6_HEADER;
                }
            }
            if (st == ST6_STORED) {
                while (stored_left && acc < MAXOUT6) {
                    br.refill();
                    sts_u8(s_ring + ((opos + acc) & RING6_MASK), br.take(8));
                    acc++;
                    stored_left--;
                }
                if (!stored_left) st = ST6_HEADER;
            } else if (st == ST6_HEADER) {
                if (bfinal) {
                    flags |= F6_FIN;
                } else {
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
                        st = len ? ST6_STORED : ST6_HEADER;
                    } else if (btype == 1) {
                        for (int i = 0; i < 144; i++) D.lens[i] = 8;
                        for (int i = 144; i < 256; i++) D.lens[i] = 9;
                        for (int i = 256; i < 280; i++) D.lens[i] = 7;
                        for (int i = 280; i < 288; i++) D.lens[i] = 8;
                        for (int i = 0; i < 32; i++) D.lens[288 + i] = 5;
                        D.hlit = 288;
                        D.hdist = 32;
                        flags |= F6_BUILD;
                        st = ST6_HUFF;
                    } else if (btype == 2) {
                        br.refill();
                        const uint32_t hlit = br.take(5) + 257, hdist = br.take(5) + 1, hclen = br.take(4) + 4;
                        uint8_t cl[19];
                        for (int i = 0; i < 19; i++) cl[i] = 0;
                        for (uint32_t i = 0; i < hclen; i++) {
                            br.refill();
                            cl[MBF_TBL(CL_ORDER, i)] = (uint8_t)br.take(3);
                        }
                        // the code-length code's 16-bit LUT lives in d_lut until the real table is built
                        uint16_t* cl_lut = (uint16_t*)D.d_lut;
                        int rc = (hlit > 286 || hdist > 30) ? (int)INF_BAD_HEADER
                                                            : build_table(cl, 19, CL_ROOT, cl_lut, D.d_sym, &D.d_hd, true);
                        const uint32_t nn = hlit + hdist;
                        uint32_t i = 0;
                        while (!rc && i < nn) {
                            br.refill();
                            const uint32_t e = cl_lut[br.peek(CL_ROOT)];
                            if (e >= LUT_SLOW) { rc = INF_BAD_HEADER; break; }
                            br.consume(e & 15);
                            const uint32_t s2 = e >> 4;
                            if (s2 < 16) { D.lens[i++] = (uint8_t)s2; continue; }
                            uint32_t rep, val = 0;
                            if (s2 == 16) {
                                if (i == 0) { rc = INF_BAD_HEADER; break; }
                                val = D.lens[i - 1];
                                rep = 3 + br.take(2);
                            } else if (s2 == 17) {
                                rep = 3 + br.take(3);
                            } else {
                                rep = 11 + br.take(7);
                            }
                            if (i + rep > nn) { rc = INF_BAD_HEADER; break; }
                            for (uint32_t k = 0; k < rep; k++) D.lens[i++] = (uint8_t)val;
                        }
                        if (!rc && D.lens[256] == 0) rc = INF_BAD_HEADER;
                        D.hlit = hlit;
                        D.hdist = hdist;
                        if (rc) err = rc;
                        else flags |= F6_BUILD;
                        st = ST6_HUFF;
                    } else {
                        err = INF_BAD_BTYPE;
                    }
                }
            }
        }
        // ================================================================ phase A1: SIMT decode
        // The loop is warp-uniform (every lane iterates, the vote at its top is the reconvergence
        // point); only active leaders do work, one symbol per trip, no early exits inside.
        bool act = sub == 0 && st == ST6_HUFF && !(flags & F6_BUILD) && !err;
        if (act && held) {
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(s_mq), "r"(held), "r"(0u) : "memory");
            nm = 1;
            acc = held >> 16;
            held = 0;
        }
        for (uint32_t it = 0; it < Q6; it++) {
            if (!__any_sync(0xffffffffu, act)) break;
            if (act) {
                br.refill();
                uint32_t e = lds_u32(s_ll + (((uint32_t)br.buf & ((1u << LL_ROOT6) - 1)) << 2));
                if (e >= K_SPECIAL) {
                    if (e == E_SLOW) e = slow_entry<false>(&D.ll_hd, D.ll_sym, br.peek(15), LL_ROOT6);
                    if (e >= K_SPECIAL) err = INF_BAD_SYMBOL;
                }
                if (err) {
                    act = false;
                } else if (e < K_LEN) {  // literal: straight into the ring
                    br.consume(e & 15);
                    sts_u8(s_ring + ((opos + acc) & RING6_MASK), (e >> 8) & 0xffu);
                    acc++;
                } else if (e >= K_EOB) {
                    br.consume(e & 15);
                    st = ST6_HEADER;
                    act = false;
                } else {
                    const uint32_t cl = e & 15, xb = (e >> 4) & 15;
                    const uint32_t len = ((e >> 8) & 0xffffu) + bfe_u32((uint32_t)br.buf, cl, xb);
                    br.consume(cl + xb);
                    br.refill();
                    uint32_t d = lds_u32(s_d + (((uint32_t)br.buf & ((1u << D_ROOT6) - 1)) << 2));
                    if (d >= K_SPECIAL) {
                        if (d == E_SLOW) d = slow_entry<true>(&D.d_hd, D.d_sym, br.peek(15), D_ROOT6);
                        if (d >= K_SPECIAL) err = INF_BAD_DISTANCE;
                    }
                    const uint32_t dcl = d & 15, dxb = (d >> 4) & 15;
                    const uint32_t dist = ((d >> 8) & 0xffffu) + bfe_u32((uint32_t)br.buf, dcl, dxb);
                    if (!err && dist > opos + acc - ostart) err = INF_BAD_DISTANCE;
                    if (err) {
                        act = false;
                    } else {
                        br.consume(dcl + dxb);
                        const uint32_t packed = (len << 16) | (dist - 1);
                        if (acc + len > MAXOUT6) {  // does not fit this batch: first entry of the next one
                            held = packed;
                            act = false;
                        } else {
                            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(s_mq + 8 * nm), "r"(packed), "r"(acc) : "memory");
                            nm++;
                            acc += len;
                        }
                    }
                }
                if (nm >= MQ6 || acc >= MAXOUT6) act = false;
            }
        }
        if (sub == 0) {
            if (!err && st != ST6_DONE && opos + acc > oend) err = INF_OUTPUT_OVERRUN;
            if (err) {
                set_err6(a, WERR_INFLATE + (uint32_t)err, b);
                st = ST6_IDLE;  // drop this member
                nm = 0;
                acc = 0;
                held = 0;
                flags &= F6_NEW;
            }
        }
        __syncwarp();
        // ---- leader -> group
        nm = __shfl_sync(0xffffffffu, nm, leader_lane);
        acc = __shfl_sync(0xffffffffu, acc, leader_lane);
        flags = __shfl_sync(0xffffffffu, flags, leader_lane);
        {
            const uint32_t os = __shfl_sync(0xffffffffu, new_ostart, leader_lane);
            const uint32_t oe = __shfl_sync(0xffffffffu, new_oend, leader_lane);
            if (flags & F6_NEW) {
                ostart = os;
                oend = oe;
                opos = flushed = os;
            }
        }
        // ================================================================ phase B: copies
        if (acc) {
            const uint32_t upto = min(opos & ~15u, oend);
            group_flush(D.ring, U, flushed, upto, sub);
            if (upto > flushed) flushed = upto;
        }
        __syncwarp();
        const uint32_t max_nm = __reduce_max_sync(0xffffffffu, nm);
        bool has_dep = false;
        // pass 1: matches that read nothing of this batch -- one whole match per lane, 16 bytes per step
        for (uint32_t r = 0; r * 4 < max_nm; r++) {
            const uint32_t j = r * 4 + sub;
            uint32_t len = 0, dist = 1, off = 0;
            if (j < nm) {
                const uint2 m = D.mq[j];
                len = m.x >> 16;
                dist = (m.x & 0x7fffu) + 1u;
                off = m.y;
                if (dist < off + len) {  // reads this batch (or overlaps itself): pass 2
                    has_dep = true;
                    len = 0;
                }
            }
            const uint32_t dst = opos + off, src = dst - dist;
            const bool near = dist <= NEAR6;
            for (uint32_t base = 0; __any_sync(0xffffffffu, base < len); base += 16) {
                if (base < len) {
                    const uint32_t rem = min(16u, len - base);
                    uint32_t w[4];
                    if (near) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            w[q] = 0;
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                if ((uint32_t)(4 * q + k) < rem) w[q] |= lds_u8(s_ring + ((src + base + 4 * q + k) & RING6_MASK)) << (8 * k);
                        }
                    } else {
                        const uint8_t* sp = U + src + base;  // flushed before this batch; read through L2
#pragma unroll
                        for (int q = 0; q < 4; q++) w[q] = (uint32_t)(4 * q) < rem ? ldu32_cg(sp + 4 * q) : 0u;
                    }
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        if ((uint32_t)k < rem) sts_u8(s_ring + ((dst + base + k) & RING6_MASK), (w[k >> 2] >> (8 * (k & 3))) & 0xffu);
                }
            }
        }
        __syncwarp();
        // pass 2: matches that read this batch's own output (or overlap themselves), in order, 4 lanes each
        if (__any_sync(0xffffffffu, has_dep)) {
            for (uint32_t j = 0; j < max_nm; j++) {
                uint32_t len = 0, dist = 1, off = 0;
                if (j < nm) {
                    const uint2 m = D.mq[j];
                    const uint32_t l = m.x >> 16, d = (m.x & 0x7fffu) + 1u;
                    if (d < m.y + l) { len = l; dist = d; off = m.y; }
                }
                if (!__any_sync(0xffffffffu, len != 0)) continue;
                const uint32_t dst = opos + off, src = dst - dist;
                for (uint32_t i0 = 0; __any_sync(0xffffffffu, i0 < len); i0 += 4) {
                    const uint32_t i = i0 + sub;
                    if (i < len) {
                        const uint32_t k = dist >= len ? i : i % dist;  // periodic source when it overlaps itself
                        sts_u8(s_ring + ((dst + i) & RING6_MASK), lds_u8(s_ring + ((src + k) & RING6_MASK)));
                    }
                }
                __syncwarp();
            }
        }
        opos += acc;
        // ================================================================ phase C: table builds
        unsigned need = __ballot_sync(0xffffffffu, sub == 0 && (flags & F6_BUILD));
        while (need) {
            const int g = (__ffs(need) - 1) >> 2;
            need &= need - 1;
            Dec6& T = S.dec[g];
            const uint32_t hlit = T.hlit, hdist = T.hdist;
            int rc = build_lut_warp<false>(T.lens, (int)hlit, LL_ROOT6, T.ll_lut, T.ll_sym, &T.ll_hd, S.run, lane);
            if (!rc) rc = build_lut_warp<true>(T.lens + hlit, (int)hdist, D_ROOT6, T.d_lut, T.d_sym, &T.d_hd, S.run, lane);
            if (rc && grp == g && sub == 0) {
                set_err6(a, WERR_INFLATE + (uint32_t)rc, b);
                st = ST6_IDLE;
            }
        }
        // ================================================================ phase D: finished members
        if (flags & F6_FIN) {
            group_flush(D.ring, U, flushed, min(opos, oend), sub);
            if (sub == 0) {
                if (opos != oend) set_err6(a, WERR_INFLATE + (uint32_t)INF_SIZE_MISMATCH, b);
                else if (br.bytes_consumed(mis) > clen_b) set_err6(a, WERR_INFLATE + (uint32_t)INF_INPUT_OVERRUN, b);
                st = ST6_IDLE;
            }
        }
        if (__ballot_sync(0xffffffffu, st != ST6_DONE) == 0) break;
    }
}

}  // namespace mbf
