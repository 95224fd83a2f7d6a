// sm_100a kernels of the read-counting path.  One "window" = a slice of the BGZF file resident
// in HBM; per window the launch sequence is
//   k_win_begin -> k_scan<count> -> k_excl_scan -> k_scan<emit> -> k_blocks_finalize -> k_excl_scan
//   -> k_par_headers -> k_tokens_par (+ k_tokens_bf for the members it hands over) -> k_lz_resolve
//                                                                        (inflate_par.cuh, inflate_twopass.cuh)
//   -> k_frame_walk<count> -> k_frame_fix -> k_excl_scan -> k_frame_emit (| k_frame_walk<emit>) -> k_carry_copy
//   -> the job's record kernel: k_count_reads<MODE> (+ k_mm_insert) | k_count_introns | k_count_reads<3> (quantify)
//      | k_dup_insert | k_barcode_hits
// Everything reads its sizes from a device-resident WinMeta, so the host never has to wait for a
// kernel before it can enqueue the next one (it reads WinMeta only for capacity decisions).
//
// Reference behaviour replaced (paths relative to the reference tree):
//   k_scan_*/k_blocks_finalize  htslib bgzf block iteration under bam.fetch/read (counters.rs:40-41)
//   k_tokens_par/_bf + k_lz_resolve  zlib inflate in htslib bgzf_read_block
//   k_frame_walk                htslib bam_read1 record framing
//   k_count_reads               count_reads_in_region_{unstranded,stranded} (counters.rs:26-202),
//                               BamRecordExtensions::blocks (bam_ext.rs:29-34), Record::aux(b"NH"),
//                               bio IntervalTree::find
//   k_mm_insert                 multimapper_dedup HashMap<u32, HashSet<Vec<u8>>> (counters.rs:36,70-73,102-104)
//   k_count_introns             count_introns (introns.rs:24-51)
//   k_count_reads<3>            quantify_gene_reads (quantify.rs:98-146)
//   k_dup_insert, k_dup_hist    py_calculate_duplicate_distribution (duplicate_distribution.rs:44-86)
//   k_barcode_hits              count_reads_primary_only_right_strand_only_by_barcode (by_barcode.rs:101-182)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "inflate_core.cuh"
#include "inflate_batched.cuh"

namespace mbf {

constexpr uint32_t U_PREFIX = 4u << 20;  // room in front of a window's data for a carried record
constexpr int NH_DENSE = 64;             // weighted mode: exact integer counters for NH 1..64
constexpr int HIST_DENSE = 64;           // introns-per-read histogram kept dense below this

enum WinErr : uint32_t {
    WERR_NONE = 0,
    WERR_CHAIN = 100,       // BGZF chain broken / signature scan disagrees
    WERR_TRUNCATED = 101,   // a BGZF block reaches past the available bytes
    WERR_ISIZE = 102,       // ISIZE > 65536
    WERR_RECORD = 103,      // malformed BAM record (block_size < 32 or inconsistent lengths)
    WERR_CARRY = 104,       // carried record larger than U_PREFIX
    WERR_NH_TYPE = 105,     // NH tag present but not an integer type
    WERR_AUX = 106,         // malformed aux area
    WERR_CAPACITY = 107,    // a device buffer was too small (host retries with a larger one)
    WERR_BARCODE_LEN = 109, // XC barcode longer than BARCODE_MAX bytes
    WERR_USIZE = 108,       // the window's blocks inflate past the 32-bit offset range (host retries with smaller windows)
    WERR_INFLATE = 200,     // + InflateError
};

struct WinMeta {
    uint32_t n_cand;        // signature candidates found
    uint32_t n_blocks;      // BGZF blocks that start in the window
    uint32_t u_total;       // sum of ISIZE
    uint32_t n_records;     // records framed
    uint32_t carry_len;     // bytes of the trailing incomplete record (copied to the next window)
    uint32_t frame_bad;     // some block did not start on a record boundary
    uint32_t n_fixups;
    uint32_t err;
    uint32_t err_info;
    uint32_t mm_count;      // multi-mapper keys appended this window (may exceed capacity)
    uint32_t owned_records;
    uint32_t cigar_ops;     // bound for the map inserts of a window's second pass: sum of n_cigar_op of owned records
                            // (introns), (read, gene) hits (quantify)
    uint32_t inflate_next;  // work queue head of the token kernel (pass 1 of K1)
    uint32_t fb_count;      // members pass 1 handed to the fallback token kernel ...
    uint32_t fb_next;       // ... and that kernel's work queue head
    uint32_t range_start;   // first window of a byte range: no carry, first record at first_rec_uoff
    uint32_t first_rec_uoff;
    unsigned long long outside;
    unsigned long long n_tokens;  // DEFLATE symbols decoded by the token pass (statistics)
};

struct ChainState {
    unsigned long long chain_cur;  // file offset where this window's first block must start
    unsigned long long chain_nxt;  // ... and the next window's
};

__device__ __forceinline__ void set_err(WinMeta* m, uint32_t code, uint32_t info) {
    if (atomicCAS(&m->err, 0u, code) == 0u) m->err_info = info;
}

// ---------------------------------------------------------------------------------------------
// unaligned little-endian loads
__device__ __forceinline__ uint32_t ldu32(const uint8_t* p) {
    uintptr_t a = (uintptr_t)p;
    uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
// same through L2 only (.cg): for bytes other lanes of this SM may have written recently
__device__ __forceinline__ uint32_t ldu32_cg(const uint8_t* p) {
    uintptr_t a = (uintptr_t)p;
    uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t lo = __ldcg(w);
    if (sh == 0) return lo;
    return __funnelshift_r(lo, __ldcg(w + 1), sh);
}
__device__ __forceinline__ uint32_t ldu16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// ---------------------------------------------------------------------------------------------
// generic single-CTA exclusive scan of u32 (n read from device memory); writes total to *total_out.
// The running sum is kept in 64 bits; with `ovf_meta` a total above `limit` is reported as WERR_USIZE
// (err_info = total in MiB) instead of wrapping silently.
__global__ void __launch_bounds__(1024) k_excl_scan(const uint32_t* in, uint32_t* out, const uint32_t* n_ptr,
                                                     uint32_t n_direct, uint32_t add, uint32_t* total_out,
                                                     WinMeta* ovf_meta, unsigned long long limit) {
    __shared__ uint32_t warp_sums[32];
    __shared__ unsigned long long carry_s;
    const uint32_t n = n_ptr ? *n_ptr : n_direct;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = add;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? in[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += y;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const unsigned long long c = carry_s;  // a tile of 1024 values < 2^22 each cannot wrap 32 bits
        uint32_t prefix = (uint32_t)c + (wid ? warp_sums[wid - 1] : 0) + x - v;
        if (i < n) out[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = c + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long total = carry_s - add;
        if (ovf_meta && total > limit) {
            set_err(ovf_meta, WERR_USIZE, (uint32_t)(total >> 20));
            if (total_out) *total_out = 0;
        } else if (total_out) {
            *total_out = (uint32_t)total;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K0: BGZF block-boundary scanner.
// A window holds file bytes [win_off, win_off + avail); blocks that START in
// [chain_cur, win_off + scan_len) belong to it.  Every 16-byte granule is tested for the fixed 12
// bytes of a BGZF header (1f 8b 08 04 .. XLEN=6 'B' 'C' 02 00); at most one header can start per
// granule (a member is >= 28 bytes).  k_blocks_finalize then proves that the candidates form the
// BSIZE chain starting at chain_cur; any disagreement is an error (no silent guess).
constexpr int SCAN_TILE = 4096;                  // bytes per CTA
constexpr int SCAN_THREADS = SCAN_TILE / 16;     // 256 threads, one granule each

struct ScanArgs {
    const uint8_t* cbuf;       // window bytes (16-byte aligned), >= 32 readable bytes of slack
    unsigned long long win_off;
    uint32_t scan_len;         // block starts are searched in [.., scan_len)
    uint32_t avail;            // bytes present in cbuf
    ChainState* chain;
    WinMeta* meta;
    uint32_t* tile_count;      // per tile
    uint32_t* tile_base;       // exclusive scan of tile_count
    uint32_t* cand_pos;        // out: offset in cbuf
    uint32_t* cand_bsize;      // out: BSIZE + 1
    uint32_t cand_cap;
};

__device__ __forceinline__ int scan_granule(const ScanArgs& a, uint32_t g_off, uint32_t lo_limit, uint32_t* bsize1) {
    // returns in-granule offset of a header start or -1
    if (g_off >= a.scan_len) return -1;
    const uint4* p = (const uint4*)(a.cbuf + g_off);
    uint4 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    uint32_t w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int wi = k >> 2, sh = (k & 3) * 8;
        auto word = [&](int i) -> uint32_t {  // 32 bits starting at byte k + 4*i
            return sh ? __funnelshift_r(w[wi + i], w[wi + i + 1], sh) : w[wi + i];
        };
        if (word(0) != 0x04088b1fu) continue;
        uint32_t x2 = word(2);  // bytes 8..11: XFL OS XLEN
        uint32_t x3 = word(3);  // bytes 12..15: 'B' 'C' 02 00
        if ((x2 >> 16) != 6u || x3 != 0x00024342u) continue;
        uint32_t pos = g_off + k;
        if (pos < lo_limit || pos >= a.scan_len) continue;
        *bsize1 = (word(4) & 0xffffu) + 1u;  // bytes 16..17
        return k;
    }
    return -1;
}

__global__ void __launch_bounds__(1) k_win_begin(ChainState* chain, WinMeta* meta, int range_start,
                                                 unsigned long long range_off, uint32_t first_rec_uoff) {
    if (range_start) chain->chain_cur = range_off;
    else chain->chain_cur = chain->chain_nxt;
    chain->chain_nxt = chain->chain_cur;
    WinMeta z = {};
    z.range_start = range_start ? 1u : 0u;
    z.first_rec_uoff = range_start ? first_rec_uoff : 0u;
    *meta = z;
}

template <bool EMIT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(ScanArgs a) {
    __shared__ uint32_t warp_cnt[SCAN_THREADS / 32];
    const unsigned long long cur = a.chain->chain_cur;
    const uint32_t lo_limit = cur > a.win_off ? (uint32_t)(cur - a.win_off) : 0u;
    const uint32_t g_off = blockIdx.x * SCAN_TILE + threadIdx.x * 16;
    uint32_t bs1 = 0;
    int k = scan_granule(a, g_off, lo_limit, &bs1);
    const unsigned hit = __ballot_sync(0xffffffffu, k >= 0);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(hit);
    __syncthreads();
    if (!EMIT) {
        if (threadIdx.x == 0) {
            uint32_t s = 0;
            for (int i = 0; i < SCAN_THREADS / 32; i++) s += warp_cnt[i];
            a.tile_count[blockIdx.x] = s;
        }
    } else if (k >= 0) {
        uint32_t idx = a.tile_base[blockIdx.x];
        for (int i = 0; i < wid; i++) idx += warp_cnt[i];
        idx += __popc(hit & ((1u << lane) - 1));
        if (idx < a.cand_cap) {
            a.cand_pos[idx] = g_off + k;
            a.cand_bsize[idx] = bs1;
        }
    }
}

struct BlockArrays {
    uint32_t* cpos;   // payload start (offset in cbuf)
    uint32_t* clen;   // payload bytes
    uint32_t* isize;
    uint32_t* uoff;   // output offset in the U buffer (exclusive scan of isize + U_PREFIX)
};

__global__ void __launch_bounds__(256) k_blocks_finalize(ScanArgs a, BlockArrays b) {
    const uint32_t n = a.meta->n_cand;
    const unsigned long long cur = a.chain->chain_cur;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (n > a.cand_cap) set_err(a.meta, WERR_CAPACITY, n);
        a.meta->n_blocks = n > a.cand_cap ? 0 : n;
        // a window with no block start: the chain must already point past its scan range
        if (n == 0 && cur < a.win_off + a.scan_len && cur < a.win_off + a.avail) set_err(a.meta, WERR_CHAIN, 0);
    }
    if (n > a.cand_cap) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t pos = a.cand_pos[i], bs1 = a.cand_bsize[i];
        if (i == 0 && a.win_off + pos != cur) set_err(a.meta, WERR_CHAIN, pos);
        if (i + 1 < n && a.cand_pos[i + 1] != pos + bs1) set_err(a.meta, WERR_CHAIN, pos);
        if (i + 1 == n) {
            // the next header (if any) lies outside the scanned range; chain continues there
            if (pos + bs1 < a.scan_len) set_err(a.meta, WERR_CHAIN, pos);
            a.chain->chain_nxt = a.win_off + pos + bs1;
        }
        uint32_t isz = 0;
        if (bs1 < 28 || pos + bs1 > a.avail) {
            set_err(a.meta, WERR_TRUNCATED, pos);
            bs1 = 28;
        } else {
            isz = ldu32(a.cbuf + pos + bs1 - 4);
            if (isz > 65536u) { set_err(a.meta, WERR_ISIZE, pos); isz = 0; }
        }
        b.cpos[i] = pos + 18;
        b.clen[i] = bs1 - 26;
        b.isize[i] = isz;
    }
}

// ---------------------------------------------------------------------------------------------
// K1 (inflate_par.cuh, inflate_twopass.cuh): helpers shared by the inflate kernels.
// explicit 32-bit shared-memory addressing for the hot loops (generic-pointer accesses make ptxas
// re-derive the shared window base, SR_CgaCtaId, on the decode chain)
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// shared-window address of a __shared__ object; volatile so that ptxas keeps the value in a register
// instead of re-deriving it (S2R SR_CgaCtaId + LEA) inside the decode loop
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    uint32_t a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

}  // namespace mbf
#include "inflate_twopass.cuh"
#include "inflate_par.cuh"
namespace mbf {

// ---------------------------------------------------------------------------------------------
// K2: record framing.  One walker thread per BGZF block follows the block_size chain of the records
// that START inside its block.  Walkers begin speculatively at their block's first byte (htslib never
// lets a record straddle blocks); k_frame_fix proves the speculation (every walker must end exactly
// where the next one began) and otherwise re-walks the affected blocks serially.
struct FrameArgs {
    const uint8_t* U;
    const BlockArrays blk;
    uint32_t* carry;      // bytes carried in front of this window's data (written by k_carry_copy)
    WinMeta* meta;
    uint32_t* blk_start;  // first record start per block (offset in U)
    uint32_t* blk_end;    // where the walker stopped
    uint32_t* blk_nrec;
    uint32_t* rec_base;   // exclusive scan of blk_nrec
    uint32_t* rec_off;    // out (EMIT)
    uint32_t rec_cap;
    uint32_t* rec_tmp;    // count walk: offsets of the records of member b at [start_b / 36 ...) (a record is >= 36 bytes,
                          // so the members' runs cannot collide); k_frame_emit compacts them into rec_off
    uint32_t tmp_cap;
};

__device__ __forceinline__ uint32_t walk_block(const uint8_t* U, uint32_t pos, uint32_t limit, uint32_t u_end,
                                               uint32_t* nrec_out, uint32_t* rec_off, uint32_t rec_cap,
                                               uint32_t rec_idx) {
    // records starting in [pos, limit); a record that does not fit before u_end stops the walk
    uint32_t n = 0;
    while (pos < limit) {
        if (pos + 4 > u_end) break;
        uint32_t bs = ldu32(U + pos);
        if (bs < 32u || bs > 0x7fffffffu) { pos = 0xffffffffu; break; }  // malformed (or bad speculation)
        unsigned long long nxt = (unsigned long long)pos + 4ull + bs;
        if (nxt > u_end) break;
        if (rec_off && rec_idx + n < rec_cap) rec_off[rec_idx + n] = pos;
        n++;
        pos = (uint32_t)nxt;
    }
    *nrec_out = n;
    return pos;
}

template <bool EMIT>
__global__ void __launch_bounds__(128) k_frame_walk(FrameArgs a) {
    const uint32_t nb = a.meta->n_blocks;
    const uint32_t u_end = U_PREFIX + a.meta->u_total;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
        uint32_t start;
        if (EMIT) start = a.blk_start[b];
        else if (b == 0) start = a.meta->range_start ? U_PREFIX + a.meta->first_rec_uoff : U_PREFIX - *a.carry;
        else start = a.blk.uoff[b];
        uint32_t limit = b + 1 < nb ? a.blk.uoff[b + 1] : u_end;
        uint32_t n = 0;
        if (EMIT && !a.meta->frame_bad) return;  // the provisional offsets are good: k_frame_emit copies them
        uint32_t end = walk_block(a.U, start, limit, u_end, &n, EMIT ? a.rec_off : a.rec_tmp, EMIT ? a.rec_cap : a.tmp_cap,
                                  EMIT ? a.rec_base[b] : start / 36u);
        if (!EMIT) {
            a.blk_start[b] = start;
            a.blk_end[b] = end;
            a.blk_nrec[b] = n;
        }
    }
}

// The count walk left member b's record offsets at rec_tmp[blk_start[b] / 36 ...]: one warp per member copies them
// to their final places (coalesced both ways) -- unless some walker had started off a record boundary
// (meta->frame_bad: htsjdk-style files), in which case k_frame_walk<emit> walks the corrected chain again.
__global__ void __launch_bounds__(128) k_frame_emit(FrameArgs a) {
    if (a.meta->frame_bad) return;
    const uint32_t nb = a.meta->n_blocks;
    const int lane = threadIdx.x & 31;
    for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nb; b += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t n = a.blk_nrec[b], from = a.blk_start[b] / 36u, to = a.rec_base[b];
        for (uint32_t i = lane; i < n; i += 32)
            if (to + i < a.rec_cap) a.rec_off[to + i] = a.rec_tmp[from + i];
    }
}

// one CTA: validate the walker chain in parallel; only when speculation failed somewhere does thread 0
// repair the affected blocks serially; then derive the carry
__global__ void __launch_bounds__(1024) k_frame_fix(FrameArgs a) {
    __shared__ uint32_t s_bad, s_malformed;
    const uint32_t nb = a.meta->n_blocks;
    const uint32_t u_end = U_PREFIX + a.meta->u_total;
    if (threadIdx.x == 0) { s_bad = 0; s_malformed = 0xffffffffu; }
    __syncthreads();
    uint32_t bad = 0, malformed = 0xffffffffu;
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
        const uint32_t e = a.blk_end[b];
        if (e == 0xffffffffu) malformed = min(malformed, b);
        if (b + 1 < nb && e != a.blk_start[b + 1]) bad = 1;
    }
    if (bad) s_bad = 1;
    if (malformed != 0xffffffffu) atomicMin(&s_malformed, malformed);
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t fixups = 0;
    if (s_bad) {
        a.meta->frame_bad = 1;
        s_malformed = 0xffffffffu;  // re-derived below: bad speculation also ends walks with the malformed mark
        for (uint32_t b = 0; b + 1 < nb; b++) {
            uint32_t e = a.blk_end[b];
            if (e == 0xffffffffu) { s_malformed = b; break; }
            if (e != a.blk_start[b + 1]) {
                uint32_t limit = b + 2 < nb ? a.blk.uoff[b + 2] : u_end;
                uint32_t n = 0;
                uint32_t e2 = e >= limit ? e : walk_block(a.U, e, limit, u_end, &n, nullptr, 0, 0);
                a.blk_start[b + 1] = e;
                a.blk_end[b + 1] = e2;
                a.blk_nrec[b + 1] = n;
                fixups++;
            }
        }
        if (s_malformed == 0xffffffffu && nb && a.blk_end[nb - 1] == 0xffffffffu) s_malformed = nb - 1;
    }
    a.meta->n_fixups = fixups;
    uint32_t carry = 0;
    if (nb) {
        uint32_t e = a.blk_end[nb - 1];
        if (s_malformed != 0xffffffffu) { set_err(a.meta, WERR_RECORD, s_malformed); e = u_end; }
        if (e < u_end) carry = u_end - e;
        if (carry > U_PREFIX) { set_err(a.meta, WERR_CARRY, carry); carry = 0; }
    } else if (!a.meta->range_start) {
        carry = *a.carry;  // nothing new in this window: keep carrying
    }
    a.meta->carry_len = carry;
}

// copies the carried tail of this window's U to the front of the next window's U and publishes it
__global__ void __launch_bounds__(256) k_carry_copy(const uint8_t* U_cur, uint8_t* U_next, uint32_t* carry_out,
                                                    const WinMeta* meta) {
    const uint32_t carry = meta->carry_len;
    const uint32_t nb = meta->n_blocks;
    const uint32_t src_end = nb ? U_PREFIX + meta->u_total : U_PREFIX;  // nb==0: carry sits in front already
    for (uint32_t i = threadIdx.x; i < carry; i += blockDim.x) U_next[U_PREFIX - carry + i] = U_cur[src_end - carry + i];
    if (threadIdx.x == 0) *carry_out = carry;
}

// ---------------------------------------------------------------------------------------------
// 128-bit CAS (sm_90+): the multi-mapper set and the intron map use 16-byte slots
__device__ __forceinline__ ulonglong2 cas128(ulonglong2* addr, ulonglong2 cmp, ulonglong2 val) {
    ulonglong2 old;
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old.x), "=l"(old.y)
        : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr)
        : "memory");
    return old;
}

__device__ __forceinline__ unsigned long long fmix64(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}
__device__ __forceinline__ unsigned long long rotl64(unsigned long long x, int r) { return (x << r) | (x >> (64 - r)); }

// 128-bit fingerprint of (chunk, gene|bucket, qname bytes): murmur3-x64-128 style absorb of 4-byte
// words.  Stands in for the exact Vec<u8> keys of the reference's HashSet; collision odds for 1e8
// keys are ~1e-23.
__device__ __forceinline__ ulonglong2 qname_fingerprint(const uint8_t* q, uint32_t len, uint32_t chunk, uint32_t key) {
    unsigned long long h1 = 0x9E3779B97F4A7C15ull ^ ((unsigned long long)chunk << 32 | key);
    unsigned long long h2 = 0xD6E8FEB86659FD93ull + len;
    const unsigned long long c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
    uint32_t i = 0;
    for (; i + 4 <= len; i += 4) {
        unsigned long long w = ldu32(q + i);
        unsigned long long k1 = w * c1;
        k1 = rotl64(k1, 31) * c2;
        h1 ^= k1;
        h1 = rotl64(h1, 27) + h2;
        h1 = h1 * 5 + 0x52dce729;
        unsigned long long k2 = (w + 0x1234567ull) * c2;
        k2 = rotl64(k2, 33) * c1;
        h2 ^= k2;
        h2 = rotl64(h2, 31) + h1;
        h2 = h2 * 5 + 0x38495ab5;
    }
    unsigned long long tail = 0;
    for (uint32_t j = 0; i + j < len; j++) tail |= (unsigned long long)q[i + j] << (8 * j);
    h1 ^= rotl64((tail | 0x100000000ull) * c1, 31) * c2;
    h2 ^= rotl64((tail + 0x77ull) * c2, 33) * c1;
    h1 ^= len;
    h2 ^= (unsigned long long)key * 0x9E3779B97F4A7C15ull;
    h1 += h2;
    h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    h1 += h2;
    h2 += h1;
    if (h1 == 0 && h2 == 0) h1 = 1;
    return make_ulonglong2(h1, h2);
}

// ---------------------------------------------------------------------------------------------
// CIGAR walk (pysam get_blocks semantics, pinned by the reference's golden records in
// src/bam_ext.rs:50-604): M/=/X emit a block and advance, D/N advance (N also emits an intron),
// I/S/H/P do nothing.
template <class FB, class FI>
__device__ __forceinline__ void cigar_walk(const uint8_t* cig, uint32_t n_ops, uint32_t pos, FB&& on_block, FI&& on_intron) {
    uint32_t p = pos;
    for (uint32_t i = 0; i < n_ops; i++) {
        uint32_t c = ldu32(cig + 4 * i);
        uint32_t op = c & 15u, len = c >> 4;
        if (op == 0u || op == 7u || op == 8u) {
            if (!on_block(p, p + len)) return;
            p += len;
        } else if (op == 3u) {
            on_intron(p, p + len);
            p += len;
        } else if (op == 2u) {
            p += len;
        }
    }
}

__global__ void k_cigar_expand(const uint32_t* cigar, int n_ops, int pos, int want_introns, uint32_t* out, int cap, int* n_out) {
    int n = 0;
    cigar_walk((const uint8_t*)cigar, (uint32_t)n_ops, (uint32_t)pos,
               [&](uint32_t s, uint32_t e) {
                   if (!want_introns) { if (n < cap) { out[2 * n] = s; out[2 * n + 1] = e; } n++; }
                   return true;
               },
               [&](uint32_t s, uint32_t e) {
                   if (want_introns) { if (n < cap) { out[2 * n] = s; out[2 * n + 1] = e; } n++; }
               });
    *n_out = n;
}

// ---------------------------------------------------------------------------------------------
// K3/K4/K5 fused: parse + CIGAR expand + overlap + per-read gene dedup + strand + counters.
struct CountCtx {
    const uint8_t* U;
    const uint32_t* rec_off;
    WinMeta* meta;
    // genome tables
    const int32_t* tid2slot;   // n_ref
    int32_t n_ref;
    const uint32_t* slot_iv_off;     // per scanned contig slot: range into the sorted interval arrays
    const uint32_t* iv_start;
    const uint32_t* iv_end;
    const uint32_t* iv_pmax;         // prefix max of iv_end (start-sorted order)
    const uint32_t* iv_gene;         // global gene slot | strand bit in bit 31 (1 = gene strand == +1)
    const uint32_t* iv_rank;         // each_read_counts_once: rank in bio's find() order (reader.hpp); else null
    const uint32_t* slot_chunk_off;  // range into chunk_stop
    const uint32_t* chunk_stop;      // ascending per slot
    uint32_t chunk_lo, chunk_hi;     // owned global chunk ids
    // outputs
    unsigned long long* counts;      // [2][n_genes]  (weighted: [NH_DENSE][2][n_genes])
    uint32_t n_genes;
    ulonglong2* mm_fp;               // appended multi-mapper keys
    uint32_t* mm_slot;               // gene | bucket<<31
    uint32_t mm_cap;
    int emit_only;                   // re-run after an mm overflow: append keys, touch no counter
    double* wspill;                  // weighted, NH > NH_DENSE: [2][n_genes] f64 (1/NH added atomically)
    // MODE 3 (quantify_gene_reads): (gene, pos) -> count per bucket in the job's 16-byte-slot map
    ulonglong2* qmap;
    uint32_t qmask;
    int q_count_only;                // pass 1: only count the (read, gene) hits of owned reads (sizes the map)
};


// key -> u32 count; slot = { key64, tag32 | count32<<32 }.  Capacity is guaranteed by the host.
__device__ __forceinline__ void map_add(ulonglong2* map, uint32_t mask, unsigned long long key, uint32_t tag, uint32_t n) {
    unsigned long long h = fmix64(key ^ ((unsigned long long)tag * 0x9E3779B97F4A7C15ull));
    uint32_t i = (uint32_t)h & mask;
    for (;;) {
        ulonglong2 cur;  // one 128-bit load: a slot is published by a 128-bit CAS, two 64-bit loads could tear
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(cur.x), "=l"(cur.y) : "l"(map + i));
        if (cur.x == ~0ull && cur.y == ~0ull) {
            ulonglong2 mine = make_ulonglong2(key, (unsigned long long)tag | ((unsigned long long)n << 32));
            ulonglong2 old = cas128(map + i, make_ulonglong2(~0ull, ~0ull), mine);
            if (old.x == ~0ull && old.y == ~0ull) return;
            cur = old;
        }
        if (cur.x == key && (uint32_t)cur.y == tag) {
            atomicAdd(((uint32_t*)(map + i)) + 3, n);
            return;
        }
        i = (i + 1) & mask;
    }
}

// NH aux tag (htslib bam_aux_get + rust-htslib Aux::Integer).  0 absent, 1 integer, <0 error.
__device__ __forceinline__ int aux_find_nh(const uint8_t* p, const uint8_t* e, long long* val) {
    while (p + 3 <= e) {
        const bool match = p[0] == 'N' && p[1] == 'H';
        const uint8_t t = p[2];
        p += 3;
        uint32_t sz;
        switch (t) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'd': sz = 8; break;
            case 'Z': case 'H':
                if (match) return -1;
                while (p < e && *p) p++;
                if (p >= e) return -2;
                p++;
                continue;
            case 'B': {
                if (match) return -1;
                if (p + 5 > e) return -2;
                const uint8_t st = p[0];
                const uint32_t cnt = ldu32(p + 1);
                const uint32_t es = (st == 'c' || st == 'C') ? 1u : (st == 's' || st == 'S') ? 2u : 4u;
                unsigned long long adv = 5ull + (unsigned long long)cnt * es;
                if (adv > (unsigned long long)(e - p)) return -2;
                p += adv;
                continue;
            }
            default: return -2;
        }
        if (p + sz > e) return -2;
        if (match) {
            switch (t) {
                case 'c': *val = (int8_t)p[0]; return 1;
                case 'C': *val = p[0]; return 1;
                case 's': *val = (int16_t)ldu16(p); return 1;
                case 'S': *val = ldu16(p); return 1;
                case 'i': *val = (int32_t)ldu32(p); return 1;
                case 'I': *val = ldu32(p); return 1;
                default: return -1;
            }
        }
        p += sz;
    }
    return 0;
}

// first index in [lo,hi) with arr[i] >= x
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* arr, uint32_t lo, uint32_t hi, uint32_t x) {
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(arr + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index in [lo,hi) with arr[i] >= x, searched outward from `hint` (galloping, then bisection): records arrive
// sorted by position, so the answer for a thread's next record lies a few entries from the last one and costs 2-4
// loads instead of the log2(n) of a cold bisection
__device__ __forceinline__ uint32_t lower_bound_near(const uint32_t* arr, uint32_t lo, uint32_t hi, uint32_t x, uint32_t hint) {
    uint32_t p = min(max(hint, lo), hi), l, h, step = 1;
    if (p < hi && __ldg(arr + p) < x) {  // everything up to p is below x
        l = p + 1;
        h = hi;
        for (;;) {
            const uint32_t q = l + step - 1;
            if (q >= hi) break;
            if (__ldg(arr + q) < x) { l = q + 1; step <<= 1; } else { h = q; break; }
        }
    } else {  // everything from p on is at or above x
        h = p;
        l = lo;
        for (;;) {
            if (h - lo < step) break;
            const uint32_t q = h - step;
            if (__ldg(arr + q) >= x) { h = q; step <<= 1; } else { l = q + 1; break; }
        }
    }
    while (l < h) {
        const uint32_t mid = l + ((h - l) >> 1);
        if (__ldg(arr + mid) < x) l = mid + 1; else h = mid;
    }
    return l;
}
// first index in [lo,hi) with arr[i] > x
__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t* arr, uint32_t lo, uint32_t hi, uint32_t x) {
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(arr + mid) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct RecView {
    const uint8_t* r;      // points at block_size
    uint32_t upos, flag, n_cig, l_qn, l_seq, bs;
    int32_t tid;
    const uint8_t* qname;
    const uint8_t* cigar;
    const uint8_t* aux;
    const uint8_t* end;
};
__device__ __forceinline__ bool rec_view(const uint8_t* U, uint32_t off, RecView& v) {
    const uint8_t* r = U + off;
    v.r = r;
    v.bs = ldu32(r);
    v.tid = (int32_t)ldu32(r + 4);
    v.upos = ldu32(r + 8);
    uint32_t w = ldu32(r + 12);  // l_read_name, mapq, bin
    v.l_qn = w & 0xffu;
    uint32_t w2 = ldu32(r + 16);  // n_cigar_op, flag
    v.n_cig = w2 & 0xffffu;
    v.flag = w2 >> 16;
    v.l_seq = ldu32(r + 20);
    v.qname = r + 36;
    v.cigar = v.qname + v.l_qn;
    v.end = r + 4 + v.bs;
    unsigned long long fixed = 32ull + v.l_qn + 4ull * v.n_cig + ((unsigned long long)v.l_seq + 1) / 2 + v.l_seq;
    if (fixed > v.bs) return false;
    v.aux = r + 4 + fixed;
    return true;
}

// Enumerate the (gene|strandbit) of every interval overlapped by every aligned block, in a fixed order.
// cb(gene_word) returns false to stop.  `filter_stop`: unstranded block filter (counters.rs:52).
// each_read_counts_once (c.iv_rank != null; counters.rs:83-92,173-182): only the first hit of the first
// block that has one counts -- "first" in bio's iteration order, i.e. the overlapping interval of least rank.
// `hint`: where the previous record's search ended (0xffffffff: no idea, plain bisection); updated.
template <class CB>
__device__ __forceinline__ void enumerate_hits(const CountCtx& c, const RecView& v, uint32_t iv_lo, uint32_t iv_hi,
                                               bool use_filter, uint32_t chunk_stop, uint32_t& hint, CB&& cb) {
    const bool once = c.iv_rank != nullptr;
    cigar_walk(v.cigar, v.n_cig, v.upos,
               [&](uint32_t b0, uint32_t b1) {
                   if (use_filter && b0 >= chunk_stop) return true;
                   const uint32_t hi = hint == 0xffffffffu ? lower_bound_u32(c.iv_start, iv_lo, iv_hi, b1)
                                                           : lower_bound_near(c.iv_start, iv_lo, iv_hi, b1, hint);  // start < b1 below hi
                   if (hint != 0xffffffffu) hint = hi;
                   uint32_t best_rank = 0xffffffffu, best_gw = 0;
                   for (uint32_t k = hi; k-- > iv_lo;) {
                       if (__ldg(c.iv_pmax + k) <= b0) break;
                       if (b0 < __ldg(c.iv_end + k)) {
                           if (once) {
                               const uint32_t rk = __ldg(c.iv_rank + k);
                               if (rk < best_rank) { best_rank = rk; best_gw = __ldg(c.iv_gene + k); }
                           } else if (!cb(__ldg(c.iv_gene + k))) {
                               return false;
                           }
                       }
                   }
                   if (once && best_rank != 0xffffffffu) {
                       cb(best_gw);
                       return false;  // the read is spent
                   }
                   return true;
               },
               [](uint32_t, uint32_t) {});
}

constexpr int SEEN_K = 6;

template <int MODE>
__global__ void __launch_bounds__(256) k_count_reads(CountCtx c) {
    const uint32_t n = c.meta->n_records;
    const int lane = threadIdx.x & 31;
    uint32_t my_outside = 0, my_owned = 0, my_hits = 0;
    // Every warp takes a contiguous run of records (positions ascend within it), so the interval and chunk searches of
    // a thread start from where its previous record's ended.
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5, w_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t per_warp = (((n + n_warps - 1) / n_warps) + 31u) & ~31u;
    const uint32_t w_beg = min(n, w_id * per_warp), w_end = min(n, w_beg + per_warp);
    uint32_t hint_iv = 0, hint_ci = 0xffffffffu;
    int hint_slot = -1;
    for (uint32_t base = w_beg; base < w_end; base += 32) {
        const uint32_t i = base + lane;
        uint32_t keys[SEEN_K];
        int nk = 0;
        bool counted_direct = false;  // nh==1 (or weighted) keys go to counters; else to the mm list
        uint32_t wclass = 0;          // weighted: NH class index
        if (i < w_end) {
            RecView v;
            bool ok = rec_view(c.U, c.rec_off[i], v);
            int slot = (v.tid >= 0 && v.tid < c.n_ref) ? c.tid2slot[v.tid] : -1;
            if (!ok) { set_err(c.meta, WERR_RECORD, i); slot = -1; }
            if (slot >= 0) {
                const uint32_t clo = c.slot_chunk_off[slot], chi = c.slot_chunk_off[slot + 1];
                if (slot != hint_slot) { hint_slot = slot; hint_iv = c.slot_iv_off[slot]; hint_ci = 0xffffffffu; }
                uint32_t ci = hint_ci;  // first stop > pos: usually the previous record's chunk
                if (!(ci >= clo && ci < chi && v.upos < __ldg(c.chunk_stop + ci) && (ci == clo || __ldg(c.chunk_stop + ci - 1) <= v.upos)))
                    ci = upper_bound_u32(c.chunk_stop, clo, chi, v.upos);
                hint_ci = ci;
                if (ci < chi && ci >= c.chunk_lo && ci < c.chunk_hi) {
                    my_owned++;
                    const uint32_t cstop = c.chunk_stop[ci];
                    const uint32_t iv_lo = c.slot_iv_off[slot], iv_hi = c.slot_iv_off[slot + 1];
                    const bool rev = (v.flag & 0x10u) != 0;
                    bool hit = false, have_nh = false;
                    long long nh = 1;
                    uint32_t h_index = 0;
                    enumerate_hits(c, v, iv_lo, iv_hi, MODE == 0, cstop, hint_iv, [&](uint32_t gw) {
                        if (MODE == 3) {  // quantify: no NH logic at all (quantify.rs:117-140)
                            have_nh = true;
                            counted_direct = true;
                        }
                        if (!have_nh) {
                            have_nh = true;
                            int rc = aux_find_nh(v.aux, v.end, &nh);
                            if (rc == 0) nh = 1;
                            else if (rc < 0) { set_err(c.meta, rc == -1 ? WERR_NH_TYPE : WERR_AUX, i); nh = 1; }
                            if (MODE == 2) {
                                if (nh < 1) nh = 1;
                                counted_direct = true;
                                wclass = nh <= NH_DENSE ? (uint32_t)(nh - 1) : 0xffffffffu;
                            } else {
                                counted_direct = nh == 1;
                            }
                        }
                        hit = true;
                        const uint32_t gene = gw & 0x7fffffffu;
                        uint32_t bucket = 0;
                        if (MODE != 0) {  // counters.rs:151
                            const bool plus = (gw >> 31) != 0;
                            bucket = ((plus && !rev) || (!plus && rev)) ? 0u : 1u;
                        }
                        const uint32_t key = gene | (bucket << 31);
                        const uint32_t my_index = h_index++;
                        for (int k = 0; k < nk; k++)
                            if (keys[k] == key) return true;
                        if (nk < SEEN_K) { keys[nk++] = key; return true; }
                        // more distinct genes than the register set holds: decide "first occurrence"
                        // by re-enumerating the earlier hits, then handle the key right here.
                        bool seen = false;
                        uint32_t j = 0;
                        uint32_t no_hint = 0xffffffffu;
                        enumerate_hits(c, v, iv_lo, iv_hi, MODE == 0, cstop, no_hint, [&](uint32_t gw2) {
                            if (j++ >= my_index) return false;
                            uint32_t b2 = 0;
                            if (MODE != 0) {
                                const bool plus2 = (gw2 >> 31) != 0;
                                b2 = ((plus2 && !rev) || (!plus2 && rev)) ? 0u : 1u;
                            }
                            if (((gw2 & 0x7fffffffu) | (b2 << 31)) == key) { seen = true; return false; }
                            return true;
                        });
                        if (seen) return true;
                        if (MODE == 3) {
                            my_hits++;
                            if (!c.q_count_only)
                                map_add(c.qmap, c.qmask, (unsigned long long)v.upos | ((unsigned long long)gene << 32), bucket, 1);
                        } else if (counted_direct) {
                            if (!c.emit_only) {
                                if (MODE == 2 && wclass == 0xffffffffu)
                                    atomicAdd(c.wspill + (size_t)bucket * c.n_genes + gene, 1.0 / (double)nh);
                                else
                                    atomicAdd(c.counts + ((size_t)(MODE == 2 ? wclass * 2 : 0) + bucket) * c.n_genes + gene, 1ull);
                            }
                        } else {
                            uint32_t idx = atomicAdd(&c.meta->mm_count, 1u);
                            if (idx < c.mm_cap) {
                                c.mm_fp[idx] = qname_fingerprint(v.qname, v.l_qn ? v.l_qn - 1 : 0, ci, key);
                                c.mm_slot[idx] = key;
                            }
                        }
                        return true;
                    });
                    if (!hit) my_outside++;
                    if (MODE == 3) {
                        my_hits += (uint32_t)nk;
                        if (!c.q_count_only)
                            for (int k = 0; k < nk; k++)
                                map_add(c.qmap, c.qmask, (unsigned long long)v.upos | ((unsigned long long)(keys[k] & 0x7fffffffu) << 32),
                                        keys[k] >> 31, 1);
                        nk = 0;
                    }
                    if (nk && !counted_direct) {
                        // multi-mapper: one key per distinct (gene,bucket) of this read
                        for (int k = 0; k < nk; k++) {
                            uint32_t idx = atomicAdd(&c.meta->mm_count, 1u);
                            if (idx < c.mm_cap) {
                                c.mm_fp[idx] = qname_fingerprint(v.qname, v.l_qn ? v.l_qn - 1 : 0, ci, keys[k]);
                                c.mm_slot[idx] = keys[k];
                            }
                        }
                        nk = 0;
                    }
                    if (MODE == 2 && nk && wclass == 0xffffffffu) {
                        if (!c.emit_only)
                            for (int k = 0; k < nk; k++)
                                atomicAdd(c.wspill + (size_t)(keys[k] >> 31) * c.n_genes + (keys[k] & 0x7fffffffu), 1.0 / (double)nh);
                        nk = 0;
                    }
                }
            }
        }
        if (c.emit_only) nk = 0;
        // warp-aggregated atomics into the per-gene counters
        const int maxk = __reduce_max_sync(0xffffffffu, nk);
        for (int k = 0; k < maxk; k++) {
            const bool valid = k < nk;
            const uint32_t key = valid ? keys[k] : 0xffffffffu;
            const unsigned long long full = valid ? ((unsigned long long)wclass << 32 | key) : ~0ull;
            const unsigned peers = __match_any_sync(0xffffffffu, full);
            if (valid && lane == __ffs(peers) - 1) {
                const uint32_t gene = key & 0x7fffffffu, bucket = key >> 31;
                atomicAdd(c.counts + ((size_t)(MODE == 2 ? wclass * 2 : 0) + bucket) * c.n_genes + gene,
                          (unsigned long long)__popc(peers));
            }
        }
    }
    if (MODE == 3) {
        if (c.q_count_only) {  // pass 1 publishes the bound for the map, pass 2 nothing
            my_hits = __reduce_add_sync(0xffffffffu, my_hits);
            my_owned = __reduce_add_sync(0xffffffffu, my_owned);
            if (lane == 0) {
                if (my_hits) atomicAdd(&c.meta->cigar_ops, my_hits);
                if (my_owned) atomicAdd(&c.meta->owned_records, my_owned);
            }
        }
    } else if (!c.emit_only) {
        my_outside = __reduce_add_sync(0xffffffffu, my_outside);
        my_owned = __reduce_add_sync(0xffffffffu, my_owned);
        if (lane == 0) {
            if (my_outside) atomicAdd(&c.meta->outside, (unsigned long long)my_outside);
            if (my_owned) atomicAdd(&c.meta->owned_records, my_owned);
        }
    }
}

// K5: insert the appended multi-mapper fingerprints into the persistent set; a key seen for the first
// time adds 1 to its gene (== HashSet::len() summed at chunk end, counters.rs:102-104).
__global__ void __launch_bounds__(256) k_mm_insert(const ulonglong2* fp, const uint32_t* slot, const WinMeta* meta, uint32_t cap,
                                                   ulonglong2* set, uint32_t set_mask, unsigned long long* counts,
                                                   uint32_t n_genes, unsigned long long* set_size) {
    uint32_t n = meta->mm_count;
    if (n > cap) n = 0;  // overflowed: the host re-runs the window with a larger buffer
    uint32_t added = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const ulonglong2 key = fp[i];
        uint32_t p = (uint32_t)key.x & set_mask;
        for (;;) {
            ulonglong2 old = cas128(set + p, make_ulonglong2(0, 0), key);
            if (old.x == 0 && old.y == 0) {
                const uint32_t s = slot[i];
                atomicAdd(counts + (size_t)(s >> 31) * n_genes + (s & 0x7fffffffu), 1ull);
                added++;
                break;
            }
            if (old.x == key.x && old.y == key.y) break;
            p = (p + 1) & set_mask;
        }
    }
    added = __reduce_add_sync(0xffffffffu, added);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(set_size, (unsigned long long)added);
}

__global__ void __launch_bounds__(256) k_set_rehash(const ulonglong2* old_set, uint32_t old_cap, ulonglong2* set, uint32_t set_mask) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < old_cap; i += gridDim.x * blockDim.x) {
        const ulonglong2 key = old_set[i];
        if (key.x == 0 && key.y == 0) continue;
        uint32_t p = (uint32_t)key.x & set_mask;
        for (;;) {
            ulonglong2 old = cas128(set + p, make_ulonglong2(0, 0), key);
            if (old.x == 0 && old.y == 0) break;
            p = (p + 1) & set_mask;
        }
    }
}

__global__ void __launch_bounds__(256) k_map_rehash(const ulonglong2* old_map, uint32_t old_cap, ulonglong2* map, uint32_t mask) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < old_cap; i += gridDim.x * blockDim.x) {
        const ulonglong2 e = old_map[i];
        if (e.x == ~0ull && e.y == ~0ull) continue;
        map_add(map, mask, e.x, (uint32_t)e.y, (uint32_t)(e.y >> 32));
    }
}

// counts live (non-empty) slots; used to size result arrays
__global__ void __launch_bounds__(256) k_map_count(const ulonglong2* map, uint32_t cap, unsigned long long* n_out) {
    uint32_t c = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        const ulonglong2 e = map[i];
        if (!(e.x == ~0ull && e.y == ~0ull)) c++;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_out, (unsigned long long)c);
}

// K6: introns.  Every owned record adds each N gap (start, stop) of its contig and one entry to the
// introns-per-read histogram (introns.rs:38-48).  Fixed 1 Mb chunks over every header contig
// (chunked_genome.rs:32-43): chunk id = pos / 1e6, bounded by ceil(len / 1e6) (at least one chunk).
struct IntronCtx {
    const uint8_t* U;
    const uint32_t* rec_off;
    WinMeta* meta;
    int32_t n_ref;
    const uint32_t* ref_chunk_off;  // n_ref + 1: global chunk id base per tid
    uint32_t chunk_lo, chunk_hi;
    ulonglong2* map;
    uint32_t map_mask;
    unsigned long long* hist;       // [n_ref][HIST_DENSE]
    int count_only;                 // pass 1: only sum cigar ops of owned records (capacity bound)
};

__global__ void __launch_bounds__(256) k_count_introns(IntronCtx c) {
    const uint32_t n = c.meta->n_records;
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t my_owned = 0, my_ops = 0;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += stride) {
        const uint32_t i = base + lane;
        bool owned = false;
        RecView v;
        if (i < n) {
            bool ok = rec_view(c.U, c.rec_off[i], v);
            if (!ok) set_err(c.meta, WERR_RECORD, i);
            if (ok && v.tid >= 0 && v.tid < c.n_ref) {
                const uint32_t nchunks = c.ref_chunk_off[v.tid + 1] - c.ref_chunk_off[v.tid];
                const uint32_t ci = v.upos / 1000000u;
                if (ci < nchunks) {
                    const uint32_t g = c.ref_chunk_off[v.tid] + ci;
                    owned = g >= c.chunk_lo && g < c.chunk_hi;
                }
            }
        }
        uint32_t k_introns = 0;
        if (owned) {
            my_owned++;
            my_ops += v.n_cig;
            if (!c.count_only) {
                cigar_walk(v.cigar, v.n_cig, v.upos, [](uint32_t, uint32_t) { return true; },
                           [&](uint32_t s, uint32_t e) {
                               map_add(c.map, c.map_mask, (unsigned long long)s | ((unsigned long long)e << 32), (uint32_t)v.tid, 1);
                               k_introns++;
                           });
            }
        }
        if (c.count_only) continue;
        // histogram of introns per read, warp aggregated
        const unsigned long long hkey = owned ? ((unsigned long long)(uint32_t)v.tid << 32 | k_introns) : ~0ull;
        const unsigned peers = __match_any_sync(0xffffffffu, hkey);
        if (owned && lane == __ffs(peers) - 1) {
            const uint32_t cnt = __popc(peers);
            if (k_introns < (uint32_t)HIST_DENSE) atomicAdd(c.hist + (size_t)v.tid * HIST_DENSE + k_introns, (unsigned long long)cnt);
            else map_add(c.map, c.map_mask, 0xffffffffull | ((unsigned long long)(0xffffffffu - k_introns) << 32), (uint32_t)v.tid, cnt);
        }
    }
    my_owned = __reduce_add_sync(0xffffffffu, my_owned);
    my_ops = __reduce_add_sync(0xffffffffu, my_ops);
    if (lane == 0 && c.count_only) {
        if (my_owned) atomicAdd(&c.meta->owned_records, my_owned);
        if (my_ops) atomicAdd(&c.meta->cigar_ops, my_ops);
    }
}

// calculate_duplicate_distribution (reference src/duplicate_distribution.rs:44-86): how often exactly k reads
// share (reference, position, strand, sequence).  Every owned record adds 1 to the map entry of a 96-bit
// fingerprint of (tid, pos, is_reverse, l_seq, packed SEQ bytes); k_dup_hist then turns the entry counts into
// the histogram {k: number of distinct (position, strand, sequence) seen exactly k times}.  The reference walks
// each contig with fetch(tid, 0, len) and groups equal `pos`: records of one group are adjacent and never
// straddle a chunk (ownership is by pos), so chunk-range shards see whole groups.
struct DupCtx {
    const uint8_t* U;
    const uint32_t* rec_off;
    WinMeta* meta;
    int32_t n_ref;
    const uint32_t* ref_chunk_off;  // n_ref + 1: global chunk id base per tid (1 Mb chunks, like the intron job)
    const uint32_t* ref_len;        // n_ref: fetch(tid, 0, len) sees pos < len only
    uint32_t chunk_lo, chunk_hi;
    ulonglong2* map;
    uint32_t map_mask;
    unsigned long long* owned_total;  // job-wide count of owned records
};

__global__ void __launch_bounds__(256) k_dup_insert(DupCtx c) {
    const uint32_t n = c.meta->n_records;
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t my_owned = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        RecView v;
        if (!rec_view(c.U, c.rec_off[i], v)) { set_err(c.meta, WERR_RECORD, i); continue; }
        if (v.tid < 0 || v.tid >= c.n_ref) continue;
        if ((int32_t)v.upos < 0 || v.upos >= c.ref_len[v.tid]) continue;  // not returned by fetch(tid, 0, len)
        const uint32_t nchunks = c.ref_chunk_off[v.tid + 1] - c.ref_chunk_off[v.tid];
        const uint32_t ci = v.upos / 1000000u;
        if (ci >= nchunks) continue;
        const uint32_t g = c.ref_chunk_off[v.tid] + ci;
        if (g < c.chunk_lo || g >= c.chunk_hi) continue;
        my_owned++;
        const uint8_t* seq = v.cigar + 4ull * v.n_cig;
        const uint32_t nb = (v.l_seq + 1u) >> 1;
        // an odd l_seq leaves the low nibble of the last byte unspecified: mask it out of the key
        const ulonglong2 fp = qname_fingerprint(seq, nb - ((v.l_seq & 1u) ? 1u : 0u), (uint32_t)v.tid, v.upos);
        unsigned long long h1 = fp.x, h2 = fp.y;
        if (v.l_seq & 1u) h1 = fmix64(h1 ^ (unsigned long long)(seq[nb - 1] & 0xf0u) * 0x9E3779B97F4A7C15ull);
        h2 = fmix64(h2 + ((v.flag & 0x10u) ? 0x51ull : 0x17ull) + ((unsigned long long)v.l_seq << 8));
        uint32_t tag = (uint32_t)h2;
        if (h1 == ~0ull && tag == 0xffffffffu) tag = 0;  // the empty-slot pattern
        map_add(c.map, c.map_mask, h1, tag, 1);
    }
    my_owned = __reduce_add_sync(0xffffffffu, my_owned);
    if (lane == 0 && my_owned) atomicAdd(c.owned_total, (unsigned long long)my_owned);
}

// live slots of a map, densely packed (order arbitrary): what the host downloads instead of the whole table
__global__ void __launch_bounds__(256) k_map_compact(const ulonglong2* map, uint32_t cap, ulonglong2* out, unsigned long long out_cap,
                                                      unsigned long long* n_out) {
    const int lane = threadIdx.x & 31;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < cap; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        ulonglong2 e = make_ulonglong2(~0ull, ~0ull);
        if (i < cap) e = map[i];
        const bool live = !(e.x == ~0ull && e.y == ~0ull);
        const unsigned b = __ballot_sync(0xffffffffu, live);
        if (!b) continue;
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(n_out, (unsigned long long)__popc(b));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(b & ((1u << lane) - 1u));
        if (live && at < out_cap) out[at] = e;
    }
}

constexpr uint32_t DUP_DENSE = 1u << 16;  // multiplicities below this are counted in a dense array
__global__ void __launch_bounds__(256) k_dup_hist(const ulonglong2* map, uint32_t cap, unsigned long long* dense,
                                                   unsigned long long* big, uint32_t big_cap, uint32_t* n_big) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        const ulonglong2 e = map[i];
        if (e.x == ~0ull && e.y == ~0ull) continue;
        const uint32_t k = (uint32_t)(e.y >> 32);
        if (k < DUP_DENSE) {
            atomicAdd(dense + k, 1ull);
        } else {
            const uint32_t j = atomicAdd(n_big, 1u);
            if (j < big_cap) big[j] = k;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// count_reads_primary_only_right_strand_only_by_barcode (reference src/count_reads/by_barcode.rs:101-182).
// Per chunk the reference keeps {(gene_no, XC barcode, pos, is_reverse) -> set of XM UMIs} over the reads that are
// not secondary and lie on the gene's strand, then emits (gene_no, barcode, sum of the set sizes).  Here:
//   * every (gene, barcode, pos, is_reverse, UMI) becomes a 128-bit fingerprint in the job-wide set (`set`, the one
//     the multi-mapper keys use); `pos` fixes the chunk, so the set needs no chunk id
//   * a fingerprint seen for the first time adds 1 to the entry (chunk, gene, barcode) of the 16-byte-slot map
//     (keyed by a 96-bit fingerprint of the triple)
//   * a triple seen for the first time appends one BarcodeRec {fingerprint, gene, chunk, barcode bytes} to a list, from
//     which the host recovers what the map's fingerprints stand for
//   * a right-strand hit of a read without a string XC / XM makes the reference drop the whole chunk
//     (by_barcode.rs:46-68 `_ => {}`): chunk_err[chunk] = 1, the host drops the chunk's entries
// Two passes per window like the intron job: pass 1 counts the right-strand hits (the bound for all three
// structures), pass 2 inserts.
constexpr uint32_t BARCODE_MAX = 44;  // bytes of a barcode kept in a BarcodeRec; longer ones: WERR_BARCODE_LEN
struct BarcodeRec {
    unsigned long long key;   // map key
    uint32_t tag;             // map tag
    uint32_t gene;            // global gene slot
    uint32_t chunk;           // global chunk id of the job's chunk table
    uint32_t len;
    uint8_t bc[BARCODE_MAX];
};
static_assert(sizeof(BarcodeRec) == 72, "BarcodeRec layout");

// String aux tag (rust-htslib Aux::String: types Z and H).  1 found, 0 absent, -1 present with another type, -2 malformed.
__device__ __forceinline__ int aux_find_str(const uint8_t* p, const uint8_t* e, uint8_t t0, uint8_t t1, const uint8_t** sp,
                                            uint32_t* slen) {
    while (p + 3 <= e) {
        const bool match = p[0] == t0 && p[1] == t1;
        const uint8_t t = p[2];
        p += 3;
        uint32_t sz;
        switch (t) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'd': sz = 8; break;
            case 'Z': case 'H': {
                const uint8_t* s0 = p;
                while (p < e && *p) p++;
                if (p >= e) return -2;
                if (match) { *sp = s0; *slen = (uint32_t)(p - s0); return 1; }
                p++;
                continue;
            }
            case 'B': {
                if (match) return -1;
                if (p + 5 > e) return -2;
                const uint8_t st = p[0];
                const uint32_t cnt = ldu32(p + 1);
                const uint32_t es = (st == 'c' || st == 'C') ? 1u : (st == 's' || st == 'S') ? 2u : 4u;
                unsigned long long adv = 5ull + (unsigned long long)cnt * es;
                if (adv > (unsigned long long)(e - p)) return -2;
                p += adv;
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

// insert into the open-addressing set (empty = 0,0); true if the key was not there
__device__ __forceinline__ bool set_insert(ulonglong2* set, uint32_t mask, ulonglong2 key) {
    if (key.x == 0 && key.y == 0) key.y = 1;
    uint32_t p = (uint32_t)key.x & mask;
    for (;;) {
        const ulonglong2 old = cas128(set + p, make_ulonglong2(0, 0), key);
        if (old.x == 0 && old.y == 0) return true;
        if (old.x == key.x && old.y == key.y) return false;
        p = (p + 1) & mask;
    }
}

// map_add that tells whether it created the entry
__device__ __forceinline__ bool map_add_new(ulonglong2* map, uint32_t mask, unsigned long long key, uint32_t tag, uint32_t n) {
    unsigned long long h = fmix64(key ^ ((unsigned long long)tag * 0x9E3779B97F4A7C15ull));
    uint32_t i = (uint32_t)h & mask;
    for (;;) {
        ulonglong2 cur;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(cur.x), "=l"(cur.y) : "l"(map + i));
        if (cur.x == ~0ull && cur.y == ~0ull) {
            ulonglong2 mine = make_ulonglong2(key, (unsigned long long)tag | ((unsigned long long)n << 32));
            ulonglong2 old = cas128(map + i, make_ulonglong2(~0ull, ~0ull), mine);
            if (old.x == ~0ull && old.y == ~0ull) return true;
            cur = old;
        }
        if (cur.x == key && (uint32_t)cur.y == tag) {
            atomicAdd(((uint32_t*)(map + i)) + 3, n);
            return false;
        }
        i = (i + 1) & mask;
    }
}

struct BarcodeCtx {
    CountCtx c;                  // record source, interval tables, chunk table, owned chunk range
    int count_only;              // pass 1
    ulonglong2* set;
    uint32_t set_mask;
    unsigned long long* set_size;
    ulonglong2* map;
    uint32_t map_mask;
    BarcodeRec* list;
    unsigned long long* list_n;  // records appended so far (job wide)
    unsigned long long list_cap;
    uint32_t* chunk_err;         // per global chunk
    uint32_t* job_err;           // first WinErr of a pass 2 (the window's own error word is read before pass 2 runs)
};

__global__ void __launch_bounds__(256) k_barcode_hits(BarcodeCtx bc) {
    const CountCtx& c = bc.c;
    const uint32_t n = c.meta->n_records;
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t my_hits = 0, my_owned = 0, my_new = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        RecView v;
        if (!rec_view(c.U, c.rec_off[i], v)) { set_err(c.meta, WERR_RECORD, i); continue; }
        const int slot = (v.tid >= 0 && v.tid < c.n_ref) ? c.tid2slot[v.tid] : -1;
        if (slot < 0) continue;
        const uint32_t clo = c.slot_chunk_off[slot], chi = c.slot_chunk_off[slot + 1];
        const uint32_t ci = upper_bound_u32(c.chunk_stop, clo, chi, v.upos);  // first stop > pos
        if (!(ci < chi && ci >= c.chunk_lo && ci < c.chunk_hi)) continue;
        my_owned++;
        if (v.flag & 0x100u) continue;  // by_barcode.rs:127 is_secondary
        const uint32_t cstop = c.chunk_stop[ci];
        const uint32_t iv_lo = c.slot_iv_off[slot], iv_hi = c.slot_iv_off[slot + 1];
        const bool rev = (v.flag & 0x10u) != 0;
        int have = 0;  // 0: tags not looked up yet, 1: ok, -1: chunk failed
        const uint8_t *xc = nullptr, *xm = nullptr;
        uint32_t xc_len = 0, xm_len = 0;
        // by_barcode.rs:131: with pos >= chunk.start only `block start >= chunk.stop` can drop a block
        uint32_t no_hint = 0xffffffffu;
        enumerate_hits(c, v, iv_lo, iv_hi, true, cstop, no_hint, [&](uint32_t gw) {
            const bool plus = (gw >> 31) != 0;
            if (!((plus && !rev) || (!plus && rev))) return true;  // by_barcode.rs:139
            my_hits++;
            if (bc.count_only) return true;
            if (have == 0) {
                have = 1;
                const int r1 = aux_find_str(v.aux, v.end, 'X', 'C', &xc, &xc_len);
                const int r2 = r1 == 1 ? aux_find_str(v.aux, v.end, 'X', 'M', &xm, &xm_len) : 0;
                if (r1 == -2 || r2 == -2) atomicCAS(bc.job_err, 0u, (uint32_t)WERR_AUX);
                if (r1 != 1 || r2 != 1) have = -1;
                else if (xc_len > BARCODE_MAX) { atomicCAS(bc.job_err, 0u, (uint32_t)WERR_BARCODE_LEN); have = -1; }
            }
            if (have < 0) {
                bc.chunk_err[ci] = 1u;  // `?` leaves the chunk's function: the chunk yields nothing
                return false;
            }
            const uint32_t gene = gw & 0x7fffffffu;
            // (gene, pos, is_reverse, barcode, UMI)
            ulonglong2 f = qname_fingerprint(xc, xc_len, v.upos, gene | (rev ? 0x80000000u : 0u));
            const ulonglong2 fu = qname_fingerprint(xm, xm_len, xc_len, 0x5bd1e995u);
            f.x = fmix64(f.x ^ rotl64(fu.x, 17));
            f.y = fmix64(f.y + fu.y);
            if (!set_insert(bc.set, bc.set_mask, f)) return true;
            my_new++;
            // (chunk, gene, barcode)
            const ulonglong2 g = qname_fingerprint(xc, xc_len, ci, gene);
            unsigned long long key = g.x;
            uint32_t tag = (uint32_t)g.y;
            if (key == ~0ull && tag == 0xffffffffu) tag = 0;
            if (map_add_new(bc.map, bc.map_mask, key, tag, 1)) {
                const unsigned long long at = atomicAdd(bc.list_n, 1ull);
                if (at < bc.list_cap) {
                    BarcodeRec& r = bc.list[at];
                    r.key = key; r.tag = tag; r.gene = gene; r.chunk = ci; r.len = xc_len;
                    for (uint32_t k = 0; k < xc_len; k++) r.bc[k] = xc[k];
                }
            }
            return true;
        });
    }
    if (bc.count_only) {
        my_hits = __reduce_add_sync(0xffffffffu, my_hits);
        my_owned = __reduce_add_sync(0xffffffffu, my_owned);
        if (lane == 0) {
            if (my_hits) atomicAdd(&c.meta->cigar_ops, my_hits);
            if (my_owned) atomicAdd(&c.meta->owned_records, my_owned);
        }
    } else {
        my_new = __reduce_add_sync(0xffffffffu, my_new);
        if (lane == 0 && my_new) atomicAdd(bc.set_size, (unsigned long long)my_new);
    }
}


}  // namespace mbf
