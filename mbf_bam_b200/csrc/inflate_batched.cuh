// Table entry formats and the warp-cooperative table build shared by the K1 kernels (inflate_par.cuh,
// inflate_twopass.cuh).  (The one-pass "decode on lane 0, materialise on 32 lanes" kernel this header was written
// for is gone: superseded by the two-pass form, see DESIGN.md for its measurements.)
//
// Measured on the synthetic BAMs (profiles/r01_symbol_stats.txt): 40 % of the DEFLATE symbols are matches, they
// produce 86 % of the output bytes, and half of them reach back more than 2 KB.
//
// Replaces zlib's inflate under htslib's bgzf_read_block (reference src/count_reads/counters.rs:41).
#pragma once
#include "inflate_core.cuh"

namespace mbf {

constexpr int RINGB = 2048;
constexpr uint32_t RINGB_MASK = RINGB - 1;
constexpr uint32_t BATCH_MAX_OUT = 512;                 // output bytes per batch
constexpr uint32_t BQ_MATCHES = 32;                     // matches per batch: one per lane in the copy pass
constexpr uint32_t BATCH_NEAR = RINGB - BATCH_MAX_OUT;  // matches with dist <= this read the ring

// 32-bit LUT entry: [3:0] code length, [7:4] extra bits, [23:8] value, [31:30] kind
constexpr uint32_t K_LIT = 0u, K_LEN = 1u << 30, K_EOB = 2u << 30, K_SPECIAL = 3u << 30;
constexpr uint32_t E_INVALID = K_SPECIAL | 0xffu, E_SLOW = K_SPECIAL;

constexpr int LL_ROOTB = 9;  // 32-bit entries; literal codes of the BAM streams are <= 7 or >= 10 bits long
constexpr int D_ROOTB = 8;
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t bfe_u32(uint32_t v, uint32_t pos, uint32_t len) {
    uint32_t r;
    asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(len));
    return r;
}

__device__ __forceinline__ uint32_t ll_entry(uint32_t sym, uint32_t codelen) {
    if (sym < 256) return K_LIT | (sym << 8) | codelen;
    if (sym == 256) return K_EOB | codelen;
    sym -= 257;
    if (sym >= 29) return E_INVALID;
    return K_LEN | ((uint32_t)MBF_TBL(LEN_BASE, sym) << 8) | ((uint32_t)MBF_TBL(LEN_EXTRA, sym) << 4) | codelen;
}
__device__ __forceinline__ uint32_t d_entry(uint32_t sym, uint32_t codelen) {
    if (sym >= 30) return E_INVALID;
    return ((uint32_t)MBF_TBL(DIST_BASE, sym) << 8) | ((uint32_t)MBF_TBL(DIST_EXTRA, sym) << 4) | codelen;
}

// Warp-cooperative table build.  lens[0..n) in shared memory.  IS_DIST selects the entry maker.
// Returns 0 or an InflateError (uniform across the warp).
// second entry format (SIMT token decoder): [4:0] bits to consume (code + extra), [8:5] code length,
// [25:10] base value, [31:30] kind.  extra = (window & ~(~0 << consume)) >> code length.  The two
// special entries consume nothing.
constexpr uint32_t E2_SLOW = K_SPECIAL, E2_INVALID = K_SPECIAL | 0x100u;
__device__ __forceinline__ uint32_t ll_entry2(uint32_t sym, uint32_t codelen) {
    if (sym < 256) return K_LIT | (sym << 10) | (codelen << 5) | codelen;
    if (sym == 256) return K_EOB | (codelen << 5) | codelen;
    sym -= 257;
    if (sym >= 29) return E2_INVALID;
    return K_LEN | ((uint32_t)MBF_TBL(LEN_BASE, sym) << 10) | (codelen << 5) | (codelen + (uint32_t)MBF_TBL(LEN_EXTRA, sym));
}
__device__ __forceinline__ uint32_t d_entry2(uint32_t sym, uint32_t codelen) {
    if (sym >= 30) return E2_INVALID;
    return ((uint32_t)MBF_TBL(DIST_BASE, sym) << 10) | (codelen << 5) | (codelen + (uint32_t)MBF_TBL(DIST_EXTRA, sym));
}
template <bool IS_DIST, int FMT>
__device__ __forceinline__ uint32_t make_entry(uint32_t sym, uint32_t codelen) {
    if (FMT == 2) return IS_DIST ? d_entry2(sym, codelen) : ll_entry2(sym, codelen);
    return IS_DIST ? d_entry(sym, codelen) : ll_entry(sym, codelen);
}

template <bool IS_DIST, int FMT = 1>
__device__ __forceinline__ int build_lut_warp(const uint8_t* lens, int n, int root, uint32_t* lut, uint16_t* sym_sorted,
                                               HuffDec* hd, uint32_t* run, int lane) {
    if (lane < 16) { hd->cnt[lane] = 0; run[lane] = 0; }
    for (int i = lane; i < (1 << root); i += 32) lut[i] = FMT == 2 ? E2_INVALID : E_INVALID;
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const uint32_t l = lens[s];
        if (l) atomicAdd(&run[l], 1u);
    }
    __syncwarp();
    if (lane == 0) {
        int left = 1, max = 0, bad = 0;
        uint32_t code = 0, off = 0;
        for (int len = 1; len <= 15; len++) {
            const uint32_t c = run[len];
            hd->cnt[len] = (uint16_t)c;
            if (c) max = len;
            left = (left << 1) - (int)c;
            if (left < 0) bad = 1;
            hd->first[len] = (uint16_t)code;
            hd->offs[len] = (uint16_t)off;
            code = (code + c) << 1;
            off += c;
        }
        if (max == 0) bad = 0;                        // no codes at all: every lookup stays invalid
        else if (left > 0 && max != 1) bad = 1;        // incomplete (zlib allows one 1-bit code)
        hd->cnt[0] = (uint16_t)bad;
    }
    __syncwarp();
    if (hd->cnt[0]) return INF_BAD_CODES;
    if (lane < 16) run[lane] = 0;
    __syncwarp();
    // canonical order = (length, symbol): rank symbols chunk by chunk
    for (int base = 0; base < n; base += 32) {
        const int s = base + lane;
        const uint32_t l = s < n ? lens[s] : 0u;
        const unsigned peers = __match_any_sync(0xffffffffu, l);
        uint32_t rank = 0;
        if (l) rank = run[l] + __popc(peers & ((1u << lane) - 1));
        __syncwarp();
        if (l && lane == __ffs(peers) - 1) run[l] += __popc(peers);
        __syncwarp();
        if (l) {
            sym_sorted[hd->offs[l] + rank] = (uint16_t)s;
            const uint32_t code = hd->first[l] + rank;
            const uint32_t r = __brev(code) >> (32 - l);
            if (l <= (uint32_t)root) {
                const uint32_t e = make_entry<IS_DIST, FMT>((uint32_t)s, l);
                for (uint32_t k = r; k < (1u << root); k += (1u << l)) lut[k] = e;
            } else {
                lut[r & ((1u << root) - 1)] = FMT == 2 ? E2_SLOW : E_SLOW;
            }
        }
    }
    __syncwarp();
    return 0;
}

// codes longer than the LUT root (lane 0 only): canonical search, then the entry of that symbol
template <bool IS_DIST>
__device__ __forceinline__ uint32_t slow_entry(const HuffDec* hd, const uint16_t* sym, uint32_t bits15, int root) {
    const uint32_t code = __brev(bits15) >> 17;
    for (int len = root + 1; len <= 15; len++) {
        const int d = (int)(code >> (15 - len)) - (int)hd->first[len];
        if (d >= 0 && d < (int)hd->cnt[len]) {
            const uint32_t s = sym[hd->offs[len] + d];
            return IS_DIST ? d_entry(s, (uint32_t)len) : ll_entry(s, (uint32_t)len);
        }
    }
    return E_INVALID;
}

}  // namespace mbf
