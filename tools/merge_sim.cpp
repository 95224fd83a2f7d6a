// Design study for the checkpoints of K1 pass 1 (CPU): where does the true token chain of a deflate block meet the path of a
// decoder that started at the first bit of a sub-sequence?  Reports, for sub-sequences of S bits: mean tokens per lane, how
// often the guess was exact, how often the paths never meet inside the sub-sequence, the distribution of the meeting point
// (in 32-bit words and in tokens), and the lock-step imbalance of a warp.
//   g++ -O2 -std=c++17 -o merge_sim merge_sim.cpp && ./merge_sim file.bam [members=500] [skip=10] [S=1312]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../mbf_bam_b200/csrc/inflate_core.cuh"
using namespace mbf;
struct Blk { const uint8_t* p; uint32_t clen; DecoderTables T; };
static uint32_t step(const Blk& b, uint32_t pos, bool* eob, uint32_t limit_bits) {
    *eob = false;
    if (pos >= limit_bits) return 0xffffffffu;
    BitReader br; br.init(b.p, b.clen + 8); br.seek(pos + ((uintptr_t)b.p & 3) * 8); br.refill();
    uint32_t e = b.T.ll_lut[br.peek(LL_ROOT)];
    if (e >= LUT_SLOW) { if (e != LUT_SLOW) return 0xffffffffu; e = slow_decode(&b.T.ll_hd, b.T.ll_sym, br.peek(15), LL_ROOT); if (!e) return 0xffffffffu; }
    uint32_t used = e & 15; br.consume(e & 15); uint32_t sym = e >> 4;
    if (sym < 256) return pos + used;
    if (sym == 256) { *eob = true; return pos + used; }
    sym -= 257; if (sym >= 29) return 0xffffffffu;
    used += H_LEN_EXTRA[sym]; br.take(H_LEN_EXTRA[sym]); br.refill();
    e = b.T.d_lut[br.peek(D_ROOT)];
    if (e >= LUT_SLOW) { if (e != LUT_SLOW) return 0xffffffffu; e = slow_decode(&b.T.d_hd, b.T.d_sym, br.peek(15), D_ROOT); if (!e) return 0xffffffffu; }
    uint32_t ds = e >> 4; if (ds >= 30) return 0xffffffffu;
    used += (e & 15) + H_DIST_EXTRA[ds];
    return pos + used;
}
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); if (!f) return 1;
    long maxm = argc > 2 ? atol(argv[2]) : 500, skip = argc > 3 ? atol(argv[3]) : 10;
    const uint32_t S = argc > 4 ? atoi(argv[4]) : 1312;
    long n = 200 << 20; std::vector<uint8_t> d(n + 64); n = fread(d.data(), 1, n, f);
    long off = 0, nmem = 0; static Blk b;
    std::vector<long> hist_bits(64, 0), hist_tok(64, 0); long nl = 0, nomerge = 0, exact = 0; double sum_tok = 0, sum_bits = 0, tot_tok_lane = 0;
    double max_tok_sum = 0, mean_tok_sum = 0; long nwarps = 0;
    while (off + 18 <= n && nmem < maxm + skip) {
        uint32_t xlen = d[off + 10] | (d[off + 11] << 8); uint32_t csize = (d[off + 16] | (d[off + 17] << 8)) + 1u;
        const uint8_t* p = &d[off + 12 + xlen]; uint32_t clen = csize - 12 - xlen - 8; off += csize; nmem++;
        if (nmem <= skip || clen < 8) continue;
        b.p = p; b.clen = clen; BitReader br; br.init(p, clen); const uint32_t mis8 = ((uintptr_t)p & 3) * 8;
        br.refill(); uint32_t bf = br.take(1), bt = br.take(2); (void)bf;
        if (bt != 2) continue;
        if (build_dynamic(&b.T, br)) return 1;
        uint32_t data_start = br.bit_position() - mis8; const uint32_t limit_bits = clen * 8 + 64;
        std::vector<uint8_t> T(limit_bits + 64, 0); std::vector<uint32_t> chain;
        uint32_t pos = data_start;
        for (;;) { bool eob; T[pos] = 1; chain.push_back(pos); uint32_t nx = step(b, pos, &eob, limit_bits); if (nx == 0xffffffffu) return 2; pos = nx; if (eob) break; }
        uint32_t data_end = pos; uint32_t base = (data_start + mis8) / 32 * 32 - mis8;  // word aligned in the GPU's frame
        std::vector<uint32_t> lane_tok;
        for (uint32_t i = 1;; i++) {
            uint32_t ni = base + i * S, ei = ni + S; if (ni >= data_end) break;
            // guess path marks
            std::vector<uint8_t> G(S + 64, 0); uint32_t q = ni; 
            while (q < ei) { bool eob; G[q - ni] = 1; uint32_t nx = step(b, q, &eob, limit_bits); if (nx == 0xffffffffu || eob) break; q = nx; }
            // true path entry
            auto it = std::lower_bound(chain.begin(), chain.end(), ni); uint32_t k = it - chain.begin(); uint32_t tk = 0; bool merged = false; uint32_t cnt_in = 0;
            for (uint32_t j = k; j < chain.size() && chain[j] < ei; j++) cnt_in++;
            lane_tok.push_back(cnt_in);
            for (uint32_t j = k; j < chain.size() && chain[j] < ei; j++, tk++) {
                if (G[chain[j] - ni]) { merged = true; uint32_t bits = chain[j] - ni; hist_bits[std::min<uint32_t>(63, bits / 32)]++; hist_tok[std::min<uint32_t>(63, tk / 4)]++; sum_tok += tk; sum_bits += bits; if (tk == 0) exact++; break; }
            }
            if (!merged) { nomerge++; sum_tok += cnt_in; }
            tot_tok_lane += cnt_in; nl++;
        }
        for (size_t w = 0; w + 32 <= lane_tok.size(); w += 32) { uint32_t mx = 0, sm = 0; for (int j = 0; j < 32; j++) { mx = std::max(mx, lane_tok[w + j]); sm += lane_tok[w + j]; } max_tok_sum += mx; mean_tok_sum += sm / 32.0; nwarps++; }
    }
    printf("S=%u lanes %ld  mean tokens/lane %.1f  exact-start %.3f  no-merge %.4f  mean sync tokens %.1f  mean merge bits %.0f\n", S, nl, tot_tok_lane / nl, (double)exact / nl, (double)nomerge / nl, sum_tok / nl, sum_bits / std::max(1L, nl - nomerge));
    printf("warp max/mean tokens per lane: %.2f\n", max_tok_sum / mean_tok_sum);
    long c = 0; printf("merge position (words from sub-sequence start) cumulative:"); for (int i = 0; i < 42; i++) { c += hist_bits[i]; printf(" %d:%.3f", i, (double)c / nl); } printf("\n");
    c = 0; printf("sync tokens (x4) cumulative:"); for (int i = 0; i < 30; i++) { c += hist_tok[i]; printf(" %d:%.3f", i * 4, (double)c / nl); } printf("\n");
}
