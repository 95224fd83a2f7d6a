// Design study for pass 1 of K1 (CPU, no GPU needed): how well does a DEFLATE token stream
// self-synchronise when 32 lanes start decoding one deflate block at guessed bit positions?
//
// For every dynamic/fixed Huffman block of the first N BGZF members of a file the tool decodes the true
// token chain, then emulates the lock-step scheme: the data bits are cut into sub-sequences of S bits,
// lane i starts at the first bit of sub-sequence i (lane 0 at the true start), decodes until it crosses
// the end of its sub-sequence and publishes the bit position where it stopped; next round every lane
// whose predecessor stopped somewhere else restarts there.  Converged <=> every lane started where its
// predecessor stopped <=> the chain is the true one.  Reported: rounds until convergence and the lock-step
// cost in symbol trips (sum over rounds of the slowest lane) against the serial trip count.
//
//   selfsync_sim file.bam [members=2000] [skip=10]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../mbf_bam_b200/csrc/inflate_core.cuh"
using namespace mbf;

struct Blk {
    const uint8_t* p;
    uint32_t clen;
    DecoderTables T;
};

// decode one token at bit position pos (litlen state); returns new position, 0xffffffff if invalid, sets *eob
static uint32_t step(const Blk& b, uint32_t pos, bool* eob, uint32_t limit_bits) {
    *eob = false;
    if (pos >= limit_bits) return 0xffffffffu;
    BitReader br;
    br.init(b.p, b.clen + 8);
    br.seek(pos + ((uintptr_t)b.p & 3) * 8);
    br.refill();
    uint32_t e = b.T.ll_lut[br.peek(LL_ROOT)];
    if (e >= LUT_SLOW) {
        if (e != LUT_SLOW) return 0xffffffffu;
        e = slow_decode(&b.T.ll_hd, b.T.ll_sym, br.peek(15), LL_ROOT);
        if (!e) return 0xffffffffu;
    }
    uint32_t used = e & 15;
    br.consume(e & 15);
    uint32_t sym = e >> 4;
    if (sym < 256) return pos + used;
    if (sym == 256) { *eob = true; return pos + used; }
    sym -= 257;
    if (sym >= 29) return 0xffffffffu;
    used += H_LEN_EXTRA[sym];
    br.take(H_LEN_EXTRA[sym]);
    br.refill();
    e = b.T.d_lut[br.peek(D_ROOT)];
    if (e >= LUT_SLOW) {
        if (e != LUT_SLOW) return 0xffffffffu;
        e = slow_decode(&b.T.d_hd, b.T.d_sym, br.peek(15), D_ROOT);
        if (!e) return 0xffffffffu;
    }
    uint32_t ds = e >> 4;
    if (ds >= 30) return 0xffffffffu;
    used += (e & 15) + H_DIST_EXTRA[ds];
    return pos + used;
}

struct LaneRes {
    uint32_t exit;   // bit position where the lane stopped (>= end of its sub-sequence), or the EOB end, or invalid
    uint32_t trips;
    bool eob, invalid;
};

static LaneRes run_lane(const Blk& b, uint32_t start, uint32_t end, uint32_t limit_bits) {
    LaneRes r{start, 0, false, false};
    uint32_t pos = start;
    while (pos < end) {
        bool eob;
        uint32_t nx = step(b, pos, &eob, limit_bits);
        r.trips++;
        if (nx == 0xffffffffu) { r.invalid = true; break; }
        pos = nx;
        if (eob) { r.eob = true; break; }
    }
    r.exit = pos;
    return r;
}

int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> d(n + 64);
    if (fread(d.data(), 1, n, f) != (size_t)n) return 1;
    long maxm = argc > 2 ? atol(argv[2]) : 2000, skip = argc > 3 ? atol(argv[3]) : 10;
    const int SS[] = {128, 256, 512, 1024, 2048, 4096};
    const int NS = 6;
    double serial_trips = 0;
    double par_trips[NS] = {0}, rounds_sum[NS] = {0}, waves[NS] = {0}, final_trips[NS] = {0};
    long rounds_hist[NS][12] = {{0}};
    long nblk = 0, nmem = 0, ntok_total = 0, data_bits_total = 0, hdr_bits_total = 0;
    long off = 0;
    static Blk b;
    std::vector<uint8_t> out(70000 + 512);
    while (off + 18 <= n && nmem < maxm + skip) {
        uint32_t xlen = d[off + 10] | (d[off + 11] << 8);
        uint32_t csize = (d[off + 16] | (d[off + 17] << 8)) + 1u;
        const uint8_t* p = &d[off + 12 + xlen];
        uint32_t clen = csize - 12 - xlen - 8;
        off += csize;
        nmem++;
        if (nmem <= skip || clen < 8) continue;
        b.p = p;
        b.clen = clen;
        BitReader br;
        br.init(p, clen);
        const uint32_t mis8 = ((uintptr_t)p & 3) * 8;
        for (;;) {
            br.refill();
            uint32_t hdr_start = br.bit_position() - mis8;
            uint32_t bf = br.take(1), bt = br.take(2);
            if (bt == 0) {
                br.consume(br.cnt & 7);
                br.refill();
                uint32_t len = br.take(16);
                br.refill();
                br.take(16);
                for (uint32_t i = 0; i < len; i++) { br.refill(); br.take(8); }
            } else {
                if (bt == 1) build_fixed(&b.T); else if (build_dynamic(&b.T, br)) { fprintf(stderr, "bad header\n"); return 1; }
                uint32_t data_start = br.bit_position() - mis8;
                hdr_bits_total += data_start - hdr_start;
                // true chain
                uint32_t pos = data_start, ntok = 0;
                const uint32_t limit_bits = clen * 8;
                for (;;) {
                    bool eob;
                    uint32_t nx = step(b, pos, &eob, limit_bits + 64);
                    if (nx == 0xffffffffu) { fprintf(stderr, "true chain invalid?\n"); return 1; }
                    ntok++;
                    pos = nx;
                    if (eob) break;
                }
                const uint32_t data_end = pos;
                serial_trips += ntok;
                ntok_total += ntok;
                data_bits_total += data_end - data_start;
                nblk++;
                for (int si = 0; si < NS; si++) {
                    const uint32_t S = SS[si];
                    // waves of 32 sub-sequences; wave w starts at the true position reached by wave w-1
                    uint32_t wave_start = data_start;
                    bool done = false;
                    while (!done) {
                        waves[si]++;
                        uint32_t st[33];
                        LaneRes res[32];
                        for (int i = 0; i < 32; i++) st[i] = wave_start + i * S;
                        auto sub_end = [&](int i) { return wave_start + (i + 1) * S; };
                        // round 0
                        uint32_t mx = 0;
                        for (int i = 0; i < 32; i++) {
                            res[i] = run_lane(b, st[i], sub_end(i), limit_bits + 64);
                            mx = std::max(mx, res[i].trips);
                        }
                        par_trips[si] += mx;
                        int rounds = 1;
                        for (;;) {
                            bool changed = false;
                            mx = 0;
                            // lanes after a (currently believed) EOB or invalid predecessor are idle
                            std::vector<int> todo;
                            for (int i = 1; i < 32; i++) {
                                if (res[i - 1].eob || res[i - 1].invalid) { continue; }
                                if (res[i - 1].exit != st[i]) todo.push_back(i);
                            }
                            // all restarts of a round use the predecessors' results of the previous round
                            std::vector<LaneRes> nr;
                            for (int i : todo) nr.push_back(run_lane(b, res[i - 1].exit, sub_end(i), limit_bits + 64));
                            for (size_t k = 0; k < todo.size(); k++) {
                                int i = todo[k];
                                st[i] = res[i - 1].exit;
                                mx = std::max(mx, nr[k].trips);
                                changed = true;
                            }
                            for (size_t k = 0; k < todo.size(); k++) res[todo[k]] = nr[k];
                            if (!changed) break;
                            par_trips[si] += mx;
                            rounds++;
                        }
                        rounds_sum[si] += rounds;
                        rounds_hist[si][std::min(rounds, 11)]++;
                        // final (write) pass: every lane decodes its sub-sequence once more
                        mx = 0;
                        int last = 31;
                        for (int i = 0; i < 32; i++) {
                            mx = std::max(mx, res[i].trips);
                            if (res[i].eob) { done = true; last = i; break; }
                            if (res[i].invalid) { fprintf(stderr, "true chain invalid in sim?\n"); return 1; }
                        }
                        final_trips[si] += mx;
                        if (done) {
                            if (res[last].exit != data_end) { fprintf(stderr, "sim end mismatch %u %u\n", res[last].exit, data_end); return 1; }
                        } else {
                            wave_start = res[31].exit;
                            // keep sub-sequence boundaries aligned to the grid of the first wave? not needed
                        }
                    }
                }
                // continue the reader after the block
                br.seek(data_end + mis8);
            }
            if (bf) break;
        }
    }
    printf("members %ld, huffman blocks %ld, tokens %ld (%.0f per block), data bits/token %.2f, header bits/block %.0f\n", nmem - skip, nblk,
           ntok_total, (double)ntok_total / nblk, (double)data_bits_total / ntok_total, (double)hdr_bits_total / nblk);
    for (int si = 0; si < NS; si++) {
        printf("S=%4d bits: waves/block %.1f, rounds/wave %.2f, lockstep trips: sync %.3f + write %.3f of serial (speed-up %.1fx)  rounds hist:",
               SS[si], waves[si] / nblk, rounds_sum[si] / waves[si], par_trips[si] / serial_trips, final_trips[si] / serial_trips,
               serial_trips / (par_trips[si] + final_trips[si]));
        for (int r = 1; r < 12; r++) printf(" %d:%.3f", r, rounds_hist[si][r] / waves[si]);
        printf("\n");
    }
    return 0;
}
