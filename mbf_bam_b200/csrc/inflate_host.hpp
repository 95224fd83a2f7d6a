// Host-side driver of the shared DEFLATE core: used ONLY for the BAM header (a few KB at
// mbfbam_open, where htslib's bam_hdr_read also runs on the CPU) and by the bit-level logic
// test.  Alignment records are never inflated on the host.
#pragma once
#include <stdint.h>
#include <string.h>

#include "inflate_core.cuh"

namespace mbf {

struct LinearOut {
    uint8_t* base;
    uint32_t pos, cap;
    MBF_HD bool want_flush() const { return pos + 258 > cap; }  // treat "no room" as an error upstream
    MBF_HD void put(uint8_t b) { base[pos++] = b; }
    MBF_HD bool copy(uint32_t dist, uint32_t len) {
        if (dist > pos) return false;
        for (uint32_t i = 0; i < len; i++) base[pos + i] = base[pos - dist + i];
        pos += len;
        return true;
    }
};

// Inflate one raw DEFLATE stream (`n` bytes at `p`, 4 readable slack bytes after it) into
// out[0..isize).  `out` needs 258 bytes of slack.  Returns 0 or an InflateError.
inline int inflate_member_host(const uint8_t* p, uint32_t n, uint8_t* out, uint32_t isize,
                               DecoderTables* T) {
    BitReader br;
    br.init(p, n);
    uint32_t mis = (uint32_t)((uintptr_t)p & 3);
    LinearOut o{out, 0, isize + 258};
    for (;;) {
        br.refill();
        uint32_t bfinal = br.take(1), btype = br.take(2);
        if (btype == 0) {
            br.consume(br.cnt & 7);
            br.refill();
            uint32_t len = br.take(16);
            br.refill();
            uint32_t nlen = br.take(16);
            if ((len ^ 0xffffu) != nlen) return INF_BAD_STORED;
            if (o.pos + len > isize) return INF_OUTPUT_OVERRUN;
            for (uint32_t i = 0; i < len; i++) {
                br.refill();
                o.put((uint8_t)br.take(8));
            }
        } else if (btype == 1 || btype == 2) {
            int rc = btype == 1 ? build_fixed(T) : build_dynamic(T, br);
            if (rc) return rc;
            rc = decode_symbols(T, br, o);
            if (rc == DEC_NEED_FLUSH) return INF_OUTPUT_OVERRUN;
            if (rc < 0) return -rc;
        } else {
            return INF_BAD_BTYPE;
        }
        if (o.pos > isize) return INF_OUTPUT_OVERRUN;
        if (bfinal) break;
    }
    if (o.pos != isize) return INF_SIZE_MISMATCH;
    if (br.bytes_consumed(mis) > n) return INF_INPUT_OVERRUN;
    return 0;
}

}  // namespace mbf
