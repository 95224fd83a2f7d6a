// Bit-level logic test of the shared DEFLATE core on the CPU (g++ only, no GPU):
// inflates every BGZF member of a file with mbf::inflate_member_host and compares with zlib.
// usage: host_emu file.bam   -> prints "OK <members> <bytes>" or a failure line, exit code 0/1
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <zlib.h>

#include "../../mbf_bam_b200/csrc/inflate_host.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> d(n + 16, 0);
    if (fread(d.data(), 1, n, f) != (size_t)n) return 2;
    fclose(f);
    std::vector<uint8_t> a(65536 + 512), b(65536 + 512);
    mbf::DecoderTables T;
    long off = 0, members = 0, bytes = 0;
    while (off + 18 <= n) {
        uint32_t xlen = d[off + 10] | (d[off + 11] << 8);
        uint32_t csize = (d[off + 16] | (d[off + 17] << 8)) + 1u;
        uint32_t isize;
        memcpy(&isize, &d[off + csize - 4], 4);
        const uint8_t* p = &d[off + 12 + xlen];
        uint32_t clen = csize - 12 - xlen - 8;
        int rc = mbf::inflate_member_host(p, clen, a.data(), isize, &T);
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        inflateInit2(&zs, -15);
        zs.next_in = (Bytef*)p; zs.avail_in = clen; zs.next_out = b.data(); zs.avail_out = 65536;
        int zrc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != 0 || zrc != Z_STREAM_END || zs.total_out != isize || memcmp(a.data(), b.data(), isize)) {
            printf("FAIL member at %ld rc=%d zrc=%d isize=%u\n", off, rc, zrc, isize);
            return 1;
        }
        off += csize; members++; bytes += isize;
    }
    printf("OK %ld %ld\n", members, bytes);
    return 0;
}
