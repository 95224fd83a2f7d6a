// Host-side control plane: BAM header + BAI parsing, the ChunkedGenome table, flattened interval
// tables and the BGZF byte ranges a job has to feed to the GPU.  No alignment record is ever parsed
// here; that is the kernels' job.
//
// Reference behaviour replaced:
//   open_bam                       src/bam_ext.rs:6-21 (IndexedReader::from_path[_and_index])
//   header tid/target_len/names    src/count_reads/chunked_genome.rs:24,36-40,81-82
//   ChunkedGenome                  src/count_reads/chunked_genome.rs:17-43,72-119
//   build_tree                     src/count_reads/mod.rs:27-45 (flattened instead of an AVL tree)
#pragma once
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>

#include <list>
#include <memory>
#include <mutex>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mbfbam.h"
#include "inflate_host.hpp"

namespace mbf {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

constexpr uint32_t CHUNK_SIZE = 1000000u;  // chunked_genome.rs:75

struct RefIndex {
    bool has_reads = false;
    uint64_t voff_beg = 0, voff_end = 0;           // over all real bins
    std::vector<uint64_t> ioffset;                 // 16 kb linear index
    // bin -> (min beg, max end).  Filled at open only when the index lacks the samtools metadata pseudo-bin;
    // otherwise on first use (voff_upper of a range that ends inside the contig), from the mapped index.
    mutable std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> bins;
    mutable bool bins_parsed = false;
    const uint8_t* bins_raw = nullptr;             // first bin record of this reference in the mapped .bai
    size_t bins_raw_len = 0;
    int32_t n_bin = 0;
};

// the mapped .bai (kept for the life of the reader: a 50 M-read file has a 10 MB index, and walking all of
// its chunk lists on every open costs more than everything else the open does)
struct BaiMap {
    const uint8_t* p = nullptr;
    size_t n = 0;
    int fd = -1;
    BaiMap() = default;
    BaiMap(const BaiMap&) = delete;
    BaiMap& operator=(const BaiMap&) = delete;
    ~BaiMap() {
        if (p && n) munmap((void*)p, n);
        if (fd >= 0) ::close(fd);
    }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
};

// Everything parsed from one index file.  Shared (through IndexCache) by every Reader that opens the same,
// unchanged file: a 400 M-read BAM has a 40 MB index, and walking it at every open costs ~10 ms -- more than the
// rest of the host side of a call.  The only thing that changes after parsing is the lazily filled RefIndex::bins,
// which `mu` guards (one call fans out over several devices, each planning its shard on its own thread).
struct IndexData {
    std::vector<RefIndex> idx;
    BaiMap bai;
    int min_shift = 14, depth = 5;  // binning scheme: BAI is fixed at 14 / 5, a CSI header says (SAM spec 5.3 / CSIv1)
    bool is_csi = false;
    std::mutex mu;
};

class IndexCache {
    std::mutex mu;
    std::list<std::pair<std::string, std::shared_ptr<IndexData>>> items;

public:
    static IndexCache& get() {
        static IndexCache c;
        return c;
    }
    std::shared_ptr<IndexData> find(const std::string& key) {
        std::lock_guard<std::mutex> lk(mu);
        for (auto it = items.begin(); it != items.end(); ++it)
            if (it->first == key) {
                items.splice(items.begin(), items, it);
                return items.front().second;
            }
        return nullptr;
    }
    void put(const std::string& key, std::shared_ptr<IndexData> v) {
        std::lock_guard<std::mutex> lk(mu);
        items.emplace_front(key, std::move(v));
        while (items.size() > 8) items.pop_back();
    }
    void clear() {
        std::lock_guard<std::mutex> lk(mu);
        items.clear();
    }
};

struct Reader {
    std::string path;
    int fd = -1;
    uint64_t file_size = 0;
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    std::map<std::string, int32_t> name2tid;
    uint64_t first_rec_voff = 0;  // virtual offset of the first alignment record
    std::shared_ptr<IndexData> ix;  // the parsed index (shared between readers of the same file)
    // optional device-resident copy of the whole file (mbfbam_preload)
    int preload_device = -1;
    uint8_t* preload_ptr = nullptr;

    ~Reader() {
        if (fd >= 0) ::close(fd);
    }

    void pread_exact(void* dst, uint64_t off, size_t n) const {
        size_t got = 0;
        while (got < n) {
            ssize_t r = ::pread(fd, (char*)dst + got, n - got, (off_t)(off + got));
            if (r <= 0) throw Error(MBFBAM_ERR_IO, "Could not read bam: short read on " + path);
            got += (size_t)r;
        }
    }

    void open(const char* bam, const char* bai) {
        path = bam;
        fd = ::open(bam, O_RDONLY);
        if (fd < 0) throw Error(MBFBAM_ERR_IO, std::string("Could not read bam: cannot open ") + bam);
        struct stat st;
        fstat(fd, &st);
        file_size = (uint64_t)st.st_size;
        parse_header();
        std::string bp;
        if (bai) {
            bp = bai;
        } else {
            // htslib's sam_index_load order: x.bam.csi, x.csi, then x.bam.bai, x.bai
            const size_t n = path.size();
            const bool dotbam = n > 4 && path.compare(n - 4, 4, ".bam") == 0;
            std::vector<std::string> cands{path + ".csi"};
            if (dotbam) cands.push_back(path.substr(0, n - 4) + ".csi");
            cands.push_back(path + ".bai");
            if (dotbam) cands.push_back(path.substr(0, n - 4) + ".bai");
            bp = path + ".bai";
            for (const auto& c : cands)
                if (access(c.c_str(), R_OK) == 0) { bp = c; break; }
        }
        parse_index(bp);
    }

    // the index file decides its own format: "BAI\1" plain, or a BGZF stream that inflates to "CSI\1"
    void parse_index(const std::string& bp) {
        uint8_t m[4] = {0, 0, 0, 0};
        int f = ::open(bp.c_str(), O_RDONLY);
        if (f < 0) throw Error(MBFBAM_ERR_IO, "Could not read bam: no index " + bp);
        struct stat sb;
        if (fstat(f, &sb) != 0) { ::close(f); throw Error(MBFBAM_ERR_IO, "Could not read bam: cannot stat index " + bp); }
        ssize_t got = ::read(f, m, 4);
        ::close(f);
        // the same index file (name, inode, size, time stamp) for a header with the same targets: parsed before
        uint64_t hl = names.size();
        for (uint32_t l : lens) hl = hl * 1099511628211ull + l;
        char kb[160];
        snprintf(kb, sizeof kb, "|%llu|%llu|%lld.%09ld|%llu", (unsigned long long)sb.st_ino, (unsigned long long)sb.st_size,
                 (long long)sb.st_mtim.tv_sec, (long)sb.st_mtim.tv_nsec, (unsigned long long)hl);
        const std::string key = bp + kb;
        if ((ix = IndexCache::get().find(key))) return;
        ix = std::make_shared<IndexData>();
        if (got == 4 && m[0] == 0x1f && m[1] == 0x8b) parse_csi(bp);
        else parse_bai(bp);
        IndexCache::get().put(key, ix);
    }

    // CSI (coordinate-sorted index, htslib): BGZF-compressed; per bin a `loffset` (what the BAI linear index holds for
    // the bin's first window) instead of a separate linear index; min_shift / depth in the header.
    void parse_csi(const std::string& bp) {
        std::vector<uint8_t> u;
        {
            std::vector<uint8_t> raw;
            {
                FILE* fh = fopen(bp.c_str(), "rb");
                if (!fh) throw Error(MBFBAM_ERR_IO, "Could not read bam: no index " + bp);
                uint8_t buf[1 << 16];
                size_t k;
                while ((k = fread(buf, 1, sizeof buf, fh)) > 0) raw.insert(raw.end(), buf, buf + k);
                fclose(fh);
            }
            const uint8_t* d = raw.data();
            const size_t n = raw.size();
            std::vector<uint8_t> cb(65536 + 64), ob(65536 + 512);
            DecoderTables T;
            for (size_t off = 0; off + 28 <= n;) {
                if (d[off] != 0x1f || d[off + 1] != 0x8b || d[off + 2] != 8 || !(d[off + 3] & 4))
                    throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: index is not BGZF: " + bp);
                const uint32_t xlen = d[off + 10] | (d[off + 11] << 8);
                int32_t bsize = -1;
                for (uint32_t q = 0; q + 4 <= xlen && off + 12 + q + 6 <= n;) {
                    const uint8_t* e = d + off + 12 + q;
                    const uint32_t slen = e[2] | (e[3] << 8);
                    if (e[0] == 66 && e[1] == 67 && slen == 2) bsize = e[4] | (e[5] << 8);
                    q += 4 + slen;
                }
                if (bsize < 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: BGZF block without BC field: " + bp);
                const uint32_t csize = (uint32_t)bsize + 1, hdr = 12 + xlen;
                if (csize < hdr + 8 || off + csize > n) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: truncated index " + bp);
                const uint32_t clen = csize - hdr - 8;
                uint32_t isize;
                memcpy(&isize, d + off + csize - 4, 4);
                if (isize > 65536) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: bad ISIZE in " + bp);
                std::fill(cb.begin(), cb.end(), 0);
                memcpy(cb.data(), d + off + hdr, clen);
                if (isize && inflate_member_host(cb.data(), clen, ob.data(), isize, &T))
                    throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: corrupt index " + bp);
                u.insert(u.end(), ob.begin(), ob.begin() + isize);
                off += csize;
            }
        }
        size_t p = 0;
        auto take = [&](void* dst, size_t n) {
            if (p + n > u.size()) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: truncated index " + bp);
            memcpy(dst, u.data() + p, n);
            p += n;
        };
        char magic[4];
        take(magic, 4);
        if (memcmp(magic, "CSI\1", 4) != 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: not a CSI index: " + bp);
        int32_t min_shift, depth, l_aux, n_ref;
        take(&min_shift, 4);
        take(&depth, 4);
        take(&l_aux, 4);
        if (min_shift < 1 || min_shift > 29 || depth < 1 || depth > 9 || min_shift + 3 * depth > 32 || l_aux < 0)
            throw Error(MBFBAM_ERR_UNSUPPORTED, "Could not read bam: CSI binning scheme not supported: " + bp);
        if (p + (size_t)l_aux > u.size()) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: truncated index " + bp);
        p += (size_t)l_aux;
        take(&n_ref, 4);
        ix->is_csi = true;
        ix->min_shift = min_shift;
        ix->depth = depth;
        const uint32_t leaf0 = ((1u << (3 * depth)) - 1u) / 7u;        // first bin of the deepest level
        const uint32_t meta_bin = ((1u << (3 * depth + 3)) - 1u) / 7u + 1u;  // 37450 for depth 5
        ix->idx.assign(names.size(), RefIndex());
        for (int32_t r = 0; r < n_ref; r++) {
            RefIndex ri;
            take(&ri.n_bin, 4);
            if (ri.n_bin < 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: bad index " + bp);
            const uint32_t len = (size_t)r < lens.size() ? lens[(size_t)r] : 0;
            ri.ioffset.assign(((size_t)len >> min_shift) + 1, 0);
            bool meta = false;
            uint64_t mn = UINT64_MAX, mx = 0;
            for (int32_t i = 0; i < ri.n_bin; i++) {
                uint32_t bin;
                uint64_t loff;
                int32_t n_chunk;
                take(&bin, 4);
                take(&loff, 8);
                take(&n_chunk, 4);
                if (n_chunk < 0 || p + 16 * (size_t)n_chunk > u.size()) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: truncated index " + bp);
                if (bin == meta_bin) {
                    if (n_chunk >= 1) {
                        uint64_t be[2];
                        memcpy(be, u.data() + p, 16);
                        ri.voff_beg = be[0];
                        ri.voff_end = be[1];
                        meta = true;
                    }
                } else if (n_chunk > 0) {
                    uint64_t bmn = UINT64_MAX, bmx = 0;
                    for (int32_t c = 0; c < n_chunk; c++) {
                        uint64_t be[2];
                        memcpy(be, u.data() + p + 16 * (size_t)c, 16);
                        bmn = std::min(bmn, be[0]);
                        bmx = std::max(bmx, be[1]);
                    }
                    ri.bins.push_back({bin, {bmn, bmx}});
                    mn = std::min(mn, bmn);
                    mx = std::max(mx, bmx);
                    // loffset = smallest offset of a record overlapping the bin's FIRST window: the linear index entry
                    // of that window.  Any level gives one (a lower bound is all voff_lower needs).
                    int l = depth;
                    uint32_t lo = leaf0;
                    while (l > 0 && bin < lo) { l--; lo = ((1u << (3 * l)) - 1u) / 7u; }
                    const uint64_t w = (uint64_t)(bin - lo) << (3 * (depth - l));
                    if (w < ri.ioffset.size() && loff > ri.ioffset[(size_t)w]) ri.ioffset[(size_t)w] = loff;
                }
                p += 16 * (size_t)n_chunk;
            }
            // offsets never decrease with the window (the file is sorted): windows without an entry inherit the last one
            for (size_t w = 1; w < ri.ioffset.size(); w++) ri.ioffset[w] = std::max(ri.ioffset[w], ri.ioffset[w - 1]);
            ri.bins_parsed = true;
            ri.has_reads = mn != UINT64_MAX;
            if (!(meta && ri.has_reads && ri.voff_end > ri.voff_beg)) {
                ri.voff_beg = ri.has_reads ? mn : 0;
                ri.voff_end = mx;
            }
            if ((size_t)r < ix->idx.size()) ix->idx[(size_t)r] = std::move(ri);
        }
    }

    // BAM header: inflated on the host with the shared DEFLATE core (a few KB, once per open)
    void parse_header() {
        std::vector<uint8_t> u;
        std::vector<std::pair<uint64_t, uint32_t>> blocks;  // (coffset, uncompressed start)
        uint64_t coff = 0;
        std::vector<uint8_t> cb(65536 + 64), ob(65536 + 512);
        DecoderTables T;
        auto need = [&](size_t n) {
            while (u.size() < n) {
                if (coff + 28 > file_size) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: truncated header in " + path);
                uint8_t h[18];
                pread_exact(h, coff, 18);
                if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4))
                    throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: not a BGZF file: " + path);
                uint32_t xlen = h[10] | (h[11] << 8);
                std::vector<uint8_t> extra(xlen);
                pread_exact(extra.data(), coff + 12, xlen);
                int32_t bsize = -1;
                for (uint32_t p = 0; p + 4 <= xlen;) {
                    uint32_t slen = extra[p + 2] | (extra[p + 3] << 8);
                    if (extra[p] == 66 && extra[p + 1] == 67 && slen == 2 && p + 6 <= xlen) bsize = extra[p + 4] | (extra[p + 5] << 8);
                    p += 4 + slen;
                }
                if (bsize < 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: BGZF block without BC field: " + path);
                uint32_t csize = (uint32_t)bsize + 1, hdr = 12 + xlen;
                if (csize < hdr + 8 || coff + csize > file_size) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: truncated BGZF block: " + path);
                std::fill(cb.begin(), cb.end(), 0);
                pread_exact(cb.data(), coff + hdr, csize - hdr);
                uint32_t clen = csize - hdr - 8, isize;
                memcpy(&isize, cb.data() + clen + 4, 4);
                if (isize > 65536) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: bad ISIZE: " + path);
                int rc = inflate_member_host(cb.data(), clen, ob.data(), isize, &T);
                if (rc) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: corrupt BGZF block in header of " + path);
                blocks.push_back({coff, (uint32_t)u.size()});
                u.insert(u.end(), ob.begin(), ob.begin() + isize);
                coff += csize;
            }
        };
        auto rd32 = [&](size_t p) { need(p + 4); int32_t v; memcpy(&v, u.data() + p, 4); return v; };
        need(4);
        if (memcmp(u.data(), "BAM\1", 4) != 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: bad magic in " + path);
        int32_t l_text = rd32(4);
        if (l_text < 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: bad header");
        size_t p = 8 + (size_t)l_text;
        int32_t n_ref = rd32(p);
        p += 4;
        if (n_ref < 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: bad header");
        for (int32_t i = 0; i < n_ref; i++) {
            int32_t ln = rd32(p);
            if (ln < 1) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: bad reference name");
            need(p + 4 + (size_t)ln + 4);
            names.emplace_back((const char*)u.data() + p + 4, (size_t)ln - 1);
            p += 4 + (size_t)ln;
            lens.push_back((uint32_t)rd32(p));
            p += 4;
            name2tid.emplace(names.back(), i);  // first wins, like htslib's name hash
        }
        // p = uncompressed offset of the first record -> virtual offset
        size_t bi = blocks.size() - 1;
        while (blocks[bi].second > p) bi--;
        if (p == u.size()) {
            first_rec_voff = coff << 16;  // header ends exactly at a block end
        } else {
            first_rec_voff = (blocks[bi].first << 16) | (uint64_t)(p - blocks[bi].second);
        }
    }

    // (min chunk begin, max chunk end) of every real bin of one reference, from its raw records
    void parse_bins(const RefIndex& ri) const {
        std::lock_guard<std::mutex> lk(ix->mu);
        if (ri.bins_parsed) return;
        ri.bins.clear();
        ri.bins.reserve((size_t)std::max(0, ri.n_bin));
        const uint8_t* q = ri.bins_raw;
        for (int32_t i = 0; i < ri.n_bin; i++) {  // bounds were checked when the index was walked at open
            uint32_t bin;
            int32_t n_chunk;
            memcpy(&bin, q, 4);
            memcpy(&n_chunk, q + 4, 4);
            q += 8;
            if (bin != 37450 && n_chunk > 0) {
                uint64_t bmn = UINT64_MAX, bmx = 0;
                for (int32_t c = 0; c < n_chunk; c++) {
                    uint64_t be[2];
                    memcpy(be, q + 16 * (size_t)c, 16);
                    bmn = std::min(bmn, be[0]);
                    bmx = std::max(bmx, be[1]);
                }
                ri.bins.push_back({bin, {bmn, bmx}});
            }
            q += 16 * (size_t)n_chunk;
        }
        ri.bins_parsed = true;
    }

    void parse_bai(const std::string& bp) {
        BaiMap& d = ix->bai;
        d.fd = ::open(bp.c_str(), O_RDONLY);
        if (d.fd < 0) throw Error(MBFBAM_ERR_IO, "Could not read bam: no index " + bp);
        {
            struct stat sb;
            if (fstat(d.fd, &sb) != 0) throw Error(MBFBAM_ERR_IO, "Could not read bam: cannot stat index " + bp);
            d.n = (size_t)sb.st_size;
            if (d.n) {
                void* m = mmap(nullptr, d.n, PROT_READ, MAP_PRIVATE, d.fd, 0);
                if (m == MAP_FAILED) {
                    d.n = 0;
                    throw Error(MBFBAM_ERR_IO, "Could not read bam: cannot map index " + bp);
                }
                d.p = (const uint8_t*)m;
            }
        }
        size_t p = 0;
        auto take = [&](void* dst, size_t n) {
            if (p + n > d.size()) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: truncated index " + bp);
            memcpy(dst, d.data() + p, n);
            p += n;
        };
        char magic[4];
        take(magic, 4);
        if (memcmp(magic, "BAI\1", 4) != 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: not a BAI index: " + bp);
        int32_t n_ref;
        take(&n_ref, 4);
        ix->idx.assign(names.size(), RefIndex());
        for (int32_t r = 0; r < n_ref; r++) {
            RefIndex ri;
            take(&ri.n_bin, 4);
            if (ri.n_bin < 0) throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: bad index " + bp);
            ri.bins_raw = d.data() + p;
            // walk the bin headers only; the metadata pseudo-bin (37450, written by samtools/htslib) holds the
            // virtual offsets of the reference's first and last record
            bool meta = false, real = false;
            for (int32_t i = 0; i < ri.n_bin; i++) {
                uint32_t bin;
                int32_t n_chunk;
                take(&bin, 4);
                take(&n_chunk, 4);
                if (n_chunk < 0 || p + 16 * (size_t)n_chunk > d.size())
                    throw Error(MBFBAM_ERR_FORMAT, "Could not read bam: truncated index " + bp);
                if (bin == 37450) {
                    if (n_chunk >= 1) {
                        uint64_t be[2];
                        memcpy(be, d.data() + p, 16);
                        ri.voff_beg = be[0];
                        ri.voff_end = be[1];
                        meta = true;
                    }
                } else if (n_chunk > 0) {
                    real = true;
                }
                p += 16 * (size_t)n_chunk;
            }
            ri.bins_raw_len = (size_t)(d.data() + p - ri.bins_raw);
            int32_t n_intv;
            take(&n_intv, 4);
            ri.ioffset.resize((size_t)std::max(0, n_intv));
            if (n_intv > 0) take(ri.ioffset.data(), 8 * (size_t)n_intv);
            if (meta && real && ri.voff_end > ri.voff_beg) {
                ri.has_reads = true;
            } else {
                // no (usable) pseudo-bin: the offsets come from the chunk lists
                parse_bins(ri);
                uint64_t mn = UINT64_MAX, mx = 0;
                for (const auto& b : ri.bins) {
                    mn = std::min(mn, b.second.first);
                    mx = std::max(mx, b.second.second);
                }
                ri.has_reads = mn != UINT64_MAX;
                ri.voff_beg = ri.has_reads ? mn : 0;
                ri.voff_end = mx;
            }
            if ((size_t)r < ix->idx.size()) ix->idx[(size_t)r] = std::move(ri);
        }
    }

    // Lower bound of the virtual offset of any record on `tid` with pos >= start (linear index).
    uint64_t voff_lower(int32_t tid, uint32_t start) const {
        const RefIndex& ri = ix->idx[(size_t)tid];
        uint64_t v = ri.voff_beg;
        size_t w = start >> ix->min_shift;
        if (w < ri.ioffset.size() && ri.ioffset[w] > v) v = ri.ioffset[w];
        return v;
    }
    // Upper bound of the virtual END offset of any record on `tid` with pos < stop.  The file is sorted, so the
    // start of ANY record with pos >= stop will do; the bins whose genomic range begins at or after `stop` hold
    // only such records, and the smallest chunk begin among them is the first of them in the file (with dense
    // data: within one 16 kb bin of `stop`).  No such bin: the end of the reference.
    // (Round 1 took the largest chunk end over the bins overlapping [0, stop) -- the reg2bins walk -- which is
    // valid but loose: one read in a 64 Mb bin drags the range to wherever that read lies, up to 20 % more
    // bytes per shard on the 400M-read file.)
    uint64_t voff_upper(int32_t tid, uint32_t stop) const {
        const RefIndex& ri = ix->idx[(size_t)tid];
        if (stop >= lens[(size_t)tid] || stop == 0) return ri.voff_end;
        parse_bins(ri);
        uint64_t mn = UINT64_MAX;
        for (const auto& b : ri.bins) {
            int l = ix->depth;
            uint32_t lo = ((1u << (3 * l)) - 1u) / 7u;  // first bin of level l (0, 1, 9, 73, 585, 4681 for BAI)
            while (l > 0 && b.first < lo) { l--; lo = ((1u << (3 * l)) - 1u) / 7u; }
            const uint64_t bin_start = (uint64_t)(b.first - lo) << (ix->min_shift + 3 * (ix->depth - l));
            if (bin_start >= stop) mn = std::min(mn, b.second.first);
        }
        return mn == UINT64_MAX ? ri.voff_end : std::min(mn, ri.voff_end);
    }
};

// ------------------------------------------------------------------------------------------------
struct Chunk {
    int32_t tid;
    uint32_t start, stop;
};

struct SortedIntervals {
    std::vector<uint32_t> start, end, pmax, gene;  // gene = local gene_no
    std::vector<int8_t> strand;
    std::vector<uint32_t> rank;  // position in bio's find() iteration order (only when asked for)
    std::vector<uint32_t> orig;  // index of the interval in the contig's list (insertion order)
};

// bio 0.25 IntervalTree (Cargo.lock:100-101; source not vendored, restated from SURVEY.md App. B.2): an AVL tree
// keyed on interval.start -- insert descends LEFT when new.start <= node.start, else right, then repairs on the
// way up with the usual single/double rotations -- whose find() iterator pops a node, pushes its left child,
// then its right child, and yields the node: hits come in the order node, right subtree, left subtree.
// Pruning skips whole subtrees but never reorders, so every interval has a static rank in that traversal and
// "the first hit" (each_read_counts_once, reference src/count_reads/counters.rs:83-92,173-182) is the
// overlapping interval of minimum rank.  Returns rank[i] for the i-th inserted interval.
inline std::vector<uint32_t> bio_tree_ranks(const uint32_t* starts, size_t n) {
    struct Node { int32_t l = -1, r = -1, h = 1; };
    std::vector<Node> nd(n);
    auto height = [&](int32_t x) { return x < 0 ? 0 : nd[(size_t)x].h; };
    auto update = [&](int32_t x) { nd[(size_t)x].h = 1 + std::max(height(nd[(size_t)x].l), height(nd[(size_t)x].r)); };
    auto rot_right = [&](int32_t x) {  // returns the new subtree root
        int32_t y = nd[(size_t)x].l;
        nd[(size_t)x].l = nd[(size_t)y].r;
        nd[(size_t)y].r = x;
        update(x);
        update(y);
        return y;
    };
    auto rot_left = [&](int32_t x) {
        int32_t y = nd[(size_t)x].r;
        nd[(size_t)x].r = nd[(size_t)y].l;
        nd[(size_t)y].l = x;
        update(x);
        update(y);
        return y;
    };
    auto repair = [&](int32_t x) {
        const int lh = height(nd[(size_t)x].l), rh = height(nd[(size_t)x].r);
        if (std::abs(lh - rh) <= 1) { update(x); return x; }
        if (rh > lh) {
            int32_t r = nd[(size_t)x].r;
            if (height(nd[(size_t)r].l) > height(nd[(size_t)r].r)) nd[(size_t)x].r = rot_right(r);
            return rot_left(x);
        }
        int32_t l = nd[(size_t)x].l;
        if (height(nd[(size_t)l].r) > height(nd[(size_t)l].l)) nd[(size_t)x].l = rot_left(l);
        return rot_right(x);
    };
    int32_t root = -1;
    std::vector<int32_t> path;
    std::vector<uint8_t> went_left;
    for (size_t i = 0; i < n; i++) {
        if (root < 0) { root = (int32_t)i; continue; }
        path.clear();
        went_left.clear();
        int32_t cur = root;
        for (;;) {
            path.push_back(cur);
            const bool left = starts[i] <= starts[(size_t)cur];
            went_left.push_back(left ? 1 : 0);
            int32_t& child = left ? nd[(size_t)cur].l : nd[(size_t)cur].r;
            if (child < 0) { child = (int32_t)i; break; }
            cur = child;
        }
        // repair bottom-up; a rotation changes the subtree root, which the parent must point to.  bio repairs every
        // node of the path; a node whose subtree keeps its root and its height leaves everything above it as it
        // was, so the walk may stop there (same tree, a fifth of the work)
        for (size_t d = path.size(); d-- > 0;) {
            const int32_t x = path[d];
            const int old_h = nd[(size_t)x].h;
            const int32_t nr = repair(x);
            if (d == 0) root = nr;
            else if (went_left[d - 1]) nd[(size_t)path[d - 1]].l = nr;
            else nd[(size_t)path[d - 1]].r = nr;
            if (nr == x && nd[(size_t)x].h == old_h) break;
        }
    }
    std::vector<uint32_t> rank(n, 0);
    std::vector<int32_t> stack;
    if (root >= 0) stack.push_back(root);
    uint32_t k = 0;
    while (!stack.empty()) {
        const int32_t x = stack.back();
        stack.pop_back();
        rank[(size_t)x] = k++;
        if (nd[(size_t)x].l >= 0) stack.push_back(nd[(size_t)x].l);
        if (nd[(size_t)x].r >= 0) stack.push_back(nd[(size_t)x].r);
    }
    return rank;
}

inline SortedIntervals sort_intervals(const mbfbam_intervals* iv, int c, bool with_rank = false) {
    const uint64_t a = iv->iv_offsets[c], b = iv->iv_offsets[c + 1];
    const size_t n = (size_t)(b - a);
    // (start, end, list index) packed into sortable keys: 64 bits of (start, end), ties by list index
    struct Key { uint64_t se; uint32_t idx; };
    std::vector<Key> keys(n);
    bool sorted = true;
    for (size_t i = 0; i < n; i++) {
        keys[i] = {((uint64_t)iv->iv_start[a + i] << 32) | iv->iv_end[a + i], (uint32_t)i};
        if (i && keys[i].se < keys[i - 1].se) sorted = false;
    }
    if (!sorted) std::sort(keys.begin(), keys.end(), [](const Key& x, const Key& y) { return x.se != y.se ? x.se < y.se : x.idx < y.idx; });
    SortedIntervals s;
    s.start.resize(n); s.end.resize(n); s.pmax.resize(n); s.gene.resize(n); s.strand.resize(n); s.orig.resize(n);
    std::vector<uint32_t> ranks;
    if (with_rank) { ranks = bio_tree_ranks(iv->iv_start + a, n); s.rank.resize(n); }
    uint32_t m = 0;
    for (size_t k = 0; k < n; k++) {
        const uint32_t o = keys[k].idx;
        const uint32_t st = iv->iv_start[a + o], en = iv->iv_end[a + o];
        if (st > en) throw Error(MBFBAM_ERR_ARG, "interval start > end");  // bio panics here
        s.start[k] = st;
        s.end[k] = en;
        s.gene[k] = iv->iv_gene[a + o];
        s.strand[k] = iv->iv_strand[a + o];
        s.orig[k] = o;
        if (with_rank) s.rank[k] = ranks[o];
        m = std::max(m, en);
        s.pmax[k] = m;
    }
    return s;
}

// chunked_genome.rs:93-109: push the chunk's right edge past a gene interval covering it.  When several
// cover it the reference takes the FIRST entry bio's find(stop..stop+1) yields, i.e. the covering interval of
// least rank in the tree's iteration order (bio_tree_ranks above).
// The ranks are only needed where two or more intervals cover an edge, so the tree is replayed (20 k inserts take
// milliseconds) only for contigs where that happens: `ranks()` fills g.rank on first use.
template <class R>
inline uint32_t extend_stop(SortedIntervals& g, uint32_t stop, R&& ranks) {
    for (;;) {
        size_t hi = (size_t)(std::upper_bound(g.start.begin(), g.start.end(), stop) - g.start.begin());  // start <= stop
        size_t n_cover = 0, only = 0;
        for (size_t i = hi; i-- > 0;) {
            if (g.pmax[i] <= stop) break;
            if (g.end[i] > stop) { n_cover++; only = i; }
        }
        if (n_cover == 0) return stop;
        uint32_t best_end = g.end[only];
        if (n_cover > 1) {
            if (g.rank.empty()) ranks();
            uint32_t best_rank = 0xffffffffu;
            for (size_t i = hi; i-- > 0;) {
                if (g.pmax[i] <= stop) break;
                if (g.end[i] > stop && g.rank[i] < best_rank) {
                    best_rank = g.rank[i];
                    best_end = g.end[i];
                }
            }
        }
        stop = best_end + 1;
    }
}

// gene_intervals == nullptr: new_without_tree (all header contigs, fixed 1 Mb chunks).
inline std::vector<Chunk> make_chunks(const Reader& r, const mbfbam_intervals* genes) {
    std::vector<Chunk> out;
    int n = genes ? genes->n_contigs : (int)r.names.size();
    for (int c = 0; c < n; c++) {
        int32_t tid = c;
        SortedIntervals g;
        if (genes) {
            auto it = r.name2tid.find(genes->contig_names[c]);
            if (it == r.name2tid.end()) continue;  // chunked_genome.rs:21-25
            tid = it->second;
            g = sort_intervals(genes, c, false);
        }
        uint32_t len = r.lens[(size_t)tid], start = 0;
        do {
            uint32_t stop = start + CHUNK_SIZE;
            if (genes)
                stop = extend_stop(g, stop, [&] {
                    const std::vector<uint32_t> rk = bio_tree_ranks(genes->iv_start + genes->iv_offsets[c], g.orig.size());
                    g.rank.resize(g.orig.size());
                    for (size_t i = 0; i < g.orig.size(); i++) g.rank[i] = rk[g.orig[i]];
                });
            out.push_back({tid, start, stop});
            start = stop;
        } while (start < len);
    }
    return out;
}

// A contiguous run of BGZF blocks to scan: file bytes [cbeg, cend), first record at first_uoff
// inside the block at cbeg.
struct ByteRange {
    uint64_t cbeg, cend;
    uint32_t first_uoff;
};

inline void add_range(std::vector<ByteRange>& v, uint64_t voff_beg, uint64_t voff_end, uint64_t file_size) {
    uint64_t cbeg = voff_beg >> 16, cend = voff_end >> 16;
    if (voff_end & 0xffff) cend += 1;  // the block holding the end offset is needed as well; its true
                                       // end is found by the scanner (scan_len reaches one byte into it)
    if (cend > file_size) cend = file_size;
    if (cbeg >= cend) return;
    if (!v.empty() && cbeg <= v.back().cend + (1u << 20)) {
        // overlapping, adjacent or separated by less than 1 MiB (e.g. the rest of the previous range's
        // last block): one range.  Blocks in between are scanned too; their records fail the
        // ownership test, so this only trades a little extra work for fewer, fuller windows.
        v.back().cend = std::max(v.back().cend, cend);
        return;
    }
    v.push_back({cbeg, cend, (uint32_t)(voff_beg & 0xffff)});
}

}  // namespace mbf
