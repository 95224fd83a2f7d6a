// Host side of the ingest path: where the compressed bytes of a BAM come from before the H2D copy.
//
// Measured on the GPU box (tools/pin_bench.cu, profiles/r02_pin_bench.txt; 16 vCPUs, file in the page cache):
//   pread -> pinned buffer        7.8 GB/s per thread, 43 GB/s with 8 threads
//   memcpy(mapping) -> pinned    13.4 GB/s per thread, 65-70 GB/s with 8 threads (first touch included)
//   H2D from page-locked memory  54-55 GB/s
//   cudaHostRegister(mapping)     9.5 GB/s in 64 MiB segments (no faster from several threads; a read-only shared
//                                 mapping is refused, a private never-written one is accepted; one copy may not
//                                 span two registrations)
// So the cold path copies out of a private mapping of the file with a few threads (FileMap), and a file that
// keeps being counted gets its mapping page-locked segment by segment (PinState) so that later jobs DMA
// straight out of the page cache with no CPU copy at all.  Both are process-wide caches keyed by
// (device, inode, size, mtime): the four counters of a suite, or the ranks' repeated jobs, share them.
//
// Replaces the file reads under htslib's bgzf_read_block (hFILE read/pread), reached from
// `bam.read(&mut read)` in the reference (src/count_reads/counters.rs:41,132; introns.rs:33).
#pragma once
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdint.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace mbf {

constexpr size_t PIN_SEG = 64u << 20;  // registration granule

struct FileMap {
    uint64_t dev = 0, ino = 0, size = 0;
    int64_t mtime_ns = 0;
    uint8_t* p = nullptr;  // private mapping (PROT_READ|PROT_WRITE so that it can be page-locked; never written)
    // page-lock state per PIN_SEG segment: 0 none, 1 being registered, 2 registered, 3 refused
    std::unique_ptr<std::atomic<uint8_t>[]> seg;
    size_t n_seg = 0;
    std::atomic<uint64_t> jobs{0};          // jobs that streamed this file (pin policy)
    std::atomic<uint64_t> pinned_bytes{0};
    std::mutex reg_mu;                      // registrations are serialised by the driver anyway
    bool refused = false;                   // the platform does not page-lock this mapping: stop trying

    ~FileMap() {
        if (!p) return;
        for (size_t i = 0; i < n_seg; i++)
            if (seg[i].load() == 2) cudaHostUnregister(p + i * PIN_SEG);
        munmap(p, (size_t)size);
    }
    bool seg_pinned(uint64_t off, uint64_t len) const {
        if (!len) return true;
        for (size_t i = off / PIN_SEG; i <= (off + len - 1) / PIN_SEG; i++)
            if (seg[i].load(std::memory_order_acquire) != 2) return false;
        return true;
    }
};

class FileMapCache {
  public:
    static FileMapCache& get() {
        static FileMapCache* c = new FileMapCache();  // never destroyed: no CUDA calls during static teardown
        return *c;
    }
    // mapping of the open file `fd` (nullptr if it cannot be mapped: the caller falls back to pread)
    std::shared_ptr<FileMap> lookup(int fd) {
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size <= 0) return nullptr;
        const int64_t mt = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
        std::lock_guard<std::mutex> lk(mu_);
        for (auto it = lru_.begin(); it != lru_.end(); ++it) {
            FileMap& m = **it;
            if (m.dev == (uint64_t)st.st_dev && m.ino == (uint64_t)st.st_ino) {
                if (m.size == (uint64_t)st.st_size && m.mtime_ns == mt) {
                    lru_.splice(lru_.begin(), lru_, it);
                    return lru_.front();
                }
                lru_.erase(it);  // the file changed: its mapping (and page locks) go when the last job lets go
                break;
            }
        }
        void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_NORESERVE, fd, 0);
        if (p == MAP_FAILED) return nullptr;
        auto m = std::make_shared<FileMap>();
        m->dev = (uint64_t)st.st_dev;
        m->ino = (uint64_t)st.st_ino;
        m->size = (uint64_t)st.st_size;
        m->mtime_ns = mt;
        m->p = (uint8_t*)p;
        m->n_seg = ((size_t)st.st_size + PIN_SEG - 1) / PIN_SEG;
        m->seg.reset(new std::atomic<uint8_t>[m->n_seg]);
        for (size_t i = 0; i < m->n_seg; i++) m->seg[i].store(0);
        lru_.push_front(m);
        while (lru_.size() > 8) lru_.pop_back();
        return m;
    }
    void clear() {
        std::lock_guard<std::mutex> lk(mu_);
        lru_.clear();
    }
    uint64_t total_pinned() {
        std::lock_guard<std::mutex> lk(mu_);
        uint64_t n = 0;
        for (auto& m : lru_) n += m->pinned_bytes.load();
        return n;
    }

  private:
    std::mutex mu_;
    std::list<std::shared_ptr<FileMap>> lru_;
};

// Page-lock every segment that overlaps [off, off+len).  Returns the milliseconds spent registering; segments the
// platform refuses stay unpinned (the job then stages those bytes through the copy path).
inline double pin_range(FileMap& m, uint64_t off, uint64_t len, uint64_t budget_bytes) {
    if (!len || m.refused) return 0.0;
    double ms = 0.0;
    for (size_t i = off / PIN_SEG; i <= (off + len - 1) / PIN_SEG && i < m.n_seg; i++) {
        if (m.seg[i].load(std::memory_order_acquire) == 2) continue;
        std::lock_guard<std::mutex> lk(m.reg_mu);
        if (m.seg[i].load() == 2 || m.refused) continue;
        const uint64_t so = (uint64_t)i * PIN_SEG, sl = std::min<uint64_t>(PIN_SEG, m.size - so);
        if (FileMapCache::get().total_pinned() + sl > budget_bytes) return ms;
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = cudaHostRegister(m.p + so, (size_t)sl, cudaHostRegisterPortable);
        ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (e != cudaSuccess) {
            cudaGetLastError();
            m.seg[i].store(3);
            m.refused = true;
            return ms;
        }
        m.pinned_bytes.fetch_add(sl);
        m.seg[i].store(2, std::memory_order_release);
    }
    return ms;
}

inline void unpin_all(FileMap& m) {
    std::lock_guard<std::mutex> lk(m.reg_mu);
    for (size_t i = 0; i < m.n_seg; i++)
        if (m.seg[i].load() == 2) {
            cudaHostUnregister(m.p + i * PIN_SEG);
            m.seg[i].store(0, std::memory_order_release);
        }
    m.pinned_bytes.store(0);
}

}  // namespace mbf
