// libmbfbam_cuda.so: window pipeline + C ABI (include/mbfbam.h).
//
// Per job the host plans BGZF byte ranges from the BAI, then streams them through the GPU in
// windows of W compressed bytes:
//
//   front stage (stream s_front):    host copy on a helper thread -> H2D (or a view into the preloaded
//                                    file) -> K0 block scan -> chain proof -> ISIZE scan -> WinMeta to the host
//   inflate     (streams s_inf[k%2]): K1 pass 1 (tokens) -> pass 2 (LZ77); consecutive windows alternate streams
//   main stage  (stream s_main):     K2 frame -> carry -> K3-5 count | K6 introns
//
// Window k+1's front stage and window k+1's K1 run while window k's K1 / main stage are executing.  All
// per-window device arrays are double buffered.  The host reads WinMeta only to size buffers (U, record offsets, the
// multi-mapper set, the intron map); it never touches alignment data.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <list>
#include <map>
#include <memory>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "reader.hpp"
#include "hostmap.hpp"

using namespace mbf;

#define MBFBAM_VERSION "0.1.0"

namespace {

std::string fmt(const char* f, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            throw Error(MBFBAM_ERR_CUDA, fmt("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), \
                                             __FILE__, __LINE__, #expr));                          \
    } while (0)

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return (uint32_t)strtoul(v, nullptr, 10);
}

// A window inflated to more than 32-bit offsets can address (the planner sizes windows from the largest
// inflate ratio seen so far; a sudden, much more compressible stretch defeats it).  The job is started
// again with windows planned for `ratio`.
struct RetryJob : Error {
    double ratio;
    RetryJob(const std::string& m, double r) : Error(MBFBAM_ERR_UNSUPPORTED, m), ratio(r) {}
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    // grows without preserving contents
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr;
        cap = 0;
        CK(cudaMalloc(&p, n));
        cap = n;
    }
    template <class T> T* as() const { return (T*)p; }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) CK(cudaFreeHost(p));
        p = nullptr;
        CK(cudaHostAlloc(&p, n, cudaHostAllocPortable));
        cap = n;
    }
    template <class T> T* as() const { return (T*)p; }
};

template <class T>
T* upload(DevBuf& b, const std::vector<T>& v, cudaStream_t s) {
    b.ensure(std::max<size_t>(16, v.size() * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return b.as<T>();
}

// Where compressed bytes come from.
struct Source {
    virtual ~Source() {}
    virtual uint64_t size() const = 0;
    virtual void read(uint64_t off, size_t n, uint8_t* dst) const = 0;
    virtual const uint8_t* resident(int device) const { return nullptr; }  // device copy of the whole file
    // page-locked host bytes of [off, off+n) if all of it is page-locked (H2D straight from there)
    virtual const uint8_t* host_pinned(uint64_t off, uint64_t n) const { return nullptr; }
};
struct FileSource : Source {
    const Reader* r;
    std::shared_ptr<FileMap> map;  // may be null: pread path
    FileSource(const Reader* r_, std::shared_ptr<FileMap> m) : r(r_), map(std::move(m)) {}
    uint64_t size() const override { return r->file_size; }
    // a mapping must not be read past the end of a file that shrank under us (SIGBUS): one fstat per piece
    bool still_there(uint64_t end) const {
        struct stat sb;
        return fstat(r->fd, &sb) == 0 && (uint64_t)sb.st_size >= end;
    }
    void read(uint64_t off, size_t n, uint8_t* dst) const override {
        if (map && off + n <= map->size) {
            if (!still_there(off + n)) throw Error(MBFBAM_ERR_IO, "Could not read bam: short read on " + r->path);
            memcpy(dst, map->p + off, n);  // hostmap.hpp: 1.7x the rate of pread
        } else {
            r->pread_exact(dst, off, n);
        }
    }
    const uint8_t* resident(int device) const override { return r->preload_device == device ? r->preload_ptr : nullptr; }
    const uint8_t* host_pinned(uint64_t off, uint64_t n) const override {
        return map && off + n <= map->size && map->seg_pinned(off, n) && still_there(off + n) ? map->p + off : nullptr;
    }
};
struct MemSource : Source {
    const uint8_t* p;
    uint64_t n;
    MemSource(const uint8_t* p_, uint64_t n_) : p(p_), n(n_) {}
    uint64_t size() const override { return n; }
    void read(uint64_t off, size_t len, uint8_t* dst) const override { memcpy(dst, p + off, len); }
};

enum JobKind { JOB_COUNT, JOB_INTRONS, JOB_INFLATE, JOB_QUANTIFY, JOB_DUPS, JOB_BARCODE };

struct WinBufs {
    PinnedBuf h_c;
    DevBuf d_c;
    DevBuf tile_count, tile_base, cand_pos, cand_bsize;
    DevBuf cpos, clen, isize, uoff;
    DevBuf blk_start, blk_end, blk_nrec, rec_base, rec_off, rec_tmp;
    DevBuf mm_fp, mm_slot;
    uint32_t mm_cap = 0;
    const uint8_t* c_ptr = nullptr;  // device pointer of this window's bytes (d_c or a view of the preload)
    bool two_pass = false;           // t[7] was recorded between the two passes of K1
    uint64_t win_off = 0;
    uint32_t scan_len = 0, avail = 0;
    cudaEvent_t ev_front = nullptr, ev_main = nullptr, ev_inf = nullptr;
    cudaEvent_t ev_cfree = nullptr;  // the compressed bytes of this window are no longer read (pass 1 of K1 is done)
    cudaEvent_t ev_h2d = nullptr;    // the window's bytes are in HBM
    cudaEvent_t t[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    WinMeta front_meta{};
    bool timed = false;
    bool in_flight = false;
    int range_start = 0;
    uint64_t range_off = 0;
    uint32_t first_uoff = 0;
};

struct Pipeline {
    int device;
    const Source* src;
    JobKind kind;
    mbfbam_stats st{};
    uint32_t W;          // window size in compressed bytes
    uint32_t cand_cap;
    cudaStream_t s_front = nullptr, s_main = nullptr, s_inf[2] = {nullptr, nullptr};
    cudaStream_t s_copy = nullptr;  // H2D of the windows: nothing else, so that window k+1's copy follows window k's at once
                                    // (behind the scan kernels on s_front the link idled 0.75 ms per window)
    cudaEvent_t ev_reuse = nullptr, ev_job_begin = nullptr, ev_job_end = nullptr;
    WinBufs wb[2];
    DevBuf U[2];
    size_t U_cap[2] = {0, 0};  // data bytes (without prefix/slack); the two buffers grow independently
    DevBuf d_meta, d_chain, d_carry, d_tscratch, d_tok[2], d_ntok[2], d_fb[2], d_pscratch[2], d_phdr[2], d_phlens[2];
    PinnedBuf h_meta_front, h_meta_main;
    int sm_count = 148;
    bool sync_debug = false;
    double ratio_hint = 1.0;  // inflate ratio the first windows are planned for (raised by a RetryJob)

    // ---- counting job state
    int mode = 0;
    CountCtx cc{};
    DevBuf d_tid2slot, d_slot_iv_off, d_iv_start, d_iv_end, d_iv_pmax, d_iv_gene, d_iv_rank, d_slot_chunk_off, d_chunk_stop;
    DevBuf d_counts, d_wspill;
    DevBuf d_set;
    uint64_t set_cap = 0, set_ub = 0;
    DevBuf d_set_size;
    size_t n_genes = 0, n_count_words = 0;
    // ---- intron job state
    IntronCtx ic{};
    DevBuf d_ref_chunk_off, d_map, d_hist, d_scalar;
    unsigned long long* hist_ptr = nullptr;
    uint64_t map_cap = 0, map_ub = 0;
    // ---- duplicate-distribution job (shares d_map / map_cap with the intron job)
    DupCtx dc{};
    DevBuf d_ref_len, d_dup_dense, d_dup_big, d_dup_nbig;
    // ---- by-barcode job (uses d_set / d_map too)
    DevBuf d_bclist, d_bclist_n, d_chunk_err;
    uint64_t bclist_cap = 0, bclist_ub = 0;
    size_t bc_n_chunks = 0;  // d_chunk_err holds bc_n_chunks flags, then the job's error word
    // ---- inflate-only job
    uint8_t* out_host = nullptr;
    uint64_t out_cap = 0, out_len = 0;

    explicit Pipeline(int dev) : device(dev), src(nullptr), kind(JOB_INFLATE) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw Error(MBFBAM_ERR_CUDA, fmt("no CUDA device available (%s); mbf_bam_b200 has no CPU path",
                                             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)));
        if (dev < 0 || dev >= n) throw Error(MBFBAM_ERR_ARG, fmt("device %d out of range (have %d)", dev, n));
        CK(cudaSetDevice(dev));
        CK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        // K1 launches thousands of CTAs that each live for milliseconds; the short kernels of the other
        // stages must not queue behind them, so their streams get the higher priority
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaStreamCreateWithPriority(&s_front, cudaStreamNonBlocking, prio_hi));
        CK(cudaStreamCreateWithPriority(&s_copy, cudaStreamNonBlocking, prio_hi));
        CK(cudaStreamCreateWithPriority(&s_main, cudaStreamNonBlocking, prio_hi));
        CK(cudaStreamCreateWithPriority(&s_inf[0], cudaStreamNonBlocking, prio_lo));
        CK(cudaStreamCreateWithPriority(&s_inf[1], cudaStreamNonBlocking, prio_lo));
        CK(cudaEventCreateWithFlags(&ev_reuse, cudaEventDisableTiming));
        CK(cudaEventCreate(&ev_job_begin));
        CK(cudaEventCreate(&ev_job_end));
        d_meta.ensure(2 * sizeof(WinMeta));
        d_chain.ensure(sizeof(ChainState));
        d_carry.ensure(16);
        h_meta_front.ensure(2 * sizeof(WinMeta));
        h_meta_main.ensure(2 * sizeof(WinMeta));
        for (int i = 0; i < 2; i++) {
            WinBufs& w = wb[i];
            CK(cudaEventCreateWithFlags(&w.ev_front, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&w.ev_main, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&w.ev_inf, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&w.ev_cfree, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&w.ev_h2d, cudaEventDisableTiming));
            for (auto& t : w.t) CK(cudaEventCreate(&t));
        }
    }

    // The pipeline object (streams, events, device and pinned buffers) is cached per device and reused
    // by every call; begin_job() resets the per-job state.  `job_bytes`: compressed bytes this job will
    // stream, so that small jobs do not allocate full-size windows.
    void begin_job(const Source* s, JobKind k, uint64_t job_bytes) {
        CK(cudaSetDevice(device));
        // a previous job may have ended with an exception while work was still queued
        for (cudaStream_t x : {s_front, s_copy, s_inf[0], s_inf[1], s_main}) CK(cudaStreamSynchronize(x));
        src = s;
        kind = k;
        st = mbfbam_stats{};
        uint64_t w = (uint64_t)env_u32("MBF_BAM_WINDOW_MB", 512) << 20;
        w = std::min<uint64_t>(std::max<uint64_t>(w, 1u << 20), 1024u << 20);
        uint64_t need = ((job_bytes + (1u << 20)) >> 20) << 20;
        W = (uint32_t)std::max<uint64_t>(1u << 20, std::min(w, need));
        cand_cap = W / 64 + 1024;
        sync_debug = env_u32("MBF_BAM_SYNC", 0) != 0;
        {
            const uint32_t hl = env_u32("MBF_BAM_U_HARD_LIMIT_MB", 0);
            u_hard_limit = hl ? std::min<uint64_t>((uint64_t)hl << 20, U_DATA_LIMIT) : U_DATA_LIMIT;
        }
        CK(cudaMemsetAsync(d_chain.p, 0, sizeof(ChainState), s_front));
        CK(cudaMemsetAsync(d_carry.p, 0, 16, s_main));
        uint32_t mm_env = env_u32("MBF_BAM_MM_CAP", 0);
        for (int i = 0; i < 2; i++) {
            WinBufs& wbi = wb[i];
            size_t ntile = (size_t)W / SCAN_TILE + 2;
            wbi.tile_count.ensure(ntile * 4);
            wbi.tile_base.ensure(ntile * 4);
            for (DevBuf* b : {&wbi.cand_pos, &wbi.cand_bsize, &wbi.cpos, &wbi.clen, &wbi.isize, &wbi.uoff, &wbi.blk_start,
                              &wbi.blk_end, &wbi.blk_nrec, &wbi.rec_base})
                b->ensure((size_t)(cand_cap + 8) * 4);
            wbi.in_flight = wbi.timed = false;
            if (mm_env) wbi.mm_cap = 0;  // tests force a tiny key buffer
        }
        for (int i = 0; i < 2; i++) ensure_U(i, std::min<size_t>((size_t)W * 4, (size_t)3 << 30), false);
        mode = 0;
        cc = CountCtx{};
        ic = IntronCtx{};
        // the multi-mapper set and the intron map keep their allocation between jobs; only their contents go
        set_ub = 0;
        if (set_cap) CK(cudaMemsetAsync(d_set.p, 0, set_cap * 16, s_main));
        map_ub = 0;
        if (map_cap) CK(cudaMemsetAsync(d_map.p, 0xff, map_cap * 16, s_main));
        bclist_ub = 0;
        n_genes = n_count_words = 0;
        out_host = nullptr;
        out_cap = out_len = 0;
        outside_total = 0;
    }
    ~Pipeline() {
        cudaSetDevice(device);
        cudaDeviceSynchronize();
        for (auto& w : wb) {
            if (w.ev_front) cudaEventDestroy(w.ev_front);
            if (w.ev_main) cudaEventDestroy(w.ev_main);
            if (w.ev_inf) cudaEventDestroy(w.ev_inf);
            if (w.ev_cfree) cudaEventDestroy(w.ev_cfree);
            if (w.ev_h2d) cudaEventDestroy(w.ev_h2d);
            for (auto& t : w.t) if (t) cudaEventDestroy(t);
        }
        if (ev_reuse) cudaEventDestroy(ev_reuse);
        if (ev_job_begin) cudaEventDestroy(ev_job_begin);
        if (ev_job_end) cudaEventDestroy(ev_job_end);
        for (auto& x : s_inf) if (x) cudaStreamDestroy(x);
        if (s_front) cudaStreamDestroy(s_front);
        if (s_copy) cudaStreamDestroy(s_copy);
        if (s_main) cudaStreamDestroy(s_main);
    }

    // MBF_BAM_TIMELINE=1 (debug): an event behind every launch and copy; run() prints when each completed, relative
    // to the start of the job -- the gaps between the entries of a stream are where that stream waited
    struct TlEv { cudaEvent_t ev; const char* what; char stream; };
    bool timeline = false;
    std::vector<TlEv> tl;
    std::mutex tl_mu;
    void mark(const char* what, cudaStream_t s) {
        if (!timeline) return;
        TlEv e{nullptr, what, s == s_front ? 'F' : s == s_copy ? 'C' : s == s_main ? 'M' : s == s_inf[0] ? 'a' : 'b'};
        if (cudaEventCreate(&e.ev) != cudaSuccess || cudaEventRecord(e.ev, s) != cudaSuccess) return;
        std::lock_guard<std::mutex> lk(tl_mu);
        tl.push_back(e);
    }
    void print_timeline() {
        if (!timeline) return;
        std::lock_guard<std::mutex> lk(tl_mu);
        for (const TlEv& e : tl) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ev_job_begin, e.ev) == cudaSuccess) fprintf(stderr, "[mbfbam-tl] %8.3f %c %s\n", ms, e.stream, e.what);
            cudaEventDestroy(e.ev);
        }
        tl.clear();
    }

    void launched(const char* what, cudaStream_t s) {
        st.n_kernel_launches++;
        mark(what, s);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) throw Error(MBFBAM_ERR_CUDA, fmt("launch of %s failed: %s", what, cudaGetErrorString(e)));
        if (sync_debug) {
            e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) throw Error(MBFBAM_ERR_CUDA, fmt("kernel %s failed: %s", what, cudaGetErrorString(e)));
        }
    }

    WinMeta* dmeta(int i) { return d_meta.as<WinMeta>() + i; }

    // Largest inflated size of one window: 32-bit offsets into U (prefix + data + slack < 4 GiB).
    static constexpr uint64_t U_DATA_LIMIT = 0xffff0000ull - U_PREFIX - 65536;
    uint64_t u_hard_limit = U_DATA_LIMIT;  // MBF_BAM_U_HARD_LIMIT_MB lowers it so that tests can reach the retry

    // Grow U[bi] (and what is sized from it: this parity's record offsets).  Only buffers of parity `bi` are
    // touched: window k-1 (the other parity) may still have its second half outstanding (intron pass 2, the
    // emit-only re-run after a key overflow), which reads U[bi^1] and wb[bi^1].rec_off.  The last users of
    // U[bi] are window k-2's main stage (complete once s_main has drained: its second half was queued one
    // iteration ago) and window k-1's carry copy INTO its prefix, which `keep_prefix` preserves.
    void ensure_U(int bi, size_t data_bytes, bool keep_prefix) {
        if (data_bytes <= U_cap[bi]) return;
        if (data_bytes > U_DATA_LIMIT) throw RetryJob("window inflates past 4 GiB", 0.0);
        size_t cap = std::min<size_t>(std::max(data_bytes, U_cap[bi] * 2), U_DATA_LIMIT);
        CK(cudaStreamSynchronize(s_inf[0]));
        CK(cudaStreamSynchronize(s_inf[1]));
        CK(cudaStreamSynchronize(s_main));
        DevBuf nb;
        nb.ensure(cap + U_PREFIX + 65536);
        if (keep_prefix && U[bi].p) CK(cudaMemcpy(nb.p, U[bi].p, U_PREFIX, cudaMemcpyDeviceToDevice));
        std::swap(U[bi].p, nb.p);
        std::swap(U[bi].cap, nb.cap);
        U_cap[bi] = cap;
        wb[bi].rec_off.ensure((cap / 36 + 64) * 4);
        wb[bi].rec_tmp.ensure(((cap + U_PREFIX + 65536) / 36 + 64) * 4);
    }

    // ---------------------------------------------------------------- front stage
    // front stage, first half: bytes of the window into HBM.  Only the compressed-byte buffers of this
    // parity are touched, so this may start as soon as K1 of window k-2 (their last reader) is done.
    void stage_front(int bi, uint64_t win_off, uint32_t scan_len, uint64_t src_end, int range_start,
                     uint64_t range_off, uint32_t first_uoff) {
        WinBufs& w = wb[bi];
        CK(cudaStreamWaitEvent(s_copy, w.ev_cfree, 0));  // the previous user of this C buffer: pass 1 of window k-2
        uint64_t want = (uint64_t)scan_len + 65536 + 64;
        uint64_t avail = std::min<uint64_t>(want, src_end - win_off);
        w.win_off = win_off;
        w.scan_len = scan_len;
        w.avail = (uint32_t)avail;
        w.range_start = range_start;
        w.range_off = range_off;
        w.first_uoff = first_uoff;
        const uint8_t* res = src->resident(device);
        if (res) {
            w.c_ptr = res + win_off;
        } else {
            w.d_c.ensure((size_t)W + 65536 + 256);
            stage_window(w, win_off, (size_t)avail);
            st.h2d_bytes += avail;
            w.c_ptr = w.d_c.as<uint8_t>();
        }
        CK(cudaEventRecord(w.ev_h2d, s_copy));
        st.compressed_bytes += scan_len;
    }

    // front stage, second half: K0 + block table + WinMeta.  Writes the per-window arrays that window k-2's
    // main stage may still be reading, hence the wait for ev_reuse.
    void scan_front(int bi) {
        WinBufs& w = wb[bi];
        const uint64_t win_off = w.win_off;
        const uint32_t scan_len = w.scan_len;
        const uint64_t avail = w.avail;
        const int range_start = w.range_start;
        const uint64_t range_off = w.range_off;
        const uint32_t first_uoff = w.first_uoff;
        CK(cudaStreamWaitEvent(s_front, ev_reuse, 0));
        CK(cudaStreamWaitEvent(s_front, w.ev_h2d, 0));
        CK(cudaEventRecord(w.t[0], s_front));
        k_win_begin<<<1, 1, 0, s_front>>>(d_chain.as<ChainState>(), dmeta(bi), range_start, range_off, first_uoff);
        launched("k_win_begin", s_front);
        ScanArgs a{};
        a.cbuf = w.c_ptr;
        a.win_off = win_off;
        a.scan_len = scan_len;
        a.avail = (uint32_t)avail;
        a.chain = d_chain.as<ChainState>();
        a.meta = dmeta(bi);
        a.tile_count = w.tile_count.as<uint32_t>();
        a.tile_base = w.tile_base.as<uint32_t>();
        a.cand_pos = w.cand_pos.as<uint32_t>();
        a.cand_bsize = w.cand_bsize.as<uint32_t>();
        a.cand_cap = cand_cap;
        uint32_t ntile = (scan_len + SCAN_TILE - 1) / SCAN_TILE;
        if (ntile) {
            k_scan<false><<<ntile, SCAN_THREADS, 0, s_front>>>(a);
            launched("k_scan<count>", s_front);
        }
        k_excl_scan<<<1, 1024, 0, s_front>>>(a.tile_count, a.tile_base, nullptr, ntile, 0, &dmeta(bi)->n_cand, nullptr, 0);
        launched("k_excl_scan(tiles)", s_front);
        if (ntile) {
            k_scan<true><<<ntile, SCAN_THREADS, 0, s_front>>>(a);
            launched("k_scan<emit>", s_front);
        }
        BlockArrays b{w.cpos.as<uint32_t>(), w.clen.as<uint32_t>(), w.isize.as<uint32_t>(), w.uoff.as<uint32_t>()};
        k_blocks_finalize<<<std::max(1u, std::min(64u, cand_cap / 256)), 256, 0, s_front>>>(a, b);
        launched("k_blocks_finalize", s_front);
        k_excl_scan<<<1, 1024, 0, s_front>>>(b.isize, b.uoff, &dmeta(bi)->n_blocks, 0, U_PREFIX, &dmeta(bi)->u_total, dmeta(bi),
                                             u_hard_limit - 1024);
        launched("k_excl_scan(isize)", s_front);
        CK(cudaEventRecord(w.t[1], s_front));
        CK(cudaMemcpyAsync(h_meta_front.as<WinMeta>() + bi, dmeta(bi), sizeof(WinMeta), cudaMemcpyDeviceToHost, s_front));
        st.d2h_bytes += sizeof(WinMeta);
        CK(cudaEventRecord(w.ev_front, s_front));
        w.in_flight = true;
    }

    // Host bytes -> HBM.  Page-locked source (hostmap.hpp): one DMA per registration segment, no CPU copy.
    // Otherwise the window is cut into pieces copied into a pinned staging buffer by a few threads (memcpy out of
    // the file mapping runs at ~13 GB/s per thread, pread at ~8); every piece goes to the device as soon as it and
    // its predecessors have landed, so the host copy and the H2D overlap.
    int read_threads = 0;  // 0: derived from the core count
    void stage_window(WinBufs& w, uint64_t win_off, size_t avail) {
        uint8_t* d = w.d_c.as<uint8_t>();
        if (const uint8_t* hp = src->host_pinned(win_off, avail)) {
            for (size_t off = 0; off < avail;) {
                const size_t seg_left = PIN_SEG - (size_t)((win_off + off) % PIN_SEG);
                const size_t len = std::min(seg_left, avail - off);
                CK(cudaMemcpyAsync(d + off, hp + off, len, cudaMemcpyHostToDevice, s_copy));
                off += len;
            }
            mark("h2d(window, from the page-locked mapping)", s_copy);
            CK(cudaMemsetAsync(d + avail, 0, 128, s_copy));
            st.h2d_pinned_bytes += avail;
            return;
        }
        w.h_c.ensure((size_t)W + 65536 + 256);
        uint8_t* h = w.h_c.as<uint8_t>();
        const size_t PIECE = 8u << 20;
        const size_t n_pieces = (avail + PIECE - 1) / PIECE;
        int n_threads = read_threads > 0 ? read_threads
                                         : (int)env_u32("MBF_BAM_READ_THREADS", std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2)));
        n_threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, n_pieces));
        if (n_pieces <= 1 || n_threads <= 1) {
            src->read(win_off, avail, h);
            memset(h + avail, 0, 128);
            CK(cudaMemcpyAsync(d, h, avail + 128, cudaMemcpyHostToDevice, s_copy));
            return;
        }
        std::vector<std::atomic<int>> done(n_pieces);
        for (auto& x : done) x.store(0);
        std::atomic<size_t> next{0};
        std::atomic<int> failed{0};
        std::string fail_msg;
        std::mutex fail_mu;
        std::vector<std::thread> th;
        for (int t = 0; t + 1 < n_threads; t++)  // the calling (staging helper) thread is the n-th copier
            th.emplace_back([&] {
                for (;;) {
                    size_t i = next.fetch_add(1);
                    if (i >= n_pieces) break;
                    size_t off = i * PIECE, len = std::min(PIECE, avail - off);
                    try {
                        if (!failed.load()) src->read(win_off + off, len, h + off);
                    } catch (const std::exception& e) {
                        std::lock_guard<std::mutex> lk(fail_mu);
                        fail_msg = e.what();
                        failed.store(1);
                    }
                    done[i].store(1, std::memory_order_release);
                }
            });
        cudaError_t cerr = cudaSuccess;
        for (size_t i = 0; i < n_pieces; i++) {
            // while piece i is still being copied this thread copies pieces too instead of spinning
            while (!done[i].load(std::memory_order_acquire)) {
                size_t j = next.fetch_add(1);
                if (j >= n_pieces) { std::this_thread::yield(); continue; }
                size_t off = j * PIECE, len = std::min(PIECE, avail - off);
                try {
                    if (!failed.load()) src->read(win_off + off, len, h + off);
                } catch (const std::exception& e) {
                    std::lock_guard<std::mutex> lk(fail_mu);
                    fail_msg = e.what();
                    failed.store(1);
                }
                done[j].store(1, std::memory_order_release);
            }
            size_t off = i * PIECE, len = std::min(PIECE, avail - off);
            if (i + 1 == n_pieces) {
                memset(h + avail, 0, 128);
                len += 128;
            }
            if (!failed.load() && cerr == cudaSuccess) cerr = cudaMemcpyAsync(d + off, h + off, len, cudaMemcpyHostToDevice, s_copy);
        }
        for (auto& t : th) t.join();
        if (failed.load()) throw Error(MBFBAM_ERR_IO, fail_msg);
        CK(cerr);
    }

    static std::string err_text(const WinMeta& m) {
        const char* what = "unknown";
        switch (m.err) {
            case WERR_CHAIN: what = "BGZF block chain broken"; break;
            case WERR_TRUNCATED: what = "truncated BGZF block"; break;
            case WERR_ISIZE: what = "BGZF ISIZE > 64 KiB"; break;
            case WERR_RECORD: what = "malformed BAM record"; break;
            case WERR_CARRY: what = "BAM record larger than 4 MiB"; break;
            case WERR_NH_TYPE: what = "NH tag is not an integer"; break;
            case WERR_AUX: what = "malformed aux data"; break;
            case WERR_BARCODE_LEN: what = "XC barcode longer than 44 bytes"; break;
            case WERR_CAPACITY: what = "too many BGZF blocks in one window (raise MBF_BAM_WINDOW_MB?)"; break;
            default: if (m.err >= WERR_INFLATE) what = "corrupt DEFLATE stream in BGZF block"; break;
        }
        return fmt("Could not read bam: %s (code %u, info %u)", what, m.err, m.err_info);
    }

    // ---------------------------------------------------------------- main stage
    // K1 of window k runs on its own stream (two alternate) so that it can start while the tail of window
    // k-1's K1 is still draining: a BGZF block takes milliseconds, and with one launch at a time the last
    // partial wave of every window would leave most SMs idle.
    void issue_inflate(int bi) {
        WinBufs& w = wb[bi];
        CK(cudaEventSynchronize(w.ev_front));
        WinMeta m = h_meta_front.as<WinMeta>()[bi];
        if (m.err == WERR_USIZE)  // err_info = inflated size in MiB
            throw RetryJob("window inflates past 4 GiB", ((double)m.err_info * 1048576.0) / (double)std::max<uint32_t>(1, w.scan_len));
        if (m.err) throw Error(MBFBAM_ERR_FORMAT, err_text(m));
        w.front_meta = m;
        ensure_U(bi, (size_t)m.u_total + 1024, true);
        st.n_bgzf_blocks += m.n_blocks;
        st.uncompressed_bytes += m.u_total;
        st.n_windows++;
        w.two_pass = false;
        cudaStream_t si = s_inf[bi];  // alternating streams: every per-window buffer of K1 exists twice
        CK(cudaStreamWaitEvent(si, w.ev_front, 0));
        // U[bi] was last read by window k-2 (main stage + its second half); ev_reuse was recorded behind both
        // in the previous iteration of run()
        CK(cudaStreamWaitEvent(si, ev_reuse, 0));
        BlockArrays b{w.cpos.as<uint32_t>(), w.clen.as<uint32_t>(), w.isize.as<uint32_t>(), w.uoff.as<uint32_t>()};
        CK(cudaEventRecord(w.t[2], si));
        if (m.n_blocks) launch_inflate(w, b, bi, m.n_blocks, si);
        CK(cudaEventRecord(w.t[3], si));
        if (!w.two_pass) CK(cudaEventRecord(w.ev_cfree, si));
        CK(cudaEventRecord(w.ev_inf, si));
    }

    void issue_main(int bi) {
        WinBufs& w = wb[bi];
        const WinMeta m = w.front_meta;
        CK(cudaStreamWaitEvent(s_main, w.ev_inf, 0));
        BlockArrays b{w.cpos.as<uint32_t>(), w.clen.as<uint32_t>(), w.isize.as<uint32_t>(), w.uoff.as<uint32_t>()};
        CK(cudaEventRecord(w.t[6], s_main));
        if (kind == JOB_INFLATE) {
            if (out_host) {
                if (out_len + m.u_total > out_cap) throw Error(MBFBAM_ERR_ARG, "output buffer too small");
                CK(cudaMemcpyAsync(out_host + out_len, U[bi].as<uint8_t>() + U_PREFIX, m.u_total, cudaMemcpyDeviceToHost, s_main));
                st.d2h_bytes += m.u_total;
            }
            out_len += m.u_total;
            CK(cudaEventRecord(w.t[4], s_main));
            CK(cudaEventRecord(w.t[5], s_main));
        } else {
            FrameArgs fa{U[bi].as<uint8_t>(), b, nullptr, dmeta(bi), w.blk_start.as<uint32_t>(), w.blk_end.as<uint32_t>(),
                         w.blk_nrec.as<uint32_t>(), w.rec_base.as<uint32_t>(), w.rec_off.as<uint32_t>(),
                         (uint32_t)(w.rec_off.cap / 4), w.rec_tmp.as<uint32_t>(), (uint32_t)(w.rec_tmp.cap / 4)};
            fa.carry = d_carry.as<uint32_t>();
            uint32_t fgrid = std::max(1u, (m.n_blocks + 127) / 128);
            k_frame_walk<false><<<fgrid, 128, 0, s_main>>>(fa);
            launched("k_frame_walk<count>", s_main);
            k_frame_fix<<<1, 1024, 0, s_main>>>(fa);
            launched("k_frame_fix", s_main);
            k_excl_scan<<<1, 1024, 0, s_main>>>(fa.blk_nrec, fa.rec_base, &dmeta(bi)->n_blocks, 0, 0, &dmeta(bi)->n_records, nullptr, 0);
            launched("k_excl_scan(records)", s_main);
            k_frame_emit<<<std::max(1u, (m.n_blocks + 3) / 4), 128, 0, s_main>>>(fa);
            launched("k_frame_emit", s_main);
            k_frame_walk<true><<<fgrid, 128, 0, s_main>>>(fa);  // (returns at once unless a walker was off a record boundary)
            launched("k_frame_walk<emit>", s_main);
            k_carry_copy<<<1, 256, 0, s_main>>>(U[bi].as<uint8_t>(), U[bi ^ 1].as<uint8_t>(), d_carry.as<uint32_t>(), dmeta(bi));
            launched("k_carry_copy", s_main);
            CK(cudaEventRecord(w.t[4], s_main));
            if (kind == JOB_COUNT) launch_count(bi, 0);
            else if (kind == JOB_QUANTIFY) launch_quantify(bi, 1);
            else if (kind == JOB_INTRONS) launch_introns(bi, 1);
            else if (kind == JOB_BARCODE) launch_barcode(bi, 1);
            // JOB_DUPS: one insert per record, bounded by n_records: everything happens in the second half
            CK(cudaEventRecord(w.t[5], s_main));
        }
        CK(cudaMemcpyAsync(h_meta_main.as<WinMeta>() + bi, dmeta(bi), sizeof(WinMeta), cudaMemcpyDeviceToHost, s_main));
        st.d2h_bytes += sizeof(WinMeta);
        CK(cudaEventRecord(w.ev_main, s_main));
        w.timed = true;
    }

    // K1.  Pass 1 (Huffman decode -> tokens): MBF_BAM_INFLATE_D = 400 (default) k_tokens_par<128,41>, 401
    // <256,21>, 402 <128,33>, 403 <64,41>, 410/411 its single-decode form (one CTA per member, lanes share the
    // tables; members it cannot handle go to k_tokens_bf<8>); 302/304/308/316 = k_tokens_bf for every member with 2/4/8/16 decoders per
    // warp (the round-1 production kernel).  Pass 2 (LZ77): k_lz_resolve.
    void launch_inflate(WinBufs& w, const BlockArrays& b, int bi, uint32_t n_blocks, cudaStream_t si) {
        const uint32_t mode_d = env_u32("MBF_BAM_INFLATE_D", 400);
        DevBuf& tokb = d_tok[bi];
        DevBuf& ntokb = d_ntok[bi];
        tokb.ensure((U_cap[bi] + 65536) * 4);
        ntokb.ensure((size_t)(cand_cap + 8) * 4);
        d_fb[bi].ensure((size_t)(cand_cap + 8) * 4);
        d_tscratch.ensure((size_t)2 * sm_count * 32 * 32 * sizeof(TokScratch));
        TokenArgs ta{w.c_ptr, b.cpos, b.clen, b.isize, b.uoff, tokb.as<uint32_t>(), ntokb.as<uint32_t>(), U_PREFIX,
                     &dmeta(bi)->n_blocks, &dmeta(bi)->inflate_next, &dmeta(bi)->err, &dmeta(bi)->err_info,
                     d_tscratch.as<TokScratch>() + (size_t)bi * sm_count * 32 * 32,
                     &dmeta(bi)->n_tokens, nullptr};
        auto launch_bf = [&](auto kernel, int smem, int dl, uint32_t n_members) {
            CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            int per_sm = std::max(1, std::min(32, (227 * 1024) / (smem + 1024)));
            // The persistent decoders take about 4/5 of what shared memory allows (9 of 11 CTAs per SM
            // with 8 decoders per warp): the rest is room for the other window's resolve pass, scan and
            // count kernels to run beside them (the decoders are latency bound and leave issue slots
            // free).  Measured on the 50M-read bench: 11 -> 112 ms, 10 -> 110 ms, 9 -> 108 ms per job.
            const uint32_t cap = env_u32("MBF_BAM_TOKENS_PER_SM", (uint32_t)std::max(1, per_sm * 9 / 11));
            if (cap) per_sm = std::min<int>(per_sm, (int)cap);
            uint32_t grid = std::min<uint32_t>((n_members + dl - 1) / dl, (uint32_t)(sm_count * per_sm));
            kernel<<<std::max(1u, grid), 32, smem, si>>>(ta);
        };
        auto launch_par = [&](auto kernel, int nt, int sw_max, size_t smem) {
            CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 1;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, smem));
            const uint32_t cap = env_u32("MBF_BAM_PAR_PER_SM", 0);
            if (cap) per_sm = std::min<int>(per_sm, (int)cap);
            const uint32_t grid = std::min<uint32_t>(n_blocks, (uint32_t)(sm_count * std::max(1, per_sm)));
            // lane-private token scratch of the guess and sync runs (touched: ~1 KB per lane, L2 resident)
            d_pscratch[bi].ensure((size_t)grid * nt * 2 * par_scratch_cap(sw_max) * 4 + 256);  // (sw_max 0: three-decode scheme, unused)
            // first block header of every member, one thread each
            d_phdr[bi].ensure((size_t)(n_blocks + 64) * sizeof(ParHdr));
            d_phlens[bi].ensure((size_t)(n_blocks + 64) * PAR_HDR_STRIDE);
            k_par_headers<<<(n_blocks + PAR_HDR_THREADS - 1) / PAR_HDR_THREADS, PAR_HDR_THREADS, 0, si>>>(ta, d_phdr[bi].as<ParHdr>(),
                                                                                                    d_phlens[bi].as<uint8_t>());
            launched("k_par_headers", si);
            ParArgs pa{ta, d_fb[bi].as<uint32_t>(), &dmeta(bi)->fb_count, d_pscratch[bi].as<uint32_t>(), d_phdr[bi].as<ParHdr>(),
                       d_phlens[bi].as<uint8_t>()};
            kernel<<<std::max(1u, grid), nt, smem, si>>>(pa);
            launched("k_tokens_par", si);
            // the members it handed over (stored blocks, exotic tables, corrupt streams): usually none
            ta.list = d_fb[bi].as<uint32_t>();
            ta.n_blocks = &dmeta(bi)->fb_count;
            ta.queue_head = &dmeta(bi)->fb_next;
            launch_bf(k_tokens_bf<8, true>, (int)sizeof(TokensShared32<8>), 8, (uint32_t)sm_count * 8u);
            launched("k_tokens_bf(fallback)", si);
        };
        switch (mode_d) {
            // three-decode scheme (guess, sync, emit runs all decode)
            case 400: launch_par(k_tokens_par<128, 41, false>, 128, 0, sizeof(ParShared<128, 41, false>)); break;
            case 401: launch_par(k_tokens_par<256, 21, false>, 256, 0, sizeof(ParShared<256, 21, false>)); break;
            case 402: launch_par(k_tokens_par<128, 33, false>, 128, 0, sizeof(ParShared<128, 33, false>)); break;
            case 403: launch_par(k_tokens_par<64, 41, false>, 64, 0, sizeof(ParShared<64, 41, false>)); break;
            // single-decode scheme (the guess run keeps its tokens; emit is a copy instead of a third decode)
            case 410: launch_par(k_tokens_par<128, 41, true>, 128, 41, sizeof(ParShared<128, 41, true>)); break;
            case 411: launch_par(k_tokens_par<128, 33, true>, 128, 33, sizeof(ParShared<128, 33, true>)); break;
            case 302: launch_bf(k_tokens_bf<2, true>, (int)sizeof(TokensShared32<2>), 2, n_blocks); launched("k_tokens_bf", si); break;
            case 304: launch_bf(k_tokens_bf<4, true>, (int)sizeof(TokensShared32<4>), 4, n_blocks); launched("k_tokens_bf", si); break;
            case 308: launch_bf(k_tokens_bf<8, true>, (int)sizeof(TokensShared32<8>), 8, n_blocks); launched("k_tokens_bf", si); break;
            case 316: launch_bf(k_tokens_bf<16, true>, (int)sizeof(TokensShared32<16>), 16, n_blocks); launched("k_tokens_bf", si); break;
            default:
                throw Error(MBFBAM_ERR_ARG, "MBF_BAM_INFLATE_D must be 400-403, 410, 411 or 302, 304, 308, 316");
        }
        CK(cudaEventRecord(w.t[7], si));
        CK(cudaEventRecord(w.ev_cfree, si));  // pass 2 reads tokens only: the next H2D into this buffer may start
        w.two_pass = true;
        ResolveArgs ra{tokb.as<uint32_t>(), ntokb.as<uint32_t>(), b.isize, b.uoff, U_PREFIX, U[bi].as<uint8_t>(),
                       &dmeta(bi)->n_blocks, &dmeta(bi)->err, &dmeta(bi)->err_info};
        const uint32_t rgrid = (n_blocks + RESOLVE_WARPS - 1) / RESOLVE_WARPS;
        if (env_u32("MBF_BAM_RESOLVE_RING", 2048) == 1024) k_lz_resolve<1024><<<rgrid, 32 * RESOLVE_WARPS, 0, si>>>(ra);
        else k_lz_resolve<2048><<<rgrid, 32 * RESOLVE_WARPS, 0, si>>>(ra);
        launched("k_lz_resolve", si);
    }

    int count_grid() const { return sm_count * 8; }

    void launch_count(int bi, int emit_only) {
        WinBufs& w = wb[bi];
        if (!w.mm_cap) {
            w.mm_cap = std::max<uint32_t>(env_u32("MBF_BAM_MM_CAP", 4u << 20), 16);
            w.mm_fp.ensure((size_t)w.mm_cap * 16);
            w.mm_slot.ensure((size_t)w.mm_cap * 4);
        }
        CountCtx c = cc;
        c.U = U[bi].as<uint8_t>();
        c.rec_off = w.rec_off.as<uint32_t>();
        c.meta = dmeta(bi);
        c.mm_fp = w.mm_fp.as<ulonglong2>();
        c.mm_slot = w.mm_slot.as<uint32_t>();
        c.mm_cap = w.mm_cap;
        c.emit_only = emit_only;
        if (mode == MBFBAM_MODE_UNSTRANDED) k_count_reads<0><<<count_grid(), 256, 0, s_main>>>(c);
        else if (mode == MBFBAM_MODE_STRANDED) k_count_reads<1><<<count_grid(), 256, 0, s_main>>>(c);
        else k_count_reads<2><<<count_grid(), 256, 0, s_main>>>(c);
        launched("k_count_reads", s_main);
    }

    // quantify_gene_reads: pass 1 counts the (read, gene) hits of the window, pass 2 adds them to the map
    void launch_quantify(int bi, int count_only) {
        WinBufs& w = wb[bi];
        CountCtx c = cc;
        c.U = U[bi].as<uint8_t>();
        c.rec_off = w.rec_off.as<uint32_t>();
        c.meta = dmeta(bi);
        c.qmap = d_map.as<ulonglong2>();
        c.qmask = (uint32_t)(map_cap - 1);
        c.q_count_only = count_only;
        k_count_reads<3><<<count_grid(), 256, 0, s_main>>>(c);
        launched("k_count_reads<quantify>", s_main);
    }

    void launch_dups(int bi) {
        WinBufs& w = wb[bi];
        DupCtx c = dc;
        c.U = U[bi].as<uint8_t>();
        c.rec_off = w.rec_off.as<uint32_t>();
        c.meta = dmeta(bi);
        c.map = d_map.as<ulonglong2>();
        c.map_mask = (uint32_t)(map_cap - 1);
        c.owned_total = d_scalar.as<unsigned long long>() + 1;
        k_dup_insert<<<count_grid(), 256, 0, s_main>>>(c);
        launched("k_dup_insert", s_main);
    }

    // by_barcode: pass 1 counts the right-strand hits of the window, pass 2 inserts them
    void launch_barcode(int bi, int count_only) {
        WinBufs& w = wb[bi];
        BarcodeCtx b{};
        b.c = cc;
        b.c.U = U[bi].as<uint8_t>();
        b.c.rec_off = w.rec_off.as<uint32_t>();
        b.c.meta = dmeta(bi);
        b.count_only = count_only;
        b.set = d_set.as<ulonglong2>();
        b.set_mask = (uint32_t)(set_cap - 1);
        b.set_size = d_set_size.as<unsigned long long>();
        b.map = d_map.as<ulonglong2>();
        b.map_mask = (uint32_t)(map_cap - 1);
        b.list = d_bclist.as<BarcodeRec>();
        b.list_n = d_bclist_n.as<unsigned long long>();
        b.list_cap = bclist_cap;
        b.chunk_err = d_chunk_err.as<uint32_t>();
        b.job_err = d_chunk_err.as<uint32_t>() + bc_n_chunks;
        k_barcode_hits<<<count_grid(), 256, 0, s_main>>>(b);
        launched("k_barcode_hits", s_main);
    }

    // room for `extra` more (chunk, gene, barcode) records; contents are kept
    void ensure_bclist(uint64_t extra) {
        if (bclist_ub + extra <= bclist_cap) { bclist_ub += extra; return; }
        CK(cudaStreamSynchronize(s_main));
        unsigned long long live = 0;
        CK(cudaMemcpy(&live, d_bclist_n.p, 8, cudaMemcpyDeviceToHost));
        bclist_ub = live + extra;
        if (bclist_ub <= bclist_cap) return;
        uint64_t ncap = std::max<uint64_t>(1u << 16, bclist_cap);
        while (ncap < bclist_ub) ncap <<= 1;
        DevBuf nb;
        nb.ensure(ncap * sizeof(BarcodeRec));
        if (live) CK(cudaMemcpy(nb.p, d_bclist.p, (size_t)live * sizeof(BarcodeRec), cudaMemcpyDeviceToDevice));
        std::swap(d_bclist.p, nb.p);
        std::swap(d_bclist.cap, nb.cap);
        bclist_cap = ncap;
    }

    void launch_introns(int bi, int count_only) {
        WinBufs& w = wb[bi];
        IntronCtx c = ic;
        c.U = U[bi].as<uint8_t>();
        c.rec_off = w.rec_off.as<uint32_t>();
        c.meta = dmeta(bi);
        c.map = d_map.as<ulonglong2>();
        c.map_mask = (uint32_t)(map_cap - 1);
        c.count_only = count_only;
        k_count_introns<<<count_grid(), 256, 0, s_main>>>(c);
        launched("k_count_introns", s_main);
    }

    // grow the multi-mapper set so that `extra` more keys keep the load factor <= 1/2
    void ensure_set(uint64_t extra) {
        if ((set_ub + extra) * 2 <= set_cap) { set_ub += extra; return; }
        CK(cudaStreamSynchronize(s_main));
        unsigned long long live = 0;
        if (set_cap) CK(cudaMemcpy(&live, d_set_size.p, 8, cudaMemcpyDeviceToHost));
        set_ub = live + extra;
        if (set_ub * 2 <= set_cap) return;
        uint64_t ncap = 1u << 22;
        while (ncap < set_ub * 4) ncap <<= 1;
        if (ncap > (1ull << 31)) throw Error(MBFBAM_ERR_UNSUPPORTED, "multi-mapper set exceeds 2^31 slots");
        DevBuf nb;
        nb.ensure(ncap * 16);
        CK(cudaMemsetAsync(nb.p, 0, ncap * 16, s_main));
        if (set_cap) {
            k_set_rehash<<<sm_count * 8, 256, 0, s_main>>>(d_set.as<ulonglong2>(), (uint32_t)set_cap, nb.as<ulonglong2>(), (uint32_t)(ncap - 1));
            launched("k_set_rehash", s_main);
            CK(cudaStreamSynchronize(s_main));
        }
        std::swap(d_set.p, nb.p);
        std::swap(d_set.cap, nb.cap);
        set_cap = ncap;
    }

    void ensure_map(uint64_t extra) {
        if ((map_ub + extra) * 2 <= map_cap) { map_ub += extra; return; }
        CK(cudaStreamSynchronize(s_main));
        unsigned long long live = 0;
        if (map_cap) {
            CK(cudaMemsetAsync(d_scalar.p, 0, 8, s_main));
            k_map_count<<<sm_count * 8, 256, 0, s_main>>>(d_map.as<ulonglong2>(), (uint32_t)map_cap, d_scalar.as<unsigned long long>());
            launched("k_map_count", s_main);
            CK(cudaMemcpyAsync(&live, d_scalar.p, 8, cudaMemcpyDeviceToHost, s_main));
            CK(cudaStreamSynchronize(s_main));
        }
        map_ub = live + extra;
        if (map_ub * 2 <= map_cap) return;
        uint64_t ncap = 1u << 16;
        while (ncap < map_ub * 4) ncap <<= 1;
        if (ncap > (1ull << 31)) throw Error(MBFBAM_ERR_UNSUPPORTED, "intron map exceeds 2^31 slots");
        DevBuf nb;
        nb.ensure(ncap * 16);
        CK(cudaMemsetAsync(nb.p, 0xff, ncap * 16, s_main));
        if (map_cap) {
            k_map_rehash<<<sm_count * 8, 256, 0, s_main>>>(d_map.as<ulonglong2>(), (uint32_t)map_cap, nb.as<ulonglong2>(), (uint32_t)(ncap - 1));
            launched("k_map_rehash", s_main);
            CK(cudaStreamSynchronize(s_main));
        }
        std::swap(d_map.p, nb.p);
        std::swap(d_map.cap, nb.cap);
        map_cap = ncap;
    }

    // the live entries of the 16-byte-slot map, compacted on the device before they cross PCIe
    std::vector<ulonglong2> download_map() {
        std::vector<ulonglong2> out;
        if (!map_cap) return out;
        CK(cudaMemsetAsync(d_scalar.p, 0, 8, s_main));
        k_map_count<<<sm_count * 8, 256, 0, s_main>>>(d_map.as<ulonglong2>(), (uint32_t)map_cap, d_scalar.as<unsigned long long>());
        launched("k_map_count", s_main);
        unsigned long long live = 0;
        CK(cudaMemcpyAsync(&live, d_scalar.p, 8, cudaMemcpyDeviceToHost, s_main));
        CK(cudaStreamSynchronize(s_main));
        if (!live) return out;
        DevBuf packed;
        packed.ensure((size_t)live * 16);
        CK(cudaMemsetAsync(d_scalar.p, 0, 8, s_main));
        k_map_compact<<<sm_count * 8, 256, 0, s_main>>>(d_map.as<ulonglong2>(), (uint32_t)map_cap, packed.as<ulonglong2>(), live,
                                                        d_scalar.as<unsigned long long>());
        launched("k_map_compact", s_main);
        out.resize((size_t)live);
        CK(cudaMemcpyAsync(out.data(), packed.p, (size_t)live * 16, cudaMemcpyDeviceToHost, s_main));
        CK(cudaStreamSynchronize(s_main));
        st.d2h_bytes += (size_t)live * 16;
        return out;
    }

    // second half of a window's main stage: needs numbers from the first half
    void finish_main(int bi) {
        WinBufs& w = wb[bi];
        if (!w.in_flight) return;
        const bool trace = env_u32("MBF_BAM_TRACE", 0) != 0;
        double t_a = now_ms();
        CK(cudaEventSynchronize(w.ev_main));
        double t_b = now_ms();
        WinMeta m = h_meta_main.as<WinMeta>()[bi];
        if (m.err) throw Error(MBFBAM_ERR_FORMAT, err_text(m));
        st.n_records += m.n_records;
        st.n_fixups += m.n_fixups;
        st.n_tokens += m.n_tokens;
        if (kind == JOB_COUNT) {
            st.n_records_owned += m.owned_records;
            if (m.mm_count > w.mm_cap) {
                // the key buffer overflowed: enlarge it and re-run the window in emit-only mode
                CK(cudaStreamSynchronize(s_main));
                w.mm_cap = m.mm_count + 1024;
                w.mm_fp.ensure((size_t)w.mm_cap * 16);
                w.mm_slot.ensure((size_t)w.mm_cap * 4);
                CK(cudaMemsetAsync(&dmeta(bi)->mm_count, 0, 4, s_main));
                launch_count(bi, 1);
                CK(cudaMemcpyAsync(h_meta_main.as<WinMeta>() + bi, dmeta(bi), sizeof(WinMeta), cudaMemcpyDeviceToHost, s_main));
                CK(cudaStreamSynchronize(s_main));
                m = h_meta_main.as<WinMeta>()[bi];
                if (m.mm_count > w.mm_cap) throw Error(MBFBAM_ERR_CUDA, "multi-mapper key buffer overflowed twice");
            }
            if (m.mm_count) {
                st.n_mm_keys += m.mm_count;
                double t_c = now_ms();
                ensure_set(m.mm_count);
                if (trace) fprintf(stderr, "[mbfbam]   finish: wait %.2f ms, ensure_set %.2f ms (keys %u, cap %llu)\n", t_b - t_a, now_ms() - t_c, m.mm_count, (unsigned long long)set_cap);
                k_mm_insert<<<sm_count * 4, 256, 0, s_main>>>(w.mm_fp.as<ulonglong2>(), w.mm_slot.as<uint32_t>(), dmeta(bi), w.mm_cap,
                                                             d_set.as<ulonglong2>(), (uint32_t)(set_cap - 1),
                                                             d_counts.as<unsigned long long>(), (uint32_t)n_genes,
                                                             d_set_size.as<unsigned long long>());
                launched("k_mm_insert", s_main);
            }
            outside_total += m.outside;
        } else if (kind == JOB_INTRONS) {
            st.n_records_owned += m.owned_records;
            if (m.owned_records) {
                ensure_map((uint64_t)m.cigar_ops + 64);
                launch_introns(bi, 0);
            }
        } else if (kind == JOB_QUANTIFY) {
            st.n_records_owned += m.owned_records;
            if (m.cigar_ops) {
                ensure_map((uint64_t)m.cigar_ops + 64);
                launch_quantify(bi, 0);
            }
        } else if (kind == JOB_BARCODE) {
            st.n_records_owned += m.owned_records;
            if (m.cigar_ops) {
                ensure_set((uint64_t)m.cigar_ops);
                ensure_map((uint64_t)m.cigar_ops + 64);
                ensure_bclist((uint64_t)m.cigar_ops);
                launch_barcode(bi, 0);
            }
        } else if (kind == JOB_DUPS) {
            if (m.n_records) {
                ensure_map((uint64_t)m.n_records + 64);
                launch_dups(bi);
            }
        }
        if (w.timed) {
            float ms = 0;
            cudaEventElapsedTime(&ms, w.t[0], w.t[1]); st.ms_gpu_scan += ms;
            cudaEventElapsedTime(&ms, w.t[2], w.t[3]); st.ms_gpu_inflate += ms;
            if (w.two_pass) { cudaEventElapsedTime(&ms, w.t[2], w.t[7]); st.ms_gpu_tokens += ms; }
            cudaEventElapsedTime(&ms, w.t[6], w.t[4]); st.ms_gpu_frame += ms;
            cudaEventElapsedTime(&ms, w.t[4], w.t[5]); st.ms_gpu_count += ms;
            w.timed = false;
        }
        w.in_flight = false;
    }
    uint64_t outside_total = 0;

    // ---------------------------------------------------------------- driver
    void run(const std::vector<ByteRange>& ranges) {
        struct Win { uint64_t off; uint32_t scan_len; uint64_t src_end; int range_start; uint64_t range_off; uint32_t first_uoff; };
        // Windows are planned one ahead, because their size depends on what the previous ones showed:
        //  * every job starts with a W/8 window; the inflated size of a window must stay below 4 GiB (32-bit
        //    offsets into U), so later windows are capped by the largest inflate ratio seen so far;
        //  * streamed jobs grow x1.5 per window until W: the bytes of window k+1 are staged while the GPU
        //    works on window k, so a window must not take longer to stage than its predecessor takes to
        //    process (a full-size second window would leave the GPU idle for ~10 ms).
        const bool ramp = (!src->resident(device) && !(ranges.size() && src->host_pinned(ranges[0].cbeg, 1))) ||
                          (!src->resident(device) && env_u32("MBF_BAM_RAMP_PINNED", 0) != 0);
        size_t ri = 0;
        uint64_t pos = ranges.empty() ? 0 : ranges[0].cbeg;
        bool first_of_range = true;
        auto next_win = [&](uint64_t cur_w, Win& out) -> bool {
            while (ri < ranges.size() && pos >= ranges[ri].cend) {
                ri++;
                if (ri < ranges.size()) { pos = ranges[ri].cbeg; first_of_range = true; }
            }
            if (ri >= ranges.size()) return false;
            const ByteRange& r = ranges[ri];
            uint64_t base = pos & ~(uint64_t)15;
            uint64_t scan_end = std::min<uint64_t>(r.cend, base + cur_w);
            out = {base, (uint32_t)(scan_end - base), src->size(), first_of_range ? 1 : 0, r.cbeg, first_of_range ? r.first_uoff : 0u};
            first_of_range = false;
            pos = scan_end;
            return true;
        };
        // one helper thread at a time; joined (and its exception rethrown) before anything depends on it, and
        // on every way out of this function
        struct Stager {
            std::thread th;
            std::exception_ptr err;
            void start(std::function<void()> f, int dev) {
                join();
                th = std::thread([this, f, dev] {
                    try {
                        CK(cudaSetDevice(dev));  // the current device is per thread
                        f();
                    } catch (...) {
                        err = std::current_exception();
                    }
                });
            }
            void join() {
                if (th.joinable()) th.join();
                if (err) {
                    std::exception_ptr e = err;
                    err = nullptr;
                    std::rethrow_exception(e);
                }
            }
            ~Stager() {
                if (th.joinable()) th.join();
            }
        } stager;
        uint64_t cur_w = std::max<uint64_t>(W / std::max<uint32_t>(1, env_u32("MBF_BAM_FIRST_WINDOW_DIV", 8)), 1u << 20);
        double max_ratio = std::max(1.0, ratio_hint);
        // inflated bytes a window may have (MBF_BAM_U_LIMIT_MB exists so that tests can reach the cap)
        const uint64_t u_limit = (uint64_t)std::min<uint32_t>(env_u32("MBF_BAM_U_LIMIT_MB", 3276), 3276) << 20;
        cur_w = std::max<uint64_t>(1u << 20, std::min<uint64_t>(cur_w, (uint64_t)((double)u_limit / (max_ratio * 1.25)) & ~(uint64_t)0xfffff));
        Win wk{}, wn{};
        bool have = next_win(cur_w, wk);
        double t0 = now_ms();
        timeline = env_u32("MBF_BAM_TIMELINE", 0) != 0;
        CK(cudaEventRecord(ev_job_begin, s_front));
        if (have) {
            stage_front(0, wk.off, wk.scan_len, wk.src_end, wk.range_start, wk.range_off, wk.first_uoff);
            scan_front(0);
        }
        const bool trace = env_u32("MBF_BAM_TRACE", 0) != 0;
        for (size_t k = 0; have; k++) {
            int bi = (int)(k & 1);
            double ta = now_ms();
            // The host must never wait for the GPU while bytes remain to be staged (the staging rate bounds a
            // streamed job): from the second window on, window k+1 is staged BEFORE the blocking wait for
            // window k's block table.  Window 0's K1 is launched first so that the GPU starts at once.
            if (k == 0) {
                issue_inflate(bi);
                if (wk.scan_len) max_ratio = std::max(max_ratio, (double)wb[bi].front_meta.u_total / (double)wk.scan_len);
            }
            double tb = now_ms();
            // size of the next window (ratios of the windows whose block tables the host has seen)
            uint64_t cap = (uint64_t)((double)u_limit / (max_ratio * 1.25));
            cap = std::max<uint64_t>(cap & ~(uint64_t)0xfffff, 1u << 20);
            const uint64_t grow_pct = env_u32("MBF_BAM_RAMP_PCT", 150);
            cur_w = std::min<uint64_t>(std::min<uint64_t>(W, cap), ramp ? cur_w * grow_pct / 100 : W);
            // (Halving the last windows of a streamed job, so that little is left to do when the copies end, was tried:
            // 75.2 instead of 70.6 ms -- a drained pipeline pays ~1 ms of launch and host round-trip latency per window.)
            const bool have_next = next_win(cur_w, wn);
            if (have_next) {
                // bytes of window k+1 (host read + H2D) while the GPU works on windows k-1 / k -- on a helper
                // thread, so that this thread can launch K1 of window k the moment its block table is ready
                // instead of after the 10 ms the host copy of 512 MiB takes
                stager.start([this, bi, wn] {
                    stage_front(bi ^ 1, wn.off, wn.scan_len, wn.src_end, wn.range_start, wn.range_off, wn.first_uoff);
                }, device);
            }
            if (k > 0) {
                issue_inflate(bi);  // waits for window k's block table; may overlap the tail of window k-1's K1
                if (wk.scan_len) max_ratio = std::max(max_ratio, (double)wb[bi].front_meta.u_total / (double)wk.scan_len);
            }
            double tc = now_ms();
            finish_main(bi ^ 1);  // second half of window k-1 (must precede window k's carry copy)
            // everything that touches window k-1's per-window arrays is queued on s_main now
            CK(cudaEventRecord(ev_reuse, s_main));
            double td = now_ms();
            issue_main(bi);
            stager.join();  // window k+1's copies are queued (s_copy, ev_h2d recorded): its scan may be queued behind them
            if (have_next) scan_front(bi ^ 1);
            if (trace)
                fprintf(stderr, "[mbfbam] win %zu t=%.2f issue_inflate %.2f ms, stage_next %.2f ms, finish_prev %.2f ms, rest %.2f ms (scan_len %u)\n",
                        k, ta - t0, tb - ta, tc - tb, td - tc, now_ms() - td, wk.scan_len);
            wk = wn;
            have = have_next;
        }
        finish_main(0);
        finish_main(1);
        CK(cudaStreamSynchronize(s_inf[0]));
        CK(cudaStreamSynchronize(s_inf[1]));
        CK(cudaStreamSynchronize(s_main));
        CK(cudaStreamSynchronize(s_front));
        CK(cudaStreamSynchronize(s_copy));
        CK(cudaEventRecord(ev_job_end, s_main));  // every stream has drained: this is the end of the device's work
        CK(cudaEventSynchronize(ev_job_end));
        float ms_all = 0;
        CK(cudaEventElapsedTime(&ms_all, ev_job_begin, ev_job_end));
        st.ms_gpu_all = ms_all;
        print_timeline();
    }
};

// One Pipeline per device, created on first use and reused by every job on that device.  Jobs on one device
// are serialised by the slot's mutex (a job owns the pipeline's streams and buffers); jobs on different
// devices run concurrently -- that is how one call fans out over the GPUs of a box.
struct DevSlot {
    std::mutex mu;
    std::unique_ptr<Pipeline> pipe;
};
DevSlot& dev_slot(int device) {
    static std::mutex map_mu;
    static std::map<int, std::unique_ptr<DevSlot>> slots;
    std::lock_guard<std::mutex> lk(map_mu);
    auto it = slots.find(device);
    if (it == slots.end()) it = slots.emplace(device, std::unique_ptr<DevSlot>(new DevSlot())).first;
    return *it->second;
}
// caller holds slot.mu
Pipeline& slot_pipeline(DevSlot& slot, int device) {
    if (!slot.pipe) slot.pipe.reset(new Pipeline(device));
    return *slot.pipe;
}

uint64_t range_bytes(const std::vector<ByteRange>& r) {
    uint64_t n = 0;
    for (const auto& x : r) n += x.cend - x.cbeg;
    return n;
}

void copy_err(char* err, size_t errlen, const std::string& m) {
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
}

template <class F>
int guarded(char* err, size_t errlen, F&& f) {
    try {
        f();
        return MBFBAM_OK;
    } catch (const Error& e) {
        copy_err(err, errlen, e.what());
        return e.code;
    } catch (const std::exception& e) {
        copy_err(err, errlen, e.what());
        return MBFBAM_ERR_CUDA;
    }
}

// f(k) for k in [0, n): on the calling thread when n == 1, else one thread each; the first exception is rethrown
template <class F>
void fan_out(size_t n, F&& f) {
    if (n == 1) { f((size_t)0); return; }
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> errs(n);
    for (size_t k = 0; k < n; k++)
        th.emplace_back([&, k] {
            try { f(k); } catch (...) { errs[k] = std::current_exception(); }
        });
    for (auto& t : th) t.join();
    for (auto& e : errs) if (e) std::rethrow_exception(e);
}

// sums the additive fields, keeps the largest of the times (devices work side by side)
void merge_stats(mbfbam_stats& a, const mbfbam_stats& b) {
    a.n_records += b.n_records; a.n_records_owned += b.n_records_owned; a.n_bgzf_blocks += b.n_bgzf_blocks;
    a.compressed_bytes += b.compressed_bytes; a.uncompressed_bytes += b.uncompressed_bytes;
    a.h2d_bytes += b.h2d_bytes; a.d2h_bytes += b.d2h_bytes; a.n_chunks += b.n_chunks;
    a.n_kernel_launches += b.n_kernel_launches; a.n_windows += b.n_windows; a.n_mm_keys += b.n_mm_keys;
    a.n_fixups += b.n_fixups; a.n_tokens += b.n_tokens; a.h2d_pinned_bytes += b.h2d_pinned_bytes;
    a.n_devices += b.n_devices; a.n_retries += b.n_retries;
    a.ms_total = std::max(a.ms_total, b.ms_total); a.ms_gpu_scan = std::max(a.ms_gpu_scan, b.ms_gpu_scan);
    a.ms_gpu_inflate = std::max(a.ms_gpu_inflate, b.ms_gpu_inflate); a.ms_gpu_frame = std::max(a.ms_gpu_frame, b.ms_gpu_frame);
    a.ms_gpu_count = std::max(a.ms_gpu_count, b.ms_gpu_count); a.ms_gpu_other = std::max(a.ms_gpu_other, b.ms_gpu_other);
    a.ms_gpu_all = std::max(a.ms_gpu_all, b.ms_gpu_all); a.ms_gpu_tokens = std::max(a.ms_gpu_tokens, b.ms_gpu_tokens);
    a.ms_pin = std::max(a.ms_pin, b.ms_pin);
}

struct ShardPlan {
    std::vector<Chunk> chunks;       // all chunks, in contig order
    size_t lo = 0, hi = 0;           // owned [lo, hi)
    std::vector<ByteRange> ranges;
};

// contiguous chunk range of shard i of n, then the BGZF ranges that can hold its records
void plan_shard(const Reader& r, ShardPlan& p, int shard, int nshards) {
    if (nshards < 1 || shard < 0 || shard >= nshards) throw Error(MBFBAM_ERR_ARG, "bad shard index/count");
    // balance by compressed bytes: weight of a chunk = span of its linear-index offsets
    size_t n = p.chunks.size();
    std::vector<double> w(n + 1, 0.0);
    for (size_t i = 0; i < n; i++) {
        const Chunk& c = p.chunks[i];
        double wt = 1.0;
        if (r.ix->idx[(size_t)c.tid].has_reads) {
            uint64_t a = r.voff_lower(c.tid, c.start) >> 16;
            uint64_t b = c.stop >= r.lens[(size_t)c.tid] ? (r.ix->idx[(size_t)c.tid].voff_end >> 16) : (r.voff_lower(c.tid, c.stop) >> 16);
            if (b > a) wt += (double)(b - a);
        }
        w[i + 1] = w[i] + wt;
    }
    auto cut = [&](int s) -> size_t {
        if (s <= 0) return 0;
        if (s >= nshards) return n;
        double target = w[n] * (double)s / (double)nshards;
        return (size_t)(std::lower_bound(w.begin(), w.end(), target) - w.begin());
    };
    p.lo = std::min(cut(shard), n);
    p.hi = std::min(std::max(cut(shard + 1), p.lo), n);
    if (shard == nshards - 1) p.hi = n;
    // byte ranges: per contiguous run of owned chunks on one tid
    std::vector<std::pair<uint64_t, uint64_t>> vr;
    size_t i = p.lo;
    while (i < p.hi) {
        size_t j = i;
        while (j + 1 < p.hi && p.chunks[j + 1].tid == p.chunks[i].tid) j++;
        const Chunk &a = p.chunks[i], &b = p.chunks[j];
        const RefIndex& ri = r.ix->idx[(size_t)a.tid];
        if (ri.has_reads) {
            uint64_t v0 = r.voff_lower(a.tid, a.start);
            uint64_t v1 = r.voff_upper(a.tid, b.stop);
            if (v1 > v0) vr.push_back({v0, v1});
        }
        i = j + 1;
    }
    std::sort(vr.begin(), vr.end());
    for (auto& x : vr) add_range(p.ranges, x.first, x.second, r.file_size);
}


// Devices of a job: job devices[] if given, else the single `device`.
std::vector<int> job_devices(const int32_t* devices, int32_t n_devices, int32_t device) {
    std::vector<int> d;
    if (devices && n_devices > 0) {
        for (int i = 0; i < n_devices; i++) {
            if (std::find(d.begin(), d.end(), (int)devices[i]) != d.end()) throw Error(MBFBAM_ERR_ARG, "device listed twice");
            d.push_back(devices[i]);
        }
    } else {
        d.push_back(device);
    }
    return d;
}

int pin_after_jobs() {  // a file's mapping is page-locked once this many jobs have streamed it (-1: never)
    const char* v = getenv("MBF_BAM_PIN_AFTER");
    if (!v || !*v) return 3;
    if (!strcmp(v, "never")) return -1;
    return atoi(v);
}
uint64_t pin_budget_bytes() {
    const uint32_t mb = env_u32("MBF_BAM_PIN_MAX_MB", 0);
    if (mb) return (uint64_t)mb << 20;
    const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
    return pages > 0 && psz > 0 ? (uint64_t)pages * (uint64_t)psz / 2 : (8ull << 30);
}

}  // namespace

struct mbfbam_reader {
    Reader r;
    DevBuf preload;
    std::shared_ptr<FileMap> fmap;   // process-wide mapping of the file (null: pread path)
    bool pin_now = false;            // mbfbam_pin_file was called: page-lock whatever a job touches
};

namespace {

// Page-locks the byte ranges of a job when the policy says so; returns the ms spent.
double maybe_pin(mbfbam_reader* rd, const std::vector<ByteRange>& ranges, uint64_t prior_jobs) {
    if (!rd->fmap) return 0.0;
    const int after = pin_after_jobs();
    if (!(rd->pin_now || (after >= 0 && prior_jobs >= (uint64_t)after))) return 0.0;
    double ms = 0.0;
    const uint64_t budget = pin_budget_bytes();
    for (const auto& r : ranges) ms += pin_range(*rd->fmap, r.cbeg, std::min<uint64_t>(r.cend + 65536 + 64, rd->fmap->size) - r.cbeg, budget);
    return ms;
}

// runs `setup(p)` + p.run(ranges) with the retry a RetryJob asks for
template <class S>
void run_with_retry(Pipeline& p, const std::vector<ByteRange>& ranges, S&& setup) {
    p.ratio_hint = 1.0;
    for (int attempt = 0;; attempt++) {
        try {
            setup(p);
            p.run(ranges);
            p.st.n_retries = (uint64_t)attempt;
            return;
        } catch (const RetryJob& e) {
            if (attempt >= 4) throw Error(MBFBAM_ERR_UNSUPPORTED, std::string("Could not read bam: ") + e.what() + " (after retries with smaller windows)");
            p.ratio_hint = std::max(p.ratio_hint * 2.0, e.ratio * 1.3);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// count_reads: host tables shared by all devices of a job
struct CountTables {
    size_t G = 0;
    std::vector<Chunk> chunks;
    std::vector<int32_t> tid2slot;
    std::vector<int> slot_iv;  // slot -> contig index in `intervals`
    std::vector<uint32_t> slot_iv_off, ivs, ive, ivp, ivg, ivr, slot_chunk_off, chunk_stop;
};

void build_count_tables(const Reader& r, const mbfbam_count_job* job, CountTables& T) {
    const mbfbam_intervals* iv = job->intervals;
    const mbfbam_intervals* gv = job->gene_intervals;
    std::vector<uint64_t> gene_base((size_t)iv->n_contigs + 1, 0);
    for (int c = 0; c < iv->n_contigs; c++) gene_base[(size_t)c + 1] = gene_base[(size_t)c] + iv->n_genes[c];
    T.G = (size_t)gene_base[(size_t)iv->n_contigs];
    if (T.G >= (1ull << 31)) throw Error(MBFBAM_ERR_UNSUPPORTED, "too many genes");
    std::map<std::string, int> iv_index;
    for (int c = 0; c < iv->n_contigs; c++) iv_index.emplace(iv->contig_names[c], c);
    // chunks over gene_intervals ∩ header (chunked_genome.rs:17-31); every such contig must be in `intervals`
    T.chunks = make_chunks(r, gv);
    T.tid2slot.assign(r.names.size(), -1);
    for (const Chunk& c : T.chunks) {
        if (T.tid2slot[(size_t)c.tid] >= 0) continue;
        auto it = iv_index.find(r.names[(size_t)c.tid]);
        if (it == iv_index.end())
            throw Error(MBFBAM_ERR_ARG, "contig " + r.names[(size_t)c.tid] + " is in gene_intervals but not in intervals");
        T.tid2slot[(size_t)c.tid] = (int32_t)T.slot_iv.size();
        T.slot_iv.push_back(it->second);
    }
    // chunk_stop per slot, chunk ids global in chunk-table order (slots are created in that order too)
    T.slot_chunk_off.assign(T.slot_iv.size() + 1, 0);
    {
        size_t i = 0;
        for (size_t s = 0; s < T.slot_iv.size(); s++) {
            T.slot_chunk_off[s] = (uint32_t)T.chunk_stop.size();
            int32_t tid = T.chunks[i].tid;
            while (i < T.chunks.size() && T.chunks[i].tid == tid) T.chunk_stop.push_back(T.chunks[i++].stop);
        }
        T.slot_chunk_off[T.slot_iv.size()] = (uint32_t)T.chunk_stop.size();
    }
    // flattened, start-sorted interval tables per slot
    T.slot_iv_off.assign(T.slot_iv.size() + 1, 0);
    for (size_t s = 0; s < T.slot_iv.size(); s++) {
        int c = T.slot_iv[s];
        SortedIntervals si = sort_intervals(iv, c, job->each_read_counts_once != 0);
        T.slot_iv_off[s] = (uint32_t)T.ivs.size();
        for (size_t k = 0; k < si.start.size(); k++) {
            if (si.gene[k] >= iv->n_genes[c]) throw Error(MBFBAM_ERR_ARG, "gene index out of range");
            T.ivs.push_back(si.start[k]);
            T.ive.push_back(si.end[k]);
            T.ivp.push_back(si.pmax[k]);
            T.ivg.push_back((uint32_t)(gene_base[(size_t)c] + si.gene[k]) | (si.strand[k] == 1 ? 0x80000000u : 0u));
            if (job->each_read_counts_once) T.ivr.push_back(si.rank[k]);
        }
    }
    T.slot_iv_off[T.slot_iv.size()] = (uint32_t)T.ivs.size();
}

// The tables are a pure function of (BAM header, intervals, gene_intervals, each_read_counts_once), and callers count
// many files and functions against one annotation: the last few results are kept, keyed by the CONTENT of the inputs
// (a byte-for-byte copy is compared on lookup, so a hit is exact).  Building takes ~6 ms for 20 k genes / 60 k
// intervals (the replay of bio's tree for the chunk edges is most of it); a lookup ~0.1 ms.
std::mutex g_tables_mu;
std::list<std::pair<std::string, std::shared_ptr<const CountTables>>> g_tables;
void clear_count_tables() {
    std::lock_guard<std::mutex> lk(g_tables_mu);
    g_tables.clear();
}

std::shared_ptr<const CountTables> cached_count_tables(const Reader& r, const mbfbam_count_job* job) {
    std::string key;
    auto put = [&](const void* p, size_t n) { key.append((const char*)p, n); };
    auto put_iv = [&](const mbfbam_intervals* iv) {
        put(&iv->n_contigs, 4);
        for (int c = 0; c < iv->n_contigs; c++) { key.append(iv->contig_names[c]); key.push_back('\0'); }
        const size_t n = (size_t)iv->iv_offsets[iv->n_contigs];
        put(iv->n_genes, (size_t)iv->n_contigs * 4);
        put(iv->iv_offsets, ((size_t)iv->n_contigs + 1) * 8);
        put(iv->iv_start, n * 4); put(iv->iv_end, n * 4); put(iv->iv_gene, n * 4); put(iv->iv_strand, n);
    };
    const int32_t once = job->each_read_counts_once ? 1 : 0;
    put(&once, 4);
    for (size_t t = 0; t < r.names.size(); t++) { key.append(r.names[t]); key.push_back('\0'); put(&r.lens[t], 4); }
    put_iv(job->intervals);
    put_iv(job->gene_intervals);
    std::mutex& mu = g_tables_mu;
    auto& cache = g_tables;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (auto it = cache.begin(); it != cache.end(); ++it)
            if (it->first.size() == key.size() && it->first == key) {
                cache.splice(cache.begin(), cache, it);
                return cache.front().second;
            }
    }
    auto T = std::make_shared<CountTables>();
    build_count_tables(r, job, *T);
    std::lock_guard<std::mutex> lk(mu);
    cache.emplace_front(std::move(key), T);
    while (cache.size() > 4) cache.pop_back();
    return T;
}

void upload_count_tables(Pipeline& p, const CountTables& T, const ShardPlan& plan, bool with_rank) {
    CountCtx& c = p.cc;
    c.tid2slot = upload(p.d_tid2slot, T.tid2slot, p.s_main);
    c.n_ref = (int32_t)T.tid2slot.size();
    c.slot_iv_off = upload(p.d_slot_iv_off, T.slot_iv_off, p.s_main);
    c.iv_start = upload(p.d_iv_start, T.ivs, p.s_main);
    c.iv_end = upload(p.d_iv_end, T.ive, p.s_main);
    c.iv_pmax = upload(p.d_iv_pmax, T.ivp, p.s_main);
    c.iv_gene = upload(p.d_iv_gene, T.ivg, p.s_main);
    c.iv_rank = with_rank ? upload(p.d_iv_rank, T.ivr, p.s_main) : nullptr;
    c.slot_chunk_off = upload(p.d_slot_chunk_off, T.slot_chunk_off, p.s_main);
    c.chunk_stop = upload(p.d_chunk_stop, T.chunk_stop, p.s_main);
    c.chunk_lo = (uint32_t)plan.lo;
    c.chunk_hi = (uint32_t)plan.hi;
    c.n_genes = (uint32_t)T.G;
}

struct CountPart {
    std::vector<unsigned long long> h;
    std::vector<double> spill;
    uint64_t outside = 0;
    mbfbam_stats st{};
};

void count_on_device(mbfbam_reader* rd, const CountTables& T, const mbfbam_count_job* job, int device, int shard, int nshards,
                     int read_threads, uint64_t prior_jobs, CountPart& out) {
    const Reader& r = rd->r;
    const double t_a = now_ms();
    ShardPlan plan;
    plan.chunks = T.chunks;
    plan_shard(r, plan, shard, nshards);
    const double t_b = now_ms();
    const double ms_pin = maybe_pin(rd, plan.ranges, prior_jobs);
    const double t_c = now_ms();
    double t_setup = 0, t_run = 0;
    const size_t G = T.G;
    const size_t n_words = (job->mode == MBFBAM_MODE_WEIGHTED ? (size_t)NH_DENSE * 2 : 2) * std::max<size_t>(G, 1);
    FileSource src(&r, rd->fmap);
    DevSlot& slot = dev_slot(device);
    std::lock_guard<std::mutex> lk(slot.mu);
    Pipeline& p = slot_pipeline(slot, device);
    p.read_threads = read_threads;
    run_with_retry(p, plan.ranges, [&](Pipeline& p) {
        p.begin_job(&src, JOB_COUNT, range_bytes(plan.ranges));
        p.mode = job->mode;
        p.n_genes = G;
        p.n_count_words = n_words;
        p.d_counts.ensure(n_words * 8);
        CK(cudaMemsetAsync(p.d_counts.p, 0, n_words * 8, p.s_main));
        p.d_wspill.ensure(std::max<size_t>(G, 1) * 16);
        CK(cudaMemsetAsync(p.d_wspill.p, 0, std::max<size_t>(G, 1) * 16, p.s_main));
        p.d_set_size.ensure(8);
        CK(cudaMemsetAsync(p.d_set_size.p, 0, 8, p.s_main));
        upload_count_tables(p, T, plan, job->each_read_counts_once != 0);
        CountCtx& c = p.cc;
        c.counts = p.d_counts.as<unsigned long long>();
        c.wspill = p.d_wspill.as<double>();
        CK(cudaStreamSynchronize(p.s_main));  // the tables are read by kernels on other streams too
        p.st.n_chunks = plan.hi - plan.lo;
        t_setup = now_ms();
    });
    t_run = now_ms();
    out.h.resize(n_words);
    CK(cudaMemcpy(out.h.data(), p.d_counts.p, n_words * 8, cudaMemcpyDeviceToHost));
    p.st.d2h_bytes += n_words * 8;
    if (job->mode == MBFBAM_MODE_WEIGHTED) {
        out.spill.resize(2 * std::max<size_t>(G, 1));
        CK(cudaMemcpy(out.spill.data(), p.d_wspill.p, out.spill.size() * 8, cudaMemcpyDeviceToHost));
        p.st.d2h_bytes += out.spill.size() * 8;
    }
    out.outside = p.outside_total;
    p.st.ms_pin = ms_pin;
    p.st.n_devices = 1;
    out.st = p.st;
    if (env_u32("MBF_BAM_TRACE", 0))
        fprintf(stderr, "[mbfbam] device %d: plan %.2f ms, pin check %.2f ms, lock+setup %.2f ms, run %.2f ms (gpu %.2f), results %.2f ms\n",
                device, t_b - t_a, t_c - t_b, t_setup - t_c, t_run - t_setup, p.st.ms_gpu_all, now_ms() - t_run);
}

// ------------------------------------------------------------------------------------------------
struct IntronEntry {
    int32_t tid;
    uint32_t start, stop;
    uint64_t count;
};
struct IntronPart {
    std::vector<IntronEntry> entries;
    std::vector<uint8_t> seen;
    mbfbam_stats st{};
};

void introns_on_device(mbfbam_reader* rd, const std::vector<Chunk>& chunks, const std::vector<uint32_t>& ref_chunk_off, int device,
                       int shard, int nshards, int read_threads, uint64_t prior_jobs, IntronPart& out) {
    const Reader& r = rd->r;
    ShardPlan plan;
    plan.chunks = chunks;
    plan_shard(r, plan, shard, nshards);
    const double ms_pin = maybe_pin(rd, plan.ranges, prior_jobs);
    const size_t n_ref = r.names.size();
    FileSource src(&r, rd->fmap);
    DevSlot& slot = dev_slot(device);
    std::lock_guard<std::mutex> lk(slot.mu);
    Pipeline& p = slot_pipeline(slot, device);
    p.read_threads = read_threads;
    run_with_retry(p, plan.ranges, [&](Pipeline& p) {
        p.begin_job(&src, JOB_INTRONS, range_bytes(plan.ranges));
        p.d_hist.ensure(std::max<size_t>(1, n_ref) * HIST_DENSE * 8);
        CK(cudaMemsetAsync(p.d_hist.p, 0, std::max<size_t>(1, n_ref) * HIST_DENSE * 8, p.s_main));
        p.d_scalar.ensure(16);
        p.ic.n_ref = (int32_t)n_ref;
        p.ic.ref_chunk_off = upload(p.d_ref_chunk_off, ref_chunk_off, p.s_main);
        p.ic.chunk_lo = (uint32_t)plan.lo;
        p.ic.chunk_hi = (uint32_t)plan.hi;
        p.ic.hist = p.d_hist.as<unsigned long long>();
        CK(cudaStreamSynchronize(p.s_main));
        p.ensure_map(1024);
        p.st.n_chunks = plan.hi - plan.lo;
    });
    // pull the live map entries and the dense histogram
    std::vector<ulonglong2> slots = p.download_map();
    std::vector<unsigned long long> hist(std::max<size_t>(1, n_ref) * HIST_DENSE);
    CK(cudaMemcpy(hist.data(), p.d_hist.p, hist.size() * 8, cudaMemcpyDeviceToHost));
    p.st.d2h_bytes += hist.size() * 8;
    for (const auto& s : slots) {
        out.entries.push_back({(int32_t)(uint32_t)s.y, (uint32_t)s.x, (uint32_t)(s.x >> 32), s.y >> 32});
    }
    for (size_t t = 0; t < n_ref; t++)
        for (int q = 0; q < HIST_DENSE; q++)
            if (hist[t * HIST_DENSE + q]) out.entries.push_back({(int32_t)t, 0xffffffffu, 0xffffffffu - (uint32_t)q, hist[t * HIST_DENSE + q]});
    // introns.rs:75-79: every chunk that was scanned inserts its contig key (possibly empty dict)
    out.seen.assign(n_ref + 1, 0);
    for (size_t i = plan.lo; i < plan.hi; i++) out.seen[(size_t)plan.chunks[i].tid] = 1;
    p.st.ms_pin = ms_pin;
    p.st.n_devices = 1;
    out.st = p.st;
}

// ------------------------------------------------------------------------------------------------
// quantify_gene_reads (reference src/count_reads/quantify.rs:98-146): (gene, bucket, pos) -> reads
struct QuantEntry {
    uint32_t gene, pos, count;
    uint8_t bucket;
};
struct QuantPart {
    std::vector<QuantEntry> entries;
    mbfbam_stats st{};
};

void quantify_on_device(mbfbam_reader* rd, const CountTables& T, int device, int shard, int nshards, int read_threads,
                        uint64_t prior_jobs, QuantPart& out) {
    const Reader& r = rd->r;
    ShardPlan plan;
    plan.chunks = T.chunks;
    plan_shard(r, plan, shard, nshards);
    const double ms_pin = maybe_pin(rd, plan.ranges, prior_jobs);
    FileSource src(&r, rd->fmap);
    DevSlot& slot = dev_slot(device);
    std::lock_guard<std::mutex> lk(slot.mu);
    Pipeline& p = slot_pipeline(slot, device);
    p.read_threads = read_threads;
    run_with_retry(p, plan.ranges, [&](Pipeline& p) {
        p.begin_job(&src, JOB_QUANTIFY, range_bytes(plan.ranges));
        p.mode = 3;
        p.d_scalar.ensure(16);
        upload_count_tables(p, T, plan, false);
        CK(cudaStreamSynchronize(p.s_main));
        p.ensure_map(1024);
        p.st.n_chunks = plan.hi - plan.lo;
    });
    for (const ulonglong2& s : p.download_map())
        out.entries.push_back({(uint32_t)(s.x >> 32), (uint32_t)s.x, (uint32_t)(s.y >> 32), (uint8_t)((uint32_t)s.y & 1u)});
    p.st.ms_pin = ms_pin;
    p.st.n_devices = 1;
    out.st = p.st;
}

// ------------------------------------------------------------------------------------------------
// count_reads_primary_only_right_strand_only_by_barcode (reference src/count_reads/by_barcode.rs:18-92, :101-182)
struct BarcodeEntry {
    uint32_t gene, chunk, count;
    std::string barcode;
};
struct BarcodePart {
    std::vector<BarcodeEntry> entries;
    mbfbam_stats st{};
};

void barcode_on_device(mbfbam_reader* rd, const CountTables& T, int device, int shard, int nshards, int read_threads,
                       uint64_t prior_jobs, BarcodePart& out) {
    const Reader& r = rd->r;
    ShardPlan plan;
    plan.chunks = T.chunks;
    plan_shard(r, plan, shard, nshards);
    const double ms_pin = maybe_pin(rd, plan.ranges, prior_jobs);
    FileSource src(&r, rd->fmap);
    DevSlot& slot = dev_slot(device);
    std::lock_guard<std::mutex> lk(slot.mu);
    Pipeline& p = slot_pipeline(slot, device);
    p.read_threads = read_threads;
    const size_t n_chunks = std::max<size_t>(1, T.chunks.size());
    run_with_retry(p, plan.ranges, [&](Pipeline& p) {
        p.begin_job(&src, JOB_BARCODE, range_bytes(plan.ranges));
        p.mode = MBFBAM_MODE_STRANDED;
        p.d_scalar.ensure(16);
        p.d_set_size.ensure(8);
        CK(cudaMemsetAsync(p.d_set_size.p, 0, 8, p.s_main));
        p.d_bclist_n.ensure(8);
        CK(cudaMemsetAsync(p.d_bclist_n.p, 0, 8, p.s_main));
        p.d_chunk_err.ensure((n_chunks + 1) * 4);
        CK(cudaMemsetAsync(p.d_chunk_err.p, 0, (n_chunks + 1) * 4, p.s_main));
        p.bc_n_chunks = n_chunks;
        upload_count_tables(p, T, plan, false);
        CK(cudaStreamSynchronize(p.s_main));
        p.ensure_set(1024);
        p.ensure_map(1024);
        p.ensure_bclist(1024);
        p.st.n_chunks = plan.hi - plan.lo;
    });
    // (chunk, gene, barcode) fingerprints -> counts, and what the fingerprints stand for
    const std::vector<ulonglong2> slots = p.download_map();
    unsigned long long n_rec = 0;
    CK(cudaMemcpy(&n_rec, p.d_bclist_n.p, 8, cudaMemcpyDeviceToHost));
    if (n_rec > p.bclist_cap) throw Error(MBFBAM_ERR_CUDA, "barcode list overflow");
    std::vector<BarcodeRec> recs((size_t)n_rec);
    if (n_rec) CK(cudaMemcpy(recs.data(), p.d_bclist.p, (size_t)n_rec * sizeof(BarcodeRec), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> cerr(n_chunks + 1);
    CK(cudaMemcpy(cerr.data(), p.d_chunk_err.p, (n_chunks + 1) * 4, cudaMemcpyDeviceToHost));
    if (cerr[n_chunks] == WERR_BARCODE_LEN) throw Error(MBFBAM_ERR_UNSUPPORTED, "XC barcode longer than 44 bytes");
    if (cerr[n_chunks]) throw Error(MBFBAM_ERR_FORMAT, "malformed aux data");
    p.st.d2h_bytes += (size_t)n_rec * sizeof(BarcodeRec) + n_chunks * 4 + 8;
    if (slots.size() != recs.size()) throw Error(MBFBAM_ERR_CUDA, "barcode map and list disagree");
    std::map<std::pair<unsigned long long, uint32_t>, uint32_t> cnt;
    for (const ulonglong2& s : slots) cnt[{s.x, (uint32_t)s.y}] = (uint32_t)(s.y >> 32);
    for (const BarcodeRec& b : recs) {
        if (b.chunk < n_chunks && cerr[b.chunk]) continue;  // the reference drops a chunk whose scan returned Err
        auto it = cnt.find({b.key, b.tag});
        if (it == cnt.end()) throw Error(MBFBAM_ERR_CUDA, "barcode record without a map entry");
        out.entries.push_back({b.gene, b.chunk, it->second, std::string((const char*)b.bc, std::min<uint32_t>(b.len, BARCODE_MAX))});
    }
    p.st.ms_pin = ms_pin;
    p.st.n_devices = 1;
    out.st = p.st;
}

// ------------------------------------------------------------------------------------------------
// calculate_duplicate_distribution (reference src/duplicate_distribution.rs:44-86)
struct DupPart {
    std::vector<unsigned long long> dense;  // DUP_DENSE counters
    std::vector<unsigned long long> big;    // multiplicities >= DUP_DENSE, one entry each
    mbfbam_stats st{};
};

void dups_on_device(mbfbam_reader* rd, const std::vector<Chunk>& chunks, const std::vector<uint32_t>& ref_chunk_off, int device,
                    int shard, int nshards, int read_threads, uint64_t prior_jobs, DupPart& out) {
    const Reader& r = rd->r;
    ShardPlan plan;
    plan.chunks = chunks;
    plan_shard(r, plan, shard, nshards);
    const double ms_pin = maybe_pin(rd, plan.ranges, prior_jobs);
    const size_t n_ref = r.names.size();
    FileSource src(&r, rd->fmap);
    DevSlot& slot = dev_slot(device);
    std::lock_guard<std::mutex> lk(slot.mu);
    Pipeline& p = slot_pipeline(slot, device);
    p.read_threads = read_threads;
    const uint32_t BIG_CAP = 1u << 20;
    run_with_retry(p, plan.ranges, [&](Pipeline& p) {
        p.begin_job(&src, JOB_DUPS, range_bytes(plan.ranges));
        p.d_scalar.ensure(16);
        CK(cudaMemsetAsync(p.d_scalar.p, 0, 16, p.s_main));
        p.dc.n_ref = (int32_t)n_ref;
        p.dc.ref_chunk_off = upload(p.d_ref_chunk_off, ref_chunk_off, p.s_main);
        p.dc.ref_len = upload(p.d_ref_len, r.lens, p.s_main);
        p.dc.chunk_lo = (uint32_t)plan.lo;
        p.dc.chunk_hi = (uint32_t)plan.hi;
        CK(cudaStreamSynchronize(p.s_main));
        p.ensure_map(1024);
        p.st.n_chunks = plan.hi - plan.lo;
    });
    p.d_dup_dense.ensure((size_t)DUP_DENSE * 8);
    p.d_dup_big.ensure((size_t)BIG_CAP * 8);
    p.d_dup_nbig.ensure(16);
    CK(cudaMemsetAsync(p.d_dup_dense.p, 0, (size_t)DUP_DENSE * 8, p.s_main));
    CK(cudaMemsetAsync(p.d_dup_nbig.p, 0, 16, p.s_main));
    k_dup_hist<<<p.sm_count * 8, 256, 0, p.s_main>>>(p.d_map.as<ulonglong2>(), (uint32_t)p.map_cap, p.d_dup_dense.as<unsigned long long>(),
                                                     p.d_dup_big.as<unsigned long long>(), BIG_CAP, p.d_dup_nbig.as<uint32_t>());
    p.launched("k_dup_hist", p.s_main);
    out.dense.resize(DUP_DENSE);
    uint32_t n_big = 0;
    unsigned long long owned[2] = {0, 0};
    CK(cudaMemcpyAsync(out.dense.data(), p.d_dup_dense.p, (size_t)DUP_DENSE * 8, cudaMemcpyDeviceToHost, p.s_main));
    CK(cudaMemcpyAsync(&n_big, p.d_dup_nbig.p, 4, cudaMemcpyDeviceToHost, p.s_main));
    CK(cudaMemcpyAsync(owned, p.d_scalar.p, 16, cudaMemcpyDeviceToHost, p.s_main));
    CK(cudaStreamSynchronize(p.s_main));
    if (n_big > BIG_CAP) throw Error(MBFBAM_ERR_UNSUPPORTED, "more than 2^20 distinct reads with 65536 or more copies");
    out.big.resize(n_big);
    if (n_big) CK(cudaMemcpy(out.big.data(), p.d_dup_big.p, (size_t)n_big * 8, cudaMemcpyDeviceToHost));
    p.st.d2h_bytes += (size_t)DUP_DENSE * 8 + (size_t)n_big * 8 + 20;
    p.st.n_records_owned = owned[1];
    p.st.ms_pin = ms_pin;
    p.st.n_devices = 1;
    out.st = p.st;
}

int default_read_threads(size_t n_devices) {
    const uint32_t env = env_u32("MBF_BAM_READ_THREADS", 0);
    if (env) return (int)env;
    const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    return (int)std::max<size_t>(2, std::min<size_t>(8, hw / 2 / std::max<size_t>(1, n_devices)));
}

}  // namespace

extern "C" {

const char* mbfbam_version(void) { return MBFBAM_VERSION; }

int32_t mbfbam_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int mbfbam_open(const char* bam_path, const char* bai_path, mbfbam_reader** out, char* err, size_t errlen) {
    *out = nullptr;
    return guarded(err, errlen, [&] {
        std::unique_ptr<mbfbam_reader> h(new mbfbam_reader());
        h->r.open(bam_path, bai_path);
        const char* stage = getenv("MBF_BAM_STAGE");  // "pread": never map the file
        if (!(stage && !strcmp(stage, "pread"))) h->fmap = FileMapCache::get().lookup(h->r.fd);
        *out = h.release();
    });
}

void mbfbam_close(mbfbam_reader* r) {
    if (!r) return;
    if (r->r.preload_device >= 0) cudaSetDevice(r->r.preload_device);
    delete r;  // the file mapping (and its page locks) stay in the process-wide cache
}

int32_t mbfbam_n_targets(const mbfbam_reader* r) { return (int32_t)r->r.names.size(); }
const char* mbfbam_target_name(const mbfbam_reader* r, int32_t tid) {
    return tid >= 0 && (size_t)tid < r->r.names.size() ? r->r.names[(size_t)tid].c_str() : nullptr;
}
uint32_t mbfbam_target_len(const mbfbam_reader* r, int32_t tid) {
    return tid >= 0 && (size_t)tid < r->r.lens.size() ? r->r.lens[(size_t)tid] : 0;
}
uint64_t mbfbam_file_size(const mbfbam_reader* r) { return r->r.file_size; }
int32_t mbfbam_tid(const mbfbam_reader* r, const char* name) {
    auto it = r->r.name2tid.find(name);
    return it == r->r.name2tid.end() ? -1 : it->second;
}

int mbfbam_preload(mbfbam_reader* r, int32_t device, char* err, size_t errlen) {
    return guarded(err, errlen, [&] {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) throw Error(MBFBAM_ERR_CUDA, "no such CUDA device");
        DevSlot& slot = dev_slot(device);
        std::lock_guard<std::mutex> lk(slot.mu);
        CK(cudaSetDevice(device));
        r->preload.ensure(r->r.file_size + 65536 + 4096);
        CK(cudaMemset((uint8_t*)r->preload.p + r->r.file_size, 0, 65536 + 4096));
        PinnedBuf stage;
        const size_t CH = 64u << 20;
        stage.ensure(CH);
        for (uint64_t off = 0; off < r->r.file_size; off += CH) {
            size_t n2 = (size_t)std::min<uint64_t>(CH, r->r.file_size - off);
            r->r.pread_exact(stage.p, off, n2);
            CK(cudaMemcpy((uint8_t*)r->preload.p + off, stage.p, n2, cudaMemcpyHostToDevice));
        }
        r->r.preload_device = device;
        r->r.preload_ptr = (uint8_t*)r->preload.p;
    });
}

int mbfbam_pin_file(mbfbam_reader* r, char* err, size_t errlen) {
    return guarded(err, errlen, [&] {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) throw Error(MBFBAM_ERR_CUDA, "no CUDA device available; mbf_bam_b200 has no CPU path");
        if (!r->fmap) throw Error(MBFBAM_ERR_UNSUPPORTED, "cannot page-lock the file mapping: the file is not mapped");
        pin_range(*r->fmap, 0, r->fmap->size, pin_budget_bytes());
        if (!r->fmap->seg_pinned(0, r->fmap->size))
            throw Error(MBFBAM_ERR_UNSUPPORTED, "cannot page-lock the file mapping (refused by the platform or over MBF_BAM_PIN_MAX_MB)");
        r->pin_now = true;
    });
}

void mbfbam_unpin_file(mbfbam_reader* r) {
    r->pin_now = false;
    if (r->fmap) unpin_all(*r->fmap);
}

void mbfbam_host_cache_clear(void) {
    FileMapCache::get().clear();
    IndexCache::get().clear();
    clear_count_tables();
}

void mbfbam_release_device(int32_t device) {
    DevSlot& slot = dev_slot(device);
    std::lock_guard<std::mutex> lk(slot.mu);
    slot.pipe.reset();
}

void mbfbam_drop_preload(mbfbam_reader* r) {
    if (r->r.preload_device >= 0) cudaSetDevice(r->r.preload_device);
    r->r.preload_device = -1;
    r->r.preload_ptr = nullptr;
    DevBuf tmp;
    std::swap(tmp.p, r->preload.p);
    std::swap(tmp.cap, r->preload.cap);
}

int64_t mbfbam_chunks(const mbfbam_reader* r, const mbfbam_intervals* gene_intervals, int32_t* tid, uint32_t* start,
                      uint32_t* stop, int64_t cap) {
    try {
        std::vector<Chunk> c = make_chunks(r->r, gene_intervals);
        for (size_t i = 0; i < c.size() && (int64_t)i < cap; i++) {
            tid[i] = c[i].tid;
            start[i] = c[i].start;
            stop[i] = c[i].stop;
        }
        return (int64_t)c.size();
    } catch (...) {
        return -1;
    }
}

int64_t mbfbam_plan_shard(const mbfbam_reader* r, const mbfbam_intervals* gene_intervals, int32_t shard_index,
                          int32_t shard_count, int64_t* chunk_lo, int64_t* chunk_hi, uint64_t* cbeg, uint64_t* cend,
                          int64_t cap) {
    try {
        ShardPlan plan;
        plan.chunks = make_chunks(r->r, gene_intervals);
        plan_shard(r->r, plan, shard_index, shard_count < 1 ? 1 : shard_count);
        if (chunk_lo) *chunk_lo = (int64_t)plan.lo;
        if (chunk_hi) *chunk_hi = (int64_t)plan.hi;
        for (size_t i = 0; i < plan.ranges.size() && (int64_t)i < cap; i++) {
            cbeg[i] = plan.ranges[i].cbeg;
            cend[i] = plan.ranges[i].cend;
        }
        return (int64_t)plan.ranges.size();
    } catch (...) {
        return -1;
    }
}

int mbfbam_count_reads(mbfbam_reader* rd, const mbfbam_count_job* job, mbfbam_counts* out, mbfbam_stats* stats, char* err,
                       size_t errlen) {
    if (stats) memset(stats, 0, sizeof *stats);
    return guarded(err, errlen, [&] {
        const double t0 = now_ms();
        const Reader& r = rd->r;
        const mbfbam_intervals* iv = job->intervals;
        if (!iv || !job->gene_intervals || !out) throw Error(MBFBAM_ERR_ARG, "null argument");
        if (job->mode < 0 || job->mode > 2) throw Error(MBFBAM_ERR_ARG, "bad mode");
        if (job->each_read_counts_once && job->mode == MBFBAM_MODE_WEIGHTED)
            throw Error(MBFBAM_ERR_UNSUPPORTED, "each_read_counts_once=True is not defined for count_reads_weighted");
        int ndev_have = 0;
        if (cudaGetDeviceCount(&ndev_have) != cudaSuccess || ndev_have == 0)
            throw Error(MBFBAM_ERR_CUDA, "no CUDA device available; mbf_bam_b200 has no CPU path");
        const std::vector<int> devs = job_devices(job->devices, job->n_devices, job->device);
        const std::shared_ptr<const CountTables> Tp = cached_count_tables(r, job);
        const CountTables& T = *Tp;
        const double t_tables = now_ms();
        if (out->scanned) {
            memset(out->scanned, 0, (size_t)iv->n_contigs);
            for (int c : T.slot_iv) out->scanned[c] = 1;
        }
        const int nsh = std::max(1, job->shard_count), sh = job->shard_index;
        if (sh < 0 || sh >= nsh) throw Error(MBFBAM_ERR_ARG, "bad shard index/count");
        const uint64_t prior = rd->fmap ? rd->fmap->jobs.fetch_add(1) : 0;
        // the caller's shard is cut once more, one sub-shard per device (chunk ranges, like everything else)
        std::vector<CountPart> parts(devs.size());
        const int rt = default_read_threads(devs.size());
        fan_out(devs.size(), [&](size_t k) {
            count_on_device(rd, T, job, devs[k], sh * (int)devs.size() + (int)k, nsh * (int)devs.size(), rt, prior, parts[k]);
        });
        const double t_dev = now_ms();
        // gather on the host (dense vectors are summed; multi-mapper sets are chunk-local, so shards
        // never share a set)
        const size_t G = T.G;
        std::vector<unsigned long long> h(parts[0].h.size(), 0);
        std::vector<double> sp(2 * std::max<size_t>(G, 1), 0.0);
        mbfbam_stats st{};
        out->outside = 0;
        for (const CountPart& pt : parts) {
            for (size_t i = 0; i < h.size(); i++) h[i] += pt.h[i];
            for (size_t i = 0; i < pt.spill.size(); i++) sp[i] += pt.spill[i];
            out->outside += pt.outside;
            merge_stats(st, pt.st);
        }
        if (job->mode == MBFBAM_MODE_WEIGHTED) {
            for (size_t g = 0; g < G; g++) {
                for (int b = 0; b < 2; b++) {
                    double acc = 0.0;
                    for (int k = 0; k < NH_DENSE; k++) {
                        unsigned long long n = h[((size_t)k * 2 + b) * G + g];
                        if (n) acc += (double)n / (double)(k + 1);
                    }
                    acc += sp[(size_t)b * G + g];
                    double* dst = b ? out->wrev : out->wfwd;
                    if (dst) dst[g] = acc;
                }
                if (out->fwd) out->fwd[g] = 0;
                if (out->rev) out->rev[g] = 0;
            }
        } else {
            for (size_t g = 0; g < G; g++) {
                if (out->fwd) out->fwd[g] = h[g];
                if (out->rev) out->rev[g] = h[G + g];
            }
        }
        st.ms_total = now_ms() - t0;
        if (env_u32("MBF_BAM_TRACE", 0))
            fprintf(stderr, "[mbfbam] count_reads: tables %.2f ms, devices %.2f ms, gather %.2f ms\n", t_tables - t0, t_dev - t_tables, now_ms() - t_dev);
        if (stats) *stats = st;
    });
}

int mbfbam_count_introns_multi(mbfbam_reader* rd, const int32_t* devices, int32_t n_devices, int32_t shard_index,
                               int32_t shard_count, mbfbam_introns* out, mbfbam_stats* stats, char* err, size_t errlen) {
    if (stats) memset(stats, 0, sizeof *stats);
    if (out) memset(out, 0, sizeof *out);
    return guarded(err, errlen, [&] {
        const double t0 = now_ms();
        const Reader& r = rd->r;
        if (!out) throw Error(MBFBAM_ERR_ARG, "null argument");
        int ndev_have = 0;
        if (cudaGetDeviceCount(&ndev_have) != cudaSuccess || ndev_have == 0)
            throw Error(MBFBAM_ERR_CUDA, "no CUDA device available; mbf_bam_b200 has no CPU path");
        const std::vector<int> devs = job_devices(devices, n_devices, 0);
        std::vector<Chunk> chunks = make_chunks(r, nullptr);
        std::vector<uint32_t> ref_chunk_off(r.names.size() + 1, 0);
        for (const Chunk& c : chunks) ref_chunk_off[(size_t)c.tid + 1]++;
        for (size_t i = 0; i < r.names.size(); i++) ref_chunk_off[i + 1] += ref_chunk_off[i];
        const int nsh = std::max(1, shard_count), sh = shard_index;
        if (sh < 0 || sh >= nsh) throw Error(MBFBAM_ERR_ARG, "bad shard index/count");
        const uint64_t prior = rd->fmap ? rd->fmap->jobs.fetch_add(1) : 0;
        std::vector<IntronPart> parts(devs.size());
        const int rt = default_read_threads(devs.size());
        fan_out(devs.size(), [&](size_t k) {
            introns_on_device(rd, chunks, ref_chunk_off, devs[k], sh * (int)devs.size() + (int)k, nsh * (int)devs.size(), rt, prior, parts[k]);
        });
        // merge: a contig split between two devices can yield the same (tid, start, stop) twice (introns.rs:12-22
        // adds such entries); the result holds every key once
        const size_t n_ref = r.names.size();
        std::vector<IntronEntry> all;
        mbfbam_stats st{};
        out->contig_seen = (uint8_t*)calloc(n_ref + 1, 1);
        for (const IntronPart& pt : parts) {
            all.insert(all.end(), pt.entries.begin(), pt.entries.end());
            for (size_t t = 0; t < n_ref; t++) out->contig_seen[t] |= pt.seen[t];
            merge_stats(st, pt.st);
        }
        size_t n = all.size();
        if (parts.size() > 1) {  // (one device: the map's keys are unique already)
            std::sort(all.begin(), all.end(), [](const IntronEntry& a, const IntronEntry& b) {
                if (a.tid != b.tid) return a.tid < b.tid;
                if (a.start != b.start) return a.start < b.start;
                return a.stop < b.stop;
            });
            n = 0;
            for (size_t i = 0; i < all.size(); i++) {
                if (n && all[n - 1].tid == all[i].tid && all[n - 1].start == all[i].start && all[n - 1].stop == all[i].stop) all[n - 1].count += all[i].count;
                else all[n++] = all[i];
            }
        }
        out->tid = (int32_t*)malloc((n + 1) * 4);
        out->start = (uint32_t*)malloc((n + 1) * 4);
        out->stop = (uint32_t*)malloc((n + 1) * 4);
        out->count = (uint64_t*)malloc((n + 1) * 8);
        out->n_targets = (int32_t)n_ref;
        for (size_t k = 0; k < n; k++) {
            out->tid[k] = all[k].tid;
            out->start[k] = all[k].start;
            out->stop[k] = all[k].stop;
            out->count[k] = all[k].count;
        }
        out->n = n;
        st.ms_total = now_ms() - t0;
        if (stats) *stats = st;
    });
}

int mbfbam_count_introns(mbfbam_reader* rd, int32_t device, int32_t shard_index, int32_t shard_count, mbfbam_introns* out,
                         mbfbam_stats* stats, char* err, size_t errlen) {
    return mbfbam_count_introns_multi(rd, &device, 1, shard_index, shard_count, out, stats, err, errlen);
}

int mbfbam_quantify_gene_reads(mbfbam_reader* rd, const mbfbam_count_job* job, mbfbam_gene_pos_counts* out, mbfbam_stats* stats,
                               char* err, size_t errlen) {
    if (stats) memset(stats, 0, sizeof *stats);
    if (out) memset(out, 0, sizeof *out);
    return guarded(err, errlen, [&] {
        const double t0 = now_ms();
        const Reader& r = rd->r;
        if (!job || !job->intervals || !job->gene_intervals || !out) throw Error(MBFBAM_ERR_ARG, "null argument");
        int ndev_have = 0;
        if (cudaGetDeviceCount(&ndev_have) != cudaSuccess || ndev_have == 0)
            throw Error(MBFBAM_ERR_CUDA, "no CUDA device available; mbf_bam_b200 has no CPU path");
        const std::vector<int> devs = job_devices(job->devices, job->n_devices, job->device);
        mbfbam_count_job j2 = *job;
        j2.each_read_counts_once = 0;
        const std::shared_ptr<const CountTables> Tp = cached_count_tables(r, &j2);
        const CountTables& T = *Tp;
        const int nsh = std::max(1, job->shard_count), sh = job->shard_index;
        if (sh < 0 || sh >= nsh) throw Error(MBFBAM_ERR_ARG, "bad shard index/count");
        const uint64_t prior = rd->fmap ? rd->fmap->jobs.fetch_add(1) : 0;
        std::vector<QuantPart> parts(devs.size());
        const int rt = default_read_threads(devs.size());
        fan_out(devs.size(), [&](size_t k) {
            quantify_on_device(rd, T, devs[k], sh * (int)devs.size() + (int)k, nsh * (int)devs.size(), rt, prior, parts[k]);
        });
        size_t n = 0;
        mbfbam_stats st{};
        for (const QuantPart& pt : parts) { n += pt.entries.size(); merge_stats(st, pt.st); }
        out->gene = (uint32_t*)malloc((n + 1) * 4);
        out->pos = (uint32_t*)malloc((n + 1) * 4);
        out->count = (uint32_t*)malloc((n + 1) * 4);
        out->bucket = (uint8_t*)malloc(n + 1);
        out->scanned = (uint8_t*)calloc((size_t)job->intervals->n_contigs + 1, 1);
        for (int c : T.slot_iv) out->scanned[c] = 1;
        size_t k = 0;
        for (const QuantPart& pt : parts)
            for (const QuantEntry& e : pt.entries) {  // chunk-range shards never share a (gene, pos): plain concatenation
                out->gene[k] = e.gene; out->pos[k] = e.pos; out->count[k] = e.count; out->bucket[k] = e.bucket;
                k++;
            }
        out->n = k;
        st.ms_total = now_ms() - t0;
        if (stats) *stats = st;
    });
}

void mbfbam_free_gene_pos_counts(mbfbam_gene_pos_counts* x) {
    if (!x) return;
    free(x->gene); free(x->pos); free(x->count); free(x->bucket); free(x->scanned);
    memset(x, 0, sizeof *x);
}

int mbfbam_duplicate_distribution(mbfbam_reader* rd, const int32_t* devices, int32_t n_devices, int32_t shard_index,
                                  int32_t shard_count, mbfbam_dup_hist* out, mbfbam_stats* stats, char* err, size_t errlen) {
    if (stats) memset(stats, 0, sizeof *stats);
    if (out) memset(out, 0, sizeof *out);
    return guarded(err, errlen, [&] {
        const double t0 = now_ms();
        const Reader& r = rd->r;
        if (!out) throw Error(MBFBAM_ERR_ARG, "null argument");
        int ndev_have = 0;
        if (cudaGetDeviceCount(&ndev_have) != cudaSuccess || ndev_have == 0)
            throw Error(MBFBAM_ERR_CUDA, "no CUDA device available; mbf_bam_b200 has no CPU path");
        const std::vector<int> devs = job_devices(devices, n_devices, 0);
        std::vector<Chunk> chunks = make_chunks(r, nullptr);
        std::vector<uint32_t> ref_chunk_off(r.names.size() + 1, 0);
        for (const Chunk& c : chunks) ref_chunk_off[(size_t)c.tid + 1]++;
        for (size_t i = 0; i < r.names.size(); i++) ref_chunk_off[i + 1] += ref_chunk_off[i];
        const int nsh = std::max(1, shard_count), sh = shard_index;
        if (sh < 0 || sh >= nsh) throw Error(MBFBAM_ERR_ARG, "bad shard index/count");
        const uint64_t prior = rd->fmap ? rd->fmap->jobs.fetch_add(1) : 0;
        std::vector<DupPart> parts(devs.size());
        const int rt = default_read_threads(devs.size());
        fan_out(devs.size(), [&](size_t k) {
            dups_on_device(rd, chunks, ref_chunk_off, devs[k], sh * (int)devs.size() + (int)k, nsh * (int)devs.size(), rt, prior, parts[k]);
        });
        std::map<uint32_t, uint64_t> hist;  // duplicate_distribution.rs:8-14 add_hashmaps
        mbfbam_stats st{};
        for (const DupPart& pt : parts) {
            for (uint32_t k = 0; k < DUP_DENSE; k++) if (pt.dense[k]) hist[k] += pt.dense[k];
            for (unsigned long long k : pt.big) hist[(uint32_t)k] += 1;
            merge_stats(st, pt.st);
        }
        out->n = hist.size();
        out->k = (uint32_t*)malloc((hist.size() + 1) * 4);
        out->count = (uint64_t*)malloc((hist.size() + 1) * 8);
        size_t i = 0;
        for (const auto& kv : hist) { out->k[i] = kv.first; out->count[i] = kv.second; i++; }
        st.ms_total = now_ms() - t0;
        if (stats) *stats = st;
    });
}

int mbfbam_count_by_barcode(mbfbam_reader* rd, const mbfbam_count_job* job, mbfbam_barcode_counts* out, mbfbam_stats* stats,
                            char* err, size_t errlen) {
    if (stats) memset(stats, 0, sizeof *stats);
    if (out) memset(out, 0, sizeof *out);
    return guarded(err, errlen, [&] {
        const double t0 = now_ms();
        const Reader& r = rd->r;
        if (!job || !job->intervals || !job->gene_intervals || !out) throw Error(MBFBAM_ERR_ARG, "null argument");
        int ndev_have = 0;
        if (cudaGetDeviceCount(&ndev_have) != cudaSuccess || ndev_have == 0)
            throw Error(MBFBAM_ERR_CUDA, "no CUDA device available; mbf_bam_b200 has no CPU path");
        const std::vector<int> devs = job_devices(job->devices, job->n_devices, job->device);
        mbfbam_count_job j2 = *job;
        j2.each_read_counts_once = 0;
        const std::shared_ptr<const CountTables> Tp = cached_count_tables(r, &j2);
        const CountTables& T = *Tp;
        const int nsh = std::max(1, job->shard_count), sh = job->shard_index;
        if (sh < 0 || sh >= nsh) throw Error(MBFBAM_ERR_ARG, "bad shard index/count");
        const uint64_t prior = rd->fmap ? rd->fmap->jobs.fetch_add(1) : 0;
        std::vector<BarcodePart> parts(devs.size());
        const int rt = default_read_threads(devs.size());
        fan_out(devs.size(), [&](size_t k) {
            barcode_on_device(rd, T, devs[k], sh * (int)devs.size() + (int)k, nsh * (int)devs.size(), rt, prior, parts[k]);
        });
        std::vector<const BarcodeEntry*> all;
        mbfbam_stats st{};
        size_t bytes = 0;
        for (const BarcodePart& pt : parts) {
            for (const BarcodeEntry& e : pt.entries) { all.push_back(&e); bytes += e.barcode.size(); }
            merge_stats(st, pt.st);
        }
        // chunks in order, then (gene, barcode): the reference's order is HashMap iteration, i.e. arbitrary
        std::sort(all.begin(), all.end(), [](const BarcodeEntry* a, const BarcodeEntry* b) {
            if (a->chunk != b->chunk) return a->chunk < b->chunk;
            if (a->gene != b->gene) return a->gene < b->gene;
            return a->barcode < b->barcode;
        });
        const size_t n = all.size();
        out->gene = (uint32_t*)malloc((n + 1) * 4);
        out->chunk = (uint32_t*)malloc((n + 1) * 4);
        out->count = (uint32_t*)malloc((n + 1) * 4);
        out->bc_off = (uint64_t*)malloc((n + 1) * 8);
        out->bc_bytes = (char*)malloc(bytes + 1);
        size_t off = 0;
        for (size_t i = 0; i < n; i++) {
            out->gene[i] = all[i]->gene; out->chunk[i] = all[i]->chunk; out->count[i] = all[i]->count;
            out->bc_off[i] = off;
            memcpy(out->bc_bytes + off, all[i]->barcode.data(), all[i]->barcode.size());
            off += all[i]->barcode.size();
        }
        out->bc_off[n] = off;
        out->n = n;
        st.ms_total = now_ms() - t0;
        if (stats) *stats = st;
    });
}

void mbfbam_free_barcode_counts(mbfbam_barcode_counts* x) {
    if (!x) return;
    free(x->gene); free(x->chunk); free(x->count); free(x->bc_off); free(x->bc_bytes);
    memset(x, 0, sizeof *x);
}

void mbfbam_free_dup_hist(mbfbam_dup_hist* x) {
    if (!x) return;
    free(x->k); free(x->count);
    memset(x, 0, sizeof *x);
}

void mbfbam_free_introns(mbfbam_introns* x) {
    if (!x) return;
    free(x->tid);
    free(x->start);
    free(x->stop);
    free(x->count);
    free(x->contig_seen);
    memset(x, 0, sizeof *x);
}

int mbfbam_inflate_buffer(const uint8_t* bgzf, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_len, int32_t device,
                          mbfbam_stats* stats, char* err, size_t errlen) {
    if (stats) memset(stats, 0, sizeof *stats);
    return guarded(err, errlen, [&] {
        double t0 = now_ms();
        MemSource src(bgzf, n);
        int ndev_have = 0;
        if (cudaGetDeviceCount(&ndev_have) != cudaSuccess || ndev_have == 0)
            throw Error(MBFBAM_ERR_CUDA, "no CUDA device available; mbf_bam_b200 has no CPU path");
        DevSlot& slot = dev_slot(device);
        std::lock_guard<std::mutex> lk(slot.mu);
        Pipeline& p = slot_pipeline(slot, device);
        p.read_threads = 0;
        std::vector<ByteRange> ranges;
        if (n) ranges.push_back({0, n, 0});
        run_with_retry(p, ranges, [&](Pipeline& p) {
            p.begin_job(&src, JOB_INFLATE, n);
            p.out_host = out;
            p.out_cap = out_cap;
        });
        if (out_len) *out_len = p.out_len;
        p.st.ms_total = now_ms() - t0;
        if (stats) *stats = p.st;
    });
}

int mbfbam_cigar_expand(const uint32_t* cigar, int32_t n_ops, int32_t pos, int32_t want_introns, uint32_t* out_pairs,
                        int32_t cap_pairs, int32_t* n_pairs, int32_t device, char* err, size_t errlen) {
    return guarded(err, errlen, [&] {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) throw Error(MBFBAM_ERR_CUDA, "no CUDA device available; mbf_bam_b200 has no CPU path");
        CK(cudaSetDevice(device));
        DevBuf d_c, d_o, d_n;
        d_c.ensure(((size_t)n_ops + 2) * 4);
        d_o.ensure(((size_t)cap_pairs + 1) * 8);
        d_n.ensure(4);
        if (n_ops) CK(cudaMemcpy(d_c.p, cigar, (size_t)n_ops * 4, cudaMemcpyHostToDevice));
        k_cigar_expand<<<1, 1>>>(d_c.as<uint32_t>(), n_ops, pos, want_introns, d_o.as<uint32_t>(), cap_pairs, d_n.as<int>());
        CK(cudaGetLastError());
        CK(cudaMemcpy(n_pairs, d_n.p, 4, cudaMemcpyDeviceToHost));
        int m = std::min(*n_pairs, cap_pairs);
        if (m > 0) CK(cudaMemcpy(out_pairs, d_o.p, (size_t)m * 8, cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
