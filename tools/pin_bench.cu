// Host-ingest microbenchmark (run on the GPU box): how fast can file bytes that sit in the page cache
// reach HBM?  Measures, for one scratch file:
//   * pread -> pinned staging with T threads (the cold path of Pipeline::stage_window)
//   * memcpy from a populated mmap -> pinned staging with T threads
//   * cudaHostRegister of the mapping: whole, in 64 MiB segments, and segments registered by T threads
//   * H2D from the registered mapping (per segment copies) and from a plain pinned buffer
// Usage: pin_bench <dir> [MiB=2048] [threads=8]
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : "/tmp";
    size_t mib = argc > 2 ? (size_t)atol(argv[2]) : 2048;
    int T = argc > 3 ? atoi(argv[3]) : 8;
    const size_t n = mib << 20, SEG = 64u << 20;
    std::string path = dir + "/pin_bench.dat";
    printf("== %s, %zu MiB, %d threads, %u hw threads\n", path.c_str(), mib, T, std::thread::hardware_concurrency());
    {
        int fd = open(path.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
        if (fd < 0) { perror("open"); return 1; }
        std::vector<uint8_t> buf(16u << 20);
        uint32_t x = 12345;
        for (auto& b : buf) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
        double t0 = now();
        for (size_t off = 0; off < n; off += buf.size()) if (write(fd, buf.data(), buf.size()) != (ssize_t)buf.size()) { perror("write"); return 1; }
        close(fd);
        printf("write file: %.2f s\n", now() - t0);
    }
    int fd = open(path.c_str(), O_RDONLY);
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    uint8_t *d = nullptr, *h = nullptr;
    CK(cudaMalloc(&d, n));
    double t0 = now();
    CK(cudaHostAlloc(&h, n, cudaHostAllocPortable));
    printf("cudaHostAlloc %zu MiB: %.3f s (%.2f GB/s)\n", mib, now() - t0, n / 1e9 / (now() - t0));
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    auto par = [&](int threads, auto fn) {
        std::atomic<size_t> next{0};
        std::vector<std::thread> th;
        double t = now();
        for (int i = 0; i < threads; i++) th.emplace_back([&] { for (;;) { size_t k = next.fetch_add(1); if (k * SEG >= n) break; fn(k * SEG, std::min(SEG, n - k * SEG)); } });
        for (auto& x : th) x.join();
        return now() - t;
    };
    for (int threads : {1, 2, 4, T, 2 * T}) {
        double dt = par(threads, [&](size_t off, size_t len) { size_t got = 0; while (got < len) { ssize_t r = pread(fd, h + off + got, len - got, off + got); if (r <= 0) break; got += r; } });
        printf("pread -> pinned, %2d threads: %.3f s (%.2f GB/s)\n", threads, dt, n / 1e9 / dt);
    }
    t0 = now();
    CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    printf("H2D from cudaHostAlloc buffer: %.3f s (%.2f GB/s)\n", now() - t0, n / 1e9 / (now() - t0));

    // ---- mapping
    t0 = now();
    uint8_t* m = (uint8_t*)mmap(nullptr, n, PROT_READ, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { perror("mmap"); return 1; }
    printf("mmap (no populate): %.4f s\n", now() - t0);
    for (int threads : {1, T}) {
        double dt = par(threads, [&](size_t off, size_t len) { memcpy(h + off, m + off, len); });
        printf("memcpy mmap -> pinned (first touch: faults), %2d threads: %.3f s (%.2f GB/s)\n", threads, dt, n / 1e9 / dt);
        dt = par(threads, [&](size_t off, size_t len) { memcpy(h + off, m + off, len); });
        printf("memcpy mmap -> pinned (mapped), %2d threads: %.3f s (%.2f GB/s)\n", threads, dt, n / 1e9 / dt);
    }
    munmap(m, n);
    // register: whole
    auto reg_test = [&](const char* what, int prot, int flags, unsigned rflags) {
        uint8_t* p = (uint8_t*)mmap(nullptr, n, prot, flags, fd, 0);
        if (p == MAP_FAILED) { printf("%s: mmap failed\n", what); return; }
        double t = now();
        cudaError_t e = cudaHostRegister(p, n, rflags);
        double dt = now() - t;
        if (e != cudaSuccess) { printf("%s: whole register refused: %s\n", what, cudaGetErrorString(e)); cudaGetLastError(); munmap(p, n); return; }
        printf("%s: register whole: %.3f s (%.2f GB/s)\n", what, dt, n / 1e9 / dt);
        t = now();
        CK(cudaMemcpyAsync(d, p, n, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        printf("%s: H2D one copy: %.3f s (%.2f GB/s)\n", what, now() - t, n / 1e9 / (now() - t));
        t = now();
        CK(cudaHostUnregister(p));
        printf("%s: unregister: %.3f s\n", what, now() - t);
        // segments, serial
        t = now();
        for (size_t off = 0; off < n; off += SEG) CK(cudaHostRegister(p + off, std::min(SEG, n - off), rflags));
        dt = now() - t;
        printf("%s: register 64 MiB segments serially: %.3f s (%.2f GB/s)\n", what, dt, n / 1e9 / dt);
        t = now();
        for (size_t off = 0; off < n; off += SEG) CK(cudaMemcpyAsync(d + off, p + off, std::min(SEG, n - off), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        printf("%s: H2D per segment: %.3f s (%.2f GB/s)\n", what, now() - t, n / 1e9 / (now() - t));
        t = now();
        cudaError_t e2 = cudaMemcpyAsync(d, p, 2 * SEG, cudaMemcpyHostToDevice, s);
        cudaError_t e3 = cudaStreamSynchronize(s);
        printf("%s: H2D across two segments in one copy: %s / %s, %.3f s (%.2f GB/s)\n", what, cudaGetErrorString(e2), cudaGetErrorString(e3), now() - t, 2 * SEG / 1e9 / (now() - t));
        cudaGetLastError();
        for (size_t off = 0; off < n; off += SEG) CK(cudaHostUnregister(p + off));
        // segments, T threads
        for (int threads : {2, 4, T}) {
            dt = par(threads, [&](size_t off, size_t len) { cudaSetDevice(0); cudaError_t e4 = cudaHostRegister(p + off, len, rflags); if (e4 != cudaSuccess) printf("  register failed: %s\n", cudaGetErrorString(e4)); });
            printf("%s: register 64 MiB segments with %d threads: %.3f s (%.2f GB/s)\n", what, threads, dt, n / 1e9 / dt);
            for (size_t off = 0; off < n; off += SEG) cudaHostUnregister(p + off);
        }
        munmap(p, n);
    };
    reg_test("shared/ro", PROT_READ, MAP_SHARED, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
    reg_test("shared/ro+populate", PROT_READ, MAP_SHARED | MAP_POPULATE, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
    reg_test("private/rw", PROT_READ | PROT_WRITE, MAP_PRIVATE, cudaHostRegisterPortable);
    close(fd);
    unlink(path.c_str());
    return 0;
}
