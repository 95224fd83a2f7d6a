import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for n in (8 << 20, 64 << 20, 1 << 30):
    torch.cuda.synchronize()
    t = time.perf_counter()
    reps = max(1, (2 << 30) // n)
    for i in range(reps):
        off = (i * n) % (1 << 30)
        y[off:off + n].copy_(x[off:off + n], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("H2D piece %4d MiB: %.1f GB/s" % (n >> 20, reps * n / dt / 1e9))
torch.cuda.synchronize(); t = time.perf_counter(); x.copy_(y, non_blocking=True); torch.cuda.synchronize()
print("D2H 1 GiB: %.1f GB/s" % ((1 << 30) / (time.perf_counter() - t) / 1e9))
