import torch, time
n = 1536 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, f in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(4): f()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (4 * n / (time.perf_counter() - t) / 1e9))
