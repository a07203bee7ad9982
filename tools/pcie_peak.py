import torch, time
n = 2 << 30
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for name, src, dst in (("D2H", d, h), ("H2D", h, d)):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(name, "%.1f GB/s" % (n / dt / 1e9))
# chunked D2H, 512 MB pieces
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        for o in range(0, n, 512 << 20):
            h[o:o + (512 << 20)].copy_(d[o:o + (512 << 20)], non_blocking=True)
    s.synchronize()
    print("D2H 512MB chunks %.1f GB/s" % (3 * n / (time.perf_counter() - t) / 1e9))
