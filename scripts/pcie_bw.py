"""Pinned host<->device copy bandwidth on the box (context for the e2e number)."""
import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device='cuda')
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h.copy_(d, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print('H2D %.1f GB/s  D2H %.1f GB/s  both at once: %.1f + %.1f GB/s' % (n / a / 1e9, n / b / 1e9, n / c / 1e9, n / c / 1e9))
# many small D2H copies (2.4 MB each), like one image's descriptors
m = 2400 * 1024
def small():
    for i in range(64): h[i * m:(i + 1) * m].copy_(d[i * m:(i + 1) * m], non_blocking=True)
e = t(small)
print('64 x 2.4 MB D2H: %.1f GB/s (%.2f ms)' % (64 * m / e / 1e9, e * 1e3))
