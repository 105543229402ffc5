"""tools/pcie_bw.py -- pinned H2D / D2H copy bandwidth of the box (alone and both directions at once)"""
import torch, time
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d1.copy_(h1, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print(f"H2D {n/a/1e9:.1f} GB/s  D2H {n/b/1e9:.1f} GB/s  simultaneous {n/c/1e9:.1f} GB/s each way")
m = 12441600
def frame():
    with torch.cuda.stream(s1): d1[:m].copy_(h1[:m], non_blocking=True)
    with torch.cuda.stream(s2): h2[:m].copy_(d2[:m], non_blocking=True)
c = t(frame, 50)
print(f"one 4K 8-bit frame each way at once: {c*1e6:.0f} us -> {1/c:.0f} frames/s ceiling")
