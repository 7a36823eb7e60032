"""Does PCIe run both directions at once on this box?  Raw pinned copies, then two contexts."""
import sys, time, threading
sys.path.insert(0, ".")
import torch

N = 512 << 20
h_a = torch.empty(N, dtype=torch.uint8, pin_memory=True)
h_b = torch.empty(N, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(N, dtype=torch.uint8, device="cuda")
d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D {N / a / 1e9:.1f} GB/s  D2H {N / b / 1e9:.1f} GB/s  both {2 * N / c / 1e9:.1f} GB/s aggregate ({c * 1e3:.1f} ms vs {a * 1e3:.1f}+{b * 1e3:.1f})")
