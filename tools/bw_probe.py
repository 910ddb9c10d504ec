"""dev probe: achievable HBM bandwidth for pure-write, copy and read-heavy streams (torch kernels)."""
import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
N = 1 << 29   # 2 GiB of fp32
x = torch.empty(N, device="cuda"); y = torch.empty(N, device="cuda")
ms = t(lambda: x.zero_()); print("memset  (write only)  %.0f GB/s" % (4 * N / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy    (1R:1W)       %.0f GB/s" % (8 * N / ms / 1e6))
ms = t(lambda: torch.add(x, 1.0, out=x)); print("inplace (1R:1W same)  %.0f GB/s" % (8 * N / ms / 1e6))
z = torch.empty(N // 2, device="cuda")
ms = t(lambda: torch.add(x[: N // 2], x[N // 2:], out=z)); print("add     (2R:1W)       %.0f GB/s" % (6 * N / ms / 1e6))
ms = t(lambda: x.sum()); print("sum     (read only)   %.0f GB/s" % (4 * N / ms / 1e6))
# write-heavy: 1R : 2W  (read half, write two halves)
w1 = torch.empty(N // 2, device="cuda"); w2 = torch.empty(N // 2, device="cuda")
def f():
    torch.add(z, 1.0, out=w1); w2.copy_(w1)   # not a single kernel; indicative only
