"""dev probe: stand-alone height-scan kernel and PD kernel time at 1 Mi envs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
n = 1 << 20
hp, raw = bench.build_a1(n, 0, 1, "cuda:0", want_heights=True)
out = torch.empty(n, 187, device="cuda")
def t(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / k
print("get_heights alone  ms", t(lambda: hp.get_heights(out=out)))
print("post_physics       ms", t(lambda: hp.post_physics()))
print("pd_torque          ms", t(lambda: hp.pd_torque()))
