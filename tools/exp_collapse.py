"""dev experiment: fused-kernel time when every robot stands on the same map cell (all gathers hit L1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
n = 1 << 20
for mode in ("normal", "collapse"):
    hp, raw = bench.build_a1(n, 0, 1, "cuda:0")
    if mode == "collapse":
        hp.root_state[:, 0] = hp.root_state[0, 0]; hp.root_state[:, 1] = hp.root_state[0, 1]
        hp.root_state[:, 3:7] = hp.root_state[0, 3:7]
    for _ in range(3):
        hp.post_physics()
    ts = []
    for _ in range(20):
        if mode == "collapse":
            hp.root_state[:, 0] = hp.root_state[0, 0]; hp.root_state[:, 1] = hp.root_state[0, 1]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); hp.post_physics(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(os.environ.get("SHIFU_B200_LIB"), mode, "median kernel ms", ts[len(ts) // 2], flush=True)
    del hp
