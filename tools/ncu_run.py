"""dev helper: minimal driver for an ncu capture of the fused kernel (no bench extras)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
n = int(os.environ.get("N", 1 << 20))
hp, raw = bench.build_a1(n, 0, 1, "cuda:0")
for _ in range(6):
    hp.step_resident(raw)
torch.cuda.synchronize()
