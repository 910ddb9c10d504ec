"""dev helper: minimal driver for an ncu capture of the fused kernel (no bench extras)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
n = int(os.environ.get("N", 1 << 20))
env, raw = bench.build_env(n, 0, 1, "cuda:0", use_graph=False)
for _ in range(6):
    env.step(raw)
torch.cuda.synchronize()
