"""dev experiment: fused-kernel time vs spatial spread of the robots (L1 hit rate of the table gathers)."""
import sys, statistics; sys.path.insert(0, '/root/repo')
import torch
import bench
from shifu_b200.sim.synthetic import a1_snapshot
n = 1 << 20
terrain = bench._terrain()
for spread in (3.0, 0.5, 0.0):
    hp, raw = bench.build_a1(n, 0, 1, "cuda:0", terrain)
    if spread < 3.0:
        # collapse all robots onto one spot (+- spread metres), keep the yaw distribution
        hp.root_state[:, 0] = 100.0 + (torch.rand(n, device="cuda") * 2 - 1) * spread
        hp.root_state[:, 1] = 100.0 + (torch.rand(n, device="cuda") * 2 - 1) * spread
        hp.env_origins[:, 0] = 100.0; hp.env_origins[:, 1] = 100.0
    tot, km = bench.time_resident(hp, raw, 20, 5)
    print(f"spread {spread}: fused kernel {statistics.mean(km):.4f} ms, step {tot/20:.4f} ms")
    del hp
