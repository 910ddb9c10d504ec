"""dev experiment: fused-kernel time with and without resetting envs in the tiles (B-chain variance)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
n = 1 << 20
env, raw = bench.build_env(n, 0, 1, "cuda:0", use_graph=False)
hp = env.hot
for _ in range(4):
    env.step(raw)
def t(label):
    ms = bench.time_fused_kernel(hp, 20)
    print(label, sum(ms) / len(ms), "reset frac", float(hp.reset_buf.float().mean()))
t("as benchmarked")
hp.contact_state.zero_()
hp.ep_len.zero_()
t("no resets (contact forces zero, ep_len 0)")
hp.ep_len.fill_(10 ** 6)
t("every env resets")
