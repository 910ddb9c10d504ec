import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from tests import util
from tests.test_a1_gpu import _terrain
from shifu_b200.sim.synthetic import a1_snapshot
for n in (32*296*2, 32*296*3, 1<<18):
    hs, origins, types, env_origins = _terrain(n)
    hp = util.make_cuda_a1(n, hs, origins, types, env_origins)
    snap = a1_snapshot(3, 1, n, gen_device="cuda", p_base=0.01)
    hp.ep_len.copy_(torch.randint(0, 500, (n,), device="cuda"))
    util.cuda_a1_step(hp, snap, snap.actions)
    keep = ~hp.reset_buf
    a = hp.measured_heights.clone(); b = torch.empty_like(a); hp.get_heights(out=b)
    bad = ((a != b).any(dim=1) & keep).nonzero().flatten()
    print(n, "bad envs", bad.numel(), bad[:20].tolist())
    if bad.numel():
        e = int(bad[0]); d = (a[e]!=b[e]).nonzero().flatten()
        print(" env", e, "tile", e//32, "slot", e%32, "npts", d.numel(), d[:10].tolist(), a[e,d[:4]].tolist(), b[e,d[:4]].tolist())
        tiles = (bad//32).unique(); print(" tiles", tiles.numel(), tiles[:20].tolist(), "slots", (bad%32).unique().tolist())
