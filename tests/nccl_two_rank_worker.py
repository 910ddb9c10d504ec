"""Worker of tests/test_a1_gpu.py::test_two_rank_nccl_matches_one_rank (launched with torchrun, 2 ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(world_state, sl, hs, origins, allreduce, steps=2):
    from tests import util
    hp = util.world_shard_cuda(world_state, sl, hs, origins, carry=True, want_heights=False,
                               device=f"cuda:{torch.cuda.current_device()}")
    snaps = world_state[5]
    for t in range(1, steps + 1):
        # simulator refresh of the shard, then the step with the statistics all-reduce
        n = hp.n
        snap = util.slice_snap(snaps[t], sl)
        hp.pd_torque(snap.actions.to(hp.device).contiguous())
        for i in range(1, 4):
            hp.pd_torque()
        root = snap.root_offset.to(hp.device).clone()
        root[:, 0:3] += hp.env_origins
        hp.root_state.copy_(root)
        hp.dof_state.view(n, 12, 2).copy_(snap.dof[4].to(hp.device))
        hp.contact_state.view(n, 17, 3).copy_(snap.contact.to(hp.device))
        hp.post_physics()
        hp.finalize(allreduce)
    torch.cuda.synchronize()
    ids = hp.reset_id_list().cpu().numpy()
    return dict(obs=hp.obs_buf.cpu().numpy(), rew=hp.rew_buf.cpu().numpy(), reset=hp.reset_buf.cpu().numpy(),
                terrain_levels=hp.terrain_levels.cpu().numpy(), ep_len=hp.ep_len.cpu().numpy(), reset_ids=ids,
                extras=hp.extras_arr.cpu().numpy())


def main():
    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from tests import util
    from tests.test_a1_gpu import _terrain
    n = 16384
    hs, origins, _, _ = _terrain(64)
    ws = util.world_state(n * world, 23, 2, origins)
    allreduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)
    np.savez(os.path.join(out, f"rank{rank}.npz"), **run(ws, slice(rank * n, (rank + 1) * n), hs, origins, allreduce))
    if rank == 0:
        np.savez(os.path.join(out, "world1.npz"), **run(ws, slice(0, n * world), hs, origins, None))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
