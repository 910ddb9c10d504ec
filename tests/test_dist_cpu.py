"""CPU test of the N>1 path (world_size 2, gloo): env sharding with global ids + the statistics
all-reduce reproduce the single-process result over the concatenated env set.
The per-env arithmetic on each shard is done by the oracle here (no GPU in this test); what is
under test is the host logic that makes a sharded run invariant to the GPU count."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util

N_GLOBAL, STEPS, SEED = 64, 3, 31


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _world(n_global):
    """Terrain + per-global-env initial state shared by every layout of the run."""
    z, meta = util.load_golden("a1_small")
    rs = np.random.RandomState(4)
    levels0 = rs.randint(0, meta["max_terrain_level"], size=n_global)
    ep = rs.randint(0, 500, size=n_global)
    return z, meta, levels0, ep


def _run_shard(offset, n_local, n_global, reduce_fn=None):
    from oracle import shifu_oracle as so
    from shifu_b200 import dist as sdist
    from shifu_b200.sim.synthetic import a1_snapshot
    z, meta, levels0, ep = _world(n_global)
    types = sdist.global_terrain_types(offset, n_local, n_global, meta["num_cols"])
    origins = torch.from_numpy(z["terrain_origins"]).float()
    sl = slice(offset, offset + n_local)
    lv = torch.from_numpy(levels0[sl])
    p, st = util.make_oracle_a1(n_local, z["height_samples"], origins, types, origins[lv, types],
                                border_size=int(meta["border_size"]), max_terrain_level=meta["max_terrain_level"],
                                num_cols=meta["num_cols"], env_offset=offset)
    st.ep_len[:] = torch.from_numpy(ep[sl])
    st.terrain_levels[:] = lv
    logged, obs = [], []
    extras = {}
    for t in range(1, STEPS + 1):
        full = a1_snapshot(SEED, t, n_global, p_base=0.15, offmap=False)        # global snapshot, sliced per shard
        snap = type(full)(dof=full.dof[:, sl], root_offset=full.root_offset[sl], contact=full.contact[sl],
                          actions=full.actions[sl])
        st.extras.pop("_stats", None)
        so.a1_step(p, st, snap.actions, snap)
        s = st.extras.get("_stats")
        if s is None:
            s = {"sums": [0.0] * 6, "n_reset": 0, "level_sum": int(st.terrain_levels.sum()), "n_envs": n_local}
        vec = sdist.pack_stats(s["sums"], s["n_reset"], s["level_sum"], 0.0, s["n_envs"])
        if reduce_fn is not None:
            reduce_fn(vec)
        extras = sdist.extras_from_stats(vec, so.A1_REWARD_TERMS, p.max_episode_length_s, extras)
        logged.append(dict(extras))
        obs.append(st.obs.clone())
    return logged, obs, st


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shifu_b200 import dist as sdist
    offset, n_local = sdist.shard_range(N_GLOBAL, world, rank)
    logged, obs, st = _run_shard(offset, n_local, N_GLOBAL, sdist.make_stats_allreduce())
    q.put((rank, logged, [o.numpy() for o in obs], st.terrain_levels.numpy(), st.reset_ids.numpy() + offset))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_logged, ref_obs, ref_st = _run_shard(0, N_GLOBAL, N_GLOBAL)
    # per-env results: bit-identical to the single-process run (global Philox ids / terrain types)
    for t in range(STEPS):
        got = np.concatenate([results[0][2][t], results[1][2][t]])
        assert np.array_equal(got, ref_obs[t].numpy()), f"obs differ at step {t + 1}"
    assert np.array_equal(np.concatenate([results[0][3], results[1][3]]), ref_st.terrain_levels.numpy())
    assert np.array_equal(np.concatenate([results[0][4], results[1][4]]), ref_st.reset_ids.numpy())
    # logged means: every rank holds the GLOBAL value after the all-reduce (fp reduction order only)
    for t in range(STEPS):
        for r in (0, 1):
            for k, v in ref_logged[t].items():
                assert results[r][1][t][k] == pytest.approx(v, rel=1e-5, abs=1e-6), (t, r, k)
    # and the oracle's own extras agree with the reconstruction from the statistics vector
    for k in ("tracking_lin_vel", "leg_collision"):
        assert ref_logged[-1][k] == pytest.approx(float(ref_st.extras["episode"][k]), rel=1e-5, abs=1e-6)
    assert ref_logged[-1]["terrain_levels"] == pytest.approx(float(ref_st.extras["episode"]["terrain_levels"]),
                                                             rel=1e-6)


def test_shard_range_and_types():
    from shifu_b200 import dist as sdist
    assert sdist.shard_range(8 << 20, 8, 3) == (3 << 20, 1 << 20)
    with pytest.raises(ValueError):
        sdist.shard_range(10, 4, 0)
    full = sdist.global_terrain_types(0, 4096, 4096, 20)
    parts = torch.cat([sdist.global_terrain_types(o, 1024, 4096, 20) for o in range(0, 4096, 1024)])
    assert torch.equal(full, parts) and int(full.max()) == 19
