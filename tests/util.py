"""Shared helpers of the parity tests: fixture loading and the two step drivers
(oracle on CPU, CUDA hot path through the C-ABI) that replay the same simulator snapshots in the
reference's refresh order."""
from __future__ import annotations

import json
import os
import types
from typing import Dict, List

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EXACT_KEYS = ("reset", "time_out", "contact_term", "ep_len", "terrain_levels", "measured_heights", "success")
RTOL, ATOL = 1e-5, 1e-6      # BASELINE.json: obs/rew within 1e-5 relative (fp32); abs floor 1e-6


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def golden_snap(z, t: int, kind: str = "a1"):
    pre = f"s{t}/in/"
    keys = ("dof", "root_offset", "contact", "actions") if kind == "a1" else ("root", "body", "dof", "actions")
    return types.SimpleNamespace(**{k: torch.from_numpy(z[pre + k].copy()) for k in keys})


def golden_out(z, t: int) -> Dict[str, np.ndarray]:
    pre = f"s{t}/out/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def assert_close(name: str, got, want, exact: bool):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    if exact or got.dtype.kind in "iub":
        bad = np.flatnonzero(got.reshape(-1) != want.reshape(-1))
        assert bad.size == 0, f"{name}: {bad.size} of {got.size} entries differ (first at {bad[:5]})"
    else:
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        tol = ATOL + RTOL * np.abs(want.astype(np.float64))
        bad = np.flatnonzero((err > tol).reshape(-1))
        assert bad.size == 0, (f"{name}: {bad.size} of {got.size} entries outside rtol={RTOL} atol={ATOL}; "
                               f"max err {err.max():.3e} at {int(err.argmax())}")


# ---------------------------------------------------------------------------------------------
# A1: oracle driver
# ---------------------------------------------------------------------------------------------

def make_oracle_a1(n, height_samples, terrain_origins, terrain_types, env_origins, *, border_size=25,
                   max_terrain_level=10, num_cols=20, rng_seed=0x5EED, env_offset=0, horizontal_scale=0.1,
                   points_x=None, points_y=None, curriculum=True):
    from oracle import shifu_oracle as so
    grid = {k: list(v) for k, v in (("points_x", points_x), ("points_y", points_y)) if v is not None}
    p = so.A1Params(n=n, border_size=border_size, max_terrain_level=max_terrain_level, num_cols=num_cols,
                    rng_seed=rng_seed, env_offset=env_offset, horizontal_scale=horizontal_scale, curriculum=curriculum,
                    **grid)
    st = so.a1_new_state(p, torch.as_tensor(height_samples), torch.as_tensor(terrain_origins).float(),
                         torch.as_tensor(terrain_types), torch.as_tensor(env_origins).float())
    return p, st


def oracle_a1_outputs(st) -> Dict[str, np.ndarray]:
    from oracle import shifu_oracle as so
    d = dict(obs=st.obs, rew=st.rew, reset=st.reset.to(torch.uint8), time_out=st.time_out.to(torch.uint8),
             contact_term=st.contact_term.to(torch.uint8), ep_len=st.ep_len, terrain_levels=st.terrain_levels,
             env_origins=st.env_origins, command=st.command, history=st.history, actions=st.actions,
             root_state=st.root_state, dof_state=st.dof_state, dof_targets=st.dof_targets,
             rand_force=st.rand_force, torques=st.torques, base_lin_vel=st.base_lin_vel,
             base_ang_vel=st.base_ang_vel, projected_gravity=st.projected_gravity,
             measured_heights=st.measured_heights)
    for k in so.A1_REWARD_TERMS:
        d["ep_sum/" + k] = st.ep_sums[k]
    for k, v in st.extras.get("episode", {}).items():
        d["extras/" + k] = torch.as_tensor(v)
    if "time_outs" in st.extras:
        d["extras_time_outs"] = st.extras["time_outs"].to(torch.uint8)
    return {k: v.detach().numpy().copy() for k, v in d.items()}


# ---------------------------------------------------------------------------------------------
# A1: CUDA driver (through the C-ABI via shifu_b200.hotpath)
# ---------------------------------------------------------------------------------------------

def make_cuda_a1(n, height_samples, terrain_origins, terrain_types, env_origins, *, border_size=25.,
                 max_terrain_level=10, num_cols=20, rng_seed=0x5EED, env_offset=0, carry=False, device="cuda:0",
                 horizontal_scale=0.1, want_measured_heights=True, points_x=None, points_y=None, curriculum=True):
    from shifu_b200 import hotpath
    grid = {k: tuple(v) for k, v in (("points_x", points_x), ("points_y", points_y)) if v is not None}
    dev = torch.device(device)
    root = torch.zeros(n, 13, device=dev)
    root[:, 6] = 1.0
    dof = torch.zeros(n * 12, 2, device=dev)
    contact = torch.zeros(n * 17, 3, device=dev)
    desc = hotpath.a1_desc(n, border_size=float(border_size), max_terrain_level=max_terrain_level,
                           num_terrain_types=num_cols, rng_seed=rng_seed, env_offset=env_offset,
                           horizontal_scale=horizontal_scale, curriculum=curriculum, **grid)
    hp = hotpath.A1HotPath(desc, root_state=root, dof_state=dof, contact_state=contact,
                           height_samples=torch.as_tensor(height_samples),
                           terrain_origins=torch.as_tensor(terrain_origins),
                           terrain_types=torch.as_tensor(terrain_types),
                           env_origins=torch.as_tensor(env_origins).float().to(dev).contiguous(),
                           carry_body_frame=carry, want_measured_heights=want_measured_heights)
    return hp


def cuda_a1_step(hp, snap, raw_actions):
    """One control step with the simulator refreshes interleaved exactly like the reference
    (a1_conditional.py:64-75, isaac_gym.py:139-154)."""
    dev = hp.device
    n = hp.n
    dofv = hp.dof_state.view(n, 12, 2)
    raw = raw_actions.to(dev).contiguous()
    hp.pd_torque(raw)
    dofv.copy_(snap.dof[0].to(dev))
    for i in range(1, 4):
        hp.pd_torque()
        dofv.copy_(snap.dof[i].to(dev))
    if not hp.carry_body_frame:
        hp.body_frame()                              # S_prev root (D7)
    root = snap.root_offset.to(dev).clone()
    root[:, 0:3] += hp.env_origins
    hp.root_state.copy_(root)
    dofv.copy_(snap.dof[4].to(dev))
    hp.contact_state.view(n, 17, 3).copy_(snap.contact.to(dev))
    hp.post_physics()
    hp.finalize()


def cuda_a1_outputs(hp) -> Dict[str, np.ndarray]:
    d = dict(obs=hp.obs_buf, rew=hp.rew_buf, reset=hp.reset_buf.to(torch.uint8),
             time_out=hp.time_out_buf.to(torch.uint8), contact_term=hp.contact_terminate_buf.to(torch.uint8),
             ep_len=hp.ep_len, terrain_levels=hp.terrain_levels, env_origins=hp.env_origins,
             command=hp.command, history=hp.history, actions=hp.actions, root_state=hp.root_state,
             dof_state=hp.dof_state, dof_targets=hp.dof_targets, rand_force=hp.rand_force,
             torques=hp.torques, base_lin_vel=hp.base_lin_vel, base_ang_vel=hp.base_ang_vel,
             projected_gravity=hp.projected_gravity)
    if hp.measured_heights is not None:
        d["measured_heights"] = hp.measured_heights
    for k in hp.terms:
        d["ep_sum/" + k] = hp.ep_sums[k]
    ex = hp.extras()
    for k, v in ex["episode"].items():
        d["extras/" + k] = v
    d["extras_time_outs"] = ex["time_outs"].to(torch.uint8)
    out = {k: v.detach().cpu().numpy().copy() for k, v in d.items()}
    out["reset_ids"] = hp.reset_id_list().cpu().numpy().copy()
    return out


def compare_a1(got: Dict, want: Dict, tag: str, skip=()):
    for k, w in want.items():
        if k in skip or k not in got:
            continue
        assert_close(f"{tag}:{k}", got[k], w, exact=(k in EXACT_KEYS))


# ---------------------------------------------------------------------------------------------
# A world of n_global envs cut into contiguous shards: same global initial state and
# snapshots whatever the shard layout
# ---------------------------------------------------------------------------------------------

def slice_snap(snap, sl):
    return type(snap)(dof=snap.dof[:, sl], root_offset=snap.root_offset[sl], contact=snap.contact[sl],
                      actions=snap.actions[sl])


def world_state(n_global, seed, steps, height_origins):
    """Global initial state + per-step snapshots (CPU) of a seeded world."""
    from shifu_b200 import dist as sdist
    from shifu_b200.sim.synthetic import a1_snapshot
    origins = height_origins
    types = sdist.global_terrain_types(0, n_global, n_global, 20).clamp_(max=19)
    rs = np.random.RandomState(seed)
    levels0 = torch.from_numpy(rs.randint(0, 6, size=n_global))
    ep0 = torch.from_numpy(rs.randint(0, 500, size=n_global))
    cmd0 = torch.from_numpy(rs.uniform(-1, 1, size=(n_global, 3)).astype(np.float32))
    env_origins = origins[levels0, types]
    snaps = [a1_snapshot(seed, t, n_global, p_base=0.02, offmap=False) for t in range(steps + 1)]
    root0 = snaps[0].root_offset.clone()
    root0[:, :3] += env_origins
    return types, levels0, ep0, cmd0, env_origins, snaps, root0


def world_shard_cuda(world, sl, hs, origins, *, carry, want_heights, device="cuda:0"):
    """CUDA hot path over the envs `sl` of a world (env_offset = sl.start), seeded like the world."""
    types, levels0, ep0, cmd0, env_origins, snaps, root0 = world
    n = sl.stop - sl.start
    hp = make_cuda_a1(n, hs, origins, types[sl], env_origins[sl], carry=carry, want_measured_heights=want_heights,
                      env_offset=sl.start, device=device)
    hp.ep_len.copy_(ep0[sl])
    hp.command.copy_(cmd0[sl])
    hp.terrain_levels.copy_(levels0[sl])
    hp.sync_level_sum()
    hp.root_state.copy_(root0[sl].to(hp.device))     # S_prev of the first step's body-frame velocities (D7)
    hp.body_frame()
    return hp
