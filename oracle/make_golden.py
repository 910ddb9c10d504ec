"""ORACLE TOOLING (container-only): generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

    python -m oracle.make_golden            # needs /root/reference

Every fixture stores, per control step, the synthetic simulator state that was injected and
everything the reference produced, so that both ``oracle/shifu_oracle.py`` and the CUDA path can be
replayed against it without the reference tree (which does not exist on the GPU box).

Crafted micro-cases inside ``a1_small``: identity quaternion (env 4), 90-degree yaw
(env 5), base contact force of norm exactly 1.0 (envs 6, 7: must NOT terminate), robots far off the
height map (envs 1-3, index clamps), ``ep_len`` 499/500/501 boundary (envs 8-10), terrain-level
up / down / wrap-around (random initial levels incl. max-1), all-envs reset (record 0 is
``env.reset()``), and a step with zero resets (step 5; ``extras`` must persist).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(_REPO, "tests", "golden")
SMALL_TERRAIN = dict(num_rows=3, num_cols=4, border_size=5, max_init_terrain_level=2)
ZERO_RESET_STEP = 5


def small_hook(step, snap):
    """Crafted rows, see module docstring."""
    r = snap.root_offset
    r[4, 3:7] = torch.tensor([0., 0., 0., 1.])
    s = float(np.sin(np.pi / 4))
    r[5, 3:7] = torch.tensor([0., 0., s, s])
    snap.contact[6, 0] = torch.tensor([1.0, 0.0, 0.0])
    snap.contact[7, 0] = torch.tensor([0.6, 0.8, 0.0])
    if step == ZERO_RESET_STEP:
        snap.contact[:, 0] = 0.


def _pack(prefix, d, out):
    for k, v in d.items():
        out[f"{prefix}/{k}"] = v


def gen_a1(ns, name, n, steps, terrain, seed, store_map, hook=None, snap_kw=None):
    env = rh.make_a1(ns, n, terrain=terrain)
    isg = env.isg_env
    rs = np.random.RandomState(seed)
    ep = rs.randint(0, 480, size=n).astype(np.int64)
    ep[8:11] = [499, 500, 501]
    max_level = int(isg.max_terrain_level)
    lv = rs.randint(0, max_level, size=n).astype(np.int64)
    out = {}
    hs = isg.height_samples.numpy()
    meta = dict(name=name, n=n, steps=steps, seed=seed, rng_seed=0x5EED, terrain=terrain or {},
                map_shape=list(hs.shape), map_sha256=hashlib.sha256(hs.tobytes()).hexdigest(),
                max_terrain_level=max_level, border_size=float(isg.terrain.cfg.border_size),
                num_cols=int(isg.cfg.terrain.num_cols), snap_kw=snap_kw or {}, map_seed=0,
                zero_reset_step=ZERO_RESET_STEP if hook is not None else -1)
    if store_map:
        out["height_samples"] = hs.copy()
    out["terrain_origins"] = isg.terrain_origins.numpy().copy()
    out["terrain_types"] = isg.terrain_types.numpy().copy()
    out["env_origins_init"] = isg.env_origins.numpy().copy()
    out["ep_len_init"] = ep
    out["levels_init"] = lv
    rec, inp, rid = rh.run_a1(ns, env, seed=seed, steps=steps, ep_len_init=ep, levels_init=lv,
                              snap_kw=snap_kw, snap_hook=hook)
    for t, (r, i, ids) in enumerate(zip(rec, inp, rid)):
        _pack(f"s{t}/in", i, out)
        _pack(f"s{t}/out", r, out)
        out[f"s{t}/reset_ids"] = ids[-1] if ids else np.zeros(0, dtype=np.int64)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    nres = [int(r["reset"].sum()) for r in rec]
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} KB, resets per record {nres}")


def gen_abb(ns, name, n, steps, seed):
    env = rh.make_abb(ns, n)
    rs = np.random.RandomState(seed)
    ep = rs.randint(0, 190, size=n).astype(np.int64)
    ep[:3] = [199, 200, 201]
    out = dict(ep_len_init=ep)
    rec, inp, rid = rh.run_abb(ns, env, seed=seed, steps=steps, ep_len_init=ep)
    for t, (r, i, ids) in enumerate(zip(rec, inp, rid), start=1):
        _pack(f"s{t}/in", i, out)
        _pack(f"s{t}/out", r, out)
        out[f"s{t}/reset_ids"] = ids[-1] if ids else np.zeros(0, dtype=np.int64)
    meta = dict(name=name, n=n, steps=steps, seed=seed, rng_seed=0x5EED)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} KB, resets {[int(r['reset'].sum()) for r in rec]}")


def arm_ik_inputs(seed, n):
    """Seeded inputs of the arm action path: random 6x6 Jacobians (some nearly rank deficient so the
    damping matters), end-effector poses around the ABB workspace, unit quaternions, actions."""
    g = torch.Generator().manual_seed(seed)
    j = (torch.rand(n, 6, 6, generator=g) * 2 - 1) * 0.6
    j[1::8, 2] = j[1::8, 1] * (1 + 1e-3)                 # two almost parallel rows
    j[2::8, 4] *= 1e-3                                    # one almost vanishing row
    ee_pos = (torch.rand(n, 3, generator=g) * 2 - 1) * torch.tensor([0.25, 0.25, 0.02]) + torch.tensor([0., 0., 0.125])
    q = torch.randn(n, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    q[3::8] = torch.tensor([0., 1., 0., 0.])             # exactly the target orientation -> w = 0 after conj*mul
    dof_pos = (torch.rand(n, 6, generator=g) * 2 - 1)
    actions = (torch.rand(n, 3, generator=g) * 2 - 1)
    gq = torch.randn(n, 4, generator=g)
    goal = torch.cat([ee_pos + 0.05 * torch.randn(n, 3, generator=g), gq / gq.norm(dim=1, keepdim=True)], dim=1)
    return dict(j_ee=j, ee_pose=torch.cat([ee_pos, q], dim=1), dof_pos=dof_pos, actions=actions, goal_pose=goal)


def gen_arm_ik(ns, name, n, seed):
    """Row N2: (a) AbbRobot.step of the unmodified reference env, (b) the reference's stand-alone
    shifu.utils.torch_utils.inverse_kinematics with free goal poses."""
    import importlib
    env = rh.make_abb(ns, n)
    rb = env.robot
    d = arm_ik_inputs(seed, n)
    ee_row = int(rb.ee_indices[0])
    rb.body_state[:, ee_row, :7] = d["ee_pose"]            # view into the simulator tensor
    rb.dof_pos[:] = d["dof_pos"]
    rb.j_ee = d["j_ee"].clone()
    rb.step(d["actions"].clone())                         # a_prior_stage.py:67-73
    out = {"in/" + k: v.numpy() for k, v in d.items()}
    out["out/dof_targets_step"] = rb.dof_targets.detach().numpy().copy()
    tu = importlib.import_module("shifu.utils.torch_utils")
    out["out/dof_targets_goal"] = tu.inverse_kinematics(
        d["dof_pos"], d["ee_pose"][:, :3], d["ee_pose"][:, 3:7], d["goal_pose"][:, :3], d["goal_pose"][:, 3:7],
        d["j_ee"], "cpu").numpy().copy()
    meta = dict(name=name, n=n, seed=seed, ee_velocity=float(rb.end_effector_velocity), dt=float(rb.env.dt),
                min_ee_pos=[float(x) for x in rb.min_ee_pos], max_ee_pos=[float(x) for x in rb.max_ee_pos],
                tar_quat=[0., 1., 0., 0.], damping=0.05)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} KB")


def gen_camera(ns, name, n, height, width):
    """Row N4: the unmodified CameraSensor.refresh_image_tensors, driven through
    IsaacGymEnv.refresh_sensors (isaac_gym.py:159-170), for both image_normalization settings."""
    from shifu_b200.sim.fake_isaacgym import synthetic_camera_image
    types_ = [0, 1, 2, 3]
    out = {}
    for norm in (False, True):
        env = rh.make_camera_env(ns, n, height, width, norm, types_)
        env.isg_env.refresh_sensors()
        env.isg_env.refresh_sensors()                      # second frame: the buffers are overwritten
        frame = env.isg_env.sim.camera_frame
        cam = env.camera
        tag = "norm" if norm else "raw"
        out[f"{tag}/color"] = cam.color_buf.numpy().copy()
        out[f"{tag}/depth"] = cam.depth_buf.numpy().copy()
        out[f"{tag}/seg"] = cam.segmentation_buf.numpy().copy()
        out[f"{tag}/flow"] = cam.optical_flow_buf.numpy().copy()
    for t, key in enumerate(("color", "depth", "seg", "flow")):
        out[f"in/{key}"] = np.stack([synthetic_camera_image(e, t, frame, height, width).numpy() for e in range(n)])
    meta = dict(name=name, n=n, height=height, width=width, frame=frame)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} KB, frame {frame}")


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = rh.load_reference()
    gen_a1(ns, "a1_small", n=48, steps=6, terrain=SMALL_TERRAIN, seed=11, store_map=True,
           hook=small_hook, snap_kw=dict(p_base=0.12))
    gen_a1(ns, "a1_fullmap", n=40, steps=2, terrain=None, seed=12, store_map=False,
           snap_kw=dict(p_base=0.1))
    gen_abb(ns, "abb_small", n=48, steps=4, seed=13)
    gen_arm_ik(ns, "arm_ik", n=64, seed=14)
    gen_camera(ns, "camera", n=5, height=8, width=24)


if __name__ == "__main__":
    main()
