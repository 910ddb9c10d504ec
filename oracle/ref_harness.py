"""ORACLE TOOLING (container-only): run the UNMODIFIED reference on torch CPU.

The reference (``/root/reference``) is pure Python on top of Isaac Gym; it has no
tests, golden vectors or fixtures for the hot path.  This
harness imports it *unmodified* with ``isaacgym`` / ``rsl_rl`` / ``matplotlib`` /
``pybullet`` replaced by stand-ins (``shifu_b200.sim.fake_isaacgym``), feeds it
seeded synthetic simulator state, replaces its reset-time random draws with
counter-based Philox draws (``oracle/philox_np.py``; the reference draws from
global generators) and records everything the step produces.

It exists to (1) generate the golden fixtures under ``tests/golden/`` and
(2) pin ``oracle/shifu_oracle.py`` (the restatement that travels to the GPU box).
``/root/reference`` does not exist on the GPU box, so nothing under ``-m gpu``,
``smoke()`` or ``bench.py`` imports this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from contextlib import contextmanager
from typing import Dict, List, Optional

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("SHIFU_REFERENCE_ROOT", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from oracle import philox_np as px  # noqa: E402


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "shifu", "gym"))


def _placeholder(name):
    import importlib.machinery
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []

    def _missing(attr):
        raise AttributeError(f"{name}.{attr} is a harness placeholder")

    m.__getattr__ = lambda attr: types.SimpleNamespace() if not attr.startswith("__") else _missing(attr)
    return m


def load_reference():
    """Import the reference package tree under the stand-ins; returns a namespace of modules."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    from shifu_b200.sim import fake_isaacgym
    fake_isaacgym.install(device="cpu", force=True)
    for name in ("matplotlib", "matplotlib.pyplot", "pybullet", "pybullet_data"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = _placeholder(name)
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # make `import shifu` / `import examples` resolve to the reference
    for k in [k for k in sys.modules if k == "shifu" or k.startswith("shifu.") or k == "examples"
              or k.startswith("examples.")]:
        del sys.modules[k]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.a1 = importlib.import_module("examples.a1_conditional.a1_conditional")
    ns.a1_cfg = importlib.import_module("examples.a1_conditional.task_config")
    ns.abb = importlib.import_module("examples.abb_pushbox_vision.a_prior_stage")
    ns.abb_cfg = importlib.import_module("examples.abb_pushbox_vision.task_config")
    ns.isaac_gym = importlib.import_module("shifu.gym.isaac_gym")
    ns.env = importlib.import_module("shifu.gym.env")
    ns.fake = fake_isaacgym
    assert ns.env.__file__.startswith(REFERENCE_ROOT)
    return ns


# ---------------------------------------------------------------------------
# RNG injection
# ---------------------------------------------------------------------------


class DrawContext:
    """Which (env_ids, step) the reference is currently resetting."""

    def __init__(self, seed: int, env_offset: int = 0):
        self.seed = seed
        self.env_offset = env_offset   # global id of local env 0 (multi-GPU sharding)
        self.env_ids: Optional[np.ndarray] = None
        self.step = 0
        self.cmd_lane = 0
        self.box_stream = None
        self.box_calls = 0
        self.reset_log: List[np.ndarray] = []

    def begin(self, env_ids, step):
        self.env_ids = np.asarray(env_ids.cpu().numpy(), dtype=np.int64) + self.env_offset
        self.step = int(step)
        self.cmd_lane = 0

    def u32(self, stream):
        return px.draw_u32(self.seed, self.env_ids, self.step, stream)


def _patched_rand_float(ctx: DrawContext):
    def torch_rand_float(lower, upper, shape, device):
        r, k = shape
        assert ctx.env_ids is not None and r == len(ctx.env_ids), "draw outside a tracked reset"
        if k == 2:
            u = px.u01_f32(ctx.u32(px.STREAM_XY)[:2]).T
        elif k == 3:
            u = px.u01_f32(ctx.u32(px.STREAM_FORCE)[:3]).T
        elif k == 1:
            u = px.u01_f32(ctx.u32(px.STREAM_CMD)[ctx.cmd_lane:ctx.cmd_lane + 1]).T
            ctx.cmd_lane += 1
        else:
            raise AssertionError(f"unexpected draw shape {shape}")
        # same expression as isaacgym.torch_utils.torch_rand_float with rand() injected
        return (upper - lower) * torch.from_numpy(np.ascontiguousarray(u)) + lower
    return torch_rand_float


def _patched_randint_like(ctx: DrawContext):
    def randint_like(t, high, **kw):
        assert ctx.env_ids is not None and t.shape[0] == len(ctx.env_ids)
        return torch.from_numpy(px.randint10(ctx.u32(px.STREAM_LEVEL)[0], int(high))).to(t.dtype)
    return randint_like


def _patched_np_uniform(ctx: DrawContext):
    def uniform(low=0.0, high=1.0, size=None):
        assert ctx.box_stream is not None, "np.random.uniform outside a tracked box reset"
        low = np.asarray(low, dtype=np.float64)
        high = np.asarray(high, dtype=np.float64)
        r = len(ctx.env_ids)
        i = ctx.box_calls % r
        stream = ctx.box_stream + (ctx.box_calls // r)     # first R calls: position, next R: euler
        ctx.box_calls += 1
        u = px.u01_f64(px.draw_u32(ctx.seed, ctx.env_ids[i:i + 1], ctx.step, stream)[:3, 0])
        return low + (high - low) * u
    return uniform


@contextmanager
def injected_draws(ns, ctx: DrawContext, env):
    """Patch the draw sites for the duration of a reference call."""
    a1_mod, abb_mod = ns.a1, ns.abb
    saved = (a1_mod.torch_rand_float, torch.randint_like, np.random.uniform)
    cls = type(env)
    orig_reset = cls.reset_idx
    box_cls = abb_mod.RandPosBox
    orig_box_reset = box_cls._reset_root_state

    def reset_idx(self, env_ids):
        ctx.begin(env_ids, self.common_step_counter)
        ctx.reset_log.append(np.asarray(env_ids.cpu().numpy(), dtype=np.int64).copy())
        return orig_reset(self, env_ids)

    def box_reset(self, env_ids):
        ctx.box_stream = px.STREAM_GOAL_POS if isinstance(self, abb_mod.GoalBox) else px.STREAM_CUBE_POS
        ctx.box_calls = 0
        try:
            return orig_box_reset(self, env_ids)
        finally:
            ctx.box_stream = None

    a1_mod.torch_rand_float = _patched_rand_float(ctx)
    torch.randint_like = _patched_randint_like(ctx)
    np.random.uniform = _patched_np_uniform(ctx)
    cls.reset_idx = reset_idx
    box_cls._reset_root_state = box_reset
    try:
        yield
    finally:
        a1_mod.torch_rand_float, torch.randint_like, np.random.uniform = saved
        cls.reset_idx = orig_reset
        box_cls._reset_root_state = orig_box_reset


# ---------------------------------------------------------------------------
# A1
# ---------------------------------------------------------------------------


def make_a1(ns, n: int, terrain: Optional[Dict] = None, map_seed: int = 0):
    """Construct the reference ``A1Conditional`` on CPU.  ``terrain`` overrides attributes of
    ``cfg.terrain`` (config values only — e.g. a smaller tile grid for compact fixtures)."""
    ns.fake.reset_gym()
    ns.fake.set_default_device("cpu")
    cfg = ns.a1_cfg.A1EnvConfig()
    cfg.num_envs = n
    cfg.device = "cpu"
    for k, v in (terrain or {}).items():
        setattr(cfg.terrain, k, v)
    np.random.seed(map_seed)
    torch.manual_seed(map_seed)
    env = ns.a1.A1Conditional(cfg)
    return env


def a1_state(env) -> Dict[str, np.ndarray]:
    """Everything the step left behind, as numpy copies."""
    isg, rb = env.isg_env, env.robot
    d = dict(
        obs=env.obs_buf, rew=env.rew_buf, reset=env.reset_buf.to(torch.uint8),
        time_out=env.time_out_buf.to(torch.uint8),
        contact_term=env.contact_terminate_buf.to(torch.uint8),
        ep_len=env.episode_length_buf, terrain_levels=env.terrain_levels,
        env_origins=isg.env_origins, command=env.command_buf,
        history=env.actions_recorder.history_buf, actions=env.actions,
        root_state=isg.root_state, dof_state=isg.dof_state, dof_targets=rb.dof_targets,
        rand_force=rb.rand_force_buf, torques=rb.torques,
        base_lin_vel=rb.base_lin_vel, base_ang_vel=rb.base_ang_vel,
        projected_gravity=rb.projected_gravity,
    )
    if hasattr(isg, "measured_heights"):
        d["measured_heights"] = isg.measured_heights
    for k, v in env.episode_rewards.items():
        d["ep_sum/" + k] = v
    ep = env.extras.get("episode", {})
    for k, v in ep.items():
        d["extras/" + k] = torch.as_tensor(v)
    if "time_outs" in env.extras:
        d["extras_time_outs"] = env.extras["time_outs"].to(torch.uint8)
    return {k: v.detach().cpu().numpy().copy() for k, v in d.items()}


def run_a1(ns, env, seed: int, steps: int, *, rng_seed: int = 0x5EED, do_reset: bool = True,
           ep_len_init: Optional[np.ndarray] = None, snap_kw: Optional[Dict] = None,
           env_offset: int = 0, levels_init: Optional[np.ndarray] = None, snap_hook=None):
    """Drive the reference for ``steps`` control steps on replayed synthetic state.

    Returns ``(records, inputs, reset_ids)``: per-step output dicts, per-step injected inputs, and
    the env-id lists handed to ``reset_idx``."""
    from shifu_b200.sim.synthetic import A1Replay
    n = env.num_envs
    sim = env.isg_env.sim
    replay = A1Replay(seed, n, lambda: env.isg_env.env_origins, snap_hook=snap_hook, **(snap_kw or {}))
    sim.provider = replay
    ctx = DrawContext(rng_seed, env_offset)
    records, inputs, reset_ids = [], [], []

    def _inputs(snap):
        return dict(dof=snap.dof.numpy().copy(), root_offset=snap.root_offset.numpy().copy(),
                    contact=snap.contact.numpy().copy(), actions=snap.actions.numpy().copy())

    with injected_draws(ns, ctx, env):
        t = 0
        if do_reset:
            replay.begin_step(0)
            inputs.append(_inputs(replay.snap))
            ctx.reset_log.clear()
            env.reset()
            reset_ids.append([a.copy() for a in ctx.reset_log])
            records.append(a1_state(env))
            t = 1
        if ep_len_init is not None:
            env.episode_length_buf[:] = torch.from_numpy(ep_len_init)
        if levels_init is not None:
            env.terrain_levels[:] = torch.from_numpy(levels_init)
        for s in range(t, t + steps):
            actions = replay.begin_step(s)
            inputs.append(_inputs(replay.snap))
            ctx.reset_log.clear()
            env.step(actions.clone())
            reset_ids.append([a.copy() for a in ctx.reset_log])
            records.append(a1_state(env))
    return records, inputs, reset_ids


# ---------------------------------------------------------------------------
# ABB prior stage
# ---------------------------------------------------------------------------


def make_abb(ns, n: int):
    ns.fake.reset_gym()
    ns.fake.set_default_device("cpu")
    cfg = ns.abb_cfg.PriorStageEnvConfig()
    cfg.num_envs = n
    cfg.device = "cpu"
    torch.manual_seed(0)
    np.random.seed(0)
    return ns.abb.AbbPushBox(cfg)


def make_camera_env(ns, n: int, height: int, width: int, image_normalization: bool, image_types):
    """The reference's ABB env with a CameraSensor, built like VisionAbbPushBox.__init__
    (examples/abb_pushbox_vision/b_regression_stage.py:34-50) from the reference's own classes."""
    ns.fake.reset_gym()
    ns.fake.set_default_device("cpu")
    cfg = ns.abb_cfg.PriorStageEnvConfig()
    cfg.num_envs = n
    cfg.device = "cpu"
    units = importlib.import_module("shifu.units")
    h, w, norm, types_ = height, width, image_normalization, list(image_types)

    class CamCfg(ns.abb_cfg.PushBoxCameraConfig):
        image_types = types_
        image_normalization = norm

        class camera_props(ns.abb_cfg.PushBoxCameraConfig.camera_props):
            width = w
            height = h

    abb = ns.abb

    class VisionEnv(abb.AbbPushBox):
        def __init__(self, cfg):
            super(abb.AbbPushBox, self).__init__(cfg)
            self.robot = abb.AbbRobot(ns.abb_cfg.AbbRobotConfig())
            self.table = units.Box(ns.abb_cfg.TableConfig())
            self.cube = abb.RandPosBox(ns.abb_cfg.PushBoxConfig())
            self.goal = abb.GoalBox(ns.abb_cfg.GoalBoxConfig())
            self.camera = units.CameraSensor(CamCfg())
            self.isg_env.create_envs(robot=self.robot, objects=[self.table, self.cube, self.goal],
                                     sensors=[self.camera])
            self.success_buf = torch.zeros(self.num_envs, device=self.device, dtype=torch.float)

    torch.manual_seed(0)
    np.random.seed(0)
    return VisionEnv(cfg)


def abb_state(env) -> Dict[str, np.ndarray]:
    isg = env.isg_env
    d = dict(obs=env.obs_buf, rew=env.rew_buf, reset=env.reset_buf.to(torch.uint8),
             time_out=env.time_out_buf.to(torch.uint8), success=env.success_buf.to(torch.uint8),
             ep_len=env.episode_length_buf, root_state=isg.root_state, dof_state=isg.dof_state,
             dof_targets=env.robot.dof_targets)
    for k, v in env.episode_rewards.items():
        d["ep_sum/" + k] = v
    for k, v in env.extras.get("episode", {}).items():
        d["extras/" + k] = torch.as_tensor(v)
    if "time_outs" in env.extras:
        d["extras_time_outs"] = env.extras["time_outs"].to(torch.uint8)
    return {k: v.detach().cpu().numpy().copy() for k, v in d.items()}


def run_abb(ns, env, seed: int, steps: int, *, rng_seed: int = 0x5EED,
            ep_len_init: Optional[np.ndarray] = None, snap_kw: Optional[Dict] = None):
    from shifu_b200.sim.synthetic import AbbReplay
    n = env.num_envs
    sim = env.isg_env.sim
    replay = AbbReplay(seed, n, **(snap_kw or {}))
    sim.provider = replay
    ctx = DrawContext(rng_seed)
    records, inputs, reset_ids = [], [], []
    with injected_draws(ns, ctx, env):
        if ep_len_init is not None:
            env.episode_length_buf[:] = torch.from_numpy(ep_len_init)
        for s in range(1, steps + 1):
            actions = replay.begin_step(s)
            inputs.append(dict(root=replay.snap.root.numpy().copy(), body=replay.snap.body.numpy().copy(),
                               dof=replay.snap.dof.numpy().copy(), actions=actions.numpy().copy()))
            ctx.reset_log.clear()
            env.step(actions.clone())
            reset_ids.append([a.copy() for a in ctx.reset_log])
            records.append(abb_state(env))
    return records, inputs, reset_ids
