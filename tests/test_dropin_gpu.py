"""The reference's OWN example classes on this package (BASELINE.json north_star: "examples/a1_conditional
and abb_pushbox_vision drop onto it unchanged").

``oracle/copy_ref_examples.py`` copies the two example packages verbatim into ``oracle/_ref/examples``
(git-ignored, travels to the GPU box); here they are imported through the ``shifu`` namespace
(``shifu/__init__.py`` -> ``shifu_b200``) and replay the golden fixtures recorded from the unmodified
reference — once automatically fused (``ShifuVecEnv.auto_fuse``: hooks proved equal to the kernel, then
replaced by it) and once in user-hook mode (the reference's torch hooks on CUDA tensors)."""
import importlib
import os
import sys
from contextlib import contextmanager, nullcontext

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
needs_examples = pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "examples", "MANIFEST.txt")),
                                    reason="run `python oracle/copy_ref_examples.py` where /root/reference exists")


def _examples():
    from shifu_b200.sim import fake_isaacgym
    fake_isaacgym.install("cuda:0")
    fake_isaacgym.reset_gym()
    fake_isaacgym.set_default_device("cuda:0")
    import shifu  # noqa: F401  (the drop-in namespace)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    for k in [k for k in sys.modules if k == "examples" or k.startswith("examples.")]:
        del sys.modules[k]
    a1 = importlib.import_module("examples.a1_conditional.a1_conditional")
    abb = importlib.import_module("examples.abb_pushbox_vision.a_prior_stage")
    assert a1.ShifuVecEnv.__module__.startswith("shifu_b200") and a1.__file__.startswith(REF_DIR)
    return a1, abb


@contextmanager
def _philox_draws(a1_mod, env, seed):
    """User-hook mode only: the reference draws from torch's global generator in data-dependent order;
    parity is defined on counter-based draws injected at its draw sites (same scheme as
    oracle/ref_harness.py)."""
    from shifu_b200.utils.philox import draw_randint, draw_u01
    state = {"ids": None, "step": 0, "cmd": 0}
    cls = type(env)
    orig_reset, orig_rand, orig_randint = cls.reset_idx, a1_mod.torch_rand_float, torch.randint_like

    def reset_idx(self, env_ids):
        state.update(ids=env_ids, step=int(self.common_step_counter), cmd=0)
        return orig_reset(self, env_ids)

    def torch_rand_float(lower, upper, shape, device):
        r, k = shape
        ids = state["ids"]
        assert ids is not None and r == len(ids)
        if k == 2:
            u = draw_u01(seed, ids, state["step"], 1, 2)
        elif k == 3:
            u = draw_u01(seed, ids, state["step"], 2, 3)
        else:
            lane = state["cmd"]
            state["cmd"] += 1
            u = draw_u01(seed, ids, state["step"], 3, 3)[:, lane:lane + 1]
        return (upper - lower) * u + lower

    def randint_like(t, high, **kw):
        return draw_randint(seed, state["ids"], state["step"], 0, int(high)).to(t.dtype)

    cls.reset_idx, a1_mod.torch_rand_float, torch.randint_like = reset_idx, torch_rand_float, randint_like
    try:
        yield
    finally:
        cls.reset_idx, a1_mod.torch_rand_float, torch.randint_like = orig_reset, orig_rand, orig_randint


def _a1_outputs(env):
    from tests.test_env_api_gpu import _env_outputs
    return _env_outputs(env)


@needs_examples
@pytest.mark.parametrize("fuse", [True, False])
def test_reference_a1_conditional_replays_golden(fuse):
    from shifu_b200.sim.synthetic import A1Replay
    a1, _ = _examples()
    z, meta = util.load_golden("a1_small")
    n = meta["n"]
    cfg = a1.A1EnvConfig()
    cfg.num_envs, cfg.device = n, "cuda:0"
    cfg.rng_seed = meta["rng_seed"]
    cfg.carry_body_frame = False
    cfg.store_measured_heights = True          # the fixture compares measured_heights of resetting envs too
    for k, v in meta["terrain"].items():
        setattr(cfg.terrain, k, v)
    np.random.seed(0)
    torch.manual_seed(0)
    env = a1.A1Conditional(cfg)                       # the reference's class, our ShifuVecEnv underneath
    env.auto_fuse = fuse
    isg = env.isg_env
    assert np.array_equal(isg.height_samples.cpu().numpy(), z["height_samples"])

    class Replay(A1Replay):
        def begin_step(self, step):
            self.snap = util.golden_snap(z, step)
            self._dof_i = 0
            self.enabled = True
            return self.snap.actions

    replay = Replay(0, n, lambda: isg.env_origins)
    isg.sim.provider = replay
    # fused: the kernel draws from the Philox streams itself; user-hook mode: inject them at the draw sites
    with (nullcontext() if fuse else _philox_draws(a1, env, meta["rng_seed"])):
        replay.begin_step(0)
        env.reset()
        assert env.fusion_report.startswith("fused: a1" if fuse else "not attempted"), env.fusion_report
        util.compare_a1(_a1_outputs(env), util.golden_out(z, 0), "ref-class/reset")
        env.episode_length_buf = torch.from_numpy(z["ep_len_init"]).cuda()
        env.terrain_levels[:] = torch.from_numpy(z["levels_init"]).cuda()
        if fuse:
            env.hot.sync_level_sum()
        for t in range(1, meta["steps"] + 1):
            actions = replay.begin_step(t)
            obs, priv, rew, dones, extras = env.step(actions.cuda())
            assert priv is None and obs.shape == (n, 259)
            util.compare_a1(_a1_outputs(env), util.golden_out(z, t), f"ref-class/s{t}")


@needs_examples
def test_reference_a1_with_edited_constant_or_shape():
    """A user who edits a literal inside a known term gets THAT constant in the fused kernel; a user who
    changes the shape of a term (or adds an unknown one) stays in user-hook mode, loudly."""
    a1, _ = _examples()
    z, meta = util.load_golden("a1_small")

    def make(cls):
        from shifu_b200.sim import fake_isaacgym
        fake_isaacgym.reset_gym()
        cfg = a1.A1EnvConfig()
        cfg.num_envs, cfg.device = meta["n"], "cuda:0"
        for k, v in meta["terrain"].items():
            setattr(cfg.terrain, k, v)
        np.random.seed(0)
        env = cls(cfg)
        env._maybe_fuse()
        return env

    class Edited(a1.A1Conditional):
        def tracking_lin_vel(self):
            err = torch.sum(torch.square(self.command_buf[:, :2] - self.robot.base_lin_vel[:, :2]), dim=1)
            return 1.5 * torch.exp(-err / 0.5)

        def torques_penalize(self):
            return -3e-5 * torch.sum(torch.square(self.robot.torques), dim=1)

    env = make(Edited)
    assert env.fusion_report == "fused: a1", env.fusion_report
    d = env.hot.desc
    assert (d.reward_params[0][0], d.reward_params[0][1]) == (1.5, 0.5)
    assert abs(d.reward_params[5][0] - (-3e-5)) < 1e-12

    class Reshaped(a1.A1Conditional):
        def tracking_ang_vel(self):                   # |err| instead of err^2: not the library term
            return 0.5 * torch.exp(-torch.abs(self.command_buf[:, 2] - self.robot.base_ang_vel[:, 2]) / 0.25)

    env = make(Reshaped)
    assert env.fusion_report.startswith("user-hook mode") and "tracking_ang_vel" in env.fusion_report

    class Extra(a1.A1Conditional):
        def build_reward_functions(self):
            return super().build_reward_functions() + [self.my_bonus]

        def my_bonus(self):
            return torch.ones(self.num_envs, device=self.device)

    env = make(Extra)
    assert env.fusion_report.startswith("user-hook mode") and "my_bonus" in env.fusion_report

    class LeggedGymStyle(a1.A1Conditional):           # row N1: legged_gym vocabulary through the term compiler
        def build_reward_functions(self):
            return [self.tracking_lin_vel, self._reward_orientation, self._reward_lin_vel_z, self._reward_ang_vel_xy,
                    self._reward_dof_vel, self._reward_action_rate, self._reward_base_height, self._reward_torques]

        def _reward_orientation(self):
            return -0.2 * torch.sum(torch.square(self.robot.projected_gravity[:, :2]), dim=1)

        def _reward_lin_vel_z(self):
            return -2.0 * torch.square(self.robot.base_lin_vel[:, 2])

        def _reward_ang_vel_xy(self):
            return -0.05 * torch.sum(torch.square(self.robot.base_ang_vel[:, :2]), dim=1)

        def _reward_dof_vel(self):
            return -1e-4 * torch.sum(torch.square(self.robot.dof_vel), dim=1)

        def _reward_action_rate(self):
            return -0.01 * torch.sum(torch.square(self.actions_recorder.get_last(0) - self.actions), dim=1)

        def _reward_base_height(self):
            return -1.0 * torch.square(self.robot.base_pose[:, 2] - 0.35)

        def _reward_torques(self):
            return -1e-5 * torch.sum(torch.square(self.robot.torques), dim=1)

    env = make(LeggedGymStyle)
    assert env.fusion_report == "fused: a1", env.fusion_report
    want = [(0, 1.0, 0.25), (10, -0.2, 0.0), (8, -2.0, 0.0), (9, -0.05, 0.0), (11, -1e-4, 0.0), (12, -0.01, 0.0),
            (13, -1.0, 0.35), (5, -1e-5, 0.0)]
    d = env.hot.desc
    for i, (code, p0, p1) in enumerate(want):       # the descriptor holds fp32: compare as fp32
        assert int(d.reward_terms[i]) == code
        assert np.float32(d.reward_params[i][0]) == np.float32(p0) and np.float32(d.reward_params[i][1]) == np.float32(p1)
    # and the fused step agrees with the Python hooks on a live step
    from shifu_b200.sim.synthetic import A1Replay
    replay = A1Replay(3, meta["n"], lambda: env.isg_env.env_origins)
    env.isg_env.sim.provider = replay
    env.reset()
    actions = replay.begin_step(1)
    obs, _, rew, dones, _ = env.step(actions.cuda())
    assert torch.isfinite(rew).all() and obs.shape == (meta["n"], 259)


@needs_examples
def test_reference_abb_pushbox_replays_golden():
    from shifu_b200.sim.synthetic import AbbReplay
    _, abb = _examples()
    z, meta = util.load_golden("abb_small")
    n = meta["n"]
    cfg = abb.PriorStageEnvConfig()
    cfg.num_envs, cfg.device = n, "cuda:0"
    cfg.rng_seed = meta["rng_seed"]
    env = abb.AbbPushBox(cfg)                          # the reference's class

    class Replay(AbbReplay):
        def begin_step(self, step):
            self.snap = util.golden_snap(z, step, "abb")
            self.enabled = True
            return self.snap.actions

    replay = Replay(0, n)
    env.isg_env.sim.provider = replay
    env._maybe_fuse()
    assert env.fusion_report == "fused: abb", env.fusion_report
    env.episode_length_buf = torch.from_numpy(z["ep_len_init"]).cuda()
    for t in range(1, meta["steps"] + 1):
        actions = replay.begin_step(t)
        obs, _, rew, dones, extras = env.step(actions.cuda())
        want = util.golden_out(z, t)
        got = dict(obs=obs, rew=rew, reset=dones.to(torch.uint8), time_out=env.time_out_buf.to(torch.uint8),
                   success=env.success_buf.to(torch.uint8), ep_len=env.episode_length_buf,
                   root_state=env.isg_env.root_state, dof_state=env.isg_env.dof_state)
        for k, v in env.episode_rewards.items():
            got["ep_sum/" + k] = v
        for k, v in extras["episode"].items():
            got["extras/" + k] = v
        got = {k: v.detach().cpu().numpy() for k, v in got.items()}
        util.compare_a1(got, want, f"ref abb/s{t}", skip=("dof_targets",))


@needs_examples
def test_run_policy_random_mode_on_reference_class():
    """shifu.runner.run_policy('random', ...) — the reference's manual integration check
    (policy_runner.py:33-41, README.md:32) — drives the unmodified class through reset() and step()."""
    a1, _ = _examples()
    from shifu.runner import run_policy
    z, meta = util.load_golden("a1_small")
    cfg = a1.A1EnvConfig()
    cfg.device = "cuda:0"
    for k, v in meta["terrain"].items():
        setattr(cfg.terrain, k, v)
    np.random.seed(0)
    env = run_policy("random", a1.A1Conditional, cfg, a1.A1PPOConfig(), play_num_envs=64, play_iterations=5)
    assert env.num_envs == 64 and env.common_step_counter == 6 and env.fusion_report == "fused: a1"
    assert torch.isfinite(env.obs_buf).all()
