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


def _stateful_class(a1):
    class AirTime(a1.A1Conditional):                  # legged_gym's two remaining terms: dof limits + feet air time
        def __init__(self, cfg):
            super().__init__(cfg)
            self.swing_time = torch.zeros(self.num_envs, 4, device=self.device)
            self.last_contacts = torch.zeros(self.num_envs, 4, dtype=torch.bool, device=self.device)
            lo, hi = self.robot.dof_lower_limits, self.robot.dof_upper_limits
            mid, half = (lo + hi) / 2, 0.9 * (hi - lo) / 2          # legged_gym soft_dof_pos_limit = 0.9
            self.dof_pos_limits = torch.stack([mid - half, mid + half], dim=1)

        def reset_idx(self, env_ids):
            super().reset_idx(env_ids)
            self.swing_time[env_ids] = 0.
            self.last_contacts[env_ids] = False

        def build_reward_functions(self):
            return [self.tracking_lin_vel, self.tracking_ang_vel, self._reward_feet_air_time,
                    self._reward_dof_pos_limits, self.leg_collision, self.torques_penalize]

        def _reward_dof_pos_limits(self):
            out_of_limits = -(self.robot.dof_pos - self.dof_pos_limits[:, 0]).clip(max=0.)
            out_of_limits += (self.robot.dof_pos - self.dof_pos_limits[:, 1]).clip(min=0.)
            return -10.0 * torch.sum(out_of_limits, dim=1)

        def _reward_feet_air_time(self):
            contact = self.robot.contact_forces[:, self.robot.ee_indices, 2] > 1.
            contact_filt = torch.logical_or(contact, self.last_contacts)
            self.last_contacts[:] = contact
            first_contact = (self.swing_time > 0.) * contact_filt
            self.swing_time += self.isg_env.dt
            rew = torch.sum((self.swing_time - 0.5) * first_contact, dim=1)
            rew *= torch.norm(self.command_buf[:, :2], dim=1) > 0.1
            self.swing_time *= ~contact_filt
            return 1.0 * rew

    return AirTime


@needs_examples
def test_stateful_legged_gym_terms_fused_vs_hooks():
    """Row N1: feet_air_time (swing_time / last_contacts carried across steps and zeroed by resets) and
    dof_pos_limits through the term compiler; the fused env and the same class in user-hook mode replay the
    same steps and must agree on rewards, episode sums and the air-time state at every step."""
    from shifu_b200.sim import fake_isaacgym
    from shifu_b200.sim.synthetic import A1Replay
    a1, _ = _examples()
    cls = _stateful_class(a1)
    z, meta = util.load_golden("a1_small")
    n = meta["n"]

    def make(fuse):
        fake_isaacgym.reset_gym()
        cfg = a1.A1EnvConfig()
        cfg.num_envs, cfg.device, cfg.rng_seed, cfg.carry_body_frame = n, "cuda:0", meta["rng_seed"], False
        for k, v in meta["terrain"].items():
            setattr(cfg.terrain, k, v)
        np.random.seed(0)
        env = cls(cfg)
        env.auto_fuse = fuse

        class Replay(A1Replay):
            def begin_step(self, step):
                self.snap = util.golden_snap(z, step)
                self._dof_i = 0
                self.enabled = True
                return self.snap.actions

        env.replay = Replay(0, n, lambda: env.isg_env.env_origins)
        env.isg_env.sim.provider = env.replay
        return env

    fused, hooks = make(True), make(False)
    fused.replay.begin_step(0)
    fused.reset()
    assert fused.fusion_report == "fused: a1", fused.fusion_report
    d = fused.hot.desc
    codes = [int(d.reward_terms[i]) for i in range(d.num_reward_terms)]
    assert codes == [0, 1, 15, 14, 4, 5]
    assert (np.float32(d.reward_params[2][0]), np.float32(d.reward_params[2][1])) == (np.float32(1.0), np.float32(0.5))
    assert np.float32(d.reward_params[3][0]) == np.float32(-10.0)
    assert (d.num_feet, list(d.feet_bodies)) == (4, [int(b) for b in fused.robot.ee_indices.tolist()])
    assert (np.float32(d.feet_contact_force), np.float32(d.air_time_cmd_min)) == (np.float32(1.0), np.float32(0.1))
    assert np.float32(d.air_time_dt) == np.float32(fused.isg_env.dt) and d.air_time_reset == 1
    assert np.allclose(np.array(d.dof_pos_limit_low), fused.dof_pos_limits[:, 0].cpu().numpy())
    assert fused.hot.swing_time.data_ptr() == fused.swing_time.data_ptr()
    with _philox_draws(a1, hooks, meta["rng_seed"]):
        hooks.replay.begin_step(0)
        hooks.reset()
        assert hooks.fusion_report.startswith("not attempted")
        for env in (fused, hooks):
            env.episode_length_buf = torch.from_numpy(z["ep_len_init"]).cuda()
            env.terrain_levels[:] = torch.from_numpy(z["levels_init"]).cuda()
        fused.hot.sync_level_sum()
        landed = 0
        for t in range(1, meta["steps"] + 1):
            out = []
            for env in (fused, hooks):
                actions = env.replay.begin_step(t)
                obs, _, rew, dones, _ = env.step(actions.cuda())
                out.append(dict(obs=obs, rew=rew, dones=dones, swing=env.swing_time, last=env.last_contacts,
                                **{"sum/" + k: v for k, v in env.episode_rewards.items()}))
            for k in out[0]:
                a, b = out[0][k].float(), out[1][k].float()
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (t, k, float((a - b).abs().max()))
            landed += int((out[0]["sum/_reward_feet_air_time"] != 0).sum())
        assert landed > 0 and float(fused.swing_time.max()) > 0     # the term fired and the state is live


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
