"""Reward-term compiler: ``build_reward_functions()`` lists -> fused term descriptors.

The reference's "registry" is a plain list of bound methods whose weights are literals inside each
body (``shifu/gym/env.py:78-80,160-166,180-185``; ``a1_conditional.py:152-192``).  To fuse such a
list the kernel needs, per term, an opcode of its term library (``enum ShifuRewardTerm``) and the
term's constants.  This module gets both WITHOUT reading source text:

1. **match** — the method name (``fn.__name__``, prefixes ``_reward_`` / ``reward_`` / ``rew_`` stripped)
   selects a library entry; every entry is a parametric form ``f(state; p0, p1)``;
2. **fit** — the user's Python hook is evaluated on a handful of crafted device states (all zeros
   except one input) which determine ``p0`` / ``p1`` exactly or — for the two constants that only
   come out of a logarithm — up to rounding, in which case the value is snapped to the matching
   literal in the method's ``co_consts``;
3. **verify** — the hook and the kernel (``shifu_a1_eval_terms``) are evaluated on the same seeded
   random state; any term that disagrees beyond the parity tolerance makes the compiler refuse
   (the env then stays in user-hook mode: torch hooks on CUDA tensors).

So a user who edits a literal inside a known term gets the edited constant in the fused kernel, and
a user who changes the *shape* of a term gets a loud refusal instead of silently different rewards.
"""
from __future__ import annotations

import math
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as nv

RTOL, ATOL = 2e-5, 2e-6


class TermMismatch(Exception):
    """The user's hook is not (numerically) the library term it was matched to."""


@dataclass
class CompiledTerm:
    name: str
    code: int
    p0: float
    p1: float
    extra: Optional[dict] = None        # descriptor-level constants of the term (a1_desc keyword -> value)


def _strip(name: str) -> str:
    for prefix in ("_reward_", "reward_", "rew_"):
        if name.startswith(prefix):
            return name[len(prefix):]
    return name


# name -> opcode (a1_conditional.py names + the legged_gym vocabulary)
A1_NAMES: Dict[str, int] = {
    "tracking_lin_vel": nv.REW_TRACKING_LIN_VEL, "tracking_ang_vel": nv.REW_TRACKING_ANG_VEL,
    "stabilizing_base": nv.REW_STABILIZING_BASE, "smoothing_action": nv.REW_SMOOTHING_ACTION,
    "leg_collision": nv.REW_LEG_COLLISION, "collision": nv.REW_LEG_COLLISION,
    "torques_penalize": nv.REW_TORQUES, "torques": nv.REW_TORQUES,
    "lin_vel_z": nv.REW_LIN_VEL_Z, "ang_vel_xy": nv.REW_ANG_VEL_XY, "orientation": nv.REW_ORIENTATION,
    "dof_vel": nv.REW_DOF_VEL, "action_rate": nv.REW_ACTION_RATE, "base_height": nv.REW_BASE_HEIGHT,
    "dof_pos_limits": nv.REW_DOF_POS_LIMITS, "feet_air_time": nv.REW_FEET_AIR_TIME,
}


def air_time_state(env):
    """(swing_time, last_contacts) of an env that carries the legged_gym feet-air-time state, else None."""
    swing = getattr(env, "swing_time", None)
    swing = getattr(env, "feet_air_time", None) if swing is None else swing
    last = getattr(env, "last_contacts", None)
    if torch.is_tensor(swing) and torch.is_tensor(last):
        return swing, last
    return None


def dof_limits(env) -> torch.Tensor:
    """(num_dof, 2) limits a dof-limit term reads: ``env.dof_pos_limits`` (legged_gym's soft limits) when the
    env defines it, else the asset limits of the robot (units/robot.py:39-40)."""
    lim = getattr(env, "dof_pos_limits", None)
    if torch.is_tensor(lim):
        return lim
    rb = env.robot
    return torch.stack([rb.dof_lower_limits, rb.dof_upper_limits], dim=1)


class A1Probe:
    """Crafted states for an A1-shaped env: every tensor a reward / observation / termination hook may
    read is saved, overwritten and restored.  Works on the env's OWN tensors (the hooks close over
    them), so it must run before the first real step or between steps."""

    def __init__(self, env):
        self.env, self.robot, self.isg = env, env.robot, env.isg_env
        rb, isg = self.robot, self.isg
        self.tensors = {
            "command": env.command_buf, "lin": rb.base_lin_vel, "ang": rb.base_ang_vel, "pg": rb.projected_gravity,
            "history": env.actions_recorder.history_buf, "contact": isg.contact_state, "torques": rb.torques,
            "dof": isg.dof_state, "root": isg.root_state, "actions": env.actions,
            "ep_len": env.episode_length_buf,
        }
        if getattr(isg, "measured_heights", None) is not None:
            self.tensors["heights"] = isg.measured_heights
        air = air_time_state(env)
        if air is not None:
            self.tensors["swing_time"], self.tensors["last_contacts"] = air
        self._saved = None
        self._attrs = ("obs_buf", "reset_buf", "time_out_buf", "contact_terminate_buf", "rew_buf")

    def __enter__(self):
        self._saved = {k: t.clone() for k, t in self.tensors.items()}
        self._saved_attrs = {a: getattr(self.env, a, None) for a in self._attrs}
        self._saved_torques_obj = self.robot.torques
        return self

    def __exit__(self, *exc):
        for k, t in self.tensors.items():
            t.copy_(self._saved[k])
        for a, v in self._saved_attrs.items():
            if v is not None:
                setattr(self.env, a, v)
        self.robot.torques = self._saved_torques_obj
        return False

    def zero(self):
        for k, t in self.tensors.items():
            t.zero_()
        self.isg.root_state[:, 6] = 1.0            # identity quaternion

    def randomize(self, seed: int = 1234):
        g = torch.Generator(device=self.env.device).manual_seed(seed)
        for k, t in self.tensors.items():
            if t.dtype.is_floating_point:
                t.copy_(torch.randn(t.shape, generator=g, device=t.device) * (0.6 if k != "contact" else 0.4))
            elif t.dtype == torch.bool:
                t.copy_(torch.rand(t.shape, generator=g, device=t.device) < 0.5)
            else:
                t.copy_(torch.randint(0, int(self.env.max_episode_length) + 40, t.shape, generator=g, device=t.device))
        q = self.isg.root_state[:, 3:7]
        q.copy_(q / q.norm(dim=1, keepdim=True).clamp_min(1e-6))

    def value(self, fn: Callable) -> float:
        return float(fn()[0])


def _snap(value: float, fn: Callable, rel: float = 1e-3) -> float:
    """The literal of the method body nearest to a fitted constant (exact user value when present)."""
    consts = []
    code = getattr(fn, "__code__", None) or getattr(getattr(fn, "__func__", None), "__code__", None)
    for c in (code.co_consts if code is not None else ()):
        if isinstance(c, (int, float)) and not isinstance(c, bool):
            consts += [float(c), -float(c)]
    best = min(consts, key=lambda c: abs(c - value), default=None)
    if best is not None and abs(best - value) <= rel * max(abs(value), 1e-12):
        return best
    return value


def _fit_exp(pr: A1Probe, fn, setter) -> Tuple[float, float]:
    pr.zero()
    p0 = pr.value(fn)                               # exp(0) = 1
    if p0 == 0.0:
        raise TermMismatch("term is identically zero")
    pr.zero()
    setter(0.5)                                     # squared error 0.25
    r = pr.value(fn)
    ratio = r / p0
    if not (0.0 < ratio < 1.0):
        raise TermMismatch("not of the form p0*exp(-err/p1)")
    return p0, _snap(-0.25 / math.log(ratio), fn)


def fit_a1_term(pr: A1Probe, fn: Callable, code: int) -> Tuple[float, float]:
    env, rb, t = pr.env, pr.robot, pr.tensors
    one = lambda name, idx: (pr.zero(), t[name].__setitem__(idx, 1.0), pr.value(fn))[2]
    if code == nv.REW_TRACKING_LIN_VEL:
        return _fit_exp(pr, fn, lambda v: t["command"].__setitem__((0, 0), v))
    if code == nv.REW_TRACKING_ANG_VEL:
        return _fit_exp(pr, fn, lambda v: t["command"].__setitem__((0, 2), v))
    if code == nv.REW_STABILIZING_BASE:
        return one("lin", (0, 2)), one("ang", (0, 0))
    if code == nv.REW_SMOOTHING_ACTION:
        return one("history", (0, 0, 0)) / 2.0, 0.0             # |a1-a0|^2 + |a2-2a1+a0|^2 = 1 + 1
    if code == nv.REW_TORQUES:
        return one("torques", (0, 0)), 0.0
    if code == nv.REW_LIN_VEL_Z:
        return one("lin", (0, 2)), 0.0
    if code == nv.REW_ANG_VEL_XY:
        return one("ang", (0, 0)), 0.0
    if code == nv.REW_ORIENTATION:
        return one("pg", (0, 0)), 0.0
    if code == nv.REW_DOF_VEL:
        return one("dof", (0, 1)), 0.0                           # dof_state row 0 = (pos, vel) of dof 0
    if code == nv.REW_ACTION_RATE:
        return one("actions", (0, 0)), 0.0
    if code == nv.REW_BASE_HEIGHT:
        vals = []
        for z in (0.0, 1.0, 2.0):
            pr.zero()
            rows = rb.root_indices[0]
            t["root"][rows, 2] = z
            vals.append(pr.value(fn))
        d1, d2 = vals[1] - vals[0], vals[2] - vals[1]
        p0 = (d2 - d1) / 2.0
        if p0 == 0.0:
            raise TermMismatch("not quadratic in the base height")
        return _snap(p0, fn), _snap((1.0 - d1 / p0) / 2.0, fn)
    if code == nv.REW_LEG_COLLISION:
        n_leg = int(rb.leg_indices.numel())
        cf = t["contact"].view(env.num_envs, -1, 3)

        def count_at(force):
            pr.zero()
            cf[0, :, 0] = force
            return pr.value(fn)

        p0 = count_at(1e4) / n_leg
        if p0 == 0.0:
            raise TermMismatch("term is identically zero")
        lo, hi = 0.0, 1e4                           # threshold by bisection, then snapped to the literal
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            if count_at(mid) != 0.0:
                hi = mid
            else:
                lo = mid
        return p0, _snap(0.5 * (lo + hi), fn)
    if code == nv.REW_DOF_POS_LIMITS:
        lim = dof_limits(env)
        if tuple(lim.shape) != (rb.num_dof, 2):
            raise TermMismatch("dof limits are not a (num_dof, 2) tensor")
        vals = []
        for over in (1.0, 2.0):                     # dof 0 that far above its upper limit; the rest unchanged
            pr.zero()
            t["dof"].view(env.num_envs, -1, 2)[0, 0, 0] = float(lim[0, 1]) + over
            vals.append(pr.value(fn))
        return (_snap(vals[1] - vals[0], fn), 0.0,
                {"dof_pos_limits": [(float(lo), float(hi)) for lo, hi in lim.tolist()]})
    if code == nv.REW_FEET_AIR_TIME:
        if "swing_time" not in t:
            raise TermMismatch("feet_air_time needs env.swing_time (N, feet) and env.last_contacts (N, feet) bool")
        feet = [int(b) for b in rb.ee_indices.tolist()]
        if tuple(t["swing_time"].shape) != (env.num_envs, len(feet)) or len(feet) > 4:
            raise TermMismatch("swing_time is not (num_envs, num_feet<=4)")
        cf = t["contact"].view(env.num_envs, -1, 3)

        def landing(air, force=100.0, cmd=1.0):     # foot 0 lands after `air` seconds in the air
            pr.zero()
            t["command"][0, 0] = cmd
            t["swing_time"][0, 0] = air
            cf[0, feet[0], 2] = force
            return pr.value(fn)

        r1, r2 = landing(1.0), landing(2.0)
        p0 = r2 - r1
        if p0 == 0.0:
            raise TermMismatch("term does not respond to a landing foot")
        dt = float(t["swing_time"][0, 1]) if len(feet) > 1 else float(pr.isg.dt)   # a foot in the air gained dt
        p0 = _snap(p0, fn)
        p1 = _snap(1.0 + dt - r1 / p0, fn)

        def bisect(lo, hi, responds):
            for _ in range(60):
                mid = 0.5 * (lo + hi)
                lo, hi = (lo, mid) if responds(mid) else (mid, hi)
            return 0.5 * (lo + hi)

        force_thr = _snap(bisect(0.0, 100.0, lambda f: landing(2.0, force=f) != 0.0), fn)
        cmd_thr = _snap(bisect(0.0, 1.0, lambda c: landing(2.0, cmd=c) != 0.0), fn)
        return p0, p1, {"feet_bodies": tuple(feet), "feet_contact_force": force_thr, "air_time_cmd_min": cmd_thr,
                        "air_time_dt": dt}
    raise TermMismatch(f"no fitting rule for term code {code}")


def compile_a1_terms(env, functions: Sequence[Callable], names: Optional[Dict[str, int]] = None) -> List[CompiledTerm]:
    """match + fit (steps 1-2).  Raises TermMismatch / KeyError with the offending term's name."""
    names = names or A1_NAMES
    if len(functions) == 0:
        raise TermMismatch("build_reward_functions() returned no term (env.py:161)")
    if len(functions) > nv.MAX_TERMS:
        raise TermMismatch(f"{len(functions)} terms: at most {nv.MAX_TERMS} fused reward terms are supported")
    out = []
    with A1Probe(env) as pr:
        for fn in functions:
            key = _strip(fn.__name__)
            if key not in names:
                raise KeyError(f"reward term {fn.__name__!r} has no fused implementation; known: {sorted(names)}")
            p0, p1, *extra = fit_a1_term(pr, fn, names[key])
            out.append(CompiledTerm(fn.__name__, names[key], float(p0), float(p1), extra[0] if extra else None))
    return out


def desc_extras(terms: Sequence[CompiledTerm]) -> dict:
    """The a1_desc keywords the compiled terms carry (feet, thresholds, dof limits)."""
    out = {}
    for term in terms:
        out.update(term.extra or {})
    return out


def verify_a1_terms(env, hot, functions: Sequence[Callable], terms: Sequence[CompiledTerm], seed: int = 1234):
    """Step 3: user hooks (torch) vs the kernel's term library on the same random state."""
    with A1Probe(env) as pr:
        pr.randomize(seed)
        stateful = [pr.tensors[k] for k in ("swing_time", "last_contacts") if k in pr.tensors]
        if stateful:
            pr.tensors["swing_time"].abs_()
        before = [t.clone() for t in stateful]
        want = torch.stack([fn().to(torch.float) for fn in functions])
        after_hooks = [t.clone() for t in stateful]
        for t, b in zip(stateful, before):          # a stateful term advanced its state: rewind for the kernel's turn
            t.copy_(b)
        got = hot.eval_terms()
        for t, a, name in zip(stateful, after_hooks, ("swing_time", "last_contacts")):
            if not torch.allclose(t.to(torch.float), a.to(torch.float), rtol=RTOL, atol=ATOL):
                raise TermMismatch(f"feet_air_time: {name} after the Python hook and after the fused term differ")
    for i, term in enumerate(terms):
        err = (got[i] - want[i]).abs()
        tol = ATOL + RTOL * want[i].abs()
        bad = int((err > tol).sum())
        if bad:
            raise TermMismatch(f"reward term {term.name!r}: the Python hook and the fused term (code {term.code}, "
                               f"p0={term.p0:g}, p1={term.p1:g}) differ on {bad} of {err.numel()} envs "
                               f"(max abs err {float(err.max()):.3e})")
