"""Seeded synthetic simulator state for the A1 and ABB workloads.

Stands in for what PhysX would write into the flat gym tensors between control
steps.  Two uses:

* **replay** (parity): every snapshot of step ``t`` is generated on the CPU from
  ``torch.Generator().manual_seed(seed*1000003 + t)`` and copied to the sim's
  device, so the unmodified reference (torch CPU) and the CUDA path are fed
  bit-identical inputs;
* **static** (benchmark): one snapshot is generated on the device once and the
  refresh hooks do nothing — the state tensors stay resident in HBM, which is
  what the hot path sees in production between two PhysX steps.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch


def _quat_from_rpy(roll, pitch, yaw):
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    return torch.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                        sy * cr * cp - cy * sr * sp, cy * cr * cp + sy * sr * sp], dim=-1)


@dataclass
class A1Snapshot:
    """All simulator outputs the A1 step consumes during ONE control step."""
    dof: torch.Tensor            # (5, N, 12, 2): after each of the 4 substeps + the refresh_state step
    root_offset: torch.Tensor    # (N, 13): root row with xyz RELATIVE to env_origins (z: + origin z)
    contact: torch.Tensor        # (N, 17, 3)
    actions: torch.Tensor        # (N, 12) raw policy output in [-1,1]


def a1_snapshot(seed: int, step: int, n: int, *, gen_device="cpu", p_base=0.01, p_leg=0.1,
                xy_range=5.0, n_bodies=17, n_dof=12, offmap=True) -> A1Snapshot:
    g = torch.Generator(device=gen_device).manual_seed(seed * 1000003 + step)
    kw = dict(generator=g, device=gen_device)
    q0 = torch.tensor([0.1, 0.8, -1.5, 0.1, 0.8, -1.5, -0.1, 0.8, -1.5, -0.1, 0.8, -1.5], device=gen_device)
    dof = torch.empty(5, n, n_dof, 2, device=gen_device)
    dof[..., 0] = q0 + 0.3 * torch.randn(5, n, n_dof, **kw)
    dof[..., 1] = 2.0 * torch.randn(5, n, n_dof, **kw)
    root = torch.zeros(n, 13, device=gen_device)
    root[:, 0:2] = (torch.rand(n, 2, **kw) * 2 - 1) * xy_range
    root[:, 2] = 0.30 + 0.25 * torch.rand(n, **kw)
    yaw = (torch.rand(n, **kw) * 2 - 1) * math.pi
    roll = 0.15 * torch.randn(n, **kw)
    pitch = 0.15 * torch.randn(n, **kw)
    quat = _quat_from_rpy(roll, pitch, yaw)
    # un-normalise a little: get_heights re-normalises the yaw quaternion itself
    quat = quat * (1.0 + 0.01 * torch.randn(n, 1, **kw))
    root[:, 3:7] = quat
    root[:, 7:10] = 0.5 * torch.randn(n, 3, **kw)
    root[:, 10:13] = 1.0 * torch.randn(n, 3, **kw)
    if offmap and n >= 8:
        # push a few robots far outside the height map to exercise the index clamps
        root[1, 0:2] = torch.tensor([-400.0, -400.0], device=gen_device)
        root[2, 0:2] = torch.tensor([900.0, 900.0], device=gen_device)
        root[3, 0:2] = torch.tensor([-400.0, 900.0], device=gen_device)
    p = torch.full((n, n_bodies), p_leg, device=gen_device)
    p[:, 0] = p_base
    hit = (torch.rand(n, n_bodies, **kw) < p).float() * 20.0 * torch.rand(n, n_bodies, **kw)
    contact = torch.randn(n, n_bodies, 3, **kw) * hit.unsqueeze(-1)
    actions = torch.rand(n, n_dof, **kw) * 2.4 - 1.2      # a few outside [-1,1]·2 to exercise the clip
    return A1Snapshot(dof=dof, root_offset=root, contact=contact, actions=actions)


class A1Replay:
    """Snapshot provider for :class:`fake_isaacgym.FakeSim` reproducing the refresh order of
    ``A1Robot.step`` (4x dof) + ``IsaacGymEnv.refresh_state`` (root, body, dof, contact):
    ``examples/a1_conditional/a1_conditional.py:64-75``, ``shifu/gym/isaac_gym.py:139-154``."""

    def __init__(self, seed: int, n: int, get_env_origins, snap_hook=None, **snap_kw):
        self.seed, self.n = seed, n
        self.get_env_origins = get_env_origins
        self.snap_hook = snap_hook          # optional callable(step, snap) editing the snapshot in place
        self.snap_kw = snap_kw
        self.step = 0
        self.snap: Optional[A1Snapshot] = None
        self._dof_i = 0
        self.enabled = False

    def begin_step(self, step: int) -> torch.Tensor:
        """Generate the step's snapshots; returns the raw actions for the step."""
        self.step = step
        self.snap = a1_snapshot(self.seed, step, self.n, **self.snap_kw)
        if self.snap_hook is not None:
            self.snap_hook(step, self.snap)
        self._dof_i = 0
        self.enabled = True
        return self.snap.actions

    def __call__(self, kind: str, sim) -> None:
        if not self.enabled or self.snap is None:
            return
        s = self.snap
        if kind == "dof":
            i = min(self._dof_i, 4)
            sim.dof_state.view(self.n, -1, 2).copy_(s.dof[i].to(sim.dof_state.device))
            self._dof_i += 1
        elif kind == "root":
            root = s.root_offset.to(sim.root_state.device).clone()
            root[:, 0:3] += self.get_env_origins()
            sim.root_state.copy_(root)
        elif kind == "contact":
            sim.contact_state.view(self.n, -1, 3).copy_(s.contact.to(sim.contact_state.device))


# ---------------------------------------------------------------------------
# ABB push-box prior stage: 4 actors / env (robot, table, cube, goal),
# 10 bodies / env (7 robot incl. tip0 at index 6, + 3 boxes), 6 dofs.
# ---------------------------------------------------------------------------


@dataclass
class AbbSnapshot:
    root: torch.Tensor      # (N, 4, 13)
    body: torch.Tensor      # (N, 10, 13)
    dof: torch.Tensor       # (N, 6, 2)
    actions: torch.Tensor   # (N, 3)


def abb_snapshot(seed: int, step: int, n: int, *, gen_device="cpu", p_success=0.05) -> AbbSnapshot:
    g = torch.Generator(device=gen_device).manual_seed(seed * 1000003 + step + 7919)
    kw = dict(generator=g, device=gen_device)
    root = torch.zeros(n, 4, 13, device=gen_device)
    root[..., 6] = 1.0
    root[:, 0, 0] = -0.48
    root[:, 1, 2] = 0.05
    root[:, 2, 0:2] = (torch.rand(n, 2, **kw) * 2 - 1) * 0.21      # cube xy, sometimes out of bounds
    root[:, 2, 2] = 0.125
    root[:, 3, 0:2] = (torch.rand(n, 2, **kw) * 2 - 1) * 0.12      # goal xy
    root[:, 3, 2] = 0.1
    # a fraction of cubes sits (almost) on its goal
    near = torch.rand(n, **kw) < p_success
    root[near, 2, 0:2] = root[near, 3, 0:2] + (torch.rand(int(near.sum()), 2, **kw) * 2 - 1) * 0.02
    body = torch.zeros(n, 10, 13, device=gen_device)
    body[..., 6] = 1.0
    body[:, :, 0:3] = torch.randn(n, 10, 3, **kw) * 0.3
    # the end effector hovers around the cube
    body[:, 6, 0:2] = root[:, 2, 0:2] + (torch.rand(n, 2, **kw) * 2 - 1) * 0.15
    body[:, 6, 2] = 0.125
    dof = torch.randn(n, 6, 2, **kw) * 0.2
    actions = torch.rand(n, 3, **kw) * 2.4 - 1.2
    return AbbSnapshot(root=root, body=body, dof=dof, actions=actions)


class AbbReplay:
    """Provider for the ABB step: ``ArmRobot.apply_dof_targets`` refreshes dof x decimation, then
    ``refresh_state`` refreshes root/body/dof/contact (``shifu/units/robot.py:66-72``)."""

    def __init__(self, seed: int, n: int, **snap_kw):
        self.seed, self.n, self.snap_kw = seed, n, snap_kw
        self.snap: Optional[AbbSnapshot] = None
        self.enabled = False

    def begin_step(self, step: int) -> torch.Tensor:
        self.snap = abb_snapshot(self.seed, step, self.n, **self.snap_kw)
        self.enabled = True
        return self.snap.actions

    def __call__(self, kind: str, sim) -> None:
        if not self.enabled or self.snap is None:
            return
        s = self.snap
        if kind == "root":
            # keep reset writes of cube/goal made earlier in this step? No: refresh happens before them.
            sim.root_state.view(self.n, 4, 13).copy_(s.root.to(sim.root_state.device))
        elif kind == "body":
            sim.body_state.view(self.n, 10, 13).copy_(s.body.to(sim.body_state.device))
        elif kind == "dof":
            sim.dof_state.view(self.n, 6, 2).copy_(s.dof.to(sim.dof_state.device))
