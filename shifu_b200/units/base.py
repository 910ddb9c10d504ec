"""Unit / Actor / Sensor — interface mirror of ``shifu/units/units.py``.

A *unit* is anything loaded into every env (robot, object, sensor).  The lifecycle is the
reference's (``Unit.set_env -> _init_props -> load_to (per env) -> init_buffers -> reset_idx``,
units.py:20-38); an *actor* additionally owns a row per env in the simulator's flat root-state
tensor, addressed through ``root_indices``.

Hot-path relevance: ``root_indices`` / ``base_pose`` /
``_reset_root_state``.  The fused kernels address root rows as ``root_offset + env*root_stride``
(``affine_root_layout``); the gathers below exist for user-written hooks.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from isaacgym import gymapi

from shifu_b200.configs import ActorConfig, BaseConfig, BaseSensorConfig


def _abstract(name: str):
    def method(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__}.{name}")
    method.__name__ = name
    return method


class Unit:
    """Lifecycle protocol; every stage is a hook a concrete unit overrides."""
    cfg: BaseConfig

    def __init__(self, cfg: BaseConfig):
        self.cfg, self.name = cfg, cfg.name

    def set_env(self, env):
        self.env = env
        self.gym, self.sim, self.device = env.gym, env.sim, env.device
        self._init_props()

    _init_props = _abstract("_init_props")
    load_to = _abstract("load_to")              # (env_id, env_handle, seg_id)
    init_buffers = _abstract("init_buffers")
    reset_idx = _abstract("reset_idx")          # (env_ids)


class Actor(Unit):
    cfg: ActorConfig

    def __init__(self, cfg: ActorConfig):
        super().__init__(cfg)
        self.asset_options = cfg.asset_options
        self.rigid_body_dict = {}
        self.root_indices = []                  # python list while envs are built, tensor afterwards

    # ---------------------------------------------------------------- asset + per-env loading
    # (simulator API, units.py:55-116; not on the hot path)
    def create_asset(self):
        cfg = self.cfg
        self.asset = self.gym.load_asset(self.sim, cfg.root_dir, cfg.urdf_filename, self.asset_options)

    def _init_props(self):
        pose = gymapi.Transform()
        pose.p, pose.r = gymapi.Vec3(*self.cfg.default_pos), gymapi.Quat(*self.cfg.default_quat)
        self._init_root_pose = pose
        self.create_asset()
        gym, asset = self.gym, self.asset
        self.num_bodies, self.num_dof = gym.get_asset_rigid_body_count(asset), gym.get_asset_dof_count(asset)
        self.default_rigid_shape_props = gym.get_asset_rigid_shape_properties(asset)
        self.dof_props = gym.get_asset_dof_properties(asset)

    def random_rigid_shape_props(self, env_ids, rigid_shape_props):
        """Optional domain-randomisation hook (units.py:118-119)."""
        raise NotImplementedError

    def load_to(self, env_id, env_handle, seg_id):
        self._init_root_pose.p += gymapi.Vec3(*self.env.env_origins[env_id].clone())
        try:
            randomised = self.random_rigid_shape_props(env_id, self.default_rigid_shape_props)
        except NotImplementedError:
            randomised = None
        if randomised is not None:
            self.gym.set_asset_rigid_shape_properties(self.asset, randomised)
        self.actor_handle = self.gym.create_actor(env_handle, self.asset, self._init_root_pose, self.name, env_id, 0)
        self.root_indices.append(self.gym.get_actor_index(env_handle, self.actor_handle, gymapi.DOMAIN_SIM))
        self.set_segmentation_id(env_handle, seg_id)

    def set_segmentation_id(self, env_handle, seg_id):
        self.segmentation_id = seg_id
        self.rigid_body_dict = self.gym.get_actor_rigid_body_dict(env_handle, self.actor_handle)
        for body in self.rigid_body_dict.values():
            self.gym.set_rigid_body_segmentation_id(env_handle, self.actor_handle, body, seg_id)

    def set_asset_rigid_properties(self, env_handle, mass=None, friction=None):
        gym, actor = self.gym, self.actor_handle
        if friction is not None:
            shapes = gym.get_actor_rigid_shape_properties(env_handle, actor)
            for shape in shapes:
                shape.friction = friction
            gym.set_actor_rigid_shape_properties(env_handle, actor, shapes)
        if mass is not None:
            bodies = gym.get_actor_rigid_body_properties(env_handle, actor)
            for body in bodies:
                body.mass = mass
            gym.set_actor_rigid_body_properties(env_handle, actor, bodies, recomputeInertia=True)

    # ---------------------------------------------------------------- state
    def init_buffers(self):
        self.root_indices = torch.as_tensor(self.root_indices, dtype=torch.long, device=self.device)
        self.rigid_body_dict = self.gym.get_asset_rigid_body_dict(self.asset)
        pose7 = [*self.cfg.default_pos, *self.cfg.default_quat]
        self.default_base_pose = torch.tensor(pose7, dtype=torch.float, device=self.device)

    def affine_root_layout(self) -> Optional[Tuple[int, int]]:
        """(stride, offset) when root_indices[e] == offset + e*stride (always true for envs built
        by ``IsaacGymEnv.create_envs``: every env loads the same actor list), else None."""
        idx = self.root_indices
        if idx.numel() == 0:
            return None
        offset = int(idx[0])
        stride = int(idx[1]) - offset if idx.numel() > 1 else 1
        if stride < 1:
            return None
        expected = offset + stride * torch.arange(idx.numel(), device=idx.device)
        return (stride, offset) if torch.equal(idx, expected) else None

    @property
    def base_pose(self):
        return self.env.root_state[self.root_indices, :7]

    def _reset_root_state(self, env_ids):                      # units.py:127-134
        state, rows = self.env.root_state, self.root_indices[env_ids]
        position, orientation = self.default_base_pose[:3], self.default_base_pose[3:7]
        state[rows, 0:3] = position + self.env.env_origins[env_ids]
        state[rows, 3:7] = orientation
        state[rows, 7:13] = 0.

    def reset_idx(self, env_ids):
        self._reset_root_state(env_ids)


class Sensor(Unit):
    cfg: BaseSensorConfig

    refresh = _abstract("refresh")
