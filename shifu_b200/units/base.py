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

import torch
from isaacgym import gymapi

from shifu_b200.configs import ActorConfig, BaseConfig, BaseSensorConfig


class Unit:
    cfg: BaseConfig

    def __init__(self, cfg: BaseConfig):
        self.cfg = cfg
        self.name = cfg.name

    def set_env(self, env):
        self.env, self.gym, self.sim, self.device = env, env.gym, env.sim, env.device
        self._init_props()

    def _init_props(self):
        raise NotImplementedError

    def reset_idx(self, env_ids):
        raise NotImplementedError

    def load_to(self, env_id, env_handle, seg_id):
        raise NotImplementedError

    def init_buffers(self):
        raise NotImplementedError


class Actor(Unit):
    cfg: ActorConfig

    def __init__(self, cfg: ActorConfig):
        super().__init__(cfg)
        self.asset_options = cfg.asset_options
        self.root_indices = []
        self.rigid_body_dict = {}

    # -- construction (simulator API; not on the hot path) ---------------------------------
    def create_asset(self):
        self.asset = self.gym.load_asset(self.sim, self.cfg.root_dir, self.cfg.urdf_filename, self.asset_options)

    def _init_props(self):
        self._init_root_pose = gymapi.Transform()
        self._init_root_pose.p = gymapi.Vec3(*self.cfg.default_pos)
        self._init_root_pose.r = gymapi.Quat(*self.cfg.default_quat)
        self.create_asset()
        self.num_bodies = self.gym.get_asset_rigid_body_count(self.asset)
        self.default_rigid_shape_props = self.gym.get_asset_rigid_shape_properties(self.asset)
        self.num_dof = self.gym.get_asset_dof_count(self.asset)
        self.dof_props = self.gym.get_asset_dof_properties(self.asset)

    def load_to(self, env_id, env_handle, seg_id):
        origin = self.env.env_origins[env_id].clone()
        self._init_root_pose.p += gymapi.Vec3(*origin)
        try:
            props = self.random_rigid_shape_props(env_id, self.default_rigid_shape_props)
            self.gym.set_asset_rigid_shape_properties(self.asset, props)
        except NotImplementedError:
            pass
        self.actor_handle = self.gym.create_actor(env_handle, self.asset, self._init_root_pose, self.name, env_id, 0)
        self.root_indices.append(self.gym.get_actor_index(env_handle, self.actor_handle, gymapi.DOMAIN_SIM))
        self.set_segmentation_id(env_handle, seg_id)

    def set_segmentation_id(self, env_handle, seg_id):
        self.segmentation_id = seg_id
        self.rigid_body_dict = self.gym.get_actor_rigid_body_dict(env_handle, self.actor_handle)
        for rigid_id in self.rigid_body_dict.values():
            self.gym.set_rigid_body_segmentation_id(env_handle, self.actor_handle, rigid_id, seg_id)

    def set_asset_rigid_properties(self, env_handle, mass=None, friction=None):
        if friction is not None:
            shape_props = self.gym.get_actor_rigid_shape_properties(env_handle, self.actor_handle)
            for sp in shape_props:
                sp.friction = friction
            self.gym.set_actor_rigid_shape_properties(env_handle, self.actor_handle, shape_props)
        if mass is not None:
            body_props = self.gym.get_actor_rigid_body_properties(env_handle, self.actor_handle)
            for bp in body_props:
                bp.mass = mass
            self.gym.set_actor_rigid_body_properties(env_handle, self.actor_handle, body_props, recomputeInertia=True)

    def random_rigid_shape_props(self, env_ids, rigid_shape_props):
        raise NotImplementedError

    # -- state views --------------------------------------------------------------------------
    def init_buffers(self):
        self.root_indices = torch.as_tensor(self.root_indices, dtype=torch.long, device=self.device)
        self.rigid_body_dict = self.gym.get_asset_rigid_body_dict(self.asset)
        self.default_base_pose = torch.tensor(list(self.cfg.default_pos) + list(self.cfg.default_quat),
                                              dtype=torch.float, device=self.device)

    def affine_root_layout(self):
        """(stride, offset) when root_indices[e] == offset + e*stride (always true for envs built
        by ``IsaacGymEnv.create_envs``: every env loads the same actor list), else None."""
        idx = self.root_indices
        if idx.numel() == 0:
            return None
        off = int(idx[0])
        stride = int(idx[1] - idx[0]) if idx.numel() > 1 else 1
        ok = bool(torch.equal(idx, off + stride * torch.arange(idx.numel(), device=idx.device)))
        return (stride, off) if ok and stride >= 1 else None

    def reset_idx(self, env_ids):
        self._reset_root_state(env_ids)

    def _reset_root_state(self, env_ids):
        rows = self.root_indices[env_ids]
        self.env.root_state[rows, :3] = self.default_base_pose[:3] + self.env.env_origins[env_ids]
        self.env.root_state[rows, 3:7] = self.default_base_pose[3:7]
        self.env.root_state[rows, 7:] = 0.

    @property
    def base_pose(self):
        return self.env.root_state[self.root_indices, :7]


class Sensor(Unit):
    cfg: BaseSensorConfig

    def refresh(self):
        raise NotImplementedError
