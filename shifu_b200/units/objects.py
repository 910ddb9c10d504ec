"""Object / Box — interface mirror of ``shifu/units/object.py`` (primitive actors; on the hot
path only as providers of ``root_indices`` / ``base_pose`` for the ABB scene)."""
from __future__ import annotations

from isaacgym import gymapi

from shifu_b200.configs import ActorConfig, BoxActorConfig
from .base import Actor


class Object(Actor):
    cfg: ActorConfig


class Box(Object):
    cfg: BoxActorConfig

    def create_asset(self):
        w, h, d = self.cfg.box_dim
        self.asset = self.gym.create_box(self.sim, w, h, d, self.asset_options)

    def load_to(self, env_id, env_handle, seg_id):
        super().load_to(env_id, env_handle, seg_id)
        self.set_asset_rigid_properties(env_handle, mass=self.cfg.mass, friction=self.cfg.friction)
        self.gym.set_rigid_body_color(env_handle, self.actor_handle, 0, gymapi.MESH_VISUAL_AND_COLLISION,
                                      gymapi.Vec3(*self.cfg.color))
