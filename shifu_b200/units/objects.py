"""Object / Box — interface mirror of ``shifu/units/object.py`` (primitive actors; on the hot
path only as providers of ``root_indices`` / ``base_pose`` for the ABB scene)."""
from __future__ import annotations

from isaacgym import gymapi

from shifu_b200.configs import ActorConfig, BoxActorConfig
from .base import Actor


class Object(Actor):
    """Static scene object."""
    cfg: ActorConfig


class Box(Object):
    """Procedural box: its asset comes from ``gym.create_box`` instead of a URDF."""
    cfg: BoxActorConfig

    def create_asset(self):
        width, height, depth = self.cfg.box_dim
        self.asset = self.gym.create_box(self.sim, width, height, depth, self.asset_options)

    def load_to(self, env_id, env_handle, seg_id):
        super().load_to(env_id, env_handle, seg_id)
        cfg = self.cfg
        self.set_asset_rigid_properties(env_handle, friction=cfg.friction, mass=cfg.mass)
        tint, whole_body = gymapi.Vec3(*cfg.color), 0
        self.gym.set_rigid_body_color(env_handle, self.actor_handle, whole_body, gymapi.MESH_VISUAL_AND_COLLISION, tint)
