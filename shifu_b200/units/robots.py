"""Robot / ArmRobot / LeggedRobot — interface mirror of ``shifu/units/robot.py``.

State views (``dof_pos``, ``dof_vel``, ``contact_forces``, ``body_state``, ``ee_pose`` ...) are
views / gathers of the simulator's flat tensors exactly as in the reference (robot.py:48-53,
138-154, 195-215); the per-step arithmetic that the reference does here in torch —
``LeggedRobot.post_step`` (robot.py:222-229) — runs as a CUDA kernel (``shifu_body_frame``).
"""
from __future__ import annotations

import torch
from isaacgym import gymapi, gymtorch

from shifu_b200.configs import ActorConfig, ArmRobotActorConfig, LeggedRobotActorConfig
from .base import Actor


def _t(x, device, dtype=torch.float):
    return torch.as_tensor(x, dtype=dtype, device=device)


class Robot(Actor):
    cfg: ActorConfig

    def reset_idx(self, env_ids):
        self._reset_dof_state(env_ids)
        self._reset_root_state(env_ids)

    def step(self, actions):
        self._internal_motor_step(actions)

    def _init_props(self):
        super()._init_props()
        self.dof_props['driveMode'][:] = self.asset_options.default_dof_drive_mode
        self.dof_props['stiffness'] = self.cfg.dof_stiffness
        self.dof_props['damping'] = self.cfg.dof_damping
        self.dof_lower_limits = _t(self.dof_props['lower'], self.device)
        self.dof_upper_limits = _t(self.dof_props['upper'], self.device)
        self.dof_vel_limits = _t(self.dof_props['velocity'], self.device)
        self.torque_limits = _t(self.dof_props['effort'], self.device)

    def load_to(self, env_id, env_handle, seg_id):
        super().load_to(env_id, env_handle, seg_id)
        self.gym.set_actor_dof_properties(env_handle, self.actor_handle, self.dof_props)

    def init_buffers(self):
        super().init_buffers()
        n = self.env.num_envs
        self.default_dof_pos = _t(self.cfg.default_dof_pos, self.device)
        dof = self.env.dof_state.view(n, self.num_dof, 2)
        self.dof_pos, self.dof_vel = dof[..., 0], dof[..., 1]
        self.dof_targets = torch.zeros((n, self.num_dof), dtype=torch.float, device=self.device)

    def _internal_motor_step(self, action):
        mode = self.asset_options.default_dof_drive_mode
        if mode == gymapi.DOF_MODE_EFFORT:
            self.gym.set_dof_actuation_force_tensor(self.sim, gymtorch.unwrap_tensor(action))
        elif mode == gymapi.DOF_MODE_POS:
            self.gym.set_dof_position_target_tensor(self.sim, gymtorch.unwrap_tensor(action))
        elif mode == gymapi.DOF_MODE_VEL:
            self.gym.set_dof_velocity_target_tensor(self.sim, gymtorch.unwrap_tensor(action))
        else:
            raise NotImplementedError

    def apply_dof_targets(self, dof_targets):
        for _ in range(int(self.env.decimation)):
            self.gym.set_dof_position_target_tensor(self.sim, gymtorch.unwrap_tensor(dof_targets))
            self.gym.simulate(self.sim)
            if self.device == 'cpu':
                self.gym.fetch_results(self.sim, True)
            self.gym.refresh_dof_state_tensor(self.sim)

    def push_dof_reset(self, env_ids):
        """Tell the simulator about dof rows that were rewritten in place (robot.py:78-86)."""
        rows = self.root_indices[env_ids].to(torch.int32)
        self.gym.set_dof_position_target_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self.dof_targets),
                                                        gymtorch.unwrap_tensor(rows), len(rows))
        self.gym.set_dof_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self.env.dof_state),
                                              gymtorch.unwrap_tensor(rows), len(rows))

    def _reset_dof_state(self, env_ids):
        self.dof_targets[env_ids] = self.default_dof_pos.clone()
        self.dof_pos[env_ids] = self.default_dof_pos.clone()
        self.dof_vel[env_ids] = 0.
        self.push_dof_reset(env_ids)

    @property
    def base_pose(self):
        return self.env.root_state[self.root_indices, :7]

    def get_root_state(self):
        return self.env.root_state[self.root_indices]

    def set_root_state(self, root_state):
        self.env.root_state[self.root_indices] = root_state
        self.gym.set_actor_root_state_tensor(self.sim, gymtorch.unwrap_tensor(self.env.root_state))


class ArmRobot(Robot):
    cfg: ArmRobotActorConfig

    def __init__(self, cfg: ArmRobotActorConfig):
        super().__init__(cfg)
        self.end_effector_names = cfg.end_effector_names
        self.end_effector_velocity = cfg.end_effector_velocity

    def init_buffers(self):
        super().init_buffers()
        ee = [self.rigid_body_dict[n] for n in self.cfg.end_effector_names]
        self.ee_indices = _t(ee, self.device, torch.long)
        self.num_ee = len(ee)
        self.contact_forces = self.env.contact_state.view(self.env.num_envs, -1, 3)
        self.ee_pose_targets = torch.zeros((self.env.num_envs, 7), dtype=torch.float, device=self.device)
        jac = gymtorch.wrap_tensor(self.gym.acquire_jacobian_tensor(self.sim, self.name))
        self.gym.refresh_jacobian_tensors(self.sim)
        self._jacobian, self._ee_link = jac, ee[0] - 1
        self.j_ee = jac[:, ee[0] - 1]

    def load_to(self, env_id, env_handle, seg_id):
        super().load_to(env_id, env_handle, seg_id)
        self.set_segmentation_id(env_handle, seg_id)

    def apply_target_end_positions(self, tar_pose):
        self.dof_targets[:] = self.inverse_kinematics(tar_pose)
        self.apply_dof_targets(self.dof_targets)

    @property
    def body_state(self):
        n = self.env.num_envs
        return self.env.body_state.view(n, -1, 13)[:, :self.num_bodies].view(n, self.num_bodies, -1)

    @property
    def ee_pose(self):
        return self.body_state[:, self.ee_indices, :7]

    @property
    def ee_vel(self):
        return self.body_state[:, self.ee_indices, 7:]

    @property
    def ee_forces(self):
        return self.contact_forces[:, self.ee_indices]

    @staticmethod
    def orientation_error(desired, current):
        from isaacgym.torch_utils import quat_conjugate, quat_mul
        q_r = quat_mul(desired, quat_conjugate(current))
        return q_r[:, 0:3] * torch.sign(q_r[:, 3]).unsqueeze(-1)

    def _ik_layout(self):
        n = self.env.num_envs
        bodies = self.env.body_state.shape[0] // n
        if self.env.dof_state.shape[0] // n != self.num_dof:
            raise NotImplementedError("shifu_arm_ik expects the arm to own every dof of its env")
        jac = self._jacobian if self._jacobian.is_contiguous() else self._jacobian.contiguous()
        return dict(body_state=self.env.body_state, num_bodies=bodies, ee_body=int(self.ee_indices[0]),
                    jacobian=jac, ee_link=self._ee_link, dof_state=self.env.dof_state, num_dof=self.num_dof)

    def inverse_kinematics(self, goal_pose, damping=0.05, out=None):
        """Damped least squares (robot.py:156-182) — SURVEY.md §8f row N2, one ``shifu_arm_ik`` launch."""
        out = torch.empty(self.env.num_envs, self.num_dof, device=self.device) if out is None else out
        return self.env.kernels().arm_ik(goal_pose=goal_pose, damping=damping, dof_targets=out, **self._ik_layout())


class LeggedRobot(ArmRobot):
    cfg: LeggedRobotActorConfig

    def __init__(self, cfg: LeggedRobotActorConfig):
        Robot.__init__(self, cfg)
        self.end_effector_names = cfg.end_effector_names
        self.ee_indices = []

    def init_buffers(self):
        Robot.init_buffers(self)
        n, dev = self.env.num_envs, self.device
        ee = [self.rigid_body_dict[name] for name in self.cfg.end_effector_names]
        self.ee_indices = _t(ee, dev, torch.long)
        self.num_ee = len(ee)
        self.contact_forces = self.env.contact_state.view(n, -1, 3)
        jac = gymtorch.wrap_tensor(self.gym.acquire_jacobian_tensor(self.sim, self.name))
        self.gym.refresh_jacobian_tensors(self.sim)
        self.j_ee = jac[:, ee]
        # body-frame state (robot.py:210-215): persistent tensors updated in place by the kernel
        self.gravity_vec = torch.tensor([0., 0., -1.], device=dev).repeat(n, 1)
        self.base_lin_vel = torch.zeros(n, 3, device=dev)
        self.base_ang_vel = torch.zeros(n, 3, device=dev)
        self.projected_gravity = torch.zeros(n, 3, device=dev)
        self.post_step()

    def step(self, actions):
        self.dof_targets[:] = self.dof_pos[:, :self.num_dof] + actions
        self.apply_dof_targets(self.dof_targets)
        self.post_step()

    def post_step(self):
        """robot.py:222-229 as one kernel over the CURRENT root tensor (S_prev)."""
        layout = self.affine_root_layout()
        if layout is None:
            raise NotImplementedError("non-affine root_indices are not supported by shifu_body_frame")
        self.env.kernels().body_frame(self.env.root_state, self.env.num_envs, layout[0], layout[1],
                                      self.base_lin_vel, self.base_ang_vel, self.projected_gravity,
                                      self.gravity_vec)

    def apply_force_on_base(self, force_tensor, pos_tensor=None):
        self.gym.apply_rigid_body_force_at_pos_tensors(
            self.sim, gymtorch.unwrap_tensor(force_tensor),
            gymtorch.unwrap_tensor(pos_tensor) if pos_tensor is not None else None)
