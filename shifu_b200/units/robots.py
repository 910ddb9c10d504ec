"""Robot / ArmRobot / LeggedRobot — interface mirror of ``shifu/units/robot.py``.

State views (``dof_pos``, ``dof_vel``, ``contact_forces``, ``body_state``, ``ee_pose`` ...) are
views / gathers of the simulator's flat tensors exactly as in the reference (robot.py:48-53,
138-154, 195-215); the per-step arithmetic the reference does here in torch runs as CUDA kernels:
``LeggedRobot.post_step`` (robot.py:222-229) -> ``shifu_body_frame``,
``ArmRobot.inverse_kinematics`` (robot.py:156-182) -> ``shifu_arm_ik``.
"""
from __future__ import annotations

import torch
from isaacgym import gymapi, gymtorch

from shifu_b200.configs import ActorConfig, ArmRobotActorConfig, LeggedRobotActorConfig
from .base import Actor


def _tensor(values, device, dtype=torch.float):
    return torch.as_tensor(values, dtype=dtype, device=device)


class Robot(Actor):
    cfg: ActorConfig

    # dof property column -> attribute published as a device tensor (robot.py:35-45)
    _LIMIT_COLUMNS = (("lower", "dof_lower_limits"), ("upper", "dof_upper_limits"),
                      ("velocity", "dof_vel_limits"), ("effort", "torque_limits"))

    # ---------------------------------------------------------------- construction
    def _init_props(self):
        super()._init_props()
        props = self.dof_props
        props['driveMode'][:] = self.asset_options.default_dof_drive_mode
        props['stiffness'], props['damping'] = self.cfg.dof_stiffness, self.cfg.dof_damping
        for column, attr in self._LIMIT_COLUMNS:
            setattr(self, attr, _tensor(props[column], self.device))

    def load_to(self, env_id, env_handle, seg_id):
        super().load_to(env_id, env_handle, seg_id)
        self.gym.set_actor_dof_properties(env_handle, self.actor_handle, self.dof_props)

    def init_buffers(self):
        super().init_buffers()
        envs = self.env.num_envs
        # (pos, vel) are strided views of the simulator tensor: writes go straight through
        pos_vel = self.env.dof_state.view(envs, self.num_dof, 2)
        self.dof_pos, self.dof_vel = pos_vel.select(-1, 0), pos_vel.select(-1, 1)
        self.default_dof_pos = _tensor(self.cfg.default_dof_pos, self.device)
        self.dof_targets = torch.zeros(envs, self.num_dof, dtype=torch.float, device=self.device)

    # ---------------------------------------------------------------- actuation
    def step(self, actions):
        self._internal_motor_step(actions)

    def _internal_motor_step(self, action):
        """Dispatch on the asset's drive mode (robot.py:55-64)."""
        setters = {gymapi.DOF_MODE_EFFORT: self.gym.set_dof_actuation_force_tensor,
                   gymapi.DOF_MODE_POS: self.gym.set_dof_position_target_tensor,
                   gymapi.DOF_MODE_VEL: self.gym.set_dof_velocity_target_tensor}
        try:
            setter = setters[self.asset_options.default_dof_drive_mode]
        except KeyError:
            raise NotImplementedError from None
        setter(self.sim, gymtorch.unwrap_tensor(action))

    def apply_dof_targets(self, dof_targets):
        """Position targets held for ``decimation`` simulator substeps (robot.py:66-72)."""
        wait_for_cpu_sim = self.device == 'cpu'
        for _ in range(int(self.env.decimation)):
            self.gym.set_dof_position_target_tensor(self.sim, gymtorch.unwrap_tensor(dof_targets))
            self.gym.simulate(self.sim)
            if wait_for_cpu_sim:
                self.gym.fetch_results(self.sim, True)
            self.gym.refresh_dof_state_tensor(self.sim)

    # ---------------------------------------------------------------- reset
    def reset_idx(self, env_ids):
        self._reset_dof_state(env_ids)
        self._reset_root_state(env_ids)

    def _reset_dof_state(self, env_ids):
        rest = self.default_dof_pos
        self.dof_targets[env_ids] = rest.clone()
        self.dof_pos[env_ids] = rest.clone()
        self.dof_vel[env_ids] = 0.
        self.push_dof_reset(env_ids)

    def push_dof_reset(self, env_ids):
        """Tell the simulator about dof rows that were rewritten in place (robot.py:78-86)."""
        rows = gymtorch.unwrap_tensor(self.root_indices[env_ids].to(torch.int32))
        count = len(env_ids)
        self.gym.set_dof_position_target_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self.dof_targets), rows, count)
        self.gym.set_dof_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self.env.dof_state), rows, count)

    # ---------------------------------------------------------------- root state
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

    def _bind_end_effectors(self):
        """End-effector body ids, contact view and the jacobian block of the first one (robot.py:117-128)."""
        bodies = [self.rigid_body_dict[name] for name in self.cfg.end_effector_names]
        self.ee_indices, self.num_ee = _tensor(bodies, self.device, torch.long), len(bodies)
        self.contact_forces = self.env.contact_state.view(self.env.num_envs, -1, 3)
        jacobian = gymtorch.wrap_tensor(self.gym.acquire_jacobian_tensor(self.sim, self.name))
        self.gym.refresh_jacobian_tensors(self.sim)
        return bodies, jacobian

    def init_buffers(self):
        super().init_buffers()
        bodies, jacobian = self._bind_end_effectors()
        self.ee_pose_targets = torch.zeros(self.env.num_envs, 7, dtype=torch.float, device=self.device)
        self._jacobian, self._ee_link = jacobian, bodies[0] - 1     # fixed base: link = body - 1
        self.j_ee = jacobian[:, self._ee_link]

    def load_to(self, env_id, env_handle, seg_id):
        super().load_to(env_id, env_handle, seg_id)
        self.set_segmentation_id(env_handle, seg_id)

    # ---------------------------------------------------------------- views (robot.py:138-154)
    @property
    def body_state(self):
        envs = self.env.num_envs
        own = self.env.body_state.view(envs, -1, 13)[:, :self.num_bodies]
        return own.view(envs, self.num_bodies, -1)

    @property
    def ee_pose(self):
        return self.body_state[:, self.ee_indices, :7]

    @property
    def ee_vel(self):
        return self.body_state[:, self.ee_indices, 7:]

    @property
    def ee_forces(self):
        return self.contact_forces[:, self.ee_indices]

    # ---------------------------------------------------------------- inverse kinematics (row N2)
    @staticmethod
    def orientation_error(desired, current):
        from isaacgym.torch_utils import quat_conjugate, quat_mul
        rel = quat_mul(desired, quat_conjugate(current))
        return rel[:, :3] * torch.sign(rel[:, 3]).unsqueeze(-1)

    def _ik_layout(self):
        envs = self.env.num_envs
        if self.env.dof_state.shape[0] // envs != self.num_dof:
            raise NotImplementedError("shifu_arm_ik expects the arm to own every dof of its env")
        jacobian = self._jacobian if self._jacobian.is_contiguous() else self._jacobian.contiguous()
        return dict(body_state=self.env.body_state, num_bodies=self.env.body_state.shape[0] // envs,
                    ee_body=int(self.ee_indices[0]), jacobian=jacobian, ee_link=self._ee_link,
                    dof_state=self.env.dof_state, num_dof=self.num_dof)

    def inverse_kinematics(self, goal_pose, damping=0.05, out=None):
        """Damped least squares (robot.py:156-182) — SURVEY.md §8f row N2, one ``shifu_arm_ik`` launch."""
        if out is None:
            out = torch.empty(self.env.num_envs, self.num_dof, device=self.device)
        return self.env.kernels().arm_ik(goal_pose=goal_pose, damping=damping, dof_targets=out, **self._ik_layout())

    def apply_target_end_positions(self, tar_pose):
        self.inverse_kinematics(tar_pose, out=self.dof_targets)
        self.apply_dof_targets(self.dof_targets)


class LeggedRobot(ArmRobot):
    cfg: LeggedRobotActorConfig

    def __init__(self, cfg: LeggedRobotActorConfig):
        Robot.__init__(self, cfg)
        self.end_effector_names = cfg.end_effector_names
        self.ee_indices = []

    def init_buffers(self):
        Robot.init_buffers(self)
        bodies, jacobian = self._bind_end_effectors()
        self.j_ee = jacobian[:, bodies]
        # body-frame state (robot.py:210-215): persistent tensors updated in place by the kernel
        envs, dev = self.env.num_envs, self.device
        self.gravity_vec = torch.tensor([0., 0., -1.], device=dev).repeat(envs, 1)
        self.base_lin_vel, self.base_ang_vel, self.projected_gravity = (torch.zeros(envs, 3, device=dev)
                                                                        for _ in range(3))
        self.post_step()

    def step(self, actions):
        torch.add(self.dof_pos[:, :self.num_dof], actions, out=self.dof_targets)
        self.apply_dof_targets(self.dof_targets)
        self.post_step()

    def post_step(self):
        """robot.py:222-229 as one kernel over the CURRENT root tensor (S_prev)."""
        layout = self.affine_root_layout()
        if layout is None:
            raise NotImplementedError("non-affine root_indices are not supported by shifu_body_frame")
        stride, offset = layout
        self.env.kernels().body_frame(self.env.root_state, self.env.num_envs, stride, offset, self.base_lin_vel,
                                      self.base_ang_vel, self.projected_gravity, self.gravity_vec)

    def apply_force_on_base(self, force_tensor, pos_tensor=None):
        positions = None if pos_tensor is None else gymtorch.unwrap_tensor(pos_tensor)
        self.gym.apply_rigid_body_force_at_pos_tensors(self.sim, gymtorch.unwrap_tensor(force_tensor), positions)
