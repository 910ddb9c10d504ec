"""IsaacGymEnv / TerrainGymEnv — the simulator façade (interface mirror of
``shifu/gym/isaac_gym.py``).

Owns the flat gym state tensors (``root_state``, ``dof_state``, ``body_state``,
``contact_state`` — borrowed from the simulator through ``gymtorch.wrap_tensor``), the env
origins / terrain bookkeeping and the step order ``robot.step -> refresh_state ->
post_physics_step`` (isaac_gym.py:45-49).  Viewer / rendering / lights are graphics and out of
scope (headless only).

Per-step arithmetic the reference does here in torch — ``TerrainGymEnv.get_heights``
(isaac_gym.py:393-433, 73 % of its CPU step) — runs in CUDA: stand-alone through
``shifu_get_heights`` for user-hook tasks, or folded into the fused A1 kernel.
"""
from __future__ import annotations

from typing import List, Union

import numpy as np
import torch
from isaacgym import gymapi, gymtorch, gymutil

from shifu_b200.configs import BaseEnvConfig, TerrainEnvConfig
from shifu_b200.units import Actor, Object, Robot, Sensor, Unit
from shifu_b200.utils.heightmap import Terrain


class IsaacGymEnv:
    robot: Robot
    objects: List[Object]
    sensors: List[Sensor]

    def __init__(self, cfg: BaseEnvConfig):
        self.cfg = cfg
        self.dt = cfg.sim.dt * cfg.control.decimation
        self.decimation = cfg.control.decimation
        self.num_envs = cfg.num_envs
        self.device = cfg.device
        self.spacing = cfg.spacing
        self.headless = cfg.debug.headless
        self.physics_engine = cfg.physics_engine
        self.sim_params = cfg.sim_params
        self.viewer = None
        self.env_handles = []
        self._units: List[Unit] = []
        self._actors: List[Actor] = []
        self._kernels = None
        self.init_done = False
        self.gym = gymapi.acquire_gym()
        _, dev_id = gymutil.parse_device_str(self.device)
        self.sim_params._fake_device = self.device          # read by the stand-in simulator only
        self.sim = self.gym.create_sim(dev_id, dev_id, self.physics_engine, self.sim_params)
        self.create_ground()

    # -- CUDA kernels of the task-independent rows ------------------------------------------
    def kernels(self):
        if self._kernels is None:
            from shifu_b200.hotpath import EnvKernels
            self._kernels = EnvKernels(self.device, self.num_envs)
        return self._kernels

    # -- stepping (isaac_gym.py:42-73) -------------------------------------------------------
    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))

    def step(self, action: torch.Tensor):
        self.render()
        self.robot.step(action)
        self.refresh_state()
        self.post_physics_step()

    def post_physics_step(self):
        pass

    def push_root_reset(self, env_ids, actors=None):
        """``set_actor_root_state_tensor_indexed`` for rows already rewritten in place."""
        actors = self._actors if actors is None else actors
        rows = torch.unique(torch.cat([a.root_indices[env_ids] for a in actors])).to(dtype=torch.int32)
        self.gym.set_actor_root_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self.root_state),
                                                     gymtorch.unwrap_tensor(rows), len(rows))

    def reset_idx(self, env_ids: Union[list, torch.Tensor], actors=None):
        if len(env_ids) == 0:
            return
        actors = self._actors if actors is None else actors
        for actor in actors:
            actor.reset_idx(env_ids)
        self.push_root_reset(env_ids, actors)

    # -- construction ------------------------------------------------------------------------
    def create_envs(self, robot: Robot, objects: List[Object] = (), sensors: List[Sensor] = ()):
        self.init_done = False
        self.robot, self.objects, self.sensors = robot, list(objects), list(sensors)
        self._units = [robot, *objects, *sensors]
        self._actors = [robot, *objects]
        for unit in self._units:
            unit.set_env(self)
        bulk = getattr(self.gym, "bulk_create", None)
        if bulk is not None and not sensors and self.num_envs > 8192:
            self._bulk_create(bulk)
        else:
            lo = gymapi.Vec3(-self.spacing, -self.spacing, -self.spacing)
            hi = gymapi.Vec3(self.spacing, self.spacing, self.spacing)
            per_row = int(np.sqrt(self.num_envs))
            for env_id in range(self.num_envs):
                handle = self.gym.create_env(self.sim, lo, hi, per_row)
                for seg_id, unit in enumerate(self._units, 1):
                    unit.load_to(env_id, handle, seg_id)
                self.env_handles.append(handle)
        self.gym.prepare_sim(self.sim)
        self._init_buffers()
        self.init_done = True

    def _bulk_create(self, bulk):
        """Synthetic simulator only: build N identical envs without an N-iteration Python loop."""
        handle = bulk(self.sim, self.num_envs, [a.asset for a in self._actors], [a.name for a in self._actors])
        k = len(self._actors)
        for j, actor in enumerate(self._actors):
            actor.actor_handle = j
            actor.segmentation_id = j + 1
            actor.rigid_body_dict = self.gym.get_actor_rigid_body_dict(handle, j)
            actor.root_indices = torch.arange(self.num_envs, dtype=torch.long) * k + j
        self.env_handles = [handle]

    def _init_buffers(self):
        gym, sim = self.gym, self.sim
        dof = gym.acquire_dof_state_tensor(sim)
        root = gym.acquire_actor_root_state_tensor(sim)
        body = gym.acquire_rigid_body_state_tensor(sim)
        contact = gym.acquire_net_contact_force_tensor(sim)
        for refresh in (gym.refresh_actor_root_state_tensor, gym.refresh_rigid_body_state_tensor,
                        gym.refresh_dof_state_tensor, gym.refresh_jacobian_tensors,
                        gym.refresh_net_contact_force_tensor, gym.refresh_force_sensor_tensor):
            refresh(sim)
        self.dof_state = gymtorch.wrap_tensor(dof)
        self.root_state = gymtorch.wrap_tensor(root)
        self.body_state = gymtorch.wrap_tensor(body)
        self.contact_state = gymtorch.wrap_tensor(contact)
        for unit in self._units:
            unit.init_buffers()

    def refresh_state(self):
        gym, sim = self.gym, self.sim
        gym.simulate(sim)
        if self.device == 'cpu':
            gym.fetch_results(sim, True)
        gym.refresh_actor_root_state_tensor(sim)
        gym.refresh_rigid_body_state_tensor(sim)
        gym.refresh_dof_state_tensor(sim)
        gym.refresh_jacobian_tensors(sim)
        gym.refresh_net_contact_force_tensor(sim)
        gym.refresh_force_sensor_tensor(sim)

    def refresh_sensors(self):                               # isaac_gym.py:159-170
        from shifu_b200.units.sensors import CameraSensor
        for sensor in self.sensors:
            if isinstance(sensor, CameraSensor):
                self.gym.fetch_results(self.sim, True)
                self.gym.step_graphics(self.sim)
                self.gym.render_all_camera_sensors(self.sim)
                self.gym.start_access_image_tensors(self.sim)
                sensor.refresh()
                self.gym.end_access_image_tensors(self.sim)
            else:
                sensor.refresh()

    def create_ground(self):
        self.up_axis_idx = 2
        plane = gymapi.PlaneParams()
        plane.normal = gymapi.Vec3(0.0, 0.0, 1.0)
        plane.static_friction = plane.dynamic_friction = 1.0
        plane.restitution = 0.
        self.gym.add_ground(self.sim, plane)
        self.env_origins = torch.zeros(self.num_envs, 3, device=self.device, requires_grad=False)

    def render(self, sync_frame_time=True):
        if self.viewer is not None:
            raise NotImplementedError("viewer / rendering is outside the shifu_b200 hot path (headless only)")

    def destroy(self):
        self.gym.destroy_sim(self.sim)
        self._kernels = None


class TerrainGymEnv(IsaacGymEnv):
    cfg: TerrainEnvConfig

    def __init__(self, cfg, env_offset: int = 0, num_envs_global: int = None):
        # sharded runs: this process owns global envs [env_offset, env_offset + cfg.num_envs)
        self.env_offset = int(env_offset)
        self.num_envs_global = int(num_envs_global) if num_envs_global else int(cfg.num_envs)
        super().__init__(cfg)

    def _init_buffers(self):
        super()._init_buffers()
        self.height_points = self._init_height_points()

    def create_envs(self, *args, **kwargs):
        self.spacing = 0          # custom origins
        super().create_envs(*args, **kwargs)

    def _init_height_points(self):
        """(N, P, 3) base-frame sample points, P = len(x)*len(y), x fastest (isaac_gym.py:304-318).
        Kept as a stride-0 expanded view: the kernels take the 17x11 grid as constants."""
        y = torch.tensor(self.cfg.terrain.measured_points_y, device=self.device)
        x = torch.tensor(self.cfg.terrain.measured_points_x, device=self.device)
        gx, gy = torch.meshgrid(x, y, indexing='xy')
        self.num_height_points = gx.numel()
        pts = torch.zeros(1, self.num_height_points, 3, device=self.device)
        pts[0, :, 0] = gx.flatten()
        pts[0, :, 1] = gy.flatten()
        return pts.expand(self.num_envs, -1, -1)

    def post_physics_step(self):
        if self.cfg.terrain.measure_heights:
            self.measured_heights = self.get_heights()

    # ``measured_heights`` (isaac_gym.py:320-322) is a plain attribute in user-hook mode.  A fused task
    # consumes the heights inside the kernel; unless it is asked to keep them (``store_measured_heights``,
    # +748 B/env written per step) the attribute is filled on demand by the stand-alone scan over the
    # CURRENT root rows — identical for every env that did not reset in the last step.
    @property
    def measured_heights(self):
        kept = self.__dict__.get("_measured_heights")
        if kept is None and self.__dict__.get("_heights_on_demand", False):
            return self.get_heights()
        return kept

    @measured_heights.setter
    def measured_heights(self, value):
        self.__dict__["_measured_heights"] = value

    @measured_heights.deleter
    def measured_heights(self):
        self.__dict__.pop("_measured_heights", None)

    def create_ground(self):
        assert isinstance(self.cfg, TerrainEnvConfig), "cfg must be a TerrainEnvConfig"
        self.up_axis_idx = 2
        tc = self.cfg.terrain
        if tc.mesh_type not in ('heightfield', 'trimesh'):
            raise NotImplementedError("cfg.terrain.mesh_type must be one of heightfield or trimesh")
        # Row N3: on a CUDA device the map is rasterised by shifu_terrain_generate (bit-identical to the
        # host builder, which remains the path for the closed-source Isaac Gym generators and for
        # cfg.terrain.generator = "host")
        on_device = getattr(tc, "generator", "auto") in ("auto", "device") and str(self.device).startswith("cuda")
        self.terrain = Terrain(tc, self.num_envs, device=self.device if on_device else None)
        params = gymapi.HeightFieldParams() if tc.mesh_type == 'heightfield' else gymapi.TriangleMeshParams()
        params.transform.p.x = params.transform.p.y = -tc.border_size
        params.transform.p.z = 0.0
        params.static_friction, params.dynamic_friction = tc.static_friction, tc.dynamic_friction
        params.restitution = tc.restitution
        if tc.mesh_type == 'heightfield':
            params.column_scale = params.row_scale = tc.horizontal_scale
            params.vertical_scale = tc.vertical_scale
            params.nbRows, params.nbColumns = self.terrain.tot_cols, self.terrain.tot_rows
            self.gym.add_heightfield(self.sim, self.terrain.heightsamples, params)
        else:
            params.nb_vertices = self.terrain.vertices.shape[0]
            params.nb_triangles = self.terrain.triangles.shape[0]
            self.gym.add_triangle_mesh(self.sim, self.terrain.vertices.flatten(order='C'),
                                       self.terrain.triangles.flatten(order='C'), params)
        if self.terrain.device_map is not None:
            self.height_samples = self.terrain.device_map           # built on the device: no upload
        else:
            self.height_samples = torch.tensor(self.terrain.heightsamples).view(
                self.terrain.tot_rows, self.terrain.tot_cols).to(self.device)
        # spawn origins (isaac_gym.py:336-347); terrain type from the GLOBAL env index so that a
        # sharded run reproduces the single-process assignment
        max_init = tc.max_init_terrain_level if tc.curriculum else tc.num_rows - 1
        self.terrain_levels = torch.randint(0, max_init + 1, (self.num_envs,)).to(self.device)
        gid = torch.arange(self.env_offset, self.env_offset + self.num_envs)
        self.terrain_types = torch.div(gid, (self.num_envs_global / tc.num_cols),
                                       rounding_mode='floor').to(torch.long).to(self.device)
        self.max_terrain_level = tc.num_rows
        self.terrain_origins = torch.from_numpy(self.terrain.env_origins).to(self.device).to(torch.float)
        self.env_origins = torch.zeros(self.num_envs, 3, device=self.device, requires_grad=False)
        self.env_origins[:] = self.terrain_origins[self.terrain_levels, self.terrain_types]

    def update_terrain_level(self, env_ids, levels):
        if not self.init_done:
            return
        self.terrain_levels = levels
        self.env_origins[env_ids] = self.terrain_origins[self.terrain_levels[env_ids], self.terrain_types[env_ids]]

    # -- height scan ---------------------------------------------------------------------------
    def scan_kernel(self):
        """Native context for the stand-alone height scan (17x11 grid of the reference config)."""
        if getattr(self, "_scan", None) is None:
            from shifu_b200 import hotpath
            tc = self.cfg.terrain
            layout = self.robot.affine_root_layout()
            if layout is None:
                raise NotImplementedError("non-affine root_indices are not supported by shifu_get_heights")
            desc = hotpath.a1_desc(self.num_envs, points_x=tuple(tc.measured_points_x),
                                   points_y=tuple(tc.measured_points_y), border_size=float(tc.border_size),
                                   horizontal_scale=tc.horizontal_scale, vertical_scale=tc.vertical_scale,
                                   max_terrain_level=tc.num_rows, num_terrain_types=tc.num_cols,
                                   root_stride=layout[0], root_offset=layout[1], terms=("torques_penalize",))
            self._scan = hotpath.HeightScan(desc, self.height_samples, self.root_state)
        return self._scan

    def get_heights(self, env_ids=None):
        """isaac_gym.py:393-433 — (N, P) fp32 heights under the yaw-rotated sample grid."""
        if self.cfg.terrain.mesh_type == 'plane':
            return torch.zeros(self.num_envs, self.num_height_points, device=self.device, requires_grad=False)
        if self.cfg.terrain.mesh_type == 'none':
            raise NameError("Can't measure height with terrain mesh type 'none'")
        out = self.scan_kernel().run()
        return out[env_ids] if env_ids else out
