"""A minimal stand-in for NVIDIA's closed-distribution ``isaacgym`` package.

Isaac Gym (Preview 3) is a manually downloaded tarball that is not available in
this image, so neither the reference (``/root/reference``) nor this package can
talk to PhysX here.  This module provides the *tensor API surface* both of them
use (``gymapi`` / ``gymtorch`` / ``gymutil`` / ``torch_utils`` /
``terrain_utils``) backed by plain torch tensors, with the physics replaced by a
*snapshot provider* that writes synthetic state into the flat gym tensors each
time the caller asks the simulator to refresh them.

It is simulator plumbing, not part of the hot path: the hot path only ever sees
the flat tensors (``root_state (n_actors*N,13)``, ``dof_state (dofs*N,2)``,
``rigid_body_state (bodies*N,13)``, ``net_contact_force (bodies*N,3)``), exactly
the layouts ``gymtorch.wrap_tensor(gym.acquire_*_tensor(sim))`` hands the
reference (``shifu/gym/isaac_gym.py:110-130``).

``install()`` registers the stub modules under ``sys.modules['isaacgym*']``
(plus ``rsl_rl`` / ``matplotlib`` / ``pybullet`` place-holders) **only when the
genuine packages are absent**, so the same code runs against real Isaac Gym.

The arithmetic helpers in ``torch_utils`` restate the published (BSD-3)
definitions of ``isaacgymenvs/utils/torch_jit_utils.py``; see SURVEY.md §8(c).
"""
from __future__ import annotations

import importlib.machinery
import sys
import types
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

# ---------------------------------------------------------------------------
# gymapi value types
# ---------------------------------------------------------------------------


class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __add__(self, o):
        return Vec3(self.x + o.x, self.y + o.y, self.z + o.z)

    __iadd__ = __add__

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __repr__(self):
        return f"Vec3({self.x}, {self.y}, {self.z})"


class Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = float(x), float(y), float(z), float(w)


class Transform:
    def __init__(self, p=None, r=None):
        self.p = p if p is not None else Vec3()
        self.r = r if r is not None else Quat()


class _Bag:
    """Attribute bag used for the option structs (SimParams, AssetOptions, ...)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class SimParams(_Bag):
    def __init__(self):
        super().__init__(physx=_Bag(), flex=_Bag())


class AssetOptions(_Bag):
    pass


class CameraProperties(_Bag):
    pass


class PlaneParams(_Bag):
    pass


class HeightFieldParams(_Bag):
    def __init__(self):
        super().__init__(transform=Transform())


class TriangleMeshParams(_Bag):
    def __init__(self):
        super().__init__(transform=Transform())


# ---------------------------------------------------------------------------
# Asset registry: what load_asset() would have parsed out of the URDFs.
# ---------------------------------------------------------------------------

DOF_PROP_DTYPE = np.dtype([
    ("hasLimits", np.bool_), ("lower", np.float32), ("upper", np.float32),
    ("driveMode", np.int32), ("velocity", np.float32), ("effort", np.float32),
    ("stiffness", np.float32), ("damping", np.float32), ("friction", np.float32),
    ("armature", np.float32),
], align=True)


class AssetSpec:
    def __init__(self, name, bodies: List[str], dofs: List[str], lower, upper, velocity, effort,
                 fixed_base=False):
        self.name = name
        self.bodies = list(bodies)
        self.dofs = list(dofs)
        self.lower, self.upper = list(lower), list(upper)
        self.velocity, self.effort = list(velocity), list(effort)
        self.fixed_base = fixed_base

    @property
    def num_bodies(self):
        return len(self.bodies)

    @property
    def num_dofs(self):
        return len(self.dofs)

    def dof_properties(self):
        p = np.zeros(self.num_dofs, dtype=DOF_PROP_DTYPE)
        p["hasLimits"] = True
        p["lower"], p["upper"] = self.lower, self.upper
        p["velocity"], p["effort"] = self.velocity, self.effort
        return p


def _a1_spec() -> AssetSpec:
    # asset/urdf/a1/urdf/a1.urdf: with collapse_fixed_joints the tree is base + 4 x
    # (hip, thigh, calf, foot[dont_collapse]) in depth-first order FR, FL, RR, RL
    # (a1.urdf:31-558); effort limits hip 20 / thigh 55 / calf 55 (a1.urdf:95,137,165).
    legs = ["FR", "FL", "RR", "RL"]
    bodies = ["base"] + [f"{l}_{p}" for l in legs for p in ("hip", "thigh", "calf", "foot")]
    dofs = [f"{l}_{j}_joint" for l in legs for j in ("hip", "thigh", "calf")]
    lower = [-0.802851455917, -1.0471975512, -2.69653369433] * 4
    upper = [0.802851455917, 4.18879020479, -0.916297857297] * 4
    return AssetSpec("a1", bodies, dofs, lower, upper, [52.4, 28.6, 28.6] * 4, [20.0, 55.0, 55.0] * 4)


def _abb_spec() -> AssetSpec:
    # asset/urdf/abb_rod_description/urdf/abb_rod_isaac.urdf:31-271 — link_6/flange/tool0 are
    # attached by fixed joints and collapse into link_5; 5 arm joints + the tool0-tip joint.
    bodies = ["base_link", "link_1", "link_2", "link_3", "link_4", "link_5", "tip0"]
    dofs = ["joint_1", "joint_2", "joint_3", "joint_4", "joint_5", "tool0-tip"]
    return AssetSpec("abb", bodies, dofs,
                     [-2.967, -1.745, -3.491, -4.712, -2.269, -3.141],
                     [2.967, 2.356, 1.222, 4.712, 2.269, 3.141],
                     [5.027, 4.189, 5.236, 6.981, 7.069, 10.472],
                     [90.0, 90.0, 90.0, 90.0, 90.0, 0.0], fixed_base=True)


ASSET_REGISTRY: Dict[str, Callable[[], AssetSpec]] = {
    "urdf/a1/urdf/a1.urdf": _a1_spec,
    "urdf/abb_rod_description/urdf/abb_rod_isaac.urdf": _abb_spec,
}


def _box_spec() -> AssetSpec:
    return AssetSpec("box", ["box"], [], [], [], [], [], fixed_base=False)


class _ShapeProps:
    def __init__(self):
        self.friction = 1.0
        self.rolling_friction = 0.0
        self.torsion_friction = 0.0
        self.restitution = 0.0


class _BodyProps:
    def __init__(self):
        self.mass = 1.0


# ---------------------------------------------------------------------------
# The fake Gym
# ---------------------------------------------------------------------------


class FakeSim:
    def __init__(self, device: str):
        self.device = device
        self.envs: List["FakeEnv"] = []
        self.actor_assets: List[AssetSpec] = []     # one entry per actor, sim-domain order
        self.actor_names: List[str] = []
        self.prepared = False
        # Flat state tensors (allocated in prepare_sim)
        self.root_state = self.dof_state = self.body_state = self.contact_state = None
        self.jacobians: Dict[str, torch.Tensor] = {}
        # The snapshot provider: callable(kind: str, sim: FakeSim) -> None, kinds are
        # 'simulate', 'root', 'body', 'dof', 'contact'.  None = leave tensors untouched.
        self.provider: Optional[Callable[[str, "FakeSim"], None]] = None
        self.calls: Dict[str, int] = {}
        self.last_forces = None
        self.last_dof_forces = None
        self.last_dof_targets = None
        # camera sensors: "rendered" frames are a seeded function of (env, image type, frame)
        self.camera_frame = 0
        self.camera_provider: Optional[Callable[[int, int, int, int, int], torch.Tensor]] = None


class FakeEnv:
    def __init__(self, index: int):
        self.index = index
        self.actors: List[int] = []        # sim-domain actor indices
        self.cameras: List[Dict] = []      # per camera: props, pose, lazily allocated image tensors


_IMAGE_DTYPES = {0: torch.uint8, 1: torch.float32, 2: torch.int32, 3: torch.int16}   # gymapi.IMAGE_*


def synthetic_camera_image(env_index: int, image_type: int, frame: int, height: int, width: int) -> torch.Tensor:
    """Seeded stand-in for a rendered frame (CPU tensor): RGBA uint8, NEGATIVE depth in metres (as
    Isaac Gym returns it), int32 segmentation ids, int16 optical flow."""
    g = torch.Generator().manual_seed(1_000_003 * env_index + 101 * image_type + frame + 7)
    if image_type == 0:
        return torch.randint(0, 256, (height, width, 4), generator=g, dtype=torch.uint8)
    if image_type == 1:
        d = -(0.1 + 2.9 * torch.rand(height, width, generator=g))
        d[0, 0] = 0.0                                      # -(+0.0) = -0.0 must survive the gather
        d[0, 1 % width] = float("-inf")
        return d
    if image_type == 2:
        return torch.randint(0, 5, (height, width), generator=g, dtype=torch.int32)
    return torch.randint(-32768, 32768, (height, width), generator=g, dtype=torch.int32).to(torch.int16)


class FakeGym:
    """Implements the subset of ``gymapi.Gym`` reached from the hot path's callers."""

    # the stand-in keeps no hidden simulator state: rows rewritten in place need no indexed setter
    needs_indexed_resets = False

    def __init__(self):
        self.sims: List[FakeSim] = []

    # -- sim / ground -------------------------------------------------------
    def create_sim(self, compute_device, graphics_device, physics_engine, sim_params):
        dev = getattr(sim_params, "_fake_device", None) or _DEFAULT_DEVICE[0]
        sim = FakeSim(dev)
        self.sims.append(sim)
        return sim

    def add_ground(self, sim, params):
        pass

    def add_heightfield(self, sim, samples, params):
        pass

    def add_triangle_mesh(self, sim, vertices, triangles, params):
        pass

    # -- assets -------------------------------------------------------------
    def load_asset(self, sim, root, filename, options=None):
        if filename not in ASSET_REGISTRY:
            raise FileNotFoundError(f"fake_isaacgym: no AssetSpec registered for {filename!r}")
        spec = ASSET_REGISTRY[filename]()
        if options is not None and getattr(options, "fix_base_link", False):
            spec.fixed_base = True
        return spec

    def create_box(self, sim, w, h, d, options=None):
        return _box_spec()

    def get_asset_rigid_body_count(self, asset):
        return asset.num_bodies

    def get_asset_dof_count(self, asset):
        return asset.num_dofs

    def get_asset_dof_properties(self, asset):
        return asset.dof_properties()

    def get_asset_rigid_shape_properties(self, asset):
        return [_ShapeProps() for _ in range(asset.num_bodies)]

    def set_asset_rigid_shape_properties(self, asset, props):
        pass

    def get_asset_rigid_body_dict(self, asset):
        return {n: i for i, n in enumerate(asset.bodies)}

    # -- envs / actors ------------------------------------------------------
    def create_env(self, sim, lower, upper, per_row):
        env = FakeEnv(len(sim.envs))
        sim._cur_env = env
        sim.envs.append(env)
        return env

    def create_actor(self, env, asset, pose, name, group, filt, seg_id=0):
        sim = self._sim_of(env)
        idx = len(sim.actor_assets)
        sim.actor_assets.append(asset)
        sim.actor_names.append(name)
        env.actors.append(idx)
        return len(env.actors) - 1

    def _sim_of(self, env):
        for sim in self.sims:
            if env.index < len(sim.envs) and sim.envs[env.index] is env:
                return sim
        raise RuntimeError("unknown env handle")

    def get_actor_index(self, env, actor_handle, domain):
        return env.actors[actor_handle]

    def get_actor_rigid_body_dict(self, env, actor_handle):
        sim = self._sim_of(env)
        return {n: i for i, n in enumerate(sim.actor_assets[env.actors[actor_handle]].bodies)}

    def find_actor_rigid_body_handle(self, env, actor_handle, name):
        return self.get_actor_rigid_body_dict(env, actor_handle)[name]

    def set_rigid_body_segmentation_id(self, *a, **k):
        pass

    def set_rigid_body_color(self, *a, **k):
        pass

    def set_actor_dof_properties(self, *a, **k):
        pass

    def get_actor_rigid_shape_properties(self, env, actor_handle):
        return [_ShapeProps()]

    def set_actor_rigid_shape_properties(self, *a, **k):
        pass

    def get_actor_rigid_body_properties(self, env, actor_handle):
        return [_BodyProps()]

    def set_actor_rigid_body_properties(self, *a, **k):
        pass

    # Bulk construction (not part of the real API; lets the synthetic backend build a
    # million envs without a Python loop).  ``assets`` is the per-env actor list.
    def bulk_create(self, sim, num_envs: int, assets: List[AssetSpec], names: List[str]):
        sim.bulk = (num_envs, list(assets), list(names))
        sim.envs = [FakeEnv(0)]
        sim.envs[0].actors = list(range(len(assets)))
        sim.actor_assets = list(assets)
        sim.actor_names = list(names)
        return sim.envs[0]

    # -- tensors ------------------------------------------------------------
    def prepare_sim(self, sim):
        bulk = getattr(sim, "bulk", None)
        if bulk is not None:
            n_env, assets, names = bulk
            per_env_assets = assets
        else:
            n_env = len(sim.envs)
            k = len(sim.envs[0].actors) if n_env else 0
            per_env_assets = sim.actor_assets[:k]
            names = sim.actor_names[:k]
        sim.num_envs = n_env
        sim.per_env_assets = per_env_assets
        n_act = len(per_env_assets)
        n_body = sum(a.num_bodies for a in per_env_assets)
        n_dof = sum(a.num_dofs for a in per_env_assets)
        dev = sim.device
        sim.root_state = torch.zeros(n_env * n_act, 13, device=dev)
        sim.root_state[:, 6] = 1.0
        sim.dof_state = torch.zeros(n_env * n_dof, 2, device=dev)
        sim.body_state = torch.zeros(n_env * n_body, 13, device=dev)
        sim.body_state[:, 6] = 1.0
        sim.contact_state = torch.zeros(n_env * n_body, 3, device=dev)
        for a, nm in zip(per_env_assets, names):
            if a.num_dofs:
                links = a.num_bodies - 1 if a.fixed_base else a.num_bodies
                cols = a.num_dofs if a.fixed_base else a.num_dofs + 6
                if a.fixed_base:
                    # arms read it every step (ArmRobot.inverse_kinematics, row N2): a real tensor
                    sim.jacobians[nm] = torch.zeros(n_env, links, 6, cols, device=dev)
                else:
                    # legged robots never read it on the hot path: stride-0 expand, no HBM spent
                    sim.jacobians[nm] = torch.zeros(1, links, 6, cols, device=dev).expand(n_env, links, 6, cols)
        sim.prepared = True

    def acquire_dof_state_tensor(self, sim):
        return sim.dof_state

    def acquire_actor_root_state_tensor(self, sim):
        return sim.root_state

    def acquire_rigid_body_state_tensor(self, sim):
        return sim.body_state

    def acquire_net_contact_force_tensor(self, sim):
        return sim.contact_state

    def acquire_jacobian_tensor(self, sim, name):
        return sim.jacobians[name]

    def _tick(self, sim, kind):
        sim.calls[kind] = sim.calls.get(kind, 0) + 1
        if sim.provider is not None:
            sim.provider(kind, sim)

    def simulate(self, sim):
        self._tick(sim, "simulate")

    def fetch_results(self, sim, wait):
        pass

    # -- camera sensors (shifu/units/sensors.py) --------------------------
    def create_camera_sensor(self, env, props):
        env.cameras.append(dict(props=props, images={}, location=None, transform=None))
        return len(env.cameras) - 1

    def destroy_camera_sensor(self, sim, env, cam):
        env.cameras[cam]["images"].clear()

    def set_camera_location(self, cam, env, pos, lookat):
        env.cameras[cam]["location"] = (pos, lookat)

    def set_camera_transform(self, cam, env, transform):
        env.cameras[cam]["transform"] = transform

    def get_camera_proj_matrix(self, sim, env, cam):
        return np.eye(4, dtype=np.float32)

    def get_camera_view_matrix(self, sim, env, cam):
        return np.eye(4, dtype=np.float32)

    def _render_into(self, sim, env, cam, image_type, t):
        p = env.cameras[cam]["props"]
        fn = sim.camera_provider or synthetic_camera_image
        t.copy_(fn(env.index, image_type, sim.camera_frame, int(p.height), int(p.width)))

    def get_camera_image_gpu_tensor(self, sim, env, cam, image_type):
        """One persistent tensor per (env, camera, image type), like Isaac Gym's interop buffers."""
        c = env.cameras[cam]
        if image_type not in c["images"]:
            h, w = int(c["props"].height), int(c["props"].width)
            shape = (h, w, 4) if image_type == 0 else (h, w)
            c["images"][image_type] = torch.zeros(shape, dtype=_IMAGE_DTYPES[image_type], device=sim.device)
            self._render_into(sim, env, cam, image_type, c["images"][image_type])
        return c["images"][image_type]

    def step_graphics(self, sim):
        pass

    def render_all_camera_sensors(self, sim):
        sim.camera_frame += 1
        for env in sim.envs:
            for cam, c in enumerate(env.cameras):
                for image_type, t in c["images"].items():
                    self._render_into(sim, env, cam, image_type, t)

    def start_access_image_tensors(self, sim):
        pass

    def end_access_image_tensors(self, sim):
        pass

    def refresh_actor_root_state_tensor(self, sim):
        self._tick(sim, "root")

    def refresh_rigid_body_state_tensor(self, sim):
        self._tick(sim, "body")

    def refresh_dof_state_tensor(self, sim):
        self._tick(sim, "dof")

    def refresh_net_contact_force_tensor(self, sim):
        self._tick(sim, "contact")

    def refresh_jacobian_tensors(self, sim):
        pass

    def refresh_force_sensor_tensor(self, sim):
        pass

    # setters: the callers have already written the flat tensors in place
    def set_dof_actuation_force_tensor(self, sim, t):
        sim.last_dof_forces = t
        sim.calls["set_dof_force"] = sim.calls.get("set_dof_force", 0) + 1

    def set_dof_position_target_tensor(self, sim, t):
        sim.last_dof_targets = t

    def set_dof_velocity_target_tensor(self, sim, t):
        pass

    def set_dof_position_target_tensor_indexed(self, sim, t, idx, n):
        sim.last_dof_targets = t

    def set_dof_state_tensor_indexed(self, sim, t, idx, n):
        sim.last_dof_state_idx = (idx, n)

    def set_actor_root_state_tensor(self, sim, t):
        pass

    def set_actor_root_state_tensor_indexed(self, sim, t, idx, n):
        sim.last_root_idx = (idx, n)

    def apply_rigid_body_force_at_pos_tensors(self, sim, force, pos=None, space=None):
        sim.last_forces = force

    # viewer / graphics: headless only
    def create_viewer(self, *a, **k):
        return None

    def destroy_viewer(self, *a, **k):
        pass

    def destroy_sim(self, *a, **k):
        pass

    def destroy_env(self, *a, **k):
        pass


_GYM = [None]
_DEFAULT_DEVICE = ["cpu"]


def set_default_device(device: str):
    """Device the fake sim allocates its flat tensors on ('cpu' for the oracle harness,
    'cuda:0' for the product path)."""
    _DEFAULT_DEVICE[0] = device


def acquire_gym():
    if _GYM[0] is None:
        _GYM[0] = FakeGym()
    return _GYM[0]


def reset_gym():
    _GYM[0] = None


# ---------------------------------------------------------------------------
# isaacgym.torch_utils (restated from the public BSD-3 definitions)
# ---------------------------------------------------------------------------


def _to_torch(x, dtype=torch.float, device=None, requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device or _DEFAULT_DEVICE[0], requires_grad=requires_grad)


def _normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


def _quat_apply(a, b):
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, :3]
    t = xyz.cross(b, dim=-1) * 2
    return (b + a[:, 3:] * t + xyz.cross(t, dim=-1)).view(shape)


def _quat_rotate(q, v):
    shape = q.shape
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a + b + c


def _quat_rotate_inverse(q, v):
    shape = q.shape
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def _quat_mul(a, b):
    shape = a.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    x1, y1, z1, w1 = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    x2, y2, z2, w2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return torch.stack([x, y, z, w], dim=-1).view(shape)


def _quat_conjugate(a):
    shape = a.shape
    a = a.reshape(-1, 4)
    return torch.cat((-a[:, :3], a[:, -1:]), dim=-1).view(shape)


def _quat_from_euler_xyz(roll, pitch, yaw):
    cy = torch.cos(yaw * 0.5)
    sy = torch.sin(yaw * 0.5)
    cr = torch.cos(roll * 0.5)
    sr = torch.sin(roll * 0.5)
    cp = torch.cos(pitch * 0.5)
    sp = torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qx, qy, qz, qw], dim=-1)


def _torch_rand_float(lower, upper, shape, device):
    return (upper - lower) * torch.rand(*shape, device=device) + lower


def _get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
    zs = np.zeros((n_dims,))
    assert axis_idx < n_dims
    zs[axis_idx] = 1.0
    params = np.where(zs == 1.0, value, zs)
    params[0] = x_value
    return list(params.astype(dtype))


def _tensor_clamp(t, min_t, max_t):
    return torch.max(torch.min(t, max_t), min_t)


# ---------------------------------------------------------------------------
# Module assembly
# ---------------------------------------------------------------------------


def _build_modules():
    pkg = types.ModuleType("isaacgym")
    pkg.__path__ = []  # mark as package
    pkg.__fake__ = True

    gymapi = types.ModuleType("isaacgym.gymapi")
    for k, v in dict(
        acquire_gym=acquire_gym, Vec3=Vec3, Quat=Quat, Transform=Transform, SimParams=SimParams,
        AssetOptions=AssetOptions, CameraProperties=CameraProperties, PlaneParams=PlaneParams,
        HeightFieldParams=HeightFieldParams, TriangleMeshParams=TriangleMeshParams,
        Gym=FakeGym, Sim=FakeSim, Asset=AssetSpec, Env=FakeEnv,
        SIM_PHYSX=1, SIM_FLEX=0, UP_AXIS_Y=0, UP_AXIS_Z=1,
        DOF_MODE_NONE=0, DOF_MODE_POS=1, DOF_MODE_VEL=2, DOF_MODE_EFFORT=3,
        FROM_ASSET=0, COMPUTE_PER_VERTEX=1, COMPUTE_PER_FACE=2,
        DOMAIN_ENV=0, DOMAIN_SIM=1, DOMAIN_ACTOR=2,
        IMAGE_COLOR=0, IMAGE_DEPTH=1, IMAGE_SEGMENTATION=2, IMAGE_OPTICAL_FLOW=3,
        MESH_NONE=0, MESH_COLLISION=1, MESH_VISUAL=2, MESH_VISUAL_AND_COLLISION=3,
        KEY_ESCAPE=256, KEY_V=86, ENV_SPACE=0, LOCAL_SPACE=1, GLOBAL_SPACE=2,
    ).items():
        setattr(gymapi, k, v)

    gymtorch = types.ModuleType("isaacgym.gymtorch")
    gymtorch.wrap_tensor = lambda t: t
    gymtorch.unwrap_tensor = lambda t: t

    gymutil = types.ModuleType("isaacgym.gymutil")

    def parse_device_str(s):
        s = str(s)
        if s.startswith("cuda"):
            return "cuda", int(s.split(":")[1]) if ":" in s else 0
        return "cpu", 0

    gymutil.parse_device_str = parse_device_str

    tu = types.ModuleType("isaacgym.torch_utils")
    for k, v in dict(
        torch=torch, np=np, to_torch=_to_torch, normalize=_normalize, quat_apply=_quat_apply,
        quat_rotate=_quat_rotate, quat_rotate_inverse=_quat_rotate_inverse, quat_mul=_quat_mul,
        quat_conjugate=_quat_conjugate, quat_from_euler_xyz=_quat_from_euler_xyz,
        torch_rand_float=_torch_rand_float, get_axis_params=_get_axis_params,
        tensor_clamp=_tensor_clamp,
    ).items():
        setattr(tu, k, v)

    terr = types.ModuleType("isaacgym.terrain_utils")
    from . import synthetic_terrain as _st
    for name in ("SubTerrain", "pyramid_sloped_terrain", "random_uniform_terrain", "pyramid_stairs_terrain",
                 "discrete_obstacles_terrain", "stepping_stones_terrain", "convert_heightfield_to_trimesh",
                 "pyramid_params", "random_uniform_params", "stairs_params", "obstacles_params", "stones_params"):
        setattr(terr, name, getattr(_st, name))

    pkg.gymapi, pkg.gymtorch, pkg.gymutil = gymapi, gymtorch, gymutil
    pkg.torch_utils, pkg.terrain_utils = tu, terr
    return {m.__name__: m for m in (pkg, gymapi, gymtorch, gymutil, tu, terr)}


def _placeholder(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


class VecEnv:
    """Shape of ``rsl_rl.env.VecEnv`` (an abstract base with no behaviour)."""
    num_envs: int
    num_obs: int
    num_privileged_obs: int
    num_actions: int
    max_episode_length: int


class OnPolicyRunner:
    def __init__(self, *a, **k):
        raise RuntimeError("rsl_rl is not installed; the PPO runner is out of scope here")


def install(device: str = "cpu", force: bool = False) -> bool:
    """Register the stand-in modules.  Returns True when the fakes are in use."""
    set_default_device(device)
    have_real = False
    if not force and "isaacgym" not in sys.modules:
        try:
            import importlib.util
            have_real = importlib.util.find_spec("isaacgym") is not None
        except (ImportError, ValueError):
            have_real = False
    if "isaacgym" in sys.modules and not getattr(sys.modules["isaacgym"], "__fake__", False) and not force:
        have_real = True
    if not have_real and not getattr(sys.modules.get("isaacgym"), "__fake__", False):
        sys.modules.update(_build_modules())
    # rsl_rl: only the VecEnv base class and the runner symbol are imported by the callers
    if "rsl_rl" not in sys.modules:
        try:
            import rsl_rl  # noqa: F401
        except ImportError:
            env_m = _placeholder("rsl_rl.env", VecEnv=VecEnv)
            run_m = _placeholder("rsl_rl.runners", OnPolicyRunner=OnPolicyRunner)
            sys.modules.update({"rsl_rl": _placeholder("rsl_rl", env=env_m, runners=run_m),
                                "rsl_rl.env": env_m, "rsl_rl.runners": run_m})
    return not have_real
