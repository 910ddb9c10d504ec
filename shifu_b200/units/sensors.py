"""CameraSensor — host mirror of shifu/units/sensors.py.

The reference refreshes its batched image buffers with a Python loop over the envs, one or more
small torch kernels per env and image type (sensors.py:165-188).  Here the per-env interop tensors'
addresses are collected once into device-resident pointer tables and every refresh is ONE
``shifu_camera_gather`` launch.  Rendering to a window (``render``: cv2 / matplotlib) is UI and
stays out of scope."""
import enum

import numpy as np
import torch
from isaacgym import gymapi, gymtorch

from shifu_b200.configs import CameraSensorConfig
from .base import Sensor

IMAGE_TYPE_COLOR = gymapi.IMAGE_COLOR
IMAGE_TYPE_DEPTH = gymapi.IMAGE_DEPTH
IMAGE_TYPE_SEGMENTATION = gymapi.IMAGE_SEGMENTATION
IMAGE_TYPE_OPTICAL_FLOW = gymapi.IMAGE_OPTICAL_FLOW


class CameraPose(enum.Enum):
    LocalLookat = 0
    Transform = 1
    AttachLocalTransform = 2


class CameraSensor(Sensor):
    cfg: CameraSensorConfig

    # image type -> (key of the gather call, buffer attribute, dtype, channels of the per-env tensor)
    _SPECS = ((IMAGE_TYPE_COLOR, "color", "color_buf", torch.uint8, 4),
              (IMAGE_TYPE_DEPTH, "depth", "depth_buf", torch.float32, 0),
              (IMAGE_TYPE_SEGMENTATION, "seg", "segmentation_buf", torch.int32, 0),
              (IMAGE_TYPE_OPTICAL_FLOW, "flow", "optical_flow_buf", torch.int16, 0))

    def __init__(self, cfg: CameraSensorConfig):
        super().__init__(cfg)
        props = cfg.camera_props
        self.width, self.height = props.width, props.height
        self.near_plane, self.far_plane = props.near_plane, props.far_plane
        self._tables = None

    def init_buffers(self):
        """Batched buffers (sensors.py:66-94): RGBA u8 or normalised RGB f32, depth f32, segmentation
        i32, optical flow i16."""
        known = {spec[0]: spec for spec in self._SPECS}
        shape = (self.env.num_envs, self.height, self.width)
        for image_type in self.cfg.image_types:
            if image_type not in known:
                raise NotImplementedError
            _, _, attr, dtype, channels = known[image_type]
            if image_type == IMAGE_TYPE_COLOR and self.cfg.image_normalization:
                dtype, channels = torch.float, 3
            full = shape + ((channels,) if channels else ())
            setattr(self, attr, torch.zeros(full, dtype=dtype, device=self.device))

    def _init_props(self):
        """Exactly one way of posing the camera must be configured (sensors.py:96-115)."""
        cfg = self.cfg
        if cfg.attach_local_transform is not None and cfg.local_lookat_positions is None and cfg.transform is None:
            raise NotImplementedError('Currently not support')
        if cfg.local_lookat_positions is not None:
            assert cfg.transform is None and cfg.attach_local_transform is None
            eye, target = cfg.local_lookat_positions
            self.local_lookat_position = (gymapi.Vec3(*eye), gymapi.Vec3(*target))
            self._pose_type = CameraPose.LocalLookat
        elif cfg.transform is not None:
            assert cfg.attach_local_transform is None
            position, rotation = cfg.transform
            self.transform = self._make_transform(position, rotation)
            self._pose_type = CameraPose.Transform
        else:
            raise NotImplementedError('choose one of method from local_lookat_positions and transform')

    @staticmethod
    def _make_transform(position, rotation):
        t = gymapi.Transform()
        t.p, t.r = gymapi.Vec3(*position), gymapi.Quat(*rotation)
        return t

    def reset_idx(self, env_ids):
        pass

    def _read_matrices(self, env_handle):
        query = (self.sim, env_handle, self.camera_handle)
        self.proj_matrix = np.matrix(self.gym.get_camera_proj_matrix(*query))
        self.view_matrix = np.matrix(self.gym.get_camera_view_matrix(*query))

    def _pose(self, camera_handle, env_handle):
        if self._pose_type is CameraPose.LocalLookat:
            self.gym.set_camera_location(camera_handle, env_handle, *self.local_lookat_position)
        elif self._pose_type is CameraPose.Transform:
            self.gym.set_camera_transform(camera_handle, env_handle, self.transform)
        else:
            raise NotImplementedError

    def load_to(self, env_id, env_handle, seg_id):             # sensors.py:120-137
        handle = self.gym.create_camera_sensor(env_handle, self.cfg.camera_props)
        self._pose(handle, env_handle)
        if env_id == 0:                                        # all envs share handle and intrinsics
            self.camera_handle = handle
            self._read_matrices(env_handle)

    def set_camera_transform(self, position, rotation):        # sensors.py:139-148
        """Re-pose every env's camera.  Reference quirk kept: env 0 is given the transform that was
        current before the call, the new one takes effect from env 1 on."""
        for env_id, env_handle in enumerate(self.env.env_handles):
            self.gym.set_camera_transform(self.camera_handle, env_handle, self.transform)
            if env_id == 0:
                self.transform = self._make_transform(position, rotation)
                self._read_matrices(env_handle)

    def set_camera_location(self, local_pos, lookat_pos):      # sensors.py:150-159
        eye, target = gymapi.Vec3(*local_pos), gymapi.Vec3(*lookat_pos)
        for env_id, env_handle in enumerate(self.env.env_handles):
            self.gym.set_camera_location(self.camera_handle, env_handle, eye, target)
            if env_id == 0:
                self.local_lookat_position = (local_pos, lookat_pos)
                self._read_matrices(env_handle)

    def refresh(self):
        self.refresh_image_tensors()

    # -- the batched gather ---------------------------------------------------------------------

    def _build_tables(self):
        """One pass over the envs (at the first refresh, not per step): the interop tensors Isaac Gym
        hands out per (env, camera, image type) are persistent, so their addresses are too."""
        tables, keep = {}, []
        for img_type, key, buf_name, dtype, channels in self._SPECS:
            if img_type not in self.cfg.image_types:
                continue
            shape = (self.height, self.width, channels) if channels else (self.height, self.width)
            ptrs = []
            for env_handle in self.env.env_handles:
                t = gymtorch.wrap_tensor(self.gym.get_camera_image_gpu_tensor(self.sim, env_handle,
                                                                              self.camera_handle, img_type))
                if tuple(t.shape) != shape or t.dtype != dtype or not t.is_contiguous() or t.data_ptr() % 16:
                    raise ValueError(f"camera image tensor {key}: expected contiguous 16-byte aligned {shape} {dtype}, "
                                     f"got {tuple(t.shape)} {t.dtype}")
                if t.device.type != "cuda":
                    from shifu_b200 import _native as nv
                    raise nv.ShifuNativeError(nv.E_NODEVICE, "camera image tensors are not CUDA tensors: "
                                                             "shifu_camera_gather has no CPU fallback")
                keep.append(t)
                ptrs.append(t.data_ptr())
            table = torch.tensor(ptrs, dtype=torch.int64).to(self.device)
            tables[key] = (table, getattr(self, buf_name))
        self._tables, self._image_refs = tables, keep

    def refresh_image_tensors(self):                           # sensors.py:165-188, one launch
        if self._tables is None:
            self._build_tables()
        self.env.kernels().camera_gather(height=self.height, width=self.width,
                                         normalize_color=self.cfg.image_normalization, **self._tables)

    def render(self, render_idx=0, dsize=None):
        raise NotImplementedError("CameraSensor.render (cv2 window) is UI, outside the shifu_b200 hot path")
