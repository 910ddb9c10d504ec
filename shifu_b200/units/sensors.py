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

    def __init__(self, cfg: CameraSensorConfig):
        super().__init__(cfg)
        self.width = cfg.camera_props.width
        self.height = cfg.camera_props.height
        self.near_plane = cfg.camera_props.near_plane
        self.far_plane = cfg.camera_props.far_plane
        self._tables = None

    def init_buffers(self):                                    # sensors.py:66-94
        n, h, w = self.env.num_envs, self.height, self.width
        for img_type in self.cfg.image_types:
            if img_type == IMAGE_TYPE_COLOR:
                self.color_buf = torch.zeros(n, h, w, 3 if self.cfg.image_normalization else 4,
                                             dtype=torch.float if self.cfg.image_normalization else torch.uint8,
                                             device=self.device)
            elif img_type == IMAGE_TYPE_DEPTH:
                self.depth_buf = torch.zeros(n, h, w, dtype=torch.float, device=self.device)
            elif img_type == IMAGE_TYPE_SEGMENTATION:
                self.segmentation_buf = torch.zeros(n, h, w, dtype=torch.int32, device=self.device)
            elif img_type == IMAGE_TYPE_OPTICAL_FLOW:
                self.optical_flow_buf = torch.zeros(n, h, w, dtype=torch.int16, device=self.device)
            else:
                raise NotImplementedError

    def _init_props(self):                                     # sensors.py:96-115
        if self.cfg.local_lookat_positions is not None:
            assert self.cfg.transform is None and self.cfg.attach_local_transform is None
            self.local_lookat_position = (gymapi.Vec3(*self.cfg.local_lookat_positions[0]),
                                          gymapi.Vec3(*self.cfg.local_lookat_positions[1]))
            self._pose_type = CameraPose.LocalLookat
        elif self.cfg.transform is not None:
            assert self.cfg.local_lookat_positions is None and self.cfg.attach_local_transform is None
            self.transform = gymapi.Transform()
            self.transform.p = gymapi.Vec3(*self.cfg.transform[0])
            self.transform.r = gymapi.Quat(*self.cfg.transform[1])
            self._pose_type = CameraPose.Transform
        elif self.cfg.attach_local_transform is not None:
            raise NotImplementedError('Currently not support')
        else:
            raise NotImplementedError('choose one of method from local_lookat_positions and transform')

    def reset_idx(self, env_ids):
        pass

    def _read_matrices(self, env_handle):
        self.proj_matrix = np.matrix(self.gym.get_camera_proj_matrix(self.sim, env_handle, self.camera_handle))
        self.view_matrix = np.matrix(self.gym.get_camera_view_matrix(self.sim, env_handle, self.camera_handle))

    def load_to(self, env_id, env_handle, seg_id):             # sensors.py:120-137
        camera_handle = self.gym.create_camera_sensor(env_handle, self.cfg.camera_props)
        if self._pose_type == CameraPose.LocalLookat:
            self.gym.set_camera_location(camera_handle, env_handle, *self.local_lookat_position)
        elif self._pose_type == CameraPose.Transform:
            self.gym.set_camera_transform(camera_handle, env_handle, self.transform)
        else:
            raise NotImplementedError
        if env_id == 0:
            self.camera_handle = camera_handle
            self._read_matrices(env_handle)

    def set_camera_transform(self, position, rotation):        # sensors.py:139-148
        for env_id, env_handle in enumerate(self.env.env_handles):
            transform = gymapi.Transform()
            transform.p = gymapi.Vec3(*position)
            transform.r = gymapi.Quat(*rotation)
            self.gym.set_camera_transform(self.camera_handle, env_handle, self.transform)
            if env_id == 0:
                self.transform = transform
                self._read_matrices(env_handle)

    def set_camera_location(self, local_pos, lookat_pos):      # sensors.py:150-159
        for env_id, env_handle in enumerate(self.env.env_handles):
            self.gym.set_camera_location(self.camera_handle, env_handle, gymapi.Vec3(*local_pos),
                                         gymapi.Vec3(*lookat_pos))
            if env_id == 0:
                self.local_lookat_position = (local_pos, lookat_pos)
                self._read_matrices(env_handle)

    def refresh(self):
        self.refresh_image_tensors()

    # -- the batched gather ---------------------------------------------------------------------
    _SPECS = ((IMAGE_TYPE_COLOR, "color", "color_buf", torch.uint8, 4),
              (IMAGE_TYPE_DEPTH, "depth", "depth_buf", torch.float32, 0),
              (IMAGE_TYPE_SEGMENTATION, "seg", "segmentation_buf", torch.int32, 0),
              (IMAGE_TYPE_OPTICAL_FLOW, "flow", "optical_flow_buf", torch.int16, 0))

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
