"""Row N4: shifu_camera_gather == CameraSensor.refresh_image_tensors
(shifu/units/sensors.py:165-188).  Byte / integer / one-rounding float work: bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KEYS = ("color", "depth", "seg", "flow")


def _fixture():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def _tables(images):
    """images: dict key -> list of per-env CUDA tensors.  Returns key -> int64 pointer table."""
    return {k: torch.tensor([t.data_ptr() for t in v], dtype=torch.int64).cuda() for k, v in images.items()}


def _outputs(n, h, w, normalize, keys):
    o = {}
    if "color" in keys:
        o["color"] = (torch.full((n, h, w, 3), float("nan"), device="cuda") if normalize
                      else torch.zeros(n, h, w, 4, dtype=torch.uint8, device="cuda"))
    if "depth" in keys:
        o["depth"] = torch.full((n, h, w), float("nan"), device="cuda")
    if "seg" in keys:
        o["seg"] = torch.full((n, h, w), -7, dtype=torch.int32, device="cuda")
    if "flow" in keys:
        o["flow"] = torch.full((n, h, w), -7, dtype=torch.int16, device="cuda")
    return o


def _gather(images, n, h, w, normalize):
    from shifu_b200 import hotpath
    k = hotpath.EnvKernels("cuda:0", n)
    tabs, outs = _tables(images), _outputs(n, h, w, normalize, images.keys())
    k.camera_gather(height=h, width=w, normalize_color=normalize, **{key: (tabs[key], outs[key]) for key in images})
    torch.cuda.synchronize()
    return {key: v.cpu() for key, v in outs.items()}


def _same(a: torch.Tensor, b) -> bool:
    b = torch.as_tensor(b)
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    if a.dtype.is_floating_point:            # bit patterns: -0.0 and inf must survive
        return bool(torch.equal(a.view(torch.int32), b.view(torch.int32)))
    return bool(torch.equal(a, b))


@pytest.mark.parametrize("normalize", [False, True])
def test_camera_gather_matches_reference_fixture(normalize):
    z, meta = _fixture()
    n, h, w = meta["n"], meta["height"], meta["width"]
    images = {k: [torch.from_numpy(z["in/" + k][e]).cuda() for e in range(n)] for k in KEYS}
    got = _gather(images, n, h, w, normalize)
    tag = "norm" if normalize else "raw"
    for k in KEYS:
        assert _same(got[k], z[f"{tag}/{k}"]), (k, normalize)


@pytest.mark.parametrize("keys", [("color", "depth", "seg"), ("depth",), ("flow", "color")])
def test_camera_gather_random_vs_oracle(keys):
    """128 x 128 frames (the reference's PushBoxCameraConfig), per-env tensors scattered over
    separate allocations and views, a subset of the image types."""
    from oracle import shifu_oracle as so
    from shifu_b200.sim.fake_isaacgym import synthetic_camera_image
    n, h, w = 67, 128, 128
    t_of = dict(color=0, depth=1, seg=2, flow=3)
    cpu = {k: [synthetic_camera_image(e, t_of[k], 3, h, w) for e in range(n)] for k in keys}
    pool = {k: torch.zeros((2 * n,) + tuple(cpu[k][0].shape), dtype=cpu[k][0].dtype, device="cuda") for k in keys}
    images = {}
    for k in keys:                                     # odd rows of a pool, in reverse order
        images[k] = [pool[k][2 * (n - 1 - e) + 1] for e in range(n)]
        for e in range(n):
            images[k][e].copy_(cpu[k][e])
    for normalize in (False, True):
        want = so.camera_refresh(**cpu, image_normalization=normalize)
        got = _gather(images, n, h, w, normalize)
        for k in keys:
            assert _same(got[k], want[k]), (k, normalize)


def test_camera_sensor_class_api_matches_reference_fixture():
    """The CameraSensor mirror inside an env, refreshed through IsaacGymEnv.refresh_sensors: same
    frames as the fixture (the stand-in renderer is seeded by env, image type and frame)."""
    from shifu_b200.sim import fake_isaacgym
    fake_isaacgym.install("cuda:0")
    fake_isaacgym.reset_gym()
    fake_isaacgym.set_default_device("cuda:0")
    from shifu_b200.configs import CameraSensorConfig
    from shifu_b200.tasks.abb_pushbox import PriorStageEnvConfig, VisionAbbPushBox
    z, meta = _fixture()
    n, h, w = meta["n"], meta["height"], meta["width"]

    for normalize in (False, True):
        class CamCfg(CameraSensorConfig):
            name = "rgbd_camera"
            local_lookat_positions = [[0.7, 0., 0.7], [0., 0., 0.1]]
            image_normalization = normalize

            class camera_props(CameraSensorConfig.camera_props):
                width = w
                height = h
                near_plane = 0.1
                far_plane = 3

        fake_isaacgym.reset_gym()
        cfg = PriorStageEnvConfig()
        cfg.num_envs, cfg.device = n, "cuda:0"
        env = VisionAbbPushBox(cfg, camera_cfg=CamCfg())
        env.isg_env.refresh_sensors()
        env.isg_env.refresh_sensors()
        assert env.isg_env.sim.camera_frame == meta["frame"]
        cam, tag = env.camera, "norm" if normalize else "raw"
        for k, buf in (("color", cam.color_buf), ("depth", cam.depth_buf), ("seg", cam.segmentation_buf),
                       ("flow", cam.optical_flow_buf)):
            assert _same(buf.cpu(), z[f"{tag}/{k}"]), (k, normalize)


def test_camera_gather_argument_errors():
    from shifu_b200 import hotpath, _native as nv
    k = hotpath.EnvKernels("cuda:0", 2)
    img = [torch.zeros(3, 5, device="cuda") for _ in range(2)]
    tab = torch.tensor([t.data_ptr() for t in img], dtype=torch.int64).cuda()
    with pytest.raises(nv.ShifuNativeError):                 # 15 pixels: not a multiple of 8
        k.camera_gather(height=3, width=5, normalize_color=False, depth=(tab, torch.zeros(2, 3, 5, device="cuda")))
    with pytest.raises(ValueError):
        k.camera_gather(height=3, width=5, normalize_color=False, depth=(tab.int(), torch.zeros(2, 3, 5, device="cuda")))
