"""Row N3: the Terrain height map rasterised on the device (shifu_terrain_generate)
is bit-identical to the host builder (shifu/utils/terrain.py:42-198 over the stand-in sub-terrain
generators) under the same numpy seed, and reproduces the map hashes committed with the fixtures."""
import hashlib

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _cfg(**over):
    from shifu_b200.sim import fake_isaacgym
    fake_isaacgym.install("cuda:0")
    from shifu_b200.configs import TerrainEnvConfig
    tc = TerrainEnvConfig().terrain
    for k, v in over.items():
        setattr(tc, k, v)
    return tc


def _both(seed=0, **over):
    from shifu_b200.utils.heightmap import Terrain
    np.random.seed(seed)
    host = Terrain(_cfg(**over), 64)
    np.random.seed(seed)
    dev = Terrain(_cfg(**over), 64, device="cuda:0")
    assert dev.device_map is not None and dev.device_map.is_cuda
    return host, dev


@pytest.mark.parametrize("fixture,over", [("a1_fullmap", {}), ("a1_small", None)])
def test_device_map_matches_host_and_committed_hash(fixture, over):
    z, meta = util.load_golden(fixture)
    over = meta["terrain"] if over is None else over
    host, dev = _both(meta["map_seed"], **over)
    got = dev.device_map.cpu().numpy()
    assert got.dtype == np.int16 and got.shape == tuple(meta["map_shape"])
    assert np.array_equal(got, host.heightsamples)
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == meta["map_sha256"]
    assert np.array_equal(dev.env_origins, host.env_origins)                 # float64, bit for bit
    np.testing.assert_array_equal(dev.env_origins.astype(np.float32), z["terrain_origins"])


def test_all_sub_terrain_kinds():
    """Seven proportions reach stepping stones, gap and pit as well (terrain.py:137-152)."""
    host, dev = _both(3, terrain_proportions=[0.1, 0.1, 0.2, 0.15, 0.15, 0.1, 0.1, 0.1], num_rows=4, num_cols=16,
                      border_size=3)
    assert np.array_equal(dev.device_map.cpu().numpy(), host.heightsamples)
    assert np.array_equal(dev.env_origins, host.env_origins)
    assert int(host.heightsamples.min()) <= -1000                            # the gap / stone trenches are there


def test_env_builds_its_map_on_the_device():
    from shifu_b200.sim import fake_isaacgym
    fake_isaacgym.install("cuda:0")
    fake_isaacgym.reset_gym()
    fake_isaacgym.set_default_device("cuda:0")
    from shifu_b200.tasks.a1_walking import A1Conditional, A1EnvConfig
    z, meta = util.load_golden("a1_small")
    cfg = A1EnvConfig()
    cfg.num_envs, cfg.device = meta["n"], "cuda:0"
    for k, v in meta["terrain"].items():
        setattr(cfg.terrain, k, v)
    cfg.terrain.generator = "device"
    np.random.seed(meta["map_seed"])
    env = A1Conditional(cfg)
    assert env.isg_env.terrain.device_map is env.isg_env.height_samples
    assert np.array_equal(env.isg_env.height_samples.cpu().numpy(), z["height_samples"])
