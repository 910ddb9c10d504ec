"""Height-field sub-terrain generators standing in for ``isaacgym.terrain_utils``.

The reference builds its int16 height map by calling Isaac Gym's terrain helpers
(``shifu/utils/terrain.py:37,106-152``).  Those helpers ship only inside the
Isaac Gym tarball, so this module offers generators with the same call
signatures and the same *kind* of output (an int16 ``height_field_raw`` in units
of ``vertical_scale``).  Only the map **content** depends on them; the hot path
consumes whatever int16 map it is given, so bit-compatibility
with NVIDIA's generators is neither needed nor claimed.

All randomness goes through ``numpy.random`` (legacy global state), like the
originals, so ``np.random.seed(k)`` before terrain construction pins the map.
"""
from __future__ import annotations

import numpy as np


class SubTerrain:
    def __init__(self, terrain_name="terrain", width=256, length=256, vertical_scale=1.0, horizontal_scale=1.0):
        self.terrain_name = terrain_name
        self.vertical_scale = vertical_scale
        self.horizontal_scale = horizontal_scale
        self.width = width
        self.length = length
        self.height_field_raw = np.zeros((self.width, self.length), dtype=np.int16)


def _centre_distance(terrain):
    """Chebyshev-like normalised distance from the tile border (0) to its centre (1)."""
    cx, cy = terrain.width // 2, terrain.length // 2
    x = np.arange(terrain.width).reshape(-1, 1)
    y = np.arange(terrain.length).reshape(1, -1)
    fx = (cx - np.abs(cx - x)) / max(cx, 1)
    fy = (cy - np.abs(cy - y)) / max(cy, 1)
    return fx, fy


def pyramid_sloped_terrain(terrain, slope=1.0, platform_size=1.0):
    fx, fy = _centre_distance(terrain)
    peak = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * (terrain.width / 2))
    hf = peak * fx * fy
    half = int(platform_size / terrain.horizontal_scale / 2)
    cx, cy = terrain.width // 2, terrain.length // 2
    x1, y1 = max(cx - half, 0), max(cy - half, 0)
    lo, hi = min(hf[x1, y1], 0), max(hf[x1, y1], 0)
    terrain.height_field_raw += np.clip(hf, lo, hi).astype(terrain.height_field_raw.dtype)
    return terrain


def random_uniform_terrain(terrain, min_height, max_height, step=1.0, downsampled_scale=None):
    if downsampled_scale is None:
        downsampled_scale = terrain.horizontal_scale
    lo = int(min_height / terrain.vertical_scale)
    hi = int(max_height / terrain.vertical_scale)
    st = max(int(step / terrain.vertical_scale), 1)
    levels = np.arange(lo, hi + st, st)
    nx = max(int(terrain.width * terrain.horizontal_scale / downsampled_scale), 2)
    ny = max(int(terrain.length * terrain.horizontal_scale / downsampled_scale), 2)
    coarse = np.random.choice(levels, (nx, ny)).astype(np.float64)
    # bilinear up-sampling of the coarse grid to the tile resolution
    gx = np.linspace(0, nx - 1, terrain.width)
    gy = np.linspace(0, ny - 1, terrain.length)
    x0 = np.floor(gx).astype(int).clip(0, nx - 2)
    y0 = np.floor(gy).astype(int).clip(0, ny - 2)
    tx = (gx - x0).reshape(-1, 1)
    ty = (gy - y0).reshape(1, -1)
    a = coarse[x0][:, y0]
    b = coarse[x0 + 1][:, y0]
    c = coarse[x0][:, y0 + 1]
    d = coarse[x0 + 1][:, y0 + 1]
    fine = (a * (1 - tx) + b * tx) * (1 - ty) + (c * (1 - tx) + d * tx) * ty
    terrain.height_field_raw += np.rint(fine).astype(np.int16)
    return terrain


def pyramid_stairs_terrain(terrain, step_width, step_height, platform_size=1.0):
    sw = max(int(step_width / terrain.horizontal_scale), 1)
    sh = int(step_height / terrain.vertical_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    x0, x1, y0, y1 = 0, terrain.width, 0, terrain.length
    h = 0
    while (x1 - x0) > plat and (y1 - y0) > plat:
        x0 += sw
        x1 -= sw
        y0 += sw
        y1 -= sw
        h += sh
        terrain.height_field_raw[x0:x1, y0:y1] = h
    return terrain


def discrete_obstacles_terrain(terrain, max_height, min_size, max_size, num_rects, platform_size=1.0):
    mh = int(max_height / terrain.vertical_scale)
    lo = int(min_size / terrain.horizontal_scale)
    hi = int(max_size / terrain.horizontal_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    heights = [-mh, -mh // 2, mh // 2, mh]
    for _ in range(num_rects):
        w = np.random.choice(np.arange(lo, hi, 4))
        l = np.random.choice(np.arange(lo, hi, 4))
        sx = np.random.choice(np.arange(0, terrain.width - w, 4))
        sy = np.random.choice(np.arange(0, terrain.length - l, 4))
        terrain.height_field_raw[sx:sx + w, sy:sy + l] = np.random.choice(heights)
    cx, cy = terrain.width // 2, terrain.length // 2
    terrain.height_field_raw[cx - plat // 2:cx + plat // 2, cy - plat // 2:cy + plat // 2] = 0
    return terrain


def stepping_stones_terrain(terrain, stone_size, stone_distance, max_height, platform_size=1.0, depth=-10):
    ss = max(int(stone_size / terrain.horizontal_scale), 1)
    sd = max(int(stone_distance / terrain.horizontal_scale), 1)
    mh = int(max_height / terrain.vertical_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    terrain.height_field_raw[:, :] = int(depth / terrain.vertical_scale)
    for sx in range(0, terrain.width, ss + sd):
        for sy in range(0, terrain.length, ss + sd):
            terrain.height_field_raw[sx:sx + ss, sy:sy + ss] = np.random.randint(-mh - 1, mh + 1) if mh else 0
    cx, cy = terrain.width // 2, terrain.length // 2
    terrain.height_field_raw[cx - plat // 2:cx + plat // 2, cy - plat // 2:cy + plat // 2] = 0
    return terrain


# ---------------------------------------------------------------------------------------------
# Parameter-only twins of the generators above, for the device rasteriser (shifu_terrain_generate,
# SURVEY.md 8f row N3): they consume numpy.random in EXACTLY the order of their twin, but return the
# tile record + the drawn numbers instead of touching pixels — so a device-built map is bit-identical
# to the host-built one under the same seed.  kind codes = enum ShifuTerrainKind.
# ---------------------------------------------------------------------------------------------

def pyramid_params(terrain, slope, platform_size):
    peak = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * (terrain.width / 2))
    return [peak, int(platform_size / terrain.horizontal_scale / 2)]


def random_uniform_params(terrain, min_height, max_height, step=1.0, downsampled_scale=None):
    if downsampled_scale is None:
        downsampled_scale = terrain.horizontal_scale
    lo = int(min_height / terrain.vertical_scale)
    hi = int(max_height / terrain.vertical_scale)
    st = max(int(step / terrain.vertical_scale), 1)
    levels = np.arange(lo, hi + st, st)
    nx = max(int(terrain.width * terrain.horizontal_scale / downsampled_scale), 2)
    ny = max(int(terrain.length * terrain.horizontal_scale / downsampled_scale), 2)
    coarse = np.random.choice(levels, (nx, ny)).astype(np.float64)
    return nx, ny, coarse.reshape(-1)


def stairs_params(terrain, step_width, step_height, platform_size):
    return [max(int(step_width / terrain.horizontal_scale), 1), int(step_height / terrain.vertical_scale),
            int(platform_size / terrain.horizontal_scale)]


def obstacles_params(terrain, max_height, min_size, max_size, num_rects, platform_size):
    mh = int(max_height / terrain.vertical_scale)
    lo = int(min_size / terrain.horizontal_scale)
    hi = int(max_size / terrain.horizontal_scale)
    heights = [-mh, -mh // 2, mh // 2, mh]
    rects = []
    for _ in range(num_rects):
        w = np.random.choice(np.arange(lo, hi, 4))
        l = np.random.choice(np.arange(lo, hi, 4))
        sx = np.random.choice(np.arange(0, terrain.width - w, 4))
        sy = np.random.choice(np.arange(0, terrain.length - l, 4))
        rects += [sx, sy, w, l, np.random.choice(heights)]
    return [num_rects, int(platform_size / terrain.horizontal_scale)], np.asarray(rects, dtype=np.float64)


def stones_params(terrain, stone_size, stone_distance, max_height, platform_size, depth=-10):
    ss = max(int(stone_size / terrain.horizontal_scale), 1)
    sd = max(int(stone_distance / terrain.horizontal_scale), 1)
    mh = int(max_height / terrain.vertical_scale)
    heights = [np.random.randint(-mh - 1, mh + 1) if mh else 0
               for _ in range(0, terrain.width, ss + sd) for _ in range(0, terrain.length, ss + sd)]
    return [ss, sd, int(platform_size / terrain.horizontal_scale), int(depth / terrain.vertical_scale)], \
        np.asarray(heights, dtype=np.float64)


def convert_heightfield_to_trimesh(height_field_raw, horizontal_scale, vertical_scale, slope_threshold=None):
    """The fake simulator never collides against the mesh; hand back an empty one."""
    return np.zeros((0, 3), dtype=np.float32), np.zeros((0, 3), dtype=np.uint32)
