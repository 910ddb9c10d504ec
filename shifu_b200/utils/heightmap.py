"""Height-map construction and the yaw-only rotation helper.

Interface mirror of ``shifu/utils/terrain.py`` (``Terrain`` 42-173, ``quat_apply_yaw`` 202-206):
``Terrain(cfg, num_robots)`` exposes ``env_origins (rows, cols, 3)``, ``heightsamples`` /
``height_field_raw`` (int16, ``tot_rows x tot_cols``), ``tot_rows``, ``tot_cols``, ``border``,
``env_length``, ``env_width``, ``vertices``, ``triangles``.  The map is one-time host
initialisation; the hot path only consumes the resulting int16 tensor (``shifu_b200/csrc``: the height
scan reads a banded min-of-3 copy of it).

Two builders: the host one below (numpy, tile by tile, works with whatever ``isaacgym.terrain_utils``
provides) and — ``Terrain(cfg, n, device="cuda:0")``; the default of ``TerrainGymEnv`` on a CUDA device, ``cfg.generator = "host"`` opts out — the device
rasteriser ``shifu_terrain_generate``: the host only draws each tile's few
random parameters (same numpy stream as the host builder), one launch writes the whole map and a
second one the spawn origins; bit-identical to the host builder over the stand-in generators.

Sub-terrains come from ``isaacgym.terrain_utils`` — the genuine package when installed, the
stand-in generators of ``shifu_b200.sim.synthetic_terrain`` otherwise.
"""
from __future__ import annotations

import numpy as np


def _terrain_utils():
    from isaacgym import terrain_utils
    return terrain_utils


class Terrain:
    def __init__(self, cfg, num_robots, device=None) -> None:
        self.cfg = cfg
        self.device_map = None          # int16 device tensor when built by shifu_terrain_generate
        if device is None and getattr(cfg, "generator", "host") == "device":
            device = getattr(cfg, "generator_device", "cuda:0")
        self._device = device
        self.num_robots = num_robots
        self.type = cfg.mesh_type
        if self.type in ("none", "plane"):
            return
        self.env_length = cfg.terrain_length
        self.env_width = cfg.terrain_width
        self.proportions = [np.sum(cfg.terrain_proportions[:i + 1]) for i in range(len(cfg.terrain_proportions))]
        self.cfg.num_sub_terrains = cfg.num_rows * cfg.num_cols
        self.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))

        hs = cfg.horizontal_scale
        self.width_per_env_pixels = int(self.env_width / hs)
        self.length_per_env_pixels = int(self.env_length / hs)
        self.border = int(cfg.border_size / hs)
        self.tot_cols = int(cfg.num_cols * self.width_per_env_pixels) + 2 * self.border
        self.tot_rows = int(cfg.num_rows * self.length_per_env_pixels) + 2 * self.border
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)

        if cfg.curriculum and self._device is not None and self._device_capable():
            self._build_on_device()
        elif cfg.curriculum:
            order = [(i, j, j / cfg.num_cols + 0.001, i / cfg.num_rows)
                     for j in range(cfg.num_cols) for i in range(cfg.num_rows)]
            for i, j, choice, difficulty in order:
                self._place(self.make_terrain(choice, difficulty), i, j)
        elif cfg.selected:
            self._selected()
        else:
            for k in range(self.cfg.num_sub_terrains):
                i, j = np.unravel_index(k, (cfg.num_rows, cfg.num_cols))
                choice = np.random.uniform(0, 1)
                difficulty = np.random.choice([0.5, 0.75, 0.9])
                self._place(self.make_terrain(choice, difficulty), i, j)

        self.heightsamples = self.height_field_raw
        if self.type == "trimesh":
            self.vertices, self.triangles = _terrain_utils().convert_heightfield_to_trimesh(
                self.height_field_raw, cfg.horizontal_scale, cfg.vertical_scale, cfg.slope_treshold)

    # -- device builder (row N3) ----------------------------------------------------------
    @staticmethod
    def _device_capable() -> bool:
        """The device rasteriser knows the stand-in generators; the genuine (closed) Isaac Gym ones stay
        on the host path."""
        return hasattr(_terrain_utils(), "pyramid_params")

    def tile_record(self, choice, difficulty):
        """(kind, p[4], table) of Terrain.make_terrain(choice, difficulty) — terrain.py:106-152."""
        from shifu_b200 import _native as nv
        tu = _terrain_utils()
        tile = self._new_tile()
        pr = self.proportions
        slope = difficulty * 0.4
        step_h = 0.05 + 0.18 * difficulty
        none = np.zeros(0)
        if choice < pr[0]:
            return nv.TERRAIN_PYRAMID, tu.pyramid_params(tile, -slope if choice < pr[0] / 2 else slope, 3.), none
        if choice < pr[1]:
            p = tu.pyramid_params(tile, slope, 3.)
            nx, ny, coarse = tu.random_uniform_params(tile, -0.05, 0.05, 0.005, 0.2)
            return nv.TERRAIN_PYRAMID_NOISE, p + [nx, ny], coarse
        if choice < pr[3]:
            return nv.TERRAIN_STAIRS, tu.stairs_params(tile, 0.31, -step_h if choice < pr[2] else step_h, 3.), none
        if choice < pr[4]:
            p, rects = tu.obstacles_params(tile, 0.05 + difficulty * 0.2, 1., 2., 20, 3.)
            return nv.TERRAIN_OBSTACLES, p, rects
        if len(pr) > 5 and choice < pr[5]:
            p, heights = tu.stones_params(tile, 1.5 * (1.05 - difficulty), 0.05 if difficulty == 0 else 0.1, 0., 4.)
            return nv.TERRAIN_STONES, p, heights
        hs, vs = tile.horizontal_scale, tile.vertical_scale
        if len(pr) > 6 and choice < pr[6]:
            return nv.TERRAIN_GAP, [int(1. * difficulty / hs), int(3. / hs)], none
        return nv.TERRAIN_PIT, [int(1. * difficulty / vs), int(4. / hs / 2)], none

    def _build_on_device(self):
        import ctypes as C
        import torch
        from shifu_b200 import _native as nv
        cfg = self.cfg
        tiles = (nv.TerrainTile * (cfg.num_rows * cfg.num_cols))()
        table, k = [], 0
        for j in range(cfg.num_cols):                       # the curriculum's draw order (terrain.py:93-104)
            for i in range(cfg.num_rows):
                kind, p, tab = self.tile_record(j / cfg.num_cols + 0.001, i / cfg.num_rows)
                t = tiles[k]
                t.kind, t.i, t.j, t.table_off = kind, i, j, sum(len(x) for x in table)
                for q, v in enumerate(p):
                    t.p[q] = int(v)
                table.append(np.asarray(tab, dtype=np.float64))
                k += 1
        flat = np.ascontiguousarray(np.concatenate(table)) if table else np.zeros(0)
        desc = nv.TerrainDesc(cfg.num_rows, cfg.num_cols, self.width_per_env_pixels, self.length_per_env_pixels,
                              self.border, float(self.env_length), float(self.env_width), float(cfg.horizontal_scale),
                              float(cfg.vertical_scale))
        dev = torch.device(self._device)
        lib = nv.load()
        ctx = C.c_void_p()
        nv.check(lib.shifu_ctx_create_util(dev.index or 0, 1, C.byref(ctx)))
        try:
            self.device_map = torch.empty(self.tot_rows, self.tot_cols, dtype=torch.int16, device=dev)
            origins = torch.empty(cfg.num_rows, cfg.num_cols, 3, dtype=torch.float64, device=dev)
            nv.check(lib.shifu_terrain_generate(ctx, C.byref(desc), tiles, len(tiles),
                                                flat.ctypes.data_as(C.POINTER(C.c_double)), int(flat.size),
                                                nv.ptr(self.device_map), nv.ptr(origins),
                                                torch.cuda.current_stream(dev).cuda_stream))
        finally:
            lib.shifu_ctx_destroy(ctx)
        # host mirrors for the callers that hand the map to the simulator (add_heightfield / trimesh)
        self.height_field_raw = self.device_map.cpu().numpy()
        self.env_origins = origins.cpu().numpy()

    # -- sub-terrain selection (shifu/utils/terrain.py:106-152) --------------------------
    def _new_tile(self):
        return _terrain_utils().SubTerrain("terrain", width=self.width_per_env_pixels,
                                           length=self.width_per_env_pixels,
                                           vertical_scale=self.cfg.vertical_scale,
                                           horizontal_scale=self.cfg.horizontal_scale)

    def make_terrain(self, choice, difficulty):
        tu = _terrain_utils()
        tile = self._new_tile()
        pr = self.proportions
        slope = difficulty * 0.4
        step_h = 0.05 + 0.18 * difficulty
        if choice < pr[0]:
            tu.pyramid_sloped_terrain(tile, slope=-slope if choice < pr[0] / 2 else slope, platform_size=3.)
        elif choice < pr[1]:
            tu.pyramid_sloped_terrain(tile, slope=slope, platform_size=3.)
            tu.random_uniform_terrain(tile, min_height=-0.05, max_height=0.05, step=0.005, downsampled_scale=0.2)
        elif choice < pr[3]:
            tu.pyramid_stairs_terrain(tile, step_width=0.31,
                                      step_height=-step_h if choice < pr[2] else step_h, platform_size=3.)
        elif choice < pr[4]:
            tu.discrete_obstacles_terrain(tile, 0.05 + difficulty * 0.2, 1., 2., 20, platform_size=3.)
        elif len(pr) > 5 and choice < pr[5]:
            tu.stepping_stones_terrain(tile, stone_size=1.5 * (1.05 - difficulty),
                                       stone_distance=0.05 if difficulty == 0 else 0.1, max_height=0.,
                                       platform_size=4.)
        elif len(pr) > 6 and choice < pr[6]:
            gap_terrain(tile, gap_size=1. * difficulty, platform_size=3.)
        else:
            pit_terrain(tile, depth=1. * difficulty, platform_size=4.)
        return tile

    def _selected(self):
        kind = self.cfg.terrain_kwargs.pop('type')
        fn = getattr(_terrain_utils(), kind.split('.')[-1])
        for k in range(self.cfg.num_sub_terrains):
            i, j = np.unravel_index(k, (self.cfg.num_rows, self.cfg.num_cols))
            tile = self._new_tile()
            fn(tile, **self.cfg.terrain_kwargs)
            self._place(tile, i, j)

    def _place(self, tile, i, j):
        """Copy a tile into the map and record the spawn origin of (level i, type j)
        (shifu/utils/terrain.py:154-173)."""
        L, W = self.length_per_env_pixels, self.width_per_env_pixels
        x0, y0 = self.border + i * L, self.border + j * W
        self.height_field_raw[x0:x0 + L, y0:y0 + W] = tile.height_field_raw
        hs = tile.horizontal_scale
        x1, x2 = int((self.env_length / 2. - 1) / hs), int((self.env_length / 2. + 1) / hs)
        y1, y2 = int((self.env_width / 2. - 1) / hs), int((self.env_width / 2. + 1) / hs)
        z = np.max(tile.height_field_raw[x1:x2, y1:y2]) * tile.vertical_scale
        self.env_origins[i, j] = [(i + 0.5) * self.env_length, (j + 0.5) * self.env_width, z]

    add_terrain_to_map = _place


def gap_terrain(terrain, gap_size, platform_size=1.):
    """A square moat of width ``gap_size`` around a centre platform."""
    g = int(gap_size / terrain.horizontal_scale)
    p = int(platform_size / terrain.horizontal_scale)
    cx, cy = terrain.length // 2, terrain.width // 2
    x1, y1 = (terrain.length - p) // 2, (terrain.width - p) // 2
    x2, y2 = x1 + g, y1 + g
    terrain.height_field_raw[cx - x2:cx + x2, cy - y2:cy + y2] = -1000
    terrain.height_field_raw[cx - x1:cx + x1, cy - y1:cy + y1] = 0


def pit_terrain(terrain, depth, platform_size=1.):
    d = int(depth / terrain.vertical_scale)
    h = int(platform_size / terrain.horizontal_scale / 2)
    cx, cy = terrain.length // 2, terrain.width // 2
    terrain.height_field_raw[cx - h:cx + h, cy - h:cy + h] = -d


def quat_apply_yaw(quat, vec):
    """Rotate ``vec`` by the yaw component of ``quat`` (torch tensors; utility for user code —
    the hot path does this inside the height-scan kernel)."""
    from isaacgym.torch_utils import quat_apply, normalize
    q = quat.clone().view(-1, 4)
    q[:, :2] = 0.
    return quat_apply(normalize(q), vec)


def wrap_to_pi(angles):
    angles %= 2 * np.pi
    angles -= 2 * np.pi * (angles > np.pi)
    return angles
