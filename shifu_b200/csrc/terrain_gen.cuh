// Row N3: the height map of Terrain (shifu/utils/terrain.py:42-173) rasterised on the
// device.  One-time initialisation: the map used to be assembled tile by tile in host numpy and copied
// up; here every tile is described by a small record (+ the few random parameters its generator
// drew, kept on the host so that the map is bit-identical to the host generator under the same
// numpy seed) and ONE launch writes the int16 map, a second one the spawn origins.
//
//   tile (i, j) of the curriculum lands at rows border + i*L .., cols border + j*W ..   (terrain.py:154-173)
//   origin(i, j) = ((i + .5) * env_length, (j + .5) * env_width, max(centre 2 m window) * vertical_scale)
//
// Sub-terrain kinds follow the generators Terrain.make_terrain calls (terrain.py:106-152): pyramid
// slope, slope + coarse uniform noise (bilinear up-sampled), pyramid stairs, discrete obstacles,
// stepping stones, gap and pit (terrain.py:176-198).  All arithmetic that numpy does in float64 is
// done with __dmul_rn / __dadd_rn in the same order (no contraction), conversions truncate like
// ndarray.astype.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/shifu_b200.h"

namespace shifu {

__device__ __forceinline__ double pyr_factor(int c, int x) {        // (c - |c - x|) / max(c, 1)
  const int a = c - x;
  return __ddiv_rn((double)(c - (a < 0 ? -a : a)), (double)(c > 1 ? c : 1));
}

__device__ __forceinline__ double pyramid_height(int peak, int cx, int cy, int x, int y) {
  return __dmul_rn(__dmul_rn((double)peak, pyr_factor(cx, x)), pyr_factor(cy, y));    // (peak * fx) * fy
}

__device__ __forceinline__ short pyramid_px(const ShifuTerrainTile& t, int W, int L, int x, int y) {
  const int cx = W / 2, cy = L / 2;
  const int half = t.p[1];
  const int x1 = max(cx - half, 0), y1 = max(cy - half, 0);
  const double edge = pyramid_height(t.p[0], cx, cy, x1, y1);
  const double lo = fmin(edge, 0.0), hi = fmax(edge, 0.0);
  const double h = fmin(fmax(pyramid_height(t.p[0], cx, cy, x, y), lo), hi);
  return (short)(int)h;                                            // astype(int16): toward zero
}

__global__ void __launch_bounds__(256)
terrain_raster_kernel(const ShifuTerrainTile* __restrict__ tiles, const double* __restrict__ table,
                      int W, int L, int border, int tot_cols, short* __restrict__ map) {
  const ShifuTerrainTile t = tiles[blockIdx.x];
  const int x0 = border + t.i * L, y0 = border + t.j * W;
  for (int px = blockIdx.y * blockDim.x + threadIdx.x; px < W * L; px += gridDim.y * blockDim.x) {
    const int x = px / L, y = px - x * L;                           // tile array is (width, length)
    short h = 0;
    switch (t.kind) {
      case SHIFU_TERRAIN_PYRAMID:
        h = pyramid_px(t, W, L, x, y);
        break;
      case SHIFU_TERRAIN_PYRAMID_NOISE: {                           // + rint(bilinear(coarse)), int16 wrap-add
        const int nx = t.p[2], ny = t.p[3];
        const double* c = table + t.table_off;
        const double sx = __ddiv_rn((double)(nx - 1), (double)(W - 1)), sy = __ddiv_rn((double)(ny - 1), (double)(L - 1));
        const double gx = (x == W - 1) ? (double)(nx - 1) : __dmul_rn((double)x, sx);      // np.linspace
        const double gy = (y == L - 1) ? (double)(ny - 1) : __dmul_rn((double)y, sy);
        const int ix = min(max((int)floor(gx), 0), nx - 2), iy = min(max((int)floor(gy), 0), ny - 2);
        const double tx = __dsub_rn(gx, (double)ix), ty = __dsub_rn(gy, (double)iy);
        const double a = c[ix * ny + iy], b = c[(ix + 1) * ny + iy], cc = c[ix * ny + iy + 1], d = c[(ix + 1) * ny + iy + 1];
        const double omx = __dsub_rn(1.0, tx), omy = __dsub_rn(1.0, ty);
        const double left = __dadd_rn(__dmul_rn(a, omx), __dmul_rn(b, tx));
        const double right = __dadd_rn(__dmul_rn(cc, omx), __dmul_rn(d, tx));
        const double fine = __dadd_rn(__dmul_rn(left, omy), __dmul_rn(right, ty));
        h = (short)(pyramid_px(t, W, L, x, y) + (short)(int)rint(fine));
        break;
      }
      case SHIFU_TERRAIN_STAIRS: {                                  // ring k covers [k*sw, W - k*sw)
        const int sw = t.p[0], sh = t.p[1], plat = t.p[2];
        int rings = 0, a0 = 0, a1 = W, b0 = 0, b1 = L;
        while ((a1 - a0) > plat && (b1 - b0) > plat) { a0 += sw; a1 -= sw; b0 += sw; b1 -= sw; ++rings; }
        const int kx = min(x / sw, (W - 1 - x) / sw), ky = min(y / sw, (L - 1 - y) / sw);
        h = (short)(min(min(kx, ky), rings) * sh);
        break;
      }
      case SHIFU_TERRAIN_OBSTACLES: {                               // later rectangles overwrite earlier ones
        const double* r = table + t.table_off;
        for (int q = 0; q < t.p[0]; ++q) {
          const int sx = (int)r[5 * q], sy = (int)r[5 * q + 1], w = (int)r[5 * q + 2], l = (int)r[5 * q + 3];
          if (x >= sx && x < sx + w && y >= sy && y < sy + l) h = (short)(int)r[5 * q + 4];
        }
        const int plat = t.p[1], cx = W / 2, cy = L / 2;
        if (x >= cx - plat / 2 && x < cx + plat / 2 && y >= cy - plat / 2 && y < cy + plat / 2) h = 0;
        break;
      }
      case SHIFU_TERRAIN_STONES: {
        const int ss = t.p[0], sd = t.p[1], plat = t.p[2], pitch = ss + sd;
        const int gx = x / pitch, gy = y / pitch, ny = (L + pitch - 1) / pitch;
        h = (short)t.p[3];
        if (x - gx * pitch < ss && y - gy * pitch < ss) h = (short)(int)table[t.table_off + gx * ny + gy];
        const int cx = W / 2, cy = L / 2;
        if (x >= cx - plat / 2 && x < cx + plat / 2 && y >= cy - plat / 2 && y < cy + plat / 2) h = 0;
        break;
      }
      case SHIFU_TERRAIN_GAP: {                                     // terrain.py:176-187 (length, width axes)
        const int g = t.p[0], p = t.p[1];
        const int cx = L / 2, cy = W / 2, x1 = (L - p) / 2, y1 = (W - p) / 2, x2 = x1 + g, y2 = y1 + g;
        if (x >= cx - x2 && x < cx + x2 && y >= cy - y2 && y < cy + y2) h = -1000;
        if (x >= cx - x1 && x < cx + x1 && y >= cy - y1 && y < cy + y1) h = 0;
        break;
      }
      case SHIFU_TERRAIN_PIT: {                                     // terrain.py:190-198
        const int d = t.p[0], hw = t.p[1], cx = L / 2, cy = W / 2;
        if (x >= cx - hw && x < cx + hw && y >= cy - hw && y < cy + hw) h = (short)(-d);
        break;
      }
      default:
        break;
    }
    if (x < L && y < W)        // the map window of a tile is (length_px, width_px); tiles are square
      map[(long long)(x0 + x) * tot_cols + (y0 + y)] = h;
  }
}

// origin(i, j): max of the tile's centre window (terrain.py:166-171), one warp per tile
__global__ void __launch_bounds__(32)
terrain_origins_kernel(const ShifuTerrainTile* __restrict__ tiles, const short* __restrict__ map, int W, int L, int border,
                       int tot_cols, int n_cols, double env_length, double env_width, int wx1, int wx2, int wy1, int wy2,
                       double vertical_scale, double* __restrict__ origins) {
  const ShifuTerrainTile t = tiles[blockIdx.x];
  const int x0 = border + t.i * L, y0 = border + t.j * W;
  int best = -32768;
  const int nx = wx2 - wx1, ny = wy2 - wy1;
  for (int q = threadIdx.x; q < nx * ny; q += 32) {
    const int x = wx1 + q / ny, y = wy1 + q % ny;
    best = max(best, (int)map[(long long)(x0 + x) * tot_cols + (y0 + y)]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (threadIdx.x == 0) {
    double* o = origins + ((long long)t.i * n_cols + t.j) * 3;
    o[0] = __dmul_rn(__dadd_rn((double)t.i, 0.5), env_length);
    o[1] = __dmul_rn(__dadd_rn((double)t.j, 0.5), env_width);
    o[2] = __dmul_rn((double)best, vertical_scale);
  }
}

}  // namespace shifu
