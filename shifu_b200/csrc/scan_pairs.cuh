// Stand-alone height scan, fast path of shifu_get_heights (row a5; TerrainGymEnv.get_heights,
// shifu/gym/isaac_gym.py:393-433, for tasks that keep their own Python hooks): the packed fp32x2 /
// one-rotation-per-point-pair arithmetic of the fused kernel's scan group (csrc/a1_fused_tma.cuh)
// without the rest of the step.  Preconditions (checked by the host, else get_heights_kernel runs):
// banded scan table, horizontal_scale == 0.1f (3-op constant division), measured-point grid symmetric
// about the base, no cell-index output requested.
//
// CTA = 12 warps; tile = 32 envs.  Warp w owns point pairs 32*(w%3).. (a lane = scan point pt and its
// mirror point 186 - pt, whose yaw rotation is the exact negation) for env pairs 4*(w/3) .. +3.
#pragma once
#include "a1_kernels.cuh"
#include "f32x2.cuh"

namespace shifu {

constexpr int SP_THREADS = 384;

__global__ void __launch_bounds__(SP_THREADS, 4)
get_heights_pairs_kernel(const __grid_constant__ A1K k, const float* __restrict__ root, float* __restrict__ mh) {
  // per env PAIR (e, e+1): (2z_e, 2z_e1, z_e, z_e1), (w_e, w_e1, x_e, x_e1), (y_e, y_e1); double-buffered
  __shared__ float4 sA[2][A1_TILE / 2], sB[2][A1_TILE / 2];
  __shared__ float2 sC[2][A1_TILE / 2];
  const int t = threadIdx.x, sw = t >> 5, lane = t & 31;
  const int raw = 32 * (sw % 3) + lane;
  const bool live = raw <= A1_POINTS / 2;                 // pairs 0..93; 93 is the centre, its own mirror
  const int pt = live ? raw : A1_POINTS / 2;
  const int mir = (A1_POINTS - 1) - 2 * pt;
  const float bx = k.px[pt % A1_NX], by = k.py[pt / A1_NX];
  const f2_t BX = pk(bx, bx), BY = pk(by, by), NBY = pk(-by, -by);
  const f2_t BORDER = pk(k.border, k.border), RCP = pk(k.hdiv.r, k.hdiv.r), NEGD = pk(-k.hdiv.d, -k.hdiv.d);
  const f2_t NZ = pk(k.neg_zero, k.neg_zero), DENORM = pk(__int_as_float(1), __int_as_float(1));
  const int max_px = k.trows - 1, max_py = k.tcols - 1;
  const unsigned c1 = (unsigned)(k.band_w - 1);
  const short* __restrict__ table = k.table;
  const int tiles = (k.n + A1_TILE - 1) / A1_TILE;

  auto publish = [&](int tile, int buf) {                 // threads 0..31: yaw normalisation of the tile's envs
    if (t < A1_TILE) {
      const int e = tile * A1_TILE + t;
      ScanEnv ev = {0.0f, 0.0f, 1.0f, 0.0f, 0.0f};
      if (e < k.n) ev = make_scan_env(root + ((long long)e * k.root_stride + k.root_offset) * 13);
      const int q = t >> 1, sl = t & 1;
      float* a = reinterpret_cast<float*>(&sA[buf][q]);
      float* b = reinterpret_cast<float*>(&sB[buf][q]);
      float* c = reinterpret_cast<float*>(&sC[buf][q]);
      a[sl] = ev.z2; a[2 + sl] = ev.z;
      b[sl] = ev.w; b[2 + sl] = ev.x;
      c[sl] = ev.y;
    }
  };
  // cell index of a packed pair of positions: the 3-op constant division, .long() + clip as a
  // round-toward-zero multiply by 2^-149 (the denormal's bits are the truncated integer), banded offset
  auto cell_pair = [&](f2_t ax, f2_t ay, unsigned& i0, unsigned& i1) {
    const f2_t qx = mul2(ax, RCP), qy = mul2(ay, RCP);
    const f2_t fx = fma2(fma2(NEGD, qx, ax), RCP, qx), fy = fma2(fma2(NEGD, qy, ay), RCP, qy);
    int ix0, ix1, iy0, iy1;
    upk_i(mulrz2(fx, DENORM), ix0, ix1);
    upk_i(mulrz2(fy, DENORM), iy0, iy1);
    const unsigned px0 = (unsigned)__vimin_s32_relu(ix0, max_px), px1 = (unsigned)__vimin_s32_relu(ix1, max_px);
    const unsigned py0 = (unsigned)__vimin_s32_relu(iy0, max_py), py1 = (unsigned)__vimin_s32_relu(iy1, max_py);
    i0 = ((px0 & ~7u) * c1 + px0) + (py0 << 3);
    i1 = ((px1 & ~7u) * c1 + px1) + (py1 << 3);
  };

  int buf = 0;
  if ((int)blockIdx.x < tiles) publish(blockIdx.x, 0);
  __syncthreads();
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < tiles) publish(next, buf ^ 1);             // overlaps this tile's arithmetic
    const long long e0 = (long long)tile * A1_TILE;
#pragma unroll
    for (int it = 0; it < 2; ++it) {                      // two items of 2 env pairs x (point, mirror point)
      const int q0 = 4 * (sw / 3) + 2 * it;
      unsigned idx[8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 a = sA[buf][q0 + u], bq = sB[buf][q0 + u];
        const float2 cq = sC[buf][q0 + u];
        const f2_t Z2 = pk(a.x, a.y), Z = pk(a.z, a.w), W = pk(bq.x, bq.y), X = pk(bq.z, bq.w), Y = pk(cq.x, cq.y);
        // quat_apply_yaw (shifu/utils/terrain.py:202-206); products that feed an add are fma(a, b, -0):
        // exact RN(a*b) that ptxas cannot contract with the add (see a1_fused_tma.cuh)
        const f2_t tx = mul2(Z2, NBY), ty = mul2(Z2, BX);
        const f2_t rx = sub2(add2(BX, fma2(W, tx, NZ)), fma2(Z, ty, NZ));
        const f2_t ry = add2(add2(BY, fma2(W, ty, NZ)), fma2(Z, tx, NZ));
        cell_pair(add2(add2(rx, X), BORDER), add2(add2(ry, Y), BORDER), idx[4 * u], idx[4 * u + 1]);
        cell_pair(add2(sub2(X, rx), BORDER), add2(sub2(Y, ry), BORDER), idx[4 * u + 2], idx[4 * u + 3]);
      }
      int h[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) h[u] = __ldg(table + idx[u]);
      if (live) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const long long e = e0 + 2 * (q0 + u) + r;
              if (e < k.n)
                __stcs(mh + e * A1_POINTS + pt + (m ? mir : 0), mul_rn((float)h[4 * u + 2 * m + r], k.vscale));   // :433
            }
          }
        }
      }
    }
    __syncthreads();                                      // next tile's env pairs are published / this tile's are dead
  }
}

}  // namespace shifu
