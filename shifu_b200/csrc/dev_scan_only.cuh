// DEV ONLY (-DV3_DEV, never shipped): the 187-point height scan as a stand-alone kernel, used to
// measure the scan's own floor outside the fused pipeline (tools/scan_only.py).
// mode bits: 1 = no obs stores, 2 = no table gathers, 4 = no index arithmetic, 8 = stage obs through
// shared memory + bulk store (TMA) instead of st.global.
#pragma once
#include "a1_fused_tma.cuh"

namespace shifu {

struct DevScanSmem {
  float4 sA[A1_TILE / 2], sB[A1_TILE / 2], sC[A1_TILE / 2];
};

template <int THREADS_PER_CTA, int NP>
__global__ void __launch_bounds__(THREADS_PER_CTA)
dev_scan_only_kernel(const __grid_constant__ A1K k, const float* __restrict__ root, float* __restrict__ obs,
                     int num_tiles, int mode) {
  constexpr int GROUPS = THREADS_PER_CTA / 192;
  __shared__ __align__(16) DevScanSmem sm[GROUPS];
  const int g = threadIdx.x / 192, p = threadIdx.x % 192;
  DevScanSmem& s = sm[g];
  const float bx = k.px[p % A1_NX], by = k.py[(p / A1_NX) % A1_NY];
  const float hclip = fminf(k.h_clip, k.clip_obs);
  const unsigned max_px = (unsigned)(k.trows - 1), max_py = (unsigned)(k.tcols - 1);
  const unsigned c1 = (unsigned)(k.band_w - 1);
  const short* __restrict__ table = k.table;
  const f2_t BX = pk(bx, bx), BY = pk(by, by), NBY = pk(-by, -by), BORDER = pk(k.border, k.border);
  const f2_t RCP = pk(k.hdiv.r, k.hdiv.r), NEGD = pk(-k.hdiv.d, -k.hdiv.d), VS = pk(k.vscale, k.vscale);
  const f2_t NZ = pk(k.neg_zero, k.neg_zero);
  const f2_t DENORM = pk(__int_as_float(1), __int_as_float(1));
  for (int tile = blockIdx.x * GROUPS + g; tile < num_tiles; tile += gridDim.x * GROUPS) {
    const long long e0 = (long long)tile * A1_TILE;
    pipe::named_barrier(1 + g, 192);
    if (p < A1_TILE) {
      const float* r = root + (e0 + p) * 13;
      const ScanEnv ev = make_scan_env(r);
      const int q = p >> 1, sl = p & 1;
      float* a = reinterpret_cast<float*>(&s.sA[q]);
      float* bb = reinterpret_cast<float*>(&s.sB[q]);
      float* cc = reinterpret_cast<float*>(&s.sC[q]);
      a[sl] = ev.z2; a[2 + sl] = ev.z;
      bb[sl] = ev.w; bb[2 + sl] = ev.x;
      cc[sl] = ev.y; cc[2 + sl] = sub_rn(r[2], k.h_off);
    }
    pipe::named_barrier(1 + g, 192);
    if (p < A1_POINTS) {
      float* ob = obs + e0 * A1_OBS + A1_HEAD + p;
      auto index_batch = [&](int q0, unsigned (&idx)[2 * NP]) {
#pragma unroll
        for (int u = 0; u < NP; ++u) {
          if (mode & 4) { idx[2 * u] = (unsigned)(p + q0 + u); idx[2 * u + 1] = (unsigned)(p + q0 + u + 1); continue; }
          const float4 a = s.sA[q0 + u], bq = s.sB[q0 + u];
          const float2 cq = *reinterpret_cast<const float2*>(&s.sC[q0 + u]);
          const f2_t Z2 = pk(a.x, a.y), Z = pk(a.z, a.w), W = pk(bq.x, bq.y), X = pk(bq.z, bq.w);
          const f2_t Y = pk(cq.x, cq.y);
          const f2_t tx = mul2(Z2, NBY), ty = mul2(Z2, BX);
          f2_t rx, ry;
          if (mode & 16) { rx = sub2(add2(BX, tx), ty); ry = add2(add2(BY, ty), tx); }       // what-if: no products
          else {
            rx = sub2(add2(BX, fma2(W, tx, NZ)), fma2(Z, ty, NZ));
            ry = add2(add2(BY, fma2(W, ty, NZ)), fma2(Z, tx, NZ));
          }
          const f2_t ax = add2(add2(rx, X), BORDER), ay = add2(add2(ry, Y), BORDER);
          f2_t fx, fy;
          if (mode & 32) { fx = mul2(ax, RCP); fy = mul2(ay, RCP); }                          // what-if: 1-op division
          else {
            const f2_t qx = mul2(ax, RCP), qy = mul2(ay, RCP);
            fx = fma2(fma2(NEGD, qx, ax), RCP, qx); fy = fma2(fma2(NEGD, qy, ay), RCP, qy);
          }
          int ix0, ix1, iy0, iy1;
          upk_i(mulrz2(fx, DENORM), ix0, ix1);
          upk_i(mulrz2(fy, DENORM), iy0, iy1);
          const unsigned px0 = (unsigned)__vimin_s32_relu(ix0, (int)max_px), px1 = (unsigned)__vimin_s32_relu(ix1, (int)max_px);
          const unsigned py0 = (unsigned)__vimin_s32_relu(iy0, (int)max_py), py1 = (unsigned)__vimin_s32_relu(iy1, (int)max_py);
          if (mode & 64) { idx[2 * u] = px0 + py0; idx[2 * u + 1] = px1 + py1; }             // what-if: 1-op index
          else {
            idx[2 * u] = ((px0 & ~7u) * c1 + px0) + (py0 << 3);
            idx[2 * u + 1] = ((px1 & ~7u) * c1 + px1) + (py1 << 3);
          }
        }
      };
      auto load_batch = [&](const unsigned (&idx)[2 * NP], int (&h)[2 * NP]) {
#pragma unroll
        for (int u = 0; u < 2 * NP; ++u) h[u] = (mode & 2) ? (int)(idx[u] & 1023u) : (int)__ldg(table + idx[u]);
      };
      auto store_batch = [&](int q0, const int (&h)[2 * NP]) {
#pragma unroll
        for (int u = 0; u < NP; ++u) {
          const float2 zb = *reinterpret_cast<const float2*>(&s.sC[q0 + u].z);
          const f2_t hg = fma2(pk((float)h[2 * u], (float)h[2 * u + 1]), VS, NZ);
          float v0, v1;
          upk(sub2(pk(zb.x, zb.y), hg), v0, v1);
          if (!(mode & 128)) { v0 = clampf(v0, -hclip, hclip); v1 = clampf(v1, -hclip, hclip); }   // what-if: no clamp
          if (mode & 1) {
            if (v0 == 123.456f || v1 == 123.456f) ob[0] = v0;     // keep the values alive
          } else {
            __stcs(ob + (2 * u) * A1_OBS, v0);
            __stcs(ob + (2 * u + 1) * A1_OBS, v1);
          }
        }
        ob += 2 * NP * A1_OBS;
      };
      unsigned idx[2 * NP];
      int h[2 * NP];
      index_batch(0, idx);
      load_batch(idx, h);
#pragma unroll 1
      for (int q0 = 0; q0 < A1_TILE / 2 - NP; q0 += NP) {
        index_batch(q0 + NP, idx);
        store_batch(q0, h);
        load_batch(idx, h);
      }
      store_batch(A1_TILE / 2 - NP, h);
    }
  }
}

}  // namespace shifu
