// Row N4: CameraSensor.refresh_image_tensors (shifu/units/sensors.py:165-188) as
// one launch: blockIdx.y = env, the threads of a row of CTAs stream that env's images in 16-byte
// pieces.  Pure HBM streaming (36 B per pixel with all four image types and normalised colour).
#pragma once
#include "exact_math.cuh"
#include "../../include/shifu_b200.h"

namespace shifu {

__global__ void __launch_bounds__(256)
camera_gather_kernel(const __grid_constant__ ShifuCameraGatherIO io, int n) {
  // x / 255 for the 256 possible channel values, each computed with the same IEEE division the
  // reference performs (normalize_color, image.py:12-15): a table look-up per channel afterwards
  __shared__ float lut[256];
  const bool norm = io.color_src != nullptr && io.normalize_color;
  if (norm) lut[threadIdx.x] = div_rn((float)threadIdx.x, 255.0f);      // blockDim.x == 256
  __syncthreads();
  const long long px = (long long)io.height * io.width;
  const long long quads = px / 4, oct = px / 8, tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (int e = blockIdx.y; e < n; e += gridDim.y) {
    const uchar4* cs = io.color_src ? static_cast<const uchar4*>(io.color_src[e]) : nullptr;
    const float* ds = io.depth_src ? static_cast<const float*>(io.depth_src[e]) : nullptr;
    const int* ss = io.seg_src ? static_cast<const int*>(io.seg_src[e]) : nullptr;
    const short* fs = io.flow_src ? static_cast<const short*>(io.flow_src[e]) : nullptr;
    float* cof = static_cast<float*>(io.color_out) + (long long)e * px * 3;       // normalised layout
    uchar4* cob = static_cast<uchar4*>(io.color_out) + (long long)e * px;          // verbatim layout
    float* dout = io.depth_out + (long long)e * px;
    int* sout = io.seg_out + (long long)e * px;
    short* fout = io.flow_out + (long long)e * px;
    // one pass: the 16-byte loads of every image type are issued together, then the stores
    for (long long i = tid; i < quads; i += nth) {
      uint4 cw = make_uint4(0, 0, 0, 0);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      int4 sg = make_int4(0, 0, 0, 0), fl = make_int4(0, 0, 0, 0);
      if (cs != nullptr) cw = __ldg(reinterpret_cast<const uint4*>(cs) + i);
      if (ds != nullptr) d = __ldg(reinterpret_cast<const float4*>(ds) + i);
      if (ss != nullptr) sg = __ldg(reinterpret_cast<const int4*>(ss) + i);
      if (fs != nullptr && i < oct) fl = __ldg(reinterpret_cast<const int4*>(fs) + i);
      if (cs != nullptr) {
        if (norm) {                                              // 4 pixels: 16 B in, 48 B out
          const unsigned v[4] = {cw.x, cw.y, cw.z, cw.w};
          float f[12];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            f[3 * p + 0] = lut[v[p] & 0xffu];
            f[3 * p + 1] = lut[(v[p] >> 8) & 0xffu];
            f[3 * p + 2] = lut[(v[p] >> 16) & 0xffu];
          }
          float4* o4 = reinterpret_cast<float4*>(cof) + 3 * i;
          __stcs(o4 + 0, make_float4(f[0], f[1], f[2], f[3]));
          __stcs(o4 + 1, make_float4(f[4], f[5], f[6], f[7]));
          __stcs(o4 + 2, make_float4(f[8], f[9], f[10], f[11]));
        } else {                                                 // sensors.py:171, verbatim RGBA
          __stcs(reinterpret_cast<uint4*>(cob) + i, cw);
        }
      }
      if (ds != nullptr)                                         // "Isaac gives negative depth map !"
        __stcs(reinterpret_cast<float4*>(dout) + i, make_float4(-d.x, -d.y, -d.z, -d.w));
      if (ss != nullptr) __stcs(reinterpret_cast<int4*>(sout) + i, sg);
      if (fs != nullptr && i < oct) __stcs(reinterpret_cast<int4*>(fout) + i, fl);
    }
    // ragged tails (the C-ABI currently requires px % 8 == 0, so these loops are empty)
    for (long long i = 4 * quads + tid; i < px; i += nth) {
      if (cs != nullptr) {
        const uchar4 c = cs[i];
        if (norm) { cof[3 * i + 0] = lut[c.x]; cof[3 * i + 1] = lut[c.y]; cof[3 * i + 2] = lut[c.z]; }
        else cob[i] = c;
      }
      if (ds != nullptr) dout[i] = -ds[i];
      if (ss != nullptr) sout[i] = ss[i];
    }
    if (fs != nullptr)
      for (long long i = 8 * oct + tid; i < px; i += nth) fout[i] = fs[i];
  }
}

}  // namespace shifu
