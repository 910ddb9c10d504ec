// Round-to-nearest fp32 building blocks with NO implicit FMA contraction.
//
// The parity contract (SURVEY.md §8a "numerics notes", §8d) is bit-exact height-cell indices,
// termination flags and terrain levels against the reference's torch-CPU arithmetic, so every
// operation on those chains is spelled with the rounding it has there:
//   * elementwise aten ops round after every op            -> __fmul_rn / __fadd_rn / __fdiv_rn
//   * aten's p=2 norm over 2 or 3 elements accumulates with a fused multiply-add
//     (acc = fma(x, x, acc), measured against torch 2.11 CPU)  -> norm2_fma / norm3_fma
//   * over the 4-element yaw quaternion (0, 0, z, w) it is sqrt(fl(z*z) + fl(w*w))
// The translation unit is additionally compiled with --fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace shifu {

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }

__device__ __forceinline__ float norm2_fma(float x, float y) {
  return sqrt_rn(fma_rn(y, y, mul_rn(x, x)));
}
__device__ __forceinline__ float norm3_fma(float x, float y, float z) {
  return sqrt_rn(fma_rn(z, z, fma_rn(y, y, mul_rn(x, x))));
}

// Correctly rounded x / d for a loop-invariant divisor d with r = RN(1/d):
//   q0 = RN(x*r); e = fma(-d, q0, x) (exact); q = fma(e, r, q0)
// (Markstein's theorem; verified exhaustively against IEEE division for d = 0.1f over every
// float with 1e-3 <= |x| <= 65536 in tests/test_exact_div.py.)  3 instructions instead of the
// ~8 + slow path of __fdiv_rn.
struct ConstDiv {
  float d, r;
};
__device__ __forceinline__ float div_const(float x, const ConstDiv c) {
  const float q0 = mul_rn(x, c.r);
  const float e = fma_rn(-c.d, q0, x);
  return fma_rn(e, c.r, q0);
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

}  // namespace shifu
