// Packed fp32x2 arithmetic of sm_100 (PTX add/sub/mul/fma .f32x2, SASS FADD2/FMUL2/FFMA2):
// two IEEE round-to-nearest fp32 operations per issued instruction, each lane rounded exactly
// like the scalar op — so the bit-exact height-index chain can process two envs per instruction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace shifu {

typedef unsigned long long f2_t;   // (lo, hi) pair of floats in one 64-bit register

__device__ __forceinline__ f2_t pk(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(f2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) {
  f2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) {
  f2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2_t sub2(f2_t a, f2_t b) {
  f2_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// round-toward-zero product, denormal results kept (no .ftz)
__device__ __forceinline__ f2_t mulrz2(f2_t a, f2_t b) {
  f2_t r;
  asm("mul.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void upk_i(f2_t v, int& lo, int& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) {
  f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

}  // namespace shifu
