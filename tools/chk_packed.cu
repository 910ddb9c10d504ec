// dev check: packed fp32x2 chain vs scalar chain of the height-index arithmetic
#include <cstdio>
#include <cuda_runtime.h>
#include "../shifu_b200/csrc/exact_math.cuh"
#include "../shifu_b200/csrc/f32x2.cuh"
using namespace shifu;
__global__ void k(int n, unsigned long long* bad, float* dump, float nz) {
  const f2_t NZ = pk(nz, nz);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned s = i * 2654435761u + 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };
  float yaw = (rnd() * 2 - 1) * 3.14159265f;
  float qz = sinf(yaw * 0.5f) * (1.0f + 0.01f * (rnd() - 0.5f)), qw = cosf(yaw * 0.5f) * (1.0f + 0.01f * (rnd() - 0.5f));
  float nrm = sqrt_rn(add_rn(mul_rn(qz, qz), mul_rn(qw, qw)));
  float z = div_rn(qz, nrm), w = div_rn(qw, nrm), z2 = mul_rn(z, 2.0f);
  float X = rnd() * 100.0f + 20.0f, Y = rnd() * 180.0f + 20.0f;
  float bx = -0.8f + 0.1f * (float)(s % 17), by = -0.5f + 0.1f * (float)((s >> 8) % 11);
  const float border = 25.0f, d = 0.1f, r = 1.0f / d;
  // scalar
  float tx = -mul_rn(z2, by), ty = mul_rn(z2, bx);
  float rx = add_rn(add_rn(bx, mul_rn(w, tx)), -mul_rn(z, ty));
  float ry = add_rn(add_rn(by, mul_rn(w, ty)), mul_rn(z, tx));
  float ax = add_rn(add_rn(rx, X), border), ay = add_rn(add_rn(ry, Y), border);
  ConstDiv cd{d, r};
  float fx = div_const(ax, cd), fy = div_const(ay, cd);
  // packed (both lanes same data)
  f2_t Z2 = pk(z2, z2), Z = pk(z, z), W = pk(w, w), PX = pk(X, X), PY = pk(Y, Y);
  f2_t BX = pk(bx, bx), BY = pk(by, by), NBY = pk(-by, -by), BORDER = pk(border, border), RCP = pk(r, r), NEGD = pk(-d, -d);
  f2_t ptx = mul2(Z2, NBY), pty = mul2(Z2, BX);
  f2_t prx = sub2(add2(BX, fma2(W, ptx, NZ)), fma2(Z, pty, NZ));
  f2_t pry = add2(add2(BY, fma2(W, pty, NZ)), fma2(Z, ptx, NZ));
  f2_t pax = add2(add2(prx, PX), BORDER), pay = add2(add2(pry, PY), BORDER);
  f2_t qx = mul2(pax, RCP), qy = mul2(pay, RCP);
  float gx0, gx1, gy0, gy1;
  upk(fma2(fma2(NEGD, qx, pax), RCP, qx), gx0, gx1);
  upk(fma2(fma2(NEGD, qy, pay), RCP, qy), gy0, gy1);
  float a0, a1; upk(pax, a0, a1);
  float t0, t1; upk(ptx, t0, t1);
  float r0, r1; upk(prx, r0, r1);
  if (gx0 != fx || gy0 != fy || gx1 != fx || gy1 != fy) {
    unsigned long long c = atomicAdd(bad, 1ull);
    if (c < 4) { float* o = dump + c * 12; o[0]=fx;o[1]=gx0;o[2]=fy;o[3]=gy0;o[4]=ax;o[5]=a0;o[6]=tx;o[7]=t0;o[8]=rx;o[9]=r0;o[10]=z;o[11]=w; }
  }
}
int main() {
  unsigned long long* bad; float* dump;
  cudaMallocManaged(&bad, 8); cudaMallocManaged(&dump, 4 * 12 * 4); *bad = 0;
  int n = 1 << 28;
  k<<<(n + 255) / 256, 256>>>(n, bad, dump, -0.0f);
  cudaDeviceSynchronize();
  printf("n=%d mismatches=%llu err=%s\n", n, *bad, cudaGetErrorString(cudaGetLastError()));
  for (int c = 0; c < 4 && c < (int)*bad; ++c) { float* o = dump + c * 12;
    printf("fx %a %a | fy %a %a | ax %a %a | tx %a %a | rx %a %a | z %a w %a\n", o[0],o[1],o[2],o[3],o[4],o[5],o[6],o[7],o[8],o[9],o[10],o[11]); }
  return 0;
}
