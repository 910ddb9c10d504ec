// Dev tool: per-SM issue rates of the instructions the height scan is made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each kernel runs ITER iterations of 8 independent dependency chains per thread; the rate printed is
// warp-instructions per clock per SM at 512 threads per SM-resident CTA x 2 CTAs (32 warps/SM).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2_t;
#define ITER 4096
__device__ __forceinline__ f2_t pk(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int OP> __device__ __forceinline__ void step(f2_t (&v)[8], float (&s)[8], int (&n)[8], f2_t c2, float c, int ci) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[u]) : "l"(c2));
    if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[u]) : "l"(c2));
    if (OP == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[u]) : "l"(c2));
    if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(s[u]) : "f"(c));
    if (OP == 4) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[u]) : "f"(c));
    if (OP == 5) { unsigned r; asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(r) : "f"(s[u])); s[u] = __uint_as_float(r | 0x3f800000u); }
    if (OP == 6) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(s[u]) : "r"(n[u])); n[u] = __float_as_int(s[u]) & 0xffff; }
    if (OP == 7) asm volatile("min.s32 %0, %0, %1;" : "+r"(n[u]) : "r"(ci));
    if (OP == 8) asm volatile("mul.rz.f32x2 %0, %0, %1;" : "+l"(v[u]) : "l"(c2));     // denormal results
    if (OP == 9) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(n[u]) : "r"(ci));
    if (OP == 10) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(n[u]) : "r"(ci));
    if (OP == 11) asm volatile("min.f32 %0, %0, %1;" : "+f"(s[u]) : "f"(c));
    if (OP == 12) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[u]) : "l"(c2)); asm volatile("min.s32 %0, %0, %1;" : "+r"(n[u]) : "r"(ci)); }
    if (OP == 13) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(s[u]) : "f"(c)); asm volatile("min.s32 %0, %0, %1;" : "+r"(n[u]) : "r"(ci)); }
    if (OP == 14) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[u]) : "l"(c2)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(s[u]) : "f"(c)); }
  }
}
__device__ unsigned long long g_cyc;
template <int OP> __global__ void __launch_bounds__(1024) k(float* out, float c, int ci, float init) {
  const long long t0 = clock64();
  f2_t v[8]; float s[8]; int n[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) { v[u] = pk(init + u, init - u); s[u] = init + u; n[u] = threadIdx.x + u; }
  const f2_t c2 = pk(c, c);
#pragma unroll 1
  for (int i = 0; i < ITER; ++i) step<OP>(v, s, n, c2, c, ci);
  float acc = 0;
#pragma unroll
  for (int u = 0; u < 8; ++u) acc += s[u] + (float)n[u] + __uint_as_float((unsigned)(v[u] >> 32)) + __uint_as_float((unsigned)v[u]);
  if (acc == 123.456f) out[0] = acc;
  const long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) g_cyc = (unsigned long long)(t1 - t0);
}
template <int OP> void run_low(const char* name, int per_iter, float c, float init, int threads) {
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  k<OP><<<sms, threads>>>(d, c, 7, init); cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a); k<OP><<<sms, threads>>>(d, c, 7, init); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  unsigned long long cyc = 0; cudaMemcpyFromSymbol(&cyc, g_cyc, 8);
  printf("%-26s warps/SMSP=%d  %6.2f SM cycles (clock64) per warp-instruction per warp; %5.3f warp-instr/clk/SMSP; event-clock ratio %.2f\n",
         name, threads / 128, (double)cyc / ((double)ITER * 8 * per_iter), (threads / 128) * ((double)ITER * 8 * per_iter) / (double)cyc,
         best * 1e-3 * khz * 1e3 / (double)cyc);
  cudaFree(d);
}
template <int OP> void run(const char* name, int per_iter, float c, float init) {
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  k<OP><<<sms * 2, 512>>>(d, c, 7, init); cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a); k<OP><<<sms * 2, 512>>>(d, c, 7, init); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  double winst = (double)sms * 2 * 16 * ITER * 8 * per_iter;
  double clocks = best * 1e-3 * khz * 1e3;
  printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM (at %d kHz nominal)\n", name, best, winst / clocks / sms, khz);
  cudaFree(d);
}
int main() {
  run<0>("FFMA2", 1, 0.999f, 1.0f);
  run<1>("FADD2", 1, 1e-3f, 1.0f);
  run<2>("FMUL2", 1, 0.9999f, 1.0f);
  run<3>("FFMA", 1, 0.999f, 1.0f);
  run<4>("FADD", 1, 1e-3f, 1.0f);
  run<5>("F2I.U32.TRUNC (+LOP)", 2, 0.f, 3.5f);
  run<6>("I2F(P).F32.S32 (+LOP)", 2, 0.f, 3.5f);
  run<7>("VIMNMX", 1, 0.f, 3.5f);
  run<8>("FMUL2.RZ denormal out", 1, 1.401298464e-45f, 1000.0f);
  run<9>("IMAD", 1, 0.f, 3.5f);
  run<10>("LOP3", 1, 0.f, 3.5f);
  run<11>("FMNMX", 1, 5.f, 3.5f);
  run<12>("FFMA2 + VIMNMX", 2, 0.999f, 1.0f);
  run<13>("FFMA + VIMNMX", 2, 0.999f, 1.0f);
  run<14>("FFMA2 + FFMA", 2, 0.999f, 1.0f);
  for (int thr = 128; thr <= 1024; thr *= 2) {
    run_low<0>("FFMA2", 1, 0.999f, 1.0f, thr);
    run_low<1>("FADD2", 1, 1e-3f, 1.0f, thr);
    run_low<3>("FFMA", 1, 0.999f, 1.0f, thr);
    run_low<7>("VIMNMX", 1, 0.f, 3.5f, thr);
    run_low<9>("IMAD", 1, 0.f, 3.5f, thr);
    run_low<12>("FFMA2+VIMNMX", 2, 0.999f, 1.0f, thr);
    run_low<14>("FFMA2+FFMA", 2, 0.999f, 1.0f, thr);
  }
  return 0;
}
