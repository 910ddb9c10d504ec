// Dev tool: achievable HBM bandwidth of a perfectly coalesced stream with the fused kernel's read/write
// mix (0.73 GB read + 1.27 GB written per launch) — the practical memory floor of that traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rwmix rwmix.cu && ./rwmix
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) rw(const float4* __restrict__ in, size_t n_in, float4* __restrict__ out, size_t n_out, int streaming) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  float4 acc = make_float4(0, 0, 0, 0);
  // interleave: every thread alternates loads and stores in the n_in : n_out ratio
  size_t i = tid, o = tid;
  while (i < n_in || o < n_out) {
    if (i < n_in) { const float4 v = __ldg(in + i); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; i += nth; }
    for (int r = 0; r < 2 && o < n_out; ++r) {
      if (streaming) __stcs(out + o, acc); else out[o] = acc;
      o += nth;
    }
  }
}
int main() {
  const size_t rb = 730ull << 20, wb = 1266ull << 20;
  float4 *in, *out; cudaMalloc(&in, rb); cudaMalloc(&out, wb); cudaMemset(in, 0, rb);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int streaming = 0; streaming < 2; ++streaming)
    for (int per_sm : {4, 8, 16}) {
      float best = 1e9;
      for (int r = 0; r < 8; ++r) {
        cudaEventRecord(a); rw<<<sms * per_sm, 256>>>(in, rb / 16, out, wb / 16, streaming); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (r > 1 && ms < best) best = ms;
      }
      printf("read 0.73 GiB + write 1.27 GiB, %s stores, %2d CTAs/SM: %.4f ms = %.0f GB/s\n", streaming ? "st.cs" : "plain", per_sm, best,
             (rb + wb) / (best * 1e-3) / 1e9);
    }
  // pure copy reference (50/50)
  float best = 1e9;
  for (int r = 0; r < 8; ++r) { cudaEventRecord(a); cudaMemcpyAsync(out, in, rb, cudaMemcpyDeviceToDevice); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r > 1 && ms < best) best = ms; }
  printf("cudaMemcpy D2D 0.73 GiB: %.4f ms = %.0f GB/s (read+write)\n", best, 2.0 * rb / (best * 1e-3) / 1e9);
  return 0;
}
