// Task-independent kernels: reset-id compaction (row a8), history push (a13), clip (a14),
// step-statistics collection / extras publication (row a11 logging).
#pragma once
#include "exact_math.cuh"
#include "../../include/shifu_b200.h"

namespace shifu {

// ------------------------------------------------------------------------------------------
// Row a8: env_ids = reset_buf.nonzero().flatten()  (shifu/gym/env.py:101) — ascending int64.
//
// Single pass, chained scan over CTAs: a CTA takes chunk b from an atomic ticket (so every lower
// chunk is owned by a CTA that is already running), counts the non-zero flags of its 4096-flag
// chunk (128-bit loads, 16 flags per thread), publishes the count tagged with the launch epoch,
// waits for the counts of all lower chunks and writes its ids at the exclusive prefix.  Inside the chunk the order is restored with
// a warp ballot + block prefix scan.  No host round trip: the total goes to a device int32.
// ------------------------------------------------------------------------------------------
constexpr int COMPACT_THREADS = 256;
constexpr int COMPACT_PER_THREAD = 16;
constexpr int COMPACT_CHUNK = COMPACT_THREADS * COMPACT_PER_THREAD;   // 4096

__global__ void __launch_bounds__(COMPACT_THREADS)
compact_ids_kernel(const unsigned char* __restrict__ flags, int n, long long* __restrict__ ids,
                   int* __restrict__ n_out, unsigned long long* __restrict__ chain,
                   unsigned* __restrict__ ticket, unsigned epoch) {
  __shared__ int s_warp[COMPACT_THREADS / 32];
  __shared__ int s_base, s_b;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_b = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int b = s_b;
  const long long first = (long long)b * COMPACT_CHUNK + (long long)t * COMPACT_PER_THREAD;

  // 16 flags of this thread -> bit mask
  unsigned mask = 0;
  if (first + COMPACT_PER_THREAD <= n && ((reinterpret_cast<uintptr_t>(flags) & 15u) == 0)) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(flags + first));
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((w[i] >> (8 * j)) & 0xffu) mask |= 1u << (4 * i + j);
  } else {
#pragma unroll
    for (int i = 0; i < COMPACT_PER_THREAD; ++i)
      if (first + i < n && flags[first + i]) mask |= 1u << i;
  }
  const int cnt = __popc(mask);

  // exclusive prefix of cnt inside the CTA
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int warp_base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < COMPACT_THREADS / 32; ++w) {
    if (w < warp) warp_base += s_warp[w];
    total += s_warp[w];
  }
  // publish this chunk's count: (epoch << 32) | count
  if (t == 0) {
    __threadfence();
    atomicExch(chain + b, ((unsigned long long)epoch << 32) | (unsigned)total);
  }
  // exclusive prefix over lower chunks (warp 0 polls 32 predecessors at a time)
  if (warp == 0) {
    int base = 0;
    for (int p0 = 0; p0 < b; p0 += 32) {
      const int p = p0 + lane;
      int c = 0;
      if (p < b) {
        unsigned long long v;
        do {
          v = *reinterpret_cast<volatile unsigned long long*>(chain + p);
        } while ((unsigned)(v >> 32) != epoch);
        c = (int)(unsigned)v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      base += c;
    }
    if (lane == 0) {
      s_base = base;
      if (b == gridDim.x - 1) *n_out = base + total;
    }
  }
  __syncthreads();
  long long* out = ids + s_base + warp_base + (incl - cnt);
  unsigned m = mask;
  while (m) {
    const int i = __ffs(m) - 1;
    m &= m - 1;
    *out++ = first + i;
  }
  // the last CTA to retire re-arms the ticket for the next launch (launches are stream-ordered)
  if (t == 0) {
    __threadfence();
    if (atomicAdd(ticket + 1, 1u) == gridDim.x - 1) { ticket[0] = 0u; ticket[1] = 0u; }
  }
}

// ------------------------------------------------------------------------------------------
// Row a13: HistoryRecorder.add (shifu/utils/train.py:12-14): buf[...,1:] = buf[...,:-1]; buf[...,0] = x
// One thread per (env, channel): reads its H history slots, writes them shifted.
// ------------------------------------------------------------------------------------------
__global__ void history_add_kernel(float* __restrict__ hist, const float* __restrict__ x, long long rows, int h) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
       i += (long long)gridDim.x * blockDim.x) {
    float* r = hist + i * h;
    for (int j = h - 1; j > 0; --j) r[j] = r[j - 1];
    r[0] = x[i];
  }
}

// Row a14 / a1: torch.clip(x, -c, c) (shifu/gym/env.py:87,90)
__global__ void clip_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float c) {
  const long long n4 = n >> 2;
  const bool vec = (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (vec) {
    for (long long i = tid; i < n4; i += stride) {
      float4 v = reinterpret_cast<const float4*>(in)[i];
      v.x = clampf(v.x, -c, c); v.y = clampf(v.y, -c, c); v.z = clampf(v.z, -c, c); v.w = clampf(v.w, -c, c);
      reinterpret_cast<float4*>(out)[i] = v;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += stride) out[i] = clampf(in[i], -c, c);
  } else {
    for (long long i = tid; i < n; i += stride) out[i] = clampf(in[i], -c, c);
  }
}

// ------------------------------------------------------------------------------------------
// Step statistics.  acc[] is what the post-physics kernels atomically add into during a step;
// collect moves it out (adding the persistent all-env terrain-level sum and N) and clears it.
// ------------------------------------------------------------------------------------------
// `out` is slot `slot` of a ring of `slots` vectors (slot < 0: *step_dev % slots, read before the
// increment, for graph-replayed steps); slots == 1 is the plain single-vector form.
__global__ void collect_stats_kernel(double* __restrict__ acc, double* __restrict__ level_sum,
                                     double* __restrict__ out, int slots, int slot, double n_envs, long long* step_dev) {
  const int i = threadIdx.x;
  if (slot < 0) slot = (int)(*step_dev % slots);
  __syncwarp();
  if (i == 31 && step_dev != nullptr) *step_dev += 1;      // common_step_counter += 1 (env.py:96)
  out += (size_t)slot * SHIFU_NUM_STATS;
  if (i >= SHIFU_NUM_STATS) return;
  double v = acc[i];
  acc[i] = 0.0;
  if (i == SHIFU_STAT_LEVEL_SUM) {
    v += *level_sum;       // delta of this step + running sum
    *level_sum = v;
  }
  if (i == SHIFU_STAT_NENVS) v = n_envs;
  out[i] = v;
}

// extras["episode"]: mean over reset envs / max_episode_length_s (env.py:149-153), left untouched
// when nobody reset (env.py:115-116); terrain_levels mean over all envs (a1_conditional.py:126-129);
// success_rate = mean over reset envs (a_prior_stage.py:92-93).
// The output is a ring of `slots` arrays of SHIFU_NUM_STATS floats: each step publishes into its own
// slot (slot < 0: derived from the device step counter, for graph-replayed steps), so the dicts a
// caller keeps from earlier steps (rsl_rl keeps 24) stay distinct; a step without resets copies the
// previous slot (the reference keeps the same dict alive).  slots == 1 is the in-place form.
__global__ void publish_extras_kernel(const double* __restrict__ stats, float* __restrict__ ring, int slots, int slot,
                                      const long long* __restrict__ step_dev, float max_len_s, int n_terms) {
  const int i = threadIdx.x;
  if (i >= SHIFU_NUM_STATS) return;
  if (slot < 0) {                       // graph-replayed step: `stats` is the ring shifu_collect_stats_ring fills
    slot = (int)((*step_dev - 1) % slots);
    stats += (size_t)slot * SHIFU_NUM_STATS;
  }
  const int prev = (slot + slots - 1) % slots;
  float* extras = ring + (size_t)slot * SHIFU_NUM_STATS;
  const float* old = ring + (size_t)prev * SHIFU_NUM_STATS;
  const double nreset = stats[SHIFU_STAT_NRESET];
  if (i == SHIFU_STAT_NRESET) { extras[i] = (float)nreset; return; }
  float v = old[i];
  if (nreset > 0.0) {
    if (i < SHIFU_MAX_REWARD_TERMS) {
      if (i < n_terms) v = div_rn((float)(stats[i] / nreset), max_len_s);
    } else if (i == SHIFU_STAT_LEVEL_SUM) {
      v = (float)(stats[i] / stats[SHIFU_STAT_NENVS]);
    } else if (i == SHIFU_STAT_SUCCESS) {
      v = (float)(stats[i] / nreset);
    }
  }
  extras[i] = v;
}

// set-up only (shifu_set_level_sum): grid-stride sum of the terrain levels; *out must be zeroed first
__global__ void __launch_bounds__(256) level_sum_kernel(const long long* __restrict__ levels, int n, double* __restrict__ out) {
  __shared__ double s[256];
  double a = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a += (double)levels[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, s[0]);       // integer-valued partial sums: exact in any order
}

}  // namespace shifu
