// K-main, pipelined form (the default for full 32-env tiles): persistent CTAs (2 per SM, 512 threads =
// the whole register file at 64 registers) with three warp roles that never meet at a CTA-wide barrier —
//
//   DMA warp (1 lane)    bulk-TMA loads (cp.async.bulk -> mbarrier, SASS UBLKCP) of everything the
//                        tile reads — the root / dof / contact / history / torque / action rows
//                        and the per-env scalars (ep_len, command, carried velocities, episode
//                        sums) — into a ring of three shared-memory stages, refilled as soon as the
//                        tile's head phase is over; bulk-TMA stores of the pushed history tile and of
//                        the carried body-frame rows
//   B group (3 warps; lane = env)
//                        yaw normalisation for the scan (warp 0, first: the scan group starts on
//                        e_done), the reward-term list (split over the three warps by a host-side cost
//                        balance; warp 2 does nothing else and runs ahead into the next tile's terms)
//                        + episode sums, termination, reset (curriculum, Philox draws, state rewrite
//                        and the carried body-frame velocities of the post-reset pose — warp 1,
//                        while warp 0 does the ordered accumulation and the flags), per-step log
//                        sums — the scalar game logic of ShifuVecEnv.post_step (env.py:93-106);
//                        reads only the stage (plus env origin / terrain level / type of the ~1 % of
//                        envs that reset)
//   scan group (12 warps; thread = scan point + its mirror point, warp w: point pairs 32*(w%3)..,
//                        env pairs 4*(w/3) .. +3 in two items of 4 envs)
//                        first item's index arithmetic + gathers, then the obs head from the
//                        post-reset rows (a1_conditional.py:131-144), history push (train.py:12-14),
//                        then the rest of the 187-point height scan (isaac_gym.py:393-433) with
//                        packed fp32x2 arithmetic (two envs per instruction, one yaw rotation per
//                        point pair), streamed with st.global.cs
//
// mbarriers hand a stage round DMA -> B -> scan -> DMA; every thread of the producing group
// arrives itself, so fast warps never wait for slow siblings.  The per-env scalars the scan needs
// travel through a 4-deep ring, which lets the B group run ahead of the scan group.
#pragma once
#include "a1_fused.cuh"
#include "f32x2.cuh"
#include "tma_pipe.cuh"

namespace shifu {

// B groups of 2 warps each: group g owns the tiles j = g, g + V3_B_GROUPS, ... (dev knob; 1 is fastest)
#ifndef V3_BG_WARPS_CFG
#define V3_BG_WARPS_CFG 3
#endif
#ifndef V3_B_GROUPS_CFG
#define V3_B_GROUPS_CFG 1
#endif
#ifndef V3_SCAN_WARPS
#define V3_SCAN_WARPS 12
#endif
#ifndef V3_CTAS_CFG
#if V3_SCAN_WARPS == 6
#define V3_CTAS_CFG 3
#else
#define V3_CTAS_CFG 2
#endif
#endif
#ifndef V3_POLL_B
#define V3_POLL_B 64
#endif
#ifndef V3_POLL_DMA
#define V3_POLL_DMA 64
#endif
#ifndef V3_POLL_SCAN
#define V3_POLL_SCAN 32
#endif
#ifdef V3_PARK_NS
#define V3_WAIT(ns, bar, par) pipe::mbar_wait_parked(bar, par, V3_PARK_NS)
#else
#define V3_WAIT(ns, bar, par) pipe::mbar_wait<ns>(bar, par)
#endif
#ifndef V3_STAGES_CFG
#if V3_SCAN_WARPS == 12
#define V3_STAGES_CFG 3
#else
#define V3_STAGES_CFG 2
#endif
#endif
constexpr int V3_STAGES = V3_STAGES_CFG;          // input stages per CTA (the scalar ring holds 4)
constexpr int V3_B_GROUPS = V3_B_GROUPS_CFG;
static_assert(V3_B_GROUPS == 1, "one B group per CTA (the stage ring assumes it)");   // B groups of 2 warps; group g owns tiles j = g, g + V3_B_GROUPS, ...
constexpr int V3_CTAS_PER_SM = V3_CTAS_CFG;
constexpr int V3_BG_WARPS = V3_BG_WARPS_CFG;            // warps 0, 1: terms + B2; further warps: terms only
static_assert(V3_BG_WARPS >= 2 && V3_BG_WARPS <= A1K_TERM_WARPS, "B group: 2..A1K_TERM_WARPS warps");
constexpr int V3_BG_THREADS = 32 * V3_BG_WARPS, V3_B2_THREADS = 64;
constexpr int V3_B_THREADS = V3_B_GROUPS * V3_BG_THREADS, V3_C_THREADS = 32 * V3_SCAN_WARPS;
constexpr int V3_DMA_THREADS = 32;
constexpr int V3_THREADS = V3_B_THREADS + V3_C_THREADS + V3_DMA_THREADS;
constexpr int V3_SCAN_BASE = V3_B_THREADS;
constexpr int V3_DMA_BASE = V3_B_THREADS + V3_C_THREADS;
static_assert(V3_SCAN_WARPS == 6 || V3_SCAN_WARPS == 12, "scan group: 6 or 12 warps");
constexpr uint32_t V3_OBS_CHUNK_BYTES = 8 * A1_OBS * 4;      // 8288 = 16 * 518

struct alignas(128) V3In {            // one tile of simulator/env rows, each member 16-B aligned
  float root[A1_TILE][13];            //  1664 B
  float dof[A1_TILE][A1_DOF * 2];     //  3072 B
  float contact[A1_TILE][A1_BODIES * 3];   // 6528 B
  float hist[A1_TILE][A1_DOF * A1_HIST];   // 4608 B  (pushed in place, then bulk-stored back)
  float tau[A1_TILE][A1_DOF];         //  1536 B
  float act[A1_TILE][A1_DOF];         //  1536 B
  // per-env scalars: bulk-loaded with the rows, so the B warps issue no global loads of their own
  long long ep_len[A1_TILE];          //   256 B
  float cla[3][A1_TILE][3];           //  1152 B  command (rewritten on reset), base lin vel, base ang vel
  float esum[SHIFU_MAX_REWARD_TERMS][A1_TILE];   // 1024 B  (the first n_terms rows are loaded)
  // output rows: carried body-frame velocities / projected gravity of the POST-reset pose, written by
  // the B group and bulk-stored by the DMA lane together with the pushed history
  float cout[3][A1_TILE][3];          //  1152 B
};
static_assert(V3_STAGES >= 2 && V3_STAGES <= 4, "ring depth");
static_assert(sizeof(V3In) == 22528 && offsetof(V3In, ep_len) == 18944 && offsetof(V3In, cout) % 16 == 0, "tile layout");

struct alignas(128) V3Smem {
  V3In in[V3_STAGES];
  // per-env scan scalars, laid out per env PAIR (e, e+1) so that one 128-bit broadcast load yields
  // two packed fp32x2 operands: sA = (2zq_e, 2zq_e1, zq_e, zq_e1), sB = (wq_e, wq_e1, x_e, x_e1),
  // sC = (y_e, y_e1, zb_e, zb_e1); (zq, wq) = normalised yaw quaternion of the PRE-reset pose,
  // zb = z_postreset - 0.5
  // (4-deep ring: the B groups may run ahead of the scan group without an extra hand-shake)
  float4 sA[4][A1_TILE / 2];
  float4 sB[4][A1_TILE / 2];
  float4 sC[4][A1_TILE / 2];
  // double-buffered on the tile parity: warp 1 may start the next tile's terms while warp 0 still
  // accumulates this tile's
  float rterm[2][SHIFU_MAX_REWARD_TERMS][A1_TILE];
  uint64_t full_in[V3_STAGES], e_done[V3_STAGES], b_done[V3_STAGES], h_done[V3_STAGES];
#ifdef V3_PROFILE
  long long t_issue[V3_STAGES];
#endif
};

// Optional phase timers (-DV3_PROFILE, tools/prof_phases.py): one lead thread per role adds the
// clock64() cycles it spends in each phase to v3_prof[]; never compiled into the shipped library.
#ifdef V3_PROFILE
__device__ unsigned long long v3_prof[32];
#define V3_T0(lead) const bool _pl = (lead); long long _pt = clock64()
#define V3_TICK(slot) do { const long long _n = clock64(); if (_pl) atomicAdd(&v3_prof[slot], (unsigned long long)(_n - _pt)); _pt = clock64(); } while (0)
#define V3_COUNT(slot) do { if (_pl) atomicAdd(&v3_prof[slot], 1ull); } while (0)
#define V3_STAMP(x) x = clock64()
#define V3_SINCE(slot, x) do { if (_pl) atomicAdd(&v3_prof[slot], (unsigned long long)(clock64() - (x))); } while (0)
#else
#define V3_STAMP(x)
#define V3_SINCE(slot, x)
#define V3_T0(lead)
#define V3_TICK(slot)
#define V3_COUNT(slot)
#endif

// dev knobs for the cache behaviour of the table gathers and the obs stores
#ifndef V3_LD_MODE
#define V3_LD_MODE 0
#endif
#ifndef V3_ST_MODE
#define V3_ST_MODE 0
#endif
__device__ __forceinline__ int v3_gather(const short* p) {
#ifdef V3_WI_NOGATHER
  return (int)((unsigned long long)p & 1023);     // what-if: no table access
#endif
#if V3_LD_MODE == 0
  return __ldg(p);
#elif V3_LD_MODE == 1
  short v; asm volatile("ld.global.s16 %0, [%1];" : "=h"(v) : "l"(p)); return v;
#elif V3_LD_MODE == 2
  short v; asm volatile("ld.global.nc.L1::evict_last.s16 %0, [%1];" : "=h"(v) : "l"(p)); return v;
#else
  short v; asm volatile("ld.global.L1::evict_last.s16 %0, [%1];" : "=h"(v) : "l"(p)); return v;
#endif
}
__device__ __forceinline__ void v3_store(float* p, float v) {
#ifdef V3_WI_NOSTORE
  if (v == 123.456f) *p = v;      // what-if: keep the value alive, (almost) never store
  return;
#endif
#if V3_ST_MODE == 0
  __stcs(p, v);
#elif V3_ST_MODE == 1
  *p = v;
#elif V3_ST_MODE == 2
  __stcg(p, v);
#else
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
#endif
}

// what-if: spin for N cycles (dev builds only) to find out which role is on the critical path
__device__ __forceinline__ void v3_delay(int cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) { }
}
constexpr uint32_t V3_ROW_BYTES = offsetof(V3In, ep_len);
constexpr uint32_t V3_HIST_BYTES = A1_TILE * A1_DOF * A1_HIST * 4;

// with_hist = false leaves the history rows out (their stage buffer is still being read by the
// bulk store of the pushed history); v3_issue_hist_load follows once that read has finished.
__device__ __forceinline__ void v3_issue_hist_load(V3In& in, const ShifuA1StepIO& io, long long e0, uint64_t* bar) {
  pipe::bulk_load(in.hist, io.history + e0 * (A1_DOF * A1_HIST), sizeof(in.hist), bar);
}

#if V3_SCAN_WARPS == 12 || defined(V3_INLINE)
#define V3_NOINLINE __forceinline__
#else
#define V3_NOINLINE __noinline__
#endif
__device__ V3_NOINLINE void v3_issue_loads(V3In& in, const A1K& k, const ShifuA1StepIO& io, long long e0,
                                               uint64_t* bar, bool with_hist) {
  const uint32_t scalars = sizeof(in.ep_len) + sizeof(in.cla) + k.n_terms * sizeof(in.esum[0]);
  pipe::mbar_arrive_expect_tx(bar, V3_ROW_BYTES + scalars);
  pipe::bulk_load(in.root, io.root_state + e0 * 13, sizeof(in.root), bar);
  pipe::bulk_load(in.dof, io.dof_state + e0 * (A1_DOF * 2), sizeof(in.dof), bar);
  pipe::bulk_load(in.contact, io.contact_state + e0 * (A1_BODIES * 3), sizeof(in.contact), bar);
  if (with_hist) v3_issue_hist_load(in, io, e0, bar);
  pipe::bulk_load(in.tau, io.torques + e0 * A1_DOF, sizeof(in.tau), bar);
  pipe::bulk_load(in.act, io.actions + e0 * A1_DOF, sizeof(in.act), bar);
  pipe::bulk_load(in.ep_len, io.ep_len + e0, sizeof(in.ep_len), bar);
  pipe::bulk_load(in.cla[0], io.command + e0 * 3, sizeof(in.cla[0]), bar);
  pipe::bulk_load(in.cla[1], io.base_lin_vel + e0 * 3, sizeof(in.cla[1]), bar);
  pipe::bulk_load(in.cla[2], io.base_ang_vel + e0 * 3, sizeof(in.cla[2]), bar);
  for (int q = 0; q < k.n_terms; ++q) pipe::bulk_load(in.esum[q], io.ep_sums[q] + e0, sizeof(in.esum[q]), bar);
}

// Reward term for env e reading the tile rows of `in` (same arithmetic as a1_eval_term).
template <bool HAS_EXTRA>
__device__ V3_NOINLINE float v3_eval_term(int code, int q, float p0, float p1, const A1K& k, const V3In& in, int e,
                                          const ShifuA1StepIO& io, long long ge) {
  const float* cmd = in.cla[0][e];
  const float* lin = in.cla[1][e];
  const float* ang = in.cla[2][e];
  if (HAS_EXTRA && code >= SHIFU_REW_LIN_VEL_Z) {   // legged_gym-style terms: pg / air-time state straight from global memory
    const TermCtx c{cmd, lin, ang, io.projected_gravity + ge * 3, in.dof[e], in.hist[e], in.act[e], in.contact[e],
                    in.root[e][2], io.swing_time ? io.swing_time + ge * k.n_feet : nullptr,
                    io.last_contacts ? io.last_contacts + ge * k.n_feet : nullptr};
    return a1_extra_term(code, p0, p1, k, c);
  }
  switch (code) {
    case SHIFU_REW_TRACKING_LIN_VEL: {
      const float dx = sub_rn(cmd[0], lin[0]), dy = sub_rn(cmd[1], lin[1]);
      const float ne = -add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
      return mul_rn(p0, expf(k.rp_pow2[q] ? mul_rn(ne, k.rp_inv[q]) : div_rn(ne, p1)));
    }
    case SHIFU_REW_TRACKING_ANG_VEL: {
      const float d = sub_rn(cmd[2], ang[2]);
      const float ne = -mul_rn(d, d);
      return mul_rn(p0, expf(k.rp_pow2[q] ? mul_rn(ne, k.rp_inv[q]) : div_rn(ne, p1)));
    }
    case SHIFU_REW_STABILIZING_BASE:
      return add_rn(mul_rn(p0, mul_rn(lin[2], lin[2])),
                    mul_rn(p1, add_rn(mul_rn(ang[0], ang[0]), mul_rn(ang[1], ang[1]))));
    case SHIFU_REW_SMOOTHING_ACTION: {
      float f1 = 0.0f, f2 = 0.0f;
#pragma unroll
      for (int d = 0; d < A1_DOF; ++d) {
        const float a0 = in.hist[e][d * A1_HIST + 0], a1 = in.hist[e][d * A1_HIST + 1],
                    a2 = in.hist[e][d * A1_HIST + 2];
        const float d1 = sub_rn(a1, a0);
        const float d2 = add_rn(sub_rn(a2, mul_rn(2.0f, a1)), a0);
        f1 = add_rn(f1, mul_rn(d1, d1));
        f2 = add_rn(f2, mul_rn(d2, d2));
      }
      return mul_rn(p0, add_rn(f1, f2));
    }
    case SHIFU_REW_LEG_COLLISION: {
      int cnt = 0;
      for (int b = 0; b < k.n_leg; ++b) {   // rolled on purpose: unrolled, the B warps spill (0.4376 vs 0.4289 ms)
        const float* f = &in.contact[e][k.leg[b] * 3];
        // |F| > p1 tested on the sum of squares (threshold pre-squared exactly on the host)
        cnt += (fma_rn(f[2], f[2], fma_rn(f[1], f[1], mul_rn(f[0], f[0]))) > k.rp_thr_sq[q]) ? 1 : 0;
      }
      return mul_rn(p0, (float)cnt);
    }
    case SHIFU_REW_TORQUES: {
      float acc = 0.0f;
#pragma unroll
      for (int d = 0; d < A1_DOF; ++d) acc = add_rn(acc, mul_rn(in.tau[e][d], in.tau[e][d]));
      return mul_rn(p0, acc);
    }
    default:
      return 0.0f;
  }
}

// Processes tiles [0, num_tiles) of 32 envs each (the ragged tail, if any, is a separate launch of
// the barrier-phased kernel).  Requires root_stride == 1 and 16-byte aligned tensors.
// HAS_EXTRA: the term list holds a code >= SHIFU_REW_LIN_VEL_Z.  The reference's own six terms run the
// instantiation without that code (its presence alone cost the B group 5-10 % of the kernel).
template <bool EXACT_DIV, bool HAS_MROW, bool HAS_EXTRA>
__global__ void __launch_bounds__(V3_THREADS, V3_CTAS_PER_SM)
a1_post_physics_tma_kernel(const __grid_constant__ A1K k, const __grid_constant__ ShifuA1StepIO io, int num_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  V3Smem& s = *reinterpret_cast<V3Smem*>(smem_raw);
#ifdef V3_SCAN_FIRST
  // warp order scan | B | DMA: the hardware arbiter favours high warp ids, which puts the
  // latency-critical B group ahead of the throughput-bound scan warps
  const int t = (threadIdx.x < V3_C_THREADS) ? (int)threadIdx.x + V3_B_THREADS
              : ((threadIdx.x < V3_C_THREADS + V3_B_THREADS) ? (int)threadIdx.x - V3_C_THREADS : (int)threadIdx.x);
#else
  const int t = threadIdx.x;
#endif
  const long long step = (io.step_dev != nullptr) ? *io.step_dev : io.step;
  const int first = blockIdx.x, stride = gridDim.x;
  const int my_tiles = (first < num_tiles) ? (num_tiles - first + stride - 1) / stride : 0;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < V3_STAGES; ++b) {
      pipe::mbar_init(&s.full_in[b], 1);
      // every thread of the producing group arrives itself: no group barrier, fast warps move on
      pipe::mbar_init(&s.e_done[b], 32);                     // yaw / xy of the tile's envs are in the ring
      pipe::mbar_init(&s.b_done[b], V3_B2_THREADS);
      pipe::mbar_init(&s.h_done[b], V3_C_THREADS);
    }
    pipe::fence_barrier_init();
  }
  __syncthreads();

  // =========================================================================================
  if (t >= V3_DMA_BASE && t < V3_DMA_BASE + V3_DMA_THREADS) {
    // ---------------- DMA warp ----------------
    if (t != V3_DMA_BASE) return;
    for (int j = 0; j < V3_STAGES && j < my_tiles; ++j) {
      V3_STAMP(s.t_issue[j]);
      v3_issue_loads(s.in[j], k, io, (long long)(first + j * stride) * A1_TILE, &s.full_in[j], true);
    }
    V3_T0(true);
    for (int j = 0; j < my_tiles; ++j) {
      const int b = j % V3_STAGES;
      const uint32_t par = (j / V3_STAGES) & 1;
      const long long e0 = (long long)(first + j * stride) * A1_TILE;
      // input rows are dead once the head / history phase is over: store the pushed history and
      // refill the stage right away, long before the tile's height scan finishes
      V3_WAIT(V3_POLL_DMA, &s.h_done[b], par);
      V3_TICK(10);
#ifdef V3_WI_DDELAY
      v3_delay(V3_WI_DDELAY);
#endif
      pipe::bulk_store(io.history + e0 * (A1_DOF * A1_HIST), s.in[b].hist, V3_HIST_BYTES);
      if (io.carry_body_frame) {                              // robot.py:222-229 for the next control step (D7)
        pipe::bulk_store(io.base_lin_vel + e0 * 3, s.in[b].cout[0], sizeof(s.in[b].cout[0]));
        pipe::bulk_store(io.base_ang_vel + e0 * 3, s.in[b].cout[1], sizeof(s.in[b].cout[1]));
        pipe::bulk_store(io.projected_gravity + e0 * 3, s.in[b].cout[2], sizeof(s.in[b].cout[2]));
      }
      pipe::bulk_commit();
      V3_STAMP(s.t_issue[b]);
      // every other row of the stage is dead already: refill it while the store still reads hist
      const long long en = (long long)(first + (j + V3_STAGES) * stride) * A1_TILE;
      if (j + V3_STAGES < my_tiles) v3_issue_loads(s.in[b], k, io, en, &s.full_in[b], false);
      pipe::bulk_wait_read_all();
      if (j + V3_STAGES < my_tiles) v3_issue_hist_load(s.in[b], io, en, &s.full_in[b]);
      V3_TICK(11);
      V3_COUNT(12);
    }
    pipe::bulk_wait_all();                                    // global writes done before exit
    return;
  }

  if (t < V3_B_THREADS) {
    // ---------------- B groups: lane = env ----------------
    const int g = t / V3_BG_THREADS, tg = t % V3_BG_THREADS;   // group g handles tiles j = g, g+2, ...
    const int warp = tg >> 5, lane = tg & 31;
    V3_T0(g == 0 && lane == 0 && warp < 2);

    for (int j = g; j < my_tiles; j += V3_B_GROUPS) {
      const int b = j % V3_STAGES;
      const uint32_t par = (j / V3_STAGES) & 1;
      const long long e0 = (long long)(first + j * stride) * A1_TILE;
      const long long ge = e0 + lane;
      V3In& in = s.in[b];
      const int rb = j & 3;                                   // scalar ring stage (see V3Smem)
      V3_TICK(warp == 0 ? 0 : 4);
      V3_WAIT(V3_POLL_B, &s.full_in[b], par);                 // tile rows + per-env scalars have landed
      if (warp == 1) V3_SINCE(19, s.t_issue[b]);
      V3_TICK(warp == 0 ? 1 : 5);

      // ---- B1: yaw normalisation for the scan first (warp 0; the scan group starts its index
      //          arithmetic as soon as e_done completes), then the reward terms, split over the two
      //          warps by the host-side cost balance k.term_warp
      if (warp == 0) {
        // heights are measured at the PRE-reset pose (isaac_gym.py:320-322 runs before post_step)
        const ScanEnv ev = make_scan_env(in.root[lane]);
        const int q = lane >> 1, sl = lane & 1;
        float* a = reinterpret_cast<float*>(&s.sA[rb][q]);
        float* bb = reinterpret_cast<float*>(&s.sB[rb][q]);
        float* cc = reinterpret_cast<float*>(&s.sC[rb][q]);
        a[sl] = ev.z2; a[2 + sl] = ev.z;
        bb[sl] = ev.w; bb[2 + sl] = ev.x;
        cc[sl] = ev.y;
        pipe::mbar_arrive(&s.e_done[b]);
      }
#pragma unroll 1
      for (int i = 0; i < k.term_count[warp]; ++i) {
        const int q = k.term_list[warp][i];
#ifdef V3_WI_NOB1
        s.rterm[j & 1][q][lane] = 0.0f;
#else
        s.rterm[j & 1][q][lane] = v3_eval_term<HAS_EXTRA>(k.terms[q], q, k.rp[q][0], k.rp[q][1], k, in, lane, io, ge);
#endif
      }
      V3_TICK(warp == 0 ? 2 : 6);
      pipe::named_barrier(1 + g, V3_BG_THREADS);
      V3_TICK(warp == 0 ? 3 : 7);
      if (warp >= 2) continue;                                 // a terms-only warp: on to the next tile's terms

      // ---- B2: both warps decide termination (a1_conditional.py:146-150); warp 0 does the ordered
      //          accumulation, flags and episode bookkeeping, warp 1 the state rewrite of resetting envs
      long long len = in.ep_len[lane] + 1;                                 // env.py:95
      const float* fb = &in.contact[lane][k.base_body * 3];
      const bool contact_term = fma_rn(fb[2], fb[2], fma_rn(fb[1], fb[1], mul_rn(fb[0], fb[0]))) > k.contact_thr_sq;
      const bool time_out = len > k.max_len;
      const bool reset = contact_term | time_out;
      double st_sum[SHIFU_MAX_REWARD_TERMS];
#pragma unroll
      for (int q = 0; q < SHIFU_MAX_REWARD_TERMS; ++q) st_sum[q] = 0.0;
      long long level_delta = 0;
      float esum[SHIFU_MAX_REWARD_TERMS];
      float cmd[3];
      if (warp == 0) {
#pragma unroll
        for (int q = 0; q < SHIFU_MAX_REWARD_TERMS; ++q) esum[q] = (q < k.n_terms) ? in.esum[q][lane] : 0.0f;
        float rew = 0.0f;                                                  // env.py:180-185
#pragma unroll
        for (int q = 0; q < SHIFU_MAX_REWARD_TERMS; ++q) {
          if (q < k.n_terms) {
            const float r = s.rterm[j & 1][q][lane];
            esum[q] = add_rn(esum[q], r);
            rew = add_rn(rew, r);
          }
        }
        io.rew_buf[ge] = rew;
        io.reset_buf[ge] = reset ? 1 : 0;
        io.time_out_buf[ge] = time_out ? 1 : 0;
        io.contact_term_buf[ge] = contact_term ? 1 : 0;
        if (reset)                                                         // env.py:101-102, bookkeeping part
          a1_reset_env<true, 2>(k, io, step, (int)ge, nullptr, nullptr, nullptr, cmd, esum, len, st_sum, level_delta,
                                0.0f, 0.0f, 0.0f, 0, 0);
        io.ep_len[ge] = len;
#pragma unroll
        for (int q = 0; q < SHIFU_MAX_REWARD_TERMS; ++q)
          if (q < k.n_terms) io.ep_sums[q][ge] = esum[q];
        a1_log_sums<2>(k, reset, st_sum, level_delta, lane);
      } else {
        if (reset) {                                                       // state part
          cmd[0] = in.cla[0][lane][0]; cmd[1] = in.cla[0][lane][1]; cmd[2] = in.cla[0][lane][2];
          a1_reset_env<true, 1, HAS_EXTRA>(k, io, step, (int)ge, in.root[lane], in.dof[lane], in.hist[lane], cmd, esum, len,
                                st_sum, level_delta, io.env_origins[ge * 3 + 0], io.env_origins[ge * 3 + 1],
                                io.env_origins[ge * 3 + 2], k.curriculum ? io.terrain_levels[ge] : 0,
                                k.curriculum ? io.terrain_types[ge] : 0);     // ~1 % of the envs: not worth a stage row
          in.cla[0][lane][0] = cmd[0]; in.cla[0][lane][1] = cmd[1]; in.cla[0][lane][2] = cmd[2];
        }
        reinterpret_cast<float*>(&s.sC[rb][lane >> 1])[2 + (lane & 1)] =
            sub_rn(in.root[lane][2], k.h_off);                              // post-reset base z (D8)
        if (io.carry_body_frame) {
          // carried body-frame velocities for the next control step (robot.py:222-229, D7), from the
          // post-reset root row of this lane's env; the rows leave with the DMA lane's bulk stores
          const float* r = in.root[lane];
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            float o[3];
            rotate_inverse(r + 3, v == 2 ? 0.0f : r[7 + 3 * v], v == 2 ? 0.0f : r[8 + 3 * v],
                           v == 2 ? -1.0f : r[9 + 3 * v], o);
            in.cout[v][lane][0] = o[0]; in.cout[v][lane][1] = o[1]; in.cout[v][lane][2] = o[2];
          }
          pipe::fence_proxy_async();                           // -> visible to the bulk stores
        }
        a1_log_sums<1>(k, reset, st_sum, level_delta, lane);
      }
#ifdef V3_WI_BDELAY
      v3_delay(V3_WI_BDELAY);
#endif
      // post-reset rows / command / zb are final: hand the tile to the scan group
      pipe::mbar_arrive(&s.b_done[b]);
      V3_TICK(warp == 0 ? 13 : 14);
      V3_COUNT(warp == 0 ? 15 : 31);
    }
    return;
  }

  // ---------------- scan group: thread = scan point ----------------
  {
    const int p = t - V3_SCAN_BASE;                          // index in the scan group
    const int sw = p >> 5, lane = p & 31;
    const float hclip = fminf(k.h_clip, k.clip_obs);         // clip(clip(v,+-a),+-b) == clip(v,+-min(a,b))
    const unsigned max_px = (unsigned)(k.trows - 1), max_py = (unsigned)(k.tcols - 1);
    // banded index: (px>>3)*8*W + py*8 + (px&7)  ==  (px & ~7)*(W-1) + px + 8*py    (3 integer ops)
    const unsigned c1 = (unsigned)(k.band_w - 1);
    const short* __restrict__ table = k.table;
    const f2_t BORDER = pk(k.border, k.border);
    const f2_t RCP = pk(k.hdiv.r, k.hdiv.r), NEGD = pk(-k.hdiv.d, -k.hdiv.d), VS = pk(k.vscale, k.vscale);
    const f2_t NZ = pk(k.neg_zero, k.neg_zero);
#ifdef V3_ADD_AS_FMA
    // RN(a + b) == fma(a, 1, b) and RN(a - b) == fma(b, -1, a) exactly
    const f2_t ONE = pk(k.one, k.one), NEG1 = pk(-k.one, -k.one);
#define ADD2(a, b) fma2(a, ONE, b)
#define SUB2(a, b) fma2(b, NEG1, a)
#define MUL2(a, b) fma2(a, b, NZ)
#elif defined(V3_ADD_SCALAR)
    // packed adds split into two scalar adds (the light FMA pipe) — dev experiment
    auto sadd2 = [](f2_t a, f2_t b) { float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1); return pk(add_rn(a0, b0), add_rn(a1, b1)); };
    auto ssub2 = [](f2_t a, f2_t b) { float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1); return pk(sub_rn(a0, b0), sub_rn(a1, b1)); };
#define ADD2(a, b) sadd2(a, b)
#define SUB2(a, b) ssub2(a, b)
#define MUL2(a, b) mul2(a, b)
#else
#define ADD2(a, b) add2(a, b)
#define SUB2(a, b) sub2(a, b)
#define MUL2(a, b) mul2(a, b)
#endif
#ifndef V3_F2I_CVT
    const f2_t DENORM = pk(__int_as_float(1), __int_as_float(1));     // 2^-149
#endif
    // Work items of a tile.  The measured-point grid is symmetric about the base (point 186 - pt is
    // -point pt) and round-to-nearest is sign-symmetric, so the yaw rotation of the mirror point is
    // EXACTLY the negated rotation of the point: a lane owns a point AND its mirror and rotates once.
    // Item = (group of 32 point pairs [3 groups cover the 94 pairs], batch of 4 envs): warp w owns pair
    // group w % 3 for the env pairs 2*(V3_ITEMS*(w/3) + it), +1.  (The host routes grids that are not
    // point-symmetric to the phased kernel.)
    static_assert(V3_SCAN_WARPS % 3 == 0 && 8 % (V3_SCAN_WARPS / 3) == 0, "scan warps: 3 pair groups x rows");
    constexpr int V3_ITEMS = 8 / (V3_SCAN_WARPS / 3);
    V3_T0(p == 0);
    for (int j = 0; j < my_tiles; ++j) {
      const int b = j % V3_STAGES;
      const uint32_t par = (j / V3_STAGES) & 1;
      const long long e0 = (long long)(first + j * stride) * A1_TILE;
      const int rb = j & 3;
      V3_WAIT(V3_POLL_SCAN, &s.e_done[b], par);
      V3_TICK(16);
      {
        // item -> (scan point of this lane, first env pair of the batch)
        struct Item { float bx, by; int q0, pt, it; bool live; };
        auto item = [&](int it) {
          Item r;
          const int raw = 32 * (sw % 3) + lane;
          r.live = raw <= A1_POINTS / 2;                      // pairs 0..93; 93 is the centre, its own mirror
          r.pt = r.live ? raw : A1_POINTS / 2;                // idle lanes shadow the centre, stores masked
          r.q0 = 2 * (V3_ITEMS * (sw / 3) + it);
          r.bx = k.px[r.pt % A1_NX]; r.by = k.py[r.pt / A1_NX];
          r.it = it;
          return r;
        };
        // Index arithmetic of one item in packed fp32x2 (two envs per instruction).
        // cell index of the packed pair of positions (ax, ay) [already + base xy + border]: the constant
        // division, .long() + clip and the banded table offset (isaac_gym.py:416-425)
        auto cell_pair = [&](f2_t ax, f2_t ay, unsigned& i0, unsigned& i1) {
          unsigned px0, px1, py0, py1;
          if (EXACT_DIV) {
            float a0, a1, b0, b1;
            upk(ax, a0, a1); upk(ay, b0, b1);
            // .long() + clip (isaac_gym.py:421-425): float->uint truncates and saturates at 0
            px0 = min(__float2uint_rz(div_rn(a0, k.hdiv.d)), max_px);
            px1 = min(__float2uint_rz(div_rn(a1, k.hdiv.d)), max_px);
            py0 = min(__float2uint_rz(div_rn(b0, k.hdiv.d)), max_py);
            py1 = min(__float2uint_rz(div_rn(b1, k.hdiv.d)), max_py);
          } else {                         // q0 = x*r; e = fma(-d, q0, x); q = fma(e, r, q0)
            const f2_t qx = MUL2(ax, RCP), qy = MUL2(ay, RCP);
            const f2_t fx = fma2(fma2(NEGD, qx, ax), RCP, qx), fy = fma2(fma2(NEGD, qy, ay), RCP, qy);
#ifndef V3_F2I_CVT
            // .long() + clip without the conversion unit: RZ(f * 2^-149) is the denormal whose bit
            // pattern IS trunc(f) for 0 <= f < 2^23 (negative f -> sign bit -> relu -> 0, larger f
            // -> a normal number >= 2^23 -> upper clip); one packed multiply per env pair
            int ix0, ix1, iy0, iy1;
            upk_i(mulrz2(fx, DENORM), ix0, ix1);
            upk_i(mulrz2(fy, DENORM), iy0, iy1);
            px0 = (unsigned)__vimin_s32_relu(ix0, (int)max_px); px1 = (unsigned)__vimin_s32_relu(ix1, (int)max_px);
            py0 = (unsigned)__vimin_s32_relu(iy0, (int)max_py); py1 = (unsigned)__vimin_s32_relu(iy1, (int)max_py);
#else
            float fx0, fx1, fy0, fy1;
            upk(fx, fx0, fx1); upk(fy, fy0, fy1);
            px0 = min(__float2uint_rz(fx0), max_px); px1 = min(__float2uint_rz(fx1), max_px);
            py0 = min(__float2uint_rz(fy0), max_py); py1 = min(__float2uint_rz(fy1), max_py);
#endif
          }
          i0 = ((px0 & ~7u) * c1 + px0) + (py0 << 3);
          i1 = ((px1 & ~7u) * c1 + px1) + (py1 << 3);
        };
        // Index arithmetic of one item in packed fp32x2 (two envs per instruction).
        auto index_batch = [&](const Item& w, unsigned (&idx)[8]) {
          const f2_t BX = pk(w.bx, w.bx), BY = pk(w.by, w.by), NBY = pk(-w.by, -w.by);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 a = s.sA[rb][w.q0 + u], bq = s.sB[rb][w.q0 + u];
            const float2 cq = *reinterpret_cast<const float2*>(&s.sC[rb][w.q0 + u]);
            const f2_t Z2 = pk(a.x, a.y), Z = pk(a.z, a.w), W = pk(bq.x, bq.y), X = pk(bq.z, bq.w);
            const f2_t Y = pk(cq.x, cq.y);
            // quat_apply_yaw (shifu/utils/terrain.py:202-206) on (bx, by, 0):
            //   t = 2(q x b) = (-2z*by, 2z*bx);  out = (b + w*t) + q x t,  q x t = (-z*ty, z*tx)
            // Products that feed an add are written fma(a, b, -0): ptxas contracts
            // mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (single rounding) even under --fmad=false,
            // which flips ~0.3 % of the cell indices; fma(a, b, -0) == RN(a*b) exactly and cannot be
            // contracted again (NZ comes from a kernel parameter, so it is not constant-folded).
            const f2_t tx = MUL2(Z2, NBY), ty = MUL2(Z2, BX);
            const f2_t rx = SUB2(ADD2(BX, fma2(W, tx, NZ)), fma2(Z, ty, NZ));
            const f2_t ry = ADD2(ADD2(BY, fma2(W, ty, NZ)), fma2(Z, tx, NZ));
            // + base xy, + border (isaac_gym.py:416-420)
            cell_pair(ADD2(ADD2(rx, X), BORDER), ADD2(ADD2(ry, Y), BORDER), idx[4 * u], idx[4 * u + 1]);
            // the mirror point: rotation = (-rx, -ry) exactly, and RN(-r + X) == RN(X - r)
            cell_pair(ADD2(SUB2(X, rx), BORDER), ADD2(SUB2(Y, ry), BORDER), idx[4 * u + 2], idx[4 * u + 3]);
          }
        };
        auto load_batch = [&](const unsigned (&idx)[8], int (&h)[8]) {
#pragma unroll
          for (int u = 0; u < 8; ++u) h[u] = v3_gather(table + idx[u]);      // isaac_gym.py:427-431 (folded)
        };
        auto store_batch = [&](const Item& w, const int (&h)[8]) {
          float* ob = io.obs_buf + (e0 + 2 * w.q0) * A1_OBS + A1_HEAD + w.pt;
          float* mb = HAS_MROW ? io.measured_heights + (e0 + 2 * w.q0) * A1_POINTS + w.pt : nullptr;
          // one packed pair of cells -> (z - 0.5) - h*vertical_scale, clipped, to rows r and r+1
          auto emit = [&](float2 zb, int h0, int h1, float* o, float* m) {
            const f2_t hg2 = fma2(pk((float)h0, (float)h1), VS, NZ);            // * vertical_scale, :433
            float v0, v1;
            upk(SUB2(pk(zb.x, zb.y), hg2), v0, v1);                             // a1_conditional.py:132
            v3_store(o, clampf(v0, -hclip, hclip));
            v3_store(o + A1_OBS, clampf(v1, -hclip, hclip));
            if (HAS_MROW) {
              float g0, g1;
              upk(hg2, g0, g1);
              __stcs(m, g0);
              __stcs(m + A1_POINTS, g1);
            }
          };
          if (w.live) {
            const int mir = (A1_POINTS - 1) - 2 * w.pt;                         // column of the mirror point
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float2 zb = *reinterpret_cast<const float2*>(&s.sC[rb][w.q0 + u].z);
              emit(zb, h[4 * u], h[4 * u + 1], ob + (2 * u) * A1_OBS, mb + (2 * u) * A1_POINTS);
              emit(zb, h[4 * u + 2], h[4 * u + 3], ob + (2 * u) * A1_OBS + mir, mb + (2 * u) * A1_POINTS + mir);
            }
          }
        };
        // Rolled software pipeline (the body stays small enough for the instruction cache shared with
        // the other warp roles): the gathers of item i are issued at the END of a trip and consumed
        // in the MIDDLE of the next one, after the index arithmetic of item i+1 — so their L1/L2
        // latency hides under ~200 instructions of independent math, with no register copies.  The
        // first item's gathers are issued BEFORE the head phase, which hides them as well.
        unsigned idx[8];
        int h[8];    // sign-extended int16 cells
        Item cur = item(0);
#ifndef V3_WI_NOSCAN
        index_batch(cur, idx);
        load_batch(idx, h);
#endif
        // ---- obs head (a1_conditional.py:131-144) + history push (train.py:12-14): post-reset rows
        V3_TICK(20);
        V3_WAIT(V3_POLL_SCAN, &s.b_done[b], par);
        V3_TICK(8);
        {
          // One (env, dof) item and one (env, column) item per thread.  Everything the head reads from the
          // stage is loaded FIRST and the history is pushed, then the stage is handed back (fence + arrive),
          // and only then are the 6 observation values clamped and stored: the MEMBAR behind
          // fence.proxy.async waits for every memory operation the thread has in flight, so with the
          // global stores issued before it the stage release waited for their acknowledgements.
          static_assert(A1_TILE * A1_DOF == V3_C_THREADS && A1_TILE * 12 == V3_C_THREADS, "one item per scan thread");
          V3In& in = s.in[b];
          const float c = k.clip_obs;
          const int e = p / A1_DOF, d = p - e * A1_DOF;                    // e, d serve both items (12 columns each)
          float* hrow = io.obs_buf + (e0 + e) * A1_OBS;
          const float2 qd = *reinterpret_cast<const float2*>(&in.dof[e][2 * d]);
          const float a0 = in.hist[e][d * A1_HIST + 0], a1 = in.hist[e][d * A1_HIST + 1],
                      a2 = in.hist[e][d * A1_HIST + 2];
          const float v = (d < 9) ? in.cla[d / 3][e][d % 3] : ((d == 11) ? -1.0f : 0.0f);   // command, velocities, gravity_vec
          in.hist[e][d * A1_HIST + 2] = a1;              // HistoryRecorder.add (train.py:12-14)
          in.hist[e][d * A1_HIST + 1] = a0;
          in.hist[e][d * A1_HIST + 0] = in.act[e][d];
          V3_TICK(21);
          pipe::fence_proxy_async();                            // pushed history -> visible to the TMA store
          pipe::mbar_arrive(&s.h_done[b]);
          V3_TICK(9);
#ifndef V3_WI_NOHEAD
#define V3_HEAD_ST(ptr, v) __stcs(ptr, v)
          V3_HEAD_ST(hrow + d, clampf(v, -c, c));
          V3_HEAD_ST(hrow + 12 + d, clampf(sub_rn(qd.x, k.q0[d]), -c, c));
          V3_HEAD_ST(hrow + 24 + d, clampf(qd.y, -c, c));
          V3_HEAD_ST(hrow + 36 + d, clampf(a0, -c, c));         // HistoryRecorder.flatten: slot-major
          V3_HEAD_ST(hrow + 48 + d, clampf(a1, -c, c));
          V3_HEAD_ST(hrow + 60 + d, clampf(a2, -c, c));
#endif
          V3_TICK(22);
        }
#ifdef V3_WI_SDELAY
        v3_delay(V3_WI_SDELAY);
#endif
#ifndef V3_WI_NOSCAN
#pragma unroll 1
        for (int it = 1; it < V3_ITEMS; ++it) {
          const Item nxt = item(it);
          index_batch(nxt, idx);
          store_batch(cur, h);
          load_batch(idx, h);
          cur = nxt;
        }
        store_batch(cur, h);
#endif
      }
      V3_TICK(17);
      V3_COUNT(18);
    }
  }
}

}  // namespace shifu
