// ABB push-box prior stage (row a16): termination, the two reward terms, reset with device-side
// Philox draws (replacing the per-env Python np.random.uniform loop that is 96 % of the
// reference's step) and the 6-column observation — one thread per env, one pass.
// ~90 B/env: the 65 536-env config is launch-latency-bound, not bandwidth-bound (§8d).
#pragma once
#include "exact_math.cuh"
#include "philox.cuh"
#include "../../include/shifu_b200.h"

namespace shifu {

struct AbbK {
  int n;
  long long env_offset;
  unsigned long long seed;
  int n_actors, n_bodies, n_dof, ee_body;
  int robot_actor, table_actor, cube_actor, goal_actor;
  float min_xy[2], max_xy[2];
  float q0[SHIFU_MAX_DOF];
  float robot_root[7], table_root[7];
  double pos_low[3], pos_high[3], goal_z;
  long long max_len;
  float max_len_s, clip_obs, success_dist;
  int n_terms, terms[SHIFU_MAX_REWARD_TERMS];
  float rp[SHIFU_MAX_REWARD_TERMS][2];
  double* stats;
};

// RandPosBox._reset_root_state (a_prior_stage.py:39-51): numpy float64 draws -> fp32 row
__device__ __forceinline__ void abb_box_reset(float* row, const AbbK& k, long long gid, long long step,
                                              unsigned s_pos, unsigned s_eul, double z) {
  const U4 up = draw(k.seed, gid, step, s_pos);
  const U4 ue = draw(k.seed, gid, step, s_eul);
  row[0] = (float)(k.pos_low[0] + (k.pos_high[0] - k.pos_low[0]) * u01d(up.x));
  row[1] = (float)(k.pos_low[1] + (k.pos_high[1] - k.pos_low[1]) * u01d(up.y));
  row[2] = (float)(z + (z - z) * u01d(up.z));
  const double PI = 3.141592653589793;
  const float yaw = (float)(-PI + (PI - (-PI)) * u01d(ue.z));
  // quat_from_euler_xyz(0, 0, yaw): (0, 0, sin(yaw/2), cos(yaw/2))
  const float half = mul_rn(yaw, 0.5f);
  row[3] = 0.0f; row[4] = 0.0f; row[5] = sinf(half); row[6] = cosf(half);
#pragma unroll
  for (int j = 7; j < 13; ++j) row[j] = 0.0f;
}

// ShifuVecEnv.reset_idx for one env of the 4-actor scene (env.py:114-130, robot.py:74-77,
// units.py:130-134, a_prior_stage.py:39-51): robot dofs / root, table root, random cube and goal.
__device__ __forceinline__ void abb_reset_env(const AbbK& k, const ShifuAbbStepIO& io, long long step, int e,
                                              float* cube, float* goal) {
  const long long gid = k.env_offset + e;
  for (int d = 0; d < k.n_dof; ++d) {                               // robot.py:74-77
    io.dof_state[((long long)e * k.n_dof + d) * 2 + 0] = k.q0[d];
    io.dof_state[((long long)e * k.n_dof + d) * 2 + 1] = 0.0f;
    io.dof_targets[(long long)e * k.n_dof + d] = k.q0[d];
  }
  float* rob = io.root_state + ((long long)e * k.n_actors + k.robot_actor) * 13;   // units.py:130-134
  float* tab = io.root_state + ((long long)e * k.n_actors + k.table_actor) * 13;
#pragma unroll
  for (int j = 0; j < 13; ++j) {
    rob[j] = (j < 7) ? k.robot_root[j] : 0.0f;
    tab[j] = (j < 7) ? k.table_root[j] : 0.0f;
  }
  abb_box_reset(cube, k, gid, step, STREAM_CUBE_POS, STREAM_CUBE_EUL, k.pos_low[2]);
  abb_box_reset(goal, k, gid, step, STREAM_GOAL_POS, STREAM_GOAL_EUL, k.goal_z);
}

// Warp-level reduction of the per-step log sums -> one set of atomics per warp.
__device__ __forceinline__ void abb_log_sums(const AbbK& k, bool reset, bool success,
                                             const double (&st_sum)[SHIFU_MAX_REWARD_TERMS], int lane) {
  const unsigned any = __ballot_sync(0xffffffffu, reset);
  if (!any) return;
  double cnt = reset ? 1.0 : 0.0, suc = (reset && success) ? 1.0 : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    suc += __shfl_xor_sync(0xffffffffu, suc, o);
  }
#pragma unroll
  for (int j = 0; j < SHIFU_MAX_REWARD_TERMS; ++j) {
    if (j < k.n_terms) {
      double v = st_sum[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(k.stats + SHIFU_STAT_TERM0 + j, v);
    }
  }
  if (lane == 0) {
    atomicAdd(k.stats + SHIFU_STAT_NRESET, cnt);
    atomicAdd(k.stats + SHIFU_STAT_SUCCESS, suc);
  }
}

// Stand-alone reset_idx(env_ids) (ShifuVecEnv.reset, env.py:108-112, and user calls): one thread per
// id; ids == nullptr means arange(n_ids).  The logged success rate is the mean of the CURRENT
// success_buf over the ids (a_prior_stage.py:92-93).
__global__ void __launch_bounds__(128)
abb_reset_idx_kernel(const __grid_constant__ AbbK k, const __grid_constant__ ShifuAbbStepIO io,
                     const long long* __restrict__ ids, int n_ids) {
  const int lane = threadIdx.x & 31;
  const long long step = (io.step_dev != nullptr) ? *io.step_dev : io.step;
  for (int i0 = blockIdx.x * blockDim.x; i0 < n_ids; i0 += gridDim.x * blockDim.x) {
    const int i = i0 + threadIdx.x;
    bool reset = false, success = false;
    double st_sum[SHIFU_MAX_REWARD_TERMS];
#pragma unroll
    for (int j = 0; j < SHIFU_MAX_REWARD_TERMS; ++j) st_sum[j] = 0.0;
    if (i < n_ids) {
      const int e = (ids != nullptr) ? (int)ids[i] : i;
      if (e >= 0 && e < k.n) {
        reset = true;
        success = io.success_buf[e] != 0;
        abb_reset_env(k, io, step, e, io.root_state + ((long long)e * k.n_actors + k.cube_actor) * 13,
                      io.root_state + ((long long)e * k.n_actors + k.goal_actor) * 13);
        io.ep_len[e] = 0;
        io.reset_buf[e] = 1;
#pragma unroll
        for (int j = 0; j < SHIFU_MAX_REWARD_TERMS; ++j)
          if (j < k.n_terms) { st_sum[j] = (double)io.ep_sums[j][e]; io.ep_sums[j][e] = 0.0f; }
      }
    }
    abb_log_sums(k, reset, success, st_sum, lane);
  }
}

__global__ void __launch_bounds__(128)
abb_post_physics_kernel(const __grid_constant__ AbbK k, const __grid_constant__ ShifuAbbStepIO io) {
  const int lane = threadIdx.x & 31;
  const long long step = (io.step_dev != nullptr) ? *io.step_dev : io.step;
  for (int e0 = blockIdx.x * blockDim.x; e0 < k.n; e0 += gridDim.x * blockDim.x) {
    const int e = e0 + threadIdx.x;
    bool reset = false, success = false;
    double st_sum[SHIFU_MAX_REWARD_TERMS];
#pragma unroll
    for (int j = 0; j < SHIFU_MAX_REWARD_TERMS; ++j) st_sum[j] = 0.0;
    if (e < k.n) {
      float* cube = io.root_state + ((long long)e * k.n_actors + k.cube_actor) * 13;
      float* goal = io.root_state + ((long long)e * k.n_actors + k.goal_actor) * 13;
      const float* ee = io.body_state + ((long long)e * k.n_bodies + k.ee_body) * 13;
      float cx = cube[0], cy = cube[1], gx = goal[0], gy = goal[1];
      const float ex = __ldg(ee), ey = __ldg(ee + 1);
      long long len = io.ep_len[e] + 1;                                   // env.py:95
      // compute_termination, a_prior_stage.py:102-110
      const bool time_out = len > k.max_len;
      const float goal_dist = norm2_fma(sub_rn(gx, cx), sub_rn(gy, cy));
      const float ee_dist = norm2_fma(sub_rn(ex, cx), sub_rn(ey, cy));
      const bool obj_out = (cx < k.min_xy[0]) | (cy < k.min_xy[1]) | (cx > k.max_xy[0]) | (cy > k.max_xy[1]);
      const bool ee_out = (ex < k.min_xy[0]) | (ey < k.min_xy[1]) | (ex > k.max_xy[0]) | (ey > k.max_xy[1]);
      success = goal_dist < k.success_dist;                               // is_success, :129-131
      float rew = 0.0f;
#pragma unroll
      for (int j = 0; j < SHIFU_MAX_REWARD_TERMS; ++j) {
        if (j < k.n_terms) {
          const float p0 = k.rp[j][0], p1 = k.rp[j][1];
          float r = 0.0f;
          if (k.terms[j] == SHIFU_REW_ABB_REACHING) {                      // :118-123
            const float v = expf(div_rn(-mul_rn(goal_dist, goal_dist), p1));
            r = (ee_dist < p0) ? v : mul_rn(0.0f, v);
          } else if (k.terms[j] == SHIFU_REW_ABB_SUCCESS) {                // :125-131
            r = mul_rn((goal_dist < p1) ? 1.0f : 0.0f, p0);
          }
          const float es = add_rn(io.ep_sums[j][e], r);
          io.ep_sums[j][e] = es;
          st_sum[j] = (double)es;
          rew = add_rn(rew, r);
        }
      }
      reset = time_out | obj_out | ee_out | success;
      io.rew_buf[e] = rew;
      io.reset_buf[e] = reset ? 1 : 0;
      io.time_out_buf[e] = time_out ? 1 : 0;
      io.success_buf[e] = success ? 1 : 0;
      if (reset) {                                                        // env.py:114-130
        abb_reset_env(k, io, step, e, cube, goal);
        cx = cube[0]; cy = cube[1]; gx = goal[0]; gy = goal[1];
        len = 0;
#pragma unroll
        for (int j = 0; j < SHIFU_MAX_REWARD_TERMS; ++j)
          if (j < k.n_terms) io.ep_sums[j][e] = 0.0f;
      } else {
#pragma unroll
        for (int j = 0; j < SHIFU_MAX_REWARD_TERMS; ++j) st_sum[j] = 0.0;
      }
      io.ep_len[e] = len;
      // compute_observations (a_prior_stage.py:95-100) + clip (env.py:90): post-reset cube/goal, old ee
      const float c = k.clip_obs;
      float* o = io.obs_buf + (long long)e * 6;
      o[0] = clampf(cx, -c, c); o[1] = clampf(cy, -c, c); o[2] = clampf(gx, -c, c);
      o[3] = clampf(gy, -c, c); o[4] = clampf(ex, -c, c); o[5] = clampf(ey, -c, c);
    }
    abb_log_sums(k, reset, success, st_sum, lane);
  }
}

}  // namespace shifu
