// Row N2: pre-physics arm action path — optional action -> end-effector goal, then
// damped least-squares inverse kinematics.  thread = env; 6x6 SPD system solved by Cholesky in
// registers.  ~300 B/env of traffic and ~400 flops/env: HBM-/launch-bound, no tensor-core shape.
#pragma once
#include "exact_math.cuh"
#include "../../include/shifu_b200.h"

namespace shifu {

// isaacgym.torch_utils.quat_mul == shifu/utils/torch_utils.py:12-31 (same operation order)
__device__ __forceinline__ void quat_mul_ref(const float* a, const float* b, float* o) {
  const float x1 = a[0], y1 = a[1], z1 = a[2], w1 = a[3];
  const float x2 = b[0], y2 = b[1], z2 = b[2], w2 = b[3];
  const float ww = mul_rn(add_rn(z1, x1), add_rn(x2, y2));
  const float yy = mul_rn(sub_rn(w1, y1), add_rn(w2, z2));
  const float zz = mul_rn(add_rn(w1, y1), sub_rn(w2, z2));
  const float xx = add_rn(add_rn(ww, yy), zz);
  const float qq = mul_rn(0.5f, add_rn(xx, mul_rn(sub_rn(z1, x1), sub_rn(x2, y2))));
  o[3] = add_rn(sub_rn(qq, ww), mul_rn(sub_rn(z1, y1), sub_rn(y2, z2)));
  o[0] = add_rn(sub_rn(qq, xx), mul_rn(add_rn(x1, w1), add_rn(x2, w2)));
  o[1] = add_rn(sub_rn(qq, yy), mul_rn(sub_rn(w1, x1), add_rn(y2, z2)));
  o[2] = add_rn(sub_rn(qq, zz), mul_rn(add_rn(z1, y1), sub_rn(w2, x2)));
}

__global__ void __launch_bounds__(128)
arm_ik_kernel(const __grid_constant__ ShifuArmIkIO io, int n) {
  const int nd = io.num_dof;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const float* ee = io.body_state + ((long long)e * io.num_bodies + io.ee_body) * 13;
    const float ee_pos[3] = {ee[0], ee[1], ee[2]};
    const float ee_quat[4] = {ee[3], ee[4], ee[5], ee[6]};
    float goal_pos[3], goal_quat[4];
    if (io.actions != nullptr) {                          // AbbRobot.step, a_prior_stage.py:67-71
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float d = mul_rn(mul_rn(io.actions[e * 3LL + i], io.ee_velocity), io.dt);
        goal_pos[i] = clampf(add_rn(ee_pos[i], d), io.min_ee_pos[i], io.max_ee_pos[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) goal_quat[i] = io.tar_quat[i];
    } else {
      const float* g = io.goal_pose + e * 7LL;
#pragma unroll
      for (int i = 0; i < 3; ++i) goal_pos[i] = g[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) goal_quat[i] = g[3 + i];
    }
    // dpose = [goal_pos - ee_pos ; orientation_error]            robot.py:150-173
    float b[6];
#pragma unroll
    for (int i = 0; i < 3; ++i) b[i] = sub_rn(goal_pos[i], ee_pos[i]);
    const float cc[4] = {-ee_quat[0], -ee_quat[1], -ee_quat[2], ee_quat[3]};
    float qr[4];
    quat_mul_ref(goal_quat, cc, qr);
    const float sg = (qr[3] > 0.0f) ? 1.0f : ((qr[3] < 0.0f) ? -1.0f : 0.0f);    // torch.sign
#pragma unroll
    for (int i = 0; i < 3; ++i) b[3 + i] = mul_rn(qr[i], sg);
    // A = J J^T + damping^2 I (lower triangle), J = jacobian[e, ee_link] (6, nd) row-major
    const float* J = io.jacobian + ((long long)e * io.num_links + io.ee_link) * 6LL * nd;
    float A[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) A[r][c] = 0.0f;
    for (int k = 0; k < nd; ++k) {
      float col[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) col[r] = __ldg(J + r * nd + k);
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) A[r][c] = fma_rn(col[r], col[c], A[r][c]);
    }
    const float lam = mul_rn(io.damping, io.damping);
#pragma unroll
    for (int r = 0; r < 6; ++r) A[r][r] = add_rn(A[r][r], lam);
    // Cholesky A = L L^T (A is SPD: J J^T >= 0, + damping^2 I), then L y = b, L^T x = y
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float d = A[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) d = fma_rn(-A[j][k], A[j][k], d);
      d = __fsqrt_rn(d);
      A[j][j] = d;
      const float inv = __frcp_rn(d);
#pragma unroll
      for (int i = j + 1; i < 6; ++i) {
        float v = A[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) v = fma_rn(-A[i][k], A[j][k], v);
        A[i][j] = mul_rn(v, inv);
      }
    }
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float v = b[i];
#pragma unroll
      for (int k = 0; k < i; ++k) v = fma_rn(-A[i][k], y[k], v);
      y[i] = __fdiv_rn(v, A[i][i]);
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      float v = y[i];
#pragma unroll
      for (int k = i + 1; k < 6; ++k) v = fma_rn(-A[k][i], y[k], v);
      y[i] = __fdiv_rn(v, A[i][i]);
    }
    // u = J^T x ; dof_targets = dof_pos + u                         robot.py:178-182
    for (int k = 0; k < nd; ++k) {
      float u = 0.0f;
#pragma unroll
      for (int r = 0; r < 6; ++r) u = fma_rn(__ldg(J + r * nd + k), y[r], u);
      const float q = io.dof_state[((long long)e * nd + k) * 2];
      io.dof_targets[(long long)e * nd + k] = add_rn(q, u);
    }
  }
}

}  // namespace shifu
