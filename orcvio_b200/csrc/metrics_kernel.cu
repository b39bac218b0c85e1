// Trajectory metrics on the device: the reference's logger, System::publishGroundtruth
// (ros_wrapper/src/orcvio/src/System.cpp:885-943), for a batch of trajectories at once.
//
// Per trajectory: the first estimated pose is aligned with the first ground-truth pose,
//     T_from_est_to_gt = T_gt(0) T_est(0)^-1,        T_corrected(k) = T_from_est_to_gt T_est(k),
// and every frame contributes  rmse_pos = |p_corrected - p_gt|  and
// rmse_ori = (180 / pi) 2 |vec(q_corrected (x) q_gt^-1)|  (Hamilton quaternions x, y, z, w: math_utils.hpp:79-93,
// 164-227); the logger's result file holds the running means of the two.  One CTA per trajectory, frames over the
// threads, fixed-order tree reduction.  Outputs per trajectory: mean orientation error (deg), mean position error (m),
// position RMSE (m), final position error (m).
#include "../../include/orcvio_b200.h"
#include "kernels.h"

namespace ob {

namespace {

__device__ __forceinline__ void quat_to_R(const double* q, double* R) {     // quaternionToRotation, math_utils.hpp:164-177
  const double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
  R[0] = 1 - 2 * (qy * qy + qz * qz); R[1] = 2 * (qx * qy - qw * qz); R[2] = 2 * (qx * qz + qw * qy);
  R[3] = 2 * (qx * qy + qw * qz); R[4] = 1 - 2 * (qx * qx + qz * qz); R[5] = 2 * (qy * qz - qw * qx);
  R[6] = 2 * (qx * qz - qw * qy); R[7] = 2 * (qy * qz + qw * qx); R[8] = 1 - 2 * (qx * qx + qy * qy);
}

__device__ __forceinline__ void R_to_quat(const double* R, double* q) {     // rotationToQuaternion, math_utils.hpp:188-227
  const double tr = R[0] + R[4] + R[8];
  const double score[4] = {R[0], R[4], R[8], tr};
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (score[i] > score[best]) best = i;
  if (best == 0) {
    q[0] = sqrt(1 + 2 * R[0] - tr) / 2.0;
    q[1] = (R[1] + R[3]) / (4 * q[0]);
    q[2] = (R[2] + R[6]) / (4 * q[0]);
    q[3] = (R[7] - R[5]) / (4 * q[0]);
  } else if (best == 1) {
    q[1] = sqrt(1 + 2 * R[4] - tr) / 2.0;
    q[0] = (R[1] + R[3]) / (4 * q[1]);
    q[2] = (R[5] + R[7]) / (4 * q[1]);
    q[3] = (R[2] - R[6]) / (4 * q[1]);
  } else if (best == 2) {
    q[2] = sqrt(1 + 2 * R[8] - tr) / 2.0;
    q[0] = (R[2] + R[6]) / (4 * q[2]);
    q[1] = (R[5] + R[7]) / (4 * q[2]);
    q[3] = (R[3] - R[1]) / (4 * q[2]);
  } else {
    q[3] = sqrt(1 + tr) / 2.0;
    q[0] = (R[7] - R[5]) / (4 * q[3]);
    q[1] = (R[2] - R[6]) / (4 * q[3]);
    q[2] = (R[3] - R[1]) / (4 * q[3]);
  }
  if (q[3] < 0)
    for (int i = 0; i < 4; ++i) q[i] = -q[i];
  const double n = sqrt((q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]));
  for (int i = 0; i < 4; ++i) q[i] /= n;
}

}  // namespace

// est / gt: n_traj x n_frames x 7 (px py pz qx qy qz qw); out: n_traj x 4
__global__ void __launch_bounds__(128) k_trajectory_metrics(const double* __restrict__ est, const double* __restrict__ gt,
                                                            int n_frames, double* __restrict__ out) {
  __shared__ double red[3][128];
  __shared__ double Ta[12];                              // T_from_est_to_gt: R (9) | t (3)
  const int tr = blockIdx.x, tid = threadIdx.x;
  const double* e = est + (size_t)tr * n_frames * 7;
  const double* g = gt + (size_t)tr * n_frames * 7;
  if (tid == 0) {
    double Re[9], Rg[9];
    quat_to_R(e + 3, Re);
    quat_to_R(g + 3, Rg);
    m3_mulT(Rg, Re, Ta);                                 // R_gt R_est^T
    double t[3];
    m3_vec(Ta, e, t);
    for (int i = 0; i < 3; ++i) Ta[9 + i] = g[i] - t[i];  // p_gt - R p_est
  }
  __syncthreads();
  double s_ori = 0.0, s_pos = 0.0, s_sq = 0.0;
  for (int k = tid; k < n_frames; k += 128) {
    double Re[9], Rc[9], pc[3], qc[4];
    quat_to_R(e + 7 * k + 3, Re);
    m3_mul(Ta, Re, Rc);
    m3_vec(Ta, e + 7 * k, pc);
    const double dx = pc[0] + Ta[9] - g[7 * k], dy = pc[1] + Ta[10] - g[7 * k + 1], dz = pc[2] + Ta[11] - g[7 * k + 2];
    const double ep = sqrt(dx * dx + dy * dy + dz * dz);
    R_to_quat(Rc, qc);
    // q_corrected (x) q_gt^-1 (Hamilton product, math_utils.hpp:79-93; the inverse of a unit quaternion: conjugate / |q|^2)
    const double* qg = g + 7 * k + 3;
    const double n2 = (qg[0] * qg[0] + qg[1] * qg[1]) + (qg[2] * qg[2] + qg[3] * qg[3]);
    const double b0 = -qg[0] / n2, b1 = -qg[1] / n2, b2 = -qg[2] / n2, b3 = qg[3] / n2;
    double d0 = qc[3] * b0 - qc[2] * b1 + qc[1] * b2 + qc[0] * b3;
    double d1 = qc[2] * b0 + qc[3] * b1 - qc[0] * b2 + qc[1] * b3;
    double d2 = -qc[1] * b0 + qc[0] * b1 + qc[3] * b2 + qc[2] * b3;
    double d3 = -qc[0] * b0 - qc[1] * b1 - qc[2] * b2 + qc[3] * b3;
    const double dn = sqrt((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
    d0 /= dn; d1 /= dn; d2 /= dn;
    const double eo = (180.0 / 3.14159265358979323846) * 2 * sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    s_ori += eo;
    s_pos += ep;
    s_sq += ep * ep;
    if (k == n_frames - 1) out[4 * (size_t)tr + 3] = ep;
  }
  red[0][tid] = s_ori; red[1][tid] = s_pos; red[2][tid] = s_sq;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o)
      for (int q = 0; q < 3; ++q) red[q][tid] += red[q][tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    out[4 * (size_t)tr + 0] = red[0][0] / n_frames;
    out[4 * (size_t)tr + 1] = red[1][0] / n_frames;
    out[4 * (size_t)tr + 2] = sqrt(red[2][0] / n_frames);
  }
}

int trajectory_metrics(const double* est_pose7, const double* gt_pose7, int n_traj, int n_frames, double* out4) {
  if (n_traj < 1 || n_frames < 1) return ORCVIO_ERR_ARG;
  const size_t nb = (size_t)n_traj * n_frames * 7 * sizeof(double);
  double *dE = nullptr, *dG = nullptr, *dO = nullptr;
  if (cudaMalloc(&dE, nb) != cudaSuccess || cudaMalloc(&dG, nb) != cudaSuccess ||
      cudaMalloc(&dO, (size_t)n_traj * 4 * sizeof(double)) != cudaSuccess) {
    cudaFree(dE); cudaFree(dG); cudaFree(dO);
    return ORCVIO_ERR_CUDA;
  }
  cudaMemcpy(dE, est_pose7, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(dG, gt_pose7, nb, cudaMemcpyHostToDevice);
  k_trajectory_metrics<<<n_traj, 128>>>(dE, dG, n_frames, dO);
  check_launch("k_trajectory_metrics");
  const cudaError_t e = cudaMemcpy(out4, dO, (size_t)n_traj * 4 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dE); cudaFree(dG); cudaFree(dO);
  return e == cudaSuccess ? ORCVIO_OK : ORCVIO_ERR_CUDA;
}

}  // namespace ob
