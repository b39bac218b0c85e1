// incrementState_IMUCam (reference src/orcvio.cpp:4468-4567) as a CTA-wide device routine shared by
// every kernel that ends in a state increment (k_dx, k_apply_dx, k_zupt).
#pragma once
#include "kernels.h"

namespace ob {

// dxs: delta_x (D = 22 + 6N) in shared memory, already visible to the whole CTA.
// apply_log: optional slot that receives 1.0 / 0.0 (large-update guard, :4479-4494).
// Must be called by every thread of the CTA (contains barriers).  Returns the guard's verdict.
__device__ __forceinline__ bool cta_increment_state(const double* dxs, double* imu, double* clones, int N, int flags,
                                                    double* apply_log) {
  __shared__ int s_apply;
  const int tid = threadIdx.x;
  if (tid == 0) {
    double nv = sqrt((dxs[3] * dxs[3] + dxs[4] * dxs[4]) + dxs[5] * dxs[5]);
    double np = sqrt((dxs[6] * dxs[6] + dxs[7] * dxs[7]) + dxs[8] * dxs[8]);
    int apply = 1;
    if ((nv > 1.0 || np > 1.5) && (flags & FL_DISCARD_LARGE)) {
      apply = 0;
      imu[IM_DISCARDS] += 1.0;
    }
    s_apply = apply;
    if (apply_log) *apply_log = (double)apply;
    if (apply) {
      const bool left = (flags & FL_LARVIO) || (flags & FL_LEFT);
      double Rt[9], Rn[9];
      so3_exp(dxs, Rt);
      if (left) m3_mul(Rt, imu + IM_R, Rn);
      else m3_mul(imu + IM_R, Rt, Rn);
      for (int i = 0; i < 9; ++i) imu[IM_R + i] = Rn[i];
      for (int i = 0; i < 3; ++i) {
        imu[IM_V + i] += dxs[3 + i];
        imu[IM_P + i] += dxs[6 + i];
        imu[IM_BG + i] += dxs[9 + i];
        imu[IM_BA + i] += dxs[12 + i];
      }
      // extrinsics / td (:4513-4520); dx is exactly zero there unless estimated
      double dq[3] = {dxs[15] / 2.0, dxs[16] / 2.0, dxs[17] / 2.0};
      double n2 = (dq[0] * dq[0] + dq[1] * dq[1]) + dq[2] * dq[2];
      double qw, qs = 1.0;
      if (n2 <= 1) qw = sqrt(1 - n2);
      else { qw = 1; qs = 1.0 / sqrt(1 + n2); }
      double Rq[9], Rb[9];
      quat_wxyz_to_R(qw * qs, dq[0] * qs, dq[1] * qs, dq[2] * qs, Rq);
      m3_mulT(imu + IM_RBC, Rq, Rb);
      for (int i = 0; i < 9; ++i) imu[IM_RBC + i] = Rb[i];
      for (int i = 0; i < 3; ++i) imu[IM_TCB + i] += dxs[18 + i];
      imu[IM_TD] += dxs[21];
    }
  }
  __syncthreads();
  const bool apply = s_apply != 0;
  if (apply && tid < N) {
    const bool left = (flags & FL_LARVIO) || (flags & FL_LEFT);
    double* c = clones + (size_t)tid * CL_STRIDE;
    const double* d = dxs + ORCVIO_LEG + 6 * tid;
    double Rt[9], Rn[9];
    so3_exp(d, Rt);
    if (left) m3_mul(Rt, c + CL_R, Rn);
    else m3_mul(c + CL_R, Rt, Rn);
    for (int i = 0; i < 9; ++i) c[CL_R + i] = Rn[i];
    for (int i = 0; i < 3; ++i) c[CL_P + i] += d[3 + i];
    double Rc[9], t[3];
    m3_mulT(Rn, imu + IM_RBC, Rc);     // R_b2w * R_b2c^T
    m3_vec(Rn, imu + IM_TCB, t);
    for (int i = 0; i < 9; ++i) c[CL_RC + i] = Rc[i];
    for (int i = 0; i < 3; ++i) c[CL_PC + i] = c[CL_P + i] + t[i];
  }
  return apply;
}

}  // namespace ob
