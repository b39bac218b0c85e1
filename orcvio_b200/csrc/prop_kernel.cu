// Stage 6 -- IMU propagation of the mean and of the covariance, state augmentation and
// window pruning of P.
//
// Reference: OrcVIO::processModel (src/orcvio.cpp:727-823), predictNewStateOrcVIO
// (:899-928), predictNewStateLARVIO (:825-897), calPhiClosedForm (:3980-4370, LEG_DIM 22
// branch), stateAugmentation (:930-1013) and the row/column deletion at the end of
// pruneImuStateBuffer (:2875-2956).
//
// One CTA per filter.  The only rows of P that an IMU step changes are the 15 IMU rows
// (Phi is the identity on rows 15..21 and on the clone rows) and, by symmetry, the 15 IMU
// columns.  The CTA therefore keeps the 15 x D strip P[0:15, :] in shared memory across all
// IMU samples of the frame and touches HBM once per frame instead of once per sample
// (the reference symmetrises the full D x D matrix after every sample):
//   phase 1  thread 0 integrates the mean over the samples (sequential by nature),
//   phase 2  one thread per sample builds its 15 x 15 Phi,
//   phase 3  per sample: every thread owns one column of the strip (Phi * column), then the
//            15 x 15 corner gets its right factor Phi^T and the noise Q = Phi G Qc G^T Phi^T dt,
//   phase 4  strip written back, mirrored into the 15 IMU columns.
#include "kernels.h"

namespace ob {

constexpr int PS = 15;        // propagated rows: theta, v, p, bg, ba
constexpr int SMAX = 32;      // samples per chunk
constexpr int CTX = 40;       // per-sample context doubles

__device__ __forceinline__ void m3_scale_add(double* out, const double* A, double s) {
  for (int i = 0; i < 9; ++i) out[i] += s * A[i];
}

// phase 2: Phi for one sample.  ctx: dt, gyro(3), acc(3), R0(9), v0(3), p0(3), v1(3), p1(3), gyro_old(3)
__device__ void build_phi(const double* ctx, int flags, double* Phi /*15x15 row-major*/) {
  for (int i = 0; i < PS * PS; ++i) Phi[i] = 0.0;
  for (int i = 0; i < PS; ++i) Phi[i * PS + i] = 1.0;
  const double dt = ctx[0];
  const double* gyro = ctx + 1;
  const double* acc = ctx + 4;
  const double* C = ctx + 7;
  const double* vk = ctx + 16;
  const double* pk = ctx + 19;
  const double* vk1 = ctx + 22;
  const double* pk1 = ctx + 25;
  const double* gyro_old = ctx + 28;
  auto put = [&](int r0, int c0, const double* B, double s) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Phi[(r0 + i) * PS + c0 + j] = s * B[3 * i + j];
  };
  const double g[3] = {0.0, 0.0, -9.81};
  if ((flags & FL_LARVIO) || (flags & FL_LEFT)) {
    // :3997-4037 with Ma = Tg = I, As = 0, if_FEJ = false
    double cr[3] = {gyro_old[1] * gyro[2] - gyro_old[2] * gyro[1],
                    gyro_old[2] * gyro[0] - gyro_old[0] * gyro[2],
                    gyro_old[0] * gyro[1] - gyro_old[1] * gyro[0]};
    double aa[3];
    for (int i = 0; i < 3; ++i) aa[i] = dt * (gyro_old[i] + gyro[i]) / 2 + dt * dt * cr[i] / 12;
    double AA[9], I2A[9], T1[9], T2[9], T3[9], Sk[9];
    m3_skew(aa, AA);
    for (int i = 0; i < 9; ++i) I2A[i] = AA[i];
    I2A[0] += 2; I2A[4] += 2; I2A[8] += 2;
    m3_mul(C, I2A, T1);                                   // C (2I + AA)
    put(0, 9, T1, -0.5 * dt);                             // Phi_q_bg
    put(3, 12, T1, -0.5 * dt);                            // Phi_v_ba
    double d1[3];
    for (int i = 0; i < 3; ++i) d1[i] = vk1[i] - vk[i] - g[i] * dt;
    m3_skew(d1, Sk);
    put(3, 0, Sk, -1.0);                                  // Phi_v_q
    double d2[3], d3[3];
    for (int i = 0; i < 3; ++i) {
      d2[i] = -pk1[i] + pk[i] + vk1[i] * dt - 0.5 * g[i] * dt * dt;
      d3[i] = -0.5 * pk1[i] + 0.5 * pk[i] + 0.5 * vk1[i] * dt - g[i] * dt * dt / 6;
    }
    m3_skew(d2, Sk);
    m3_mul(Sk, C, T2);
    m3_skew(d3, Sk);
    m3_mul(Sk, C, T3);
    double T4[9];
    m3_mul(T3, AA, T4);
    for (int i = 0; i < 9; ++i) T2[i] += T4[i];
    put(3, 9, T2, 1.0);                                   // Phi_v_bg
    double d4[3];
    for (int i = 0; i < 3; ++i) d4[i] = pk1[i] - pk[i] - vk[i] * dt - 0.5 * g[i] * dt * dt;
    m3_skew(d4, Sk);
    put(6, 0, Sk, -1.0);                                  // Phi_p_q
    double I3[9];
    m3_eye(I3);
    put(6, 3, I3, dt);                                    // Phi_p_v
    double gs[9], d5[3];
    m3_skew(g, gs);
    m3_mul(gs, C, T2);
    for (int i = 0; i < 3; ++i) d5[i] = pk1[i] - pk[i] - g[i] * dt * dt / 6;
    m3_skew(d5, Sk);
    m3_mul(Sk, C, T3);
    m3_mul(T3, AA, T4);
    for (int i = 0; i < 9; ++i) T2[i] = -dt * dt * dt * T2[i] / 6 + dt * T4[i] / 4;
    put(6, 9, T2, 1.0);                                   // Phi_p_bg
    double I3A[9];
    for (int i = 0; i < 9; ++i) I3A[i] = AA[i];
    I3A[0] += 3; I3A[4] += 3; I3A[8] += 3;
    m3_mul(C, I3A, T1);
    put(6, 12, T1, -dt * dt / 6);                         // Phi_p_ba
  } else {
    // OrcVIO right perturbation, :4308-4367
    double as[9], gsk[9], gg[9];
    m3_skew(acc, as);
    m3_skew(gyro, gsk);
    m3_mul(gsk, gsk, gg);
    const double gn = v3_norm(gyro);
    const double gn2 = gn * gn;
    double mg[3] = {-dt * gyro[0], -dt * gyro[1], -dt * gyro[2]};
    double pg[3] = {dt * gyro[0], dt * gyro[1], dt * gyro[2]};
    double tt[9], JLp[9], JLm[9], HLp[9], HLm[9];
    so3_exp(mg, tt);
    Jl_op(pg, JLp);
    Jl_op(mg, JLm);
    Hl_op(pg, HLp);
    Hl_op(mg, HLm);
    double I3[9];
    m3_eye(I3);
    // Delta = -(g_skew/gn2) (tt^T (dt g_skew - I) + I)
    double A1[9], A2[9], Delta[9];
    for (int i = 0; i < 9; ++i) A1[i] = dt * gsk[i] - I3[i];
    m3_Tmul(tt, A1, A2);
    for (int i = 0; i < 9; ++i) A2[i] += I3[i];
    double gs_n[9];
    for (int i = 0; i < 9; ++i) gs_n[i] = gsk[i] / gn2;
    m3_mul(gs_n, A2, Delta);
    for (int i = 0; i < 9; ++i) Delta[i] = -Delta[i];
    put(0, 0, tt, 1.0);                                   // theta_theta
    put(0, 9, JLm, -dt);                                  // theta_gyro
    double v1[3], Sk[9], T1[9], T2[9], T3[9];
    m3_vec(JLp, acc, v1);
    m3_skew(v1, Sk);
    m3_mul(C, Sk, T1);
    put(3, 0, T1, -dt);                                   // v_theta
    // v_gyro = wRi Delta a_skew (I + gg/gn2) + dt wRi JLp (a_skew g_skew/gn2)
    //          + dt wRi (g a^T/gn2) JLm - dt (a.g/gn2) I
    double IG[9];
    for (int i = 0; i < 9; ++i) IG[i] = I3[i] + gg[i] / gn2;
    double ag[9];
    m3_mul(as, gsk, ag);
    for (int i = 0; i < 9; ++i) ag[i] = ag[i] / gn2;
    double ga[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) ga[3 * i + j] = gyro[i] * acc[j] / gn2;
    const double adg = (acc[0] * gyro[0] + acc[1] * gyro[1]) + acc[2] * gyro[2];
    double vg[9];
    m3_mul(C, Delta, T1);
    m3_mul(T1, as, T2);
    m3_mul(T2, IG, vg);
    m3_mul(C, JLp, T1);
    m3_mul(T1, ag, T2);
    for (int i = 0; i < 9; ++i) vg[i] += dt * T2[i];
    m3_mul(C, ga, T1);
    m3_mul(T1, JLm, T2);
    for (int i = 0; i < 9; ++i) vg[i] += dt * T2[i];
    for (int i = 0; i < 9; ++i) vg[i] -= dt * (adg / gn2) * I3[i];
    put(3, 9, vg, 1.0);                                   // v_gyro
    m3_mul(C, JLp, T1);
    put(3, 12, T1, -dt);                                  // v_acc
    m3_vec(HLp, acc, v1);
    m3_skew(v1, Sk);
    m3_mul(C, Sk, T1);
    put(6, 0, T1, -dt * dt);                              // p_theta
    put(6, 3, I3, dt);                                    // p_v
    // p_gyro = wRi (-g_skew Delta - dt JLp + dt I) a_skew (I + gg/gn2) (g_skew/gn2)
    //          + dt^2 wRi HLp (a_skew g_skew/gn2) + dt^2 wRi (g a^T/gn2) HLm
    //          - dt^2 (a.g/(2 gn2)) wRi
    double B1[9];
    m3_mul(gsk, Delta, B1);
    for (int i = 0; i < 9; ++i) B1[i] = -B1[i] - dt * JLp[i] + dt * I3[i];
    double pgm[9];
    m3_mul(C, B1, T1);
    m3_mul(T1, as, T2);
    m3_mul(T2, IG, T3);
    m3_mul(T3, gs_n, pgm);
    m3_mul(C, HLp, T1);
    m3_mul(T1, ag, T2);
    for (int i = 0; i < 9; ++i) pgm[i] += dt * dt * T2[i];
    m3_mul(C, ga, T1);
    m3_mul(T1, HLm, T2);
    for (int i = 0; i < 9; ++i) pgm[i] += dt * dt * T2[i];
    for (int i = 0; i < 9; ++i) pgm[i] -= dt * dt * (adg / (2 * gn2)) * C[i];
    put(6, 9, pgm, 1.0);                                  // p_gyro
    m3_mul(C, HLp, T1);
    put(6, 12, T1, -dt * dt);                             // p_acc
  }
}

// Mean step (predictNewStateOrcVIO :899-928 / predictNewStateLARVIO :825-897), in two phases per chunk of samples.
// The closed-form (OrcVIO) mean step split in two: what depends on the SAMPLE only -- bias-corrected rates, dt, the
// operators Hl(dt w) a, Jl(dt w) a and exp(dt w), i.e. every transcendental of the step -- is computed by one thread
// per sample (mean_pre), and thread 0 only runs the short recursion R, v, p through them (mean_seq): the same operations
// on the same values as a single serial step, so the results are identical bit for bit, with ~0.4 us instead of ~2.2 us per sample on
// the serial thread.
constexpr int PRE = 16;       // per-sample doubles handed from mean_pre to mean_seq
__device__ void mean_pre(const double* imu, const PropSample* smp, int k, bool first, double* ctx, double* pre) {
  const PropSample& sm = smp[k];
  double acc[3], gyro[3], gyro_old[3];
  for (int i = 0; i < 3; ++i) {
    acc[i] = sm.a[i] - imu[IM_BA + i];
    gyro[i] = sm.w[i] - imu[IM_BG + i];
    gyro_old[i] = (first ? imu[IM_GOLD + i] : smp[k - 1].w[i]) - imu[IM_BG + i];
  }
  const double dt = sm.t - (first ? imu[IM_TIME] : smp[k - 1].t);
  ctx[0] = dt;
  for (int i = 0; i < 3; ++i) { ctx[1 + i] = gyro[i]; ctx[4 + i] = acc[i]; ctx[28 + i] = gyro_old[i]; }
  double dg[3] = {dt * gyro[0], dt * gyro[1], dt * gyro[2]};
  double Hl[9], Jl[9];
  Hl_op(dg, Hl);
  m3_vec(Hl, acc, pre);
  Jl_op(dg, Jl);
  m3_vec(Jl, acc, pre + 3);
  so3_exp(dg, pre + 6);
}
__device__ void mean_seq(double* imu, const PropSample& sm, double* ctx, const double* pre) {
  const double dt = ctx[0];
  for (int i = 0; i < 9; ++i) ctx[7 + i] = imu[IM_R + i];
  for (int i = 0; i < 3; ++i) { ctx[16 + i] = imu[IM_V + i]; ctx[19 + i] = imu[IM_P + i]; }
  const double g[3] = {0.0, 0.0, -9.81};
  double* R = imu + IM_R;
  double* v = imu + IM_V;
  double* p = imu + IM_P;
  double t2[3], Rn[9];
  m3_vec(R, pre, t2);
  for (int i = 0; i < 3; ++i) p[i] = p[i] + dt * v[i] + g[i] * (dt * dt / 2) + t2[i] * (dt * dt);
  m3_vec(R, pre + 3, t2);
  for (int i = 0; i < 3; ++i) v[i] = v[i] + g[i] * dt + t2[i] * dt;
  m3_mul(R, pre + 6, Rn);
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  for (int i = 0; i < 3; ++i) { ctx[22 + i] = v[i]; ctx[25 + i] = p[i]; }
  imu[IM_TIME] = sm.t;
  for (int i = 0; i < 3; ++i) { imu[IM_GOLD + i] = sm.w[i]; imu[IM_AOLD + i] = sm.a[i]; }
}

// The same split for the LARVIO step: the sample-only part is the rates, dt, |w| and the four trigonometric factors of
// the two quaternion increments; the quaternion / Runge-Kutta recursion itself stays on the serial thread, expression
// for expression as in a single serial step (bit-identical results).
__device__ void mean_pre_larvio(const double* imu, const PropSample* smp, int k, bool first, double* ctx, double* pre) {
  const PropSample& sm = smp[k];
  double gyro[3];
  for (int i = 0; i < 3; ++i) {
    ctx[4 + i] = sm.a[i] - imu[IM_BA + i];
    gyro[i] = sm.w[i] - imu[IM_BG + i];
    ctx[1 + i] = gyro[i];
    ctx[28 + i] = (first ? imu[IM_GOLD + i] : smp[k - 1].w[i]) - imu[IM_BG + i];
  }
  const double dt = sm.t - (first ? imu[IM_TIME] : smp[k - 1].t);
  ctx[0] = dt;
  const double gn = v3_norm(gyro);
  pre[0] = gn;
  pre[1] = cos(gn * dt * 0.5);
  pre[3] = cos(gn * dt * 0.25);
  if (gn > 1e-5) {
    pre[2] = 1 / gn * sin(gn * dt * 0.5);
    pre[4] = 1 / gn * sin(gn * dt * 0.25);
  }
}
__device__ void mean_seq_larvio(double* imu, const PropSample& sm, double* ctx, const double* pre) {
  const double dt = ctx[0];
  const double gyro[3] = {ctx[1], ctx[2], ctx[3]}, acc[3] = {ctx[4], ctx[5], ctx[6]};
  for (int i = 0; i < 9; ++i) ctx[7 + i] = imu[IM_R + i];
  for (int i = 0; i < 3; ++i) { ctx[16 + i] = imu[IM_V + i]; ctx[19 + i] = imu[IM_P + i]; }
  const double g[3] = {0.0, 0.0, -9.81};
  double* R = imu + IM_R;
  double* v = imu + IM_V;
  double* p = imu + IM_P;
  const double gn = pre[0];
  double q[4];
  R_to_quat_xyzw(R, q);
  auto omq = [&](const double* qq, double* o) {
    const double wx = gyro[0], wy = gyro[1], wz = gyro[2];
    o[0] = (0 * qq[0] + wz * qq[1] - wy * qq[2]) + wx * qq[3];
    o[1] = (-wz * qq[0] + 0 * qq[1] + wx * qq[2]) + wy * qq[3];
    o[2] = (wy * qq[0] - wx * qq[1] + 0 * qq[2]) + wz * qq[3];
    o[3] = -wx * qq[0] - wy * qq[1] - wz * qq[2];
  };
  double oq[4], dq1[4], dq2[4];
  omq(q, oq);
  if (gn > 1e-5) {
    const double c1 = pre[1], s1 = pre[2];
    const double c2 = pre[3], s2 = pre[4];
    for (int i = 0; i < 4; ++i) { dq1[i] = c1 * q[i] + s1 * oq[i]; dq2[i] = c2 * q[i] + s2 * oq[i]; }
  } else {
    const double c1 = pre[1], c2 = pre[3];
    for (int i = 0; i < 4; ++i) {
      dq1[i] = (q[i] + 0.5 * dt * oq[i]) * c1;
      dq2[i] = (q[i] + 0.25 * dt * oq[i]) * c2;
    }
  }
  double R1[9], R2[9], R0[9];
  quat_wxyz_to_R(dq1[3], dq1[0], dq1[1], dq1[2], R1);
  quat_wxyz_to_R(dq2[3], dq2[0], dq2[1], dq2[2], R2);
  quat_wxyz_to_R(q[3], q[0], q[1], q[2], R0);
  double k1v[3], k2v[3], k4v[3], t1[3];
  m3_vec(R0, acc, t1);
  for (int i = 0; i < 3; ++i) k1v[i] = t1[i] + g[i];
  m3_vec(R2, acc, t1);
  for (int i = 0; i < 3; ++i) k2v[i] = t1[i] + g[i];
  m3_vec(R1, acc, t1);
  for (int i = 0; i < 3; ++i) k4v[i] = t1[i] + g[i];
  double vn[3], pn[3];
  for (int i = 0; i < 3; ++i) {
    const double k1_v = v[i] + k1v[i] * dt / 2;
    const double k2_v = v[i] + k2v[i] * dt / 2;
    const double k3_v = v[i] + k2v[i] * dt;          // k3_v_dot == k2_v_dot in the reference
    vn[i] = v[i] + dt / 6 * (k1v[i] + 2 * k2v[i] + 2 * k2v[i] + k4v[i]);
    pn[i] = p[i] + dt / 6 * (v[i] + 2 * k1_v + 2 * k2_v + k3_v);
  }
  double nq = sqrt(dq1[0] * dq1[0] + dq1[1] * dq1[1] + dq1[2] * dq1[2] + dq1[3] * dq1[3]);
  double qn[4] = {dq1[0] / nq, dq1[1] / nq, dq1[2] / nq, dq1[3] / nq};
  for (int i = 0; i < 3; ++i) { v[i] = vn[i]; p[i] = pn[i]; }
  quat_xyzw_to_R(qn, R);
  for (int i = 0; i < 3; ++i) { ctx[22 + i] = v[i]; ctx[25 + i] = p[i]; }
  imu[IM_TIME] = sm.t;
  for (int i = 0; i < 3; ++i) { imu[IM_GOLD + i] = sm.w[i]; imu[IM_AOLD + i] = sm.a[i]; }
}

// One CTA per filter; every sample's transition is applied to the whole 15 x D strip, in the reference's order.
// (Applying the PRODUCT of a chunk's transitions to the columns behind the corner once per chunk -- only the 15 x 15
// corner needs the per-sample recursion -- was built and measured: 117 -> 107 us, the serial mean and the transition
// build dominate.  It was REJECTED for parity: mathematically identical, but its rounding no longer follows the
// reference's operation order, and the update that follows amplifies the 1e-16 differences in P_IC by the conditioning
// of an early, strongly correlated window: 2.5e-9 in the velocity at frame 5 of the 300-frame EuRoC soak sequence
// (scripts/soak_parity.py), above the 1e-9 per-update bound, against 1e-10 for this form.)
__global__ void __launch_bounds__(256) k_propagate(PropArgs a) {
  extern __shared__ double sm[];
  const int fi = blockIdx.x;
  const int s0 = a.samp_off[fi], s1 = a.samp_off[fi + 1];
  if (s1 <= s0) return;
  const int D = a.D[fi];
  const int ldp = a.ldp;
  double* P = a.P + (size_t)fi * a.p_stride;
  double* imu = a.imu + (size_t)fi * IM_STRIDE;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* strip = sm;                               // [PS][ldp]
  double* phi = strip + (size_t)PS * ldp;           // [SMAX][PS*PS]
  double* ctx = phi + (size_t)SMAX * PS * PS;       // [SMAX][CTX]
  double* tmp = ctx + (size_t)SMAX * CTX;           // [PS*PS]
  double* pre = tmp + PS * PS;                      // [SMAX][PRE]
  for (int e = tid; e < PS * D; e += nt) strip[(e / D) * ldp + (e % D)] = P[(size_t)(e / D) * ldp + (e % D)];
  __syncthreads();
  const double nq[PS] = {a.qc[0], a.qc[0], a.qc[0], a.qc[1], a.qc[1], a.qc[1], 0, 0, 0,
                         a.qc[2], a.qc[2], a.qc[2], a.qc[3], a.qc[3], a.qc[3]};
  for (int c0 = s0; c0 < s1; c0 += SMAX) {
    const int ns = min(SMAX, s1 - c0);
    const bool larvio = (a.flags & FL_LARVIO) != 0;
    if (tid < ns) {
      if (larvio) mean_pre_larvio(imu, a.samples, c0 + tid, tid == 0, ctx + tid * CTX, pre + tid * PRE);
      else mean_pre(imu, a.samples, c0 + tid, tid == 0, ctx + tid * CTX, pre + tid * PRE);
    }
    __syncthreads();
    if (tid == 0) {
      for (int s = 0; s < ns; ++s) {
        if (larvio) mean_seq_larvio(imu, a.samples[c0 + s], ctx + s * CTX, pre + s * PRE);
        else mean_seq(imu, a.samples[c0 + s], ctx + s * CTX, pre + s * PRE);
      }
    }
    __syncthreads();
    if (tid < ns) build_phi(ctx + tid * CTX, a.flags, phi + (size_t)tid * PS * PS);
    __syncthreads();
    for (int s = 0; s < ns; ++s) {
      const double* F = phi + (size_t)s * PS * PS;
      const double dt = ctx[s * CTX];
      for (int j = tid; j < D; j += nt) {
        double col[PS], out[PS];
        for (int i = 0; i < PS; ++i) col[i] = strip[i * ldp + j];
        for (int i = 0; i < PS; ++i) {
          double acc = 0.0;
          for (int k = 0; k < PS; ++k) acc += F[i * PS + k] * col[k];
          out[i] = acc;
        }
        for (int i = 0; i < PS; ++i) strip[i * ldp + j] = out[i];
      }
      __syncthreads();
      // corner: P11 = (Phi P11) Phi^T + Phi N Phi^T dt, N = G Qc G^T = diag(nq)
      for (int e = tid; e < PS * PS; e += nt) {
        const int i = e / PS, j = e % PS;
        double acc = 0.0, q = 0.0;
        for (int k = 0; k < PS; ++k) {
          acc += strip[i * ldp + k] * F[j * PS + k];
          q += F[i * PS + k] * nq[k] * F[j * PS + k];
        }
        tmp[e] = acc + q * dt;
      }
      __syncthreads();
      for (int e = tid; e < PS * PS; e += nt) {
        const int i = e / PS, j = e % PS;
        strip[i * ldp + j] = (tmp[i * PS + j] + tmp[j * PS + i]) / 2.0;
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < PS * D; e += nt) {
    const int i = e / D, j = e % D;
    const double v = strip[i * ldp + j];
    P[(size_t)i * ldp + j] = v;
    if (j >= PS) P[(size_t)j * ldp + i] = v;
  }
}

void launch_propagate(const PropArgs& a, cudaStream_t s) {
  size_t smem = ((size_t)PS * a.ldp + (size_t)SMAX * PS * PS + (size_t)SMAX * CTX + PS * PS + (size_t)SMAX * PRE) * sizeof(double);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_propagate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  k_propagate<<<a.n_filters, 256, smem, s>>>(a);
  check_launch("k_propagate");
}

// stateAugmentation: clone the IMU pose, P <- [[P, P J^T],[J P, J P J^T]], J selects theta, p.
__global__ void __launch_bounds__(256) k_augment(AugArgs a) {
  const int fi = blockIdx.x;
  const int N = a.N[fi];
  if (N < 0) return;                       // filter did not take part in this frame
  const int D = ORCVIO_LEG + 6 * N;          // the new block goes in at index D, in front of the feature states
  const int E = a.E ? a.E[fi] : 0;
  const int ldp = a.ldp;
  double* P = a.P + (size_t)fi * a.p_stride;
  const double* imu = a.imu + (size_t)fi * IM_STRIDE;
  double* c = a.clones + (size_t)fi * a.clone_stride + (size_t)N * CL_STRIDE;
  const int tid = threadIdx.x;
  if (E > 0) {
    // stateAugmentation :985-994: the feature rows / columns move 6 places down (descending order: in place is safe)
    const int Dt = D + E;
    for (int i = Dt - 1; i >= 0; --i) {
      const int ni = i < D ? i : i + 6;
      // row i -> row ni with the feature columns shifted; rows < D only shift their feature columns
      if (i >= D) {
        for (int j = tid; j < D; j += blockDim.x) P[(size_t)ni * ldp + j] = P[(size_t)i * ldp + j];
      }
      __syncthreads();
      // feature columns of this row, via registers (each thread one column; E <= blockDim.x)
      double keep = 0.0;
      if (tid < E) keep = P[(size_t)i * ldp + D + tid];
      __syncthreads();
      if (tid < E) P[(size_t)ni * ldp + D + 6 + tid] = keep;
      __syncthreads();
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) c[CL_R + i] = imu[IM_R + i];
    for (int i = 0; i < 3; ++i) c[CL_P + i] = imu[IM_P + i];
    double Rc[9], t[3];
    m3_mulT(imu + IM_R, imu + IM_RBC, Rc);    // (R_b2c R_b2w^T)^T = R_b2w R_b2c^T
    m3_vec(imu + IM_R, imu + IM_TCB, t);
    for (int i = 0; i < 9; ++i) c[CL_RC + i] = Rc[i];
    for (int i = 0; i < 3; ++i) c[CL_PC + i] = imu[IM_P + i] + t[i];
  }
  auto src = [](int q) { return q < 3 ? q : q + 3; };   // theta rows 0..2, p rows 6..8
  for (int e = tid; e < 6 * D; e += blockDim.x) {
    const int q = e / D, j = e % D;
    const double v = P[(size_t)src(q) * ldp + j];
    P[(size_t)(D + q) * ldp + j] = v;
    P[(size_t)j * ldp + D + q] = v;
  }
  // cross terms with the (shifted) feature states: J P over the feature columns
  for (int e = tid; e < 6 * E; e += blockDim.x) {
    const int q = e / E, j = D + 6 + e % E;
    const double v = P[(size_t)src(q) * ldp + j];
    P[(size_t)(D + q) * ldp + j] = v;
    P[(size_t)j * ldp + D + q] = v;
  }
  if (tid < 36) {
    const int q = tid / 6, r = tid % 6;
    P[(size_t)(D + q) * ldp + D + r] = P[(size_t)src(q) * ldp + src(r)];
  }
}

void launch_augment(const AugArgs& a, cudaStream_t s) {
  k_augment<<<a.n_filters, 256, 0, s>>>(a);
  check_launch("k_augment");
}

// Delete the rows/columns of up to two clones from P and compact the clone array.  In place, RM_ROWS rows at a time
// through shared memory: a row moves up and its entries move left, so a chunk only overwrites rows that were already
// staged (two barriers per chunk instead of two per row, warps over rows and lanes over columns: 116 -> 45 -> ~20 us at D = 202).
constexpr int RM_ROWS = 24, RM_THREADS = 512;
__global__ void __launch_bounds__(RM_THREADS) k_remove(RemoveArgs a) {
  extern __shared__ double rowbuf[];
  const int fi = blockIdx.x;
  const int r0 = a.rm[2 * fi], r1 = a.rm[2 * fi + 1];
  if (r0 < 0 && r1 < 0) return;
  const int N = a.N[fi];
  const int D = ORCVIO_LEG + 6 * N + (a.E ? a.E[fi] : 0);
  const int ldp = a.ldp;
  double* P = a.P + (size_t)fi * a.p_stride;
  double* clones = a.clones + (size_t)fi * a.clone_stride;
  const int tid = threadIdx.x, nt = blockDim.x;
  auto removed = [&](int idx) {           // idx = state index; true if inside a removed clone block
    if (idx < ORCVIO_LEG || idx >= ORCVIO_LEG + 6 * N) return false;
    const int c = (idx - ORCVIO_LEG) / 6;
    return c == r0 || c == r1;
  };
  auto newidx = [&](int idx) {
    if (idx < ORCVIO_LEG) return idx;
    const int c = idx >= ORCVIO_LEG + 6 * N ? N : (idx - ORCVIO_LEG) / 6;   // feature states sit behind every clone
    int shift = 0;
    if (r0 >= 0 && c > r0) shift += 6;
    if (r1 >= 0 && c > r1) shift += 6;
    return idx - shift;
  };
  // new index of every row / column (-1 = removed), once; then warps over rows, lanes over columns (no divisions)
  __shared__ short nidx[256];
  for (int j = tid; j < D; j += nt) nidx[j] = removed(j) ? (short)-1 : (short)newidx(j);
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int i0 = 0; i0 < D; i0 += RM_ROWS) {
    const int nr = min(RM_ROWS, D - i0);
    for (int r = warp; r < nr; r += nw) {
      if (nidx[i0 + r] < 0) continue;
      const double* src = P + (size_t)(i0 + r) * ldp;
      for (int j = lane; j < D; j += 32) rowbuf[r * ldp + j] = src[j];
    }
    __syncthreads();
    for (int r = warp; r < nr; r += nw) {
      const int ni = nidx[i0 + r];
      if (ni < 0) continue;
      double* dst = P + (size_t)ni * ldp;
      for (int j = lane; j < D; j += 32) {
        const int nj = nidx[j];
        if (nj >= 0) dst[nj] = rowbuf[r * ldp + j];
      }
    }
    __syncthreads();
  }
  // clone array (24 doubles per clone), ascending order so in-place is safe row by row
  for (int c = 0; c < N; ++c) {
    if (c == r0 || c == r1) continue;
    int nc = c - ((r0 >= 0 && c > r0) ? 1 : 0) - ((r1 >= 0 && c > r1) ? 1 : 0);
    double v = 0.0;
    if (tid < CL_STRIDE) v = clones[(size_t)c * CL_STRIDE + tid];
    __syncthreads();
    if (tid < CL_STRIDE) clones[(size_t)nc * CL_STRIDE + tid] = v;
    __syncthreads();
  }
}

void launch_remove(const RemoveArgs& a, cudaStream_t s) {
  size_t smem = (size_t)RM_ROWS * a.ldp * sizeof(double);
  k_remove<<<a.n_filters, RM_THREADS, smem, s>>>(a);
  check_launch("k_remove");
}

}  // namespace ob
