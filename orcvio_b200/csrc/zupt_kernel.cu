// Zero-velocity update: the IMU stationarity test of OrcVIO::checkZUPTIMU
// (reference src/orcvio.cpp:3129-3323) and the 9-row (v, delta p, delta q) update of
// OrcVIO::measurementUpdate_ZUPT_vpq (:3326-3454), one CTA per filter.
//
// checkZUPTIMU builds a (6n x 9) H over [theta, b_g, b_a], R = diag(sigma^2 / dt) and evaluates
//   chi2 = res^T (H P_m H^T + R)^-1 res                      (6n x 6n LLT in the reference).
// R is diagonal and H has nine columns, so the same number is
//   chi2 = c - u^T (I + L^T M L)^-1 u,   M = H^T R^-1 H, b = H^T R^-1 res, c = res^T R^-1 res,
//   P_m = L L^T, u = L^T b                                    (Woodbury; two 9 x 9 Cholesky factors)
// which needs no 6n x 6n matrix at all.  M, b, c are accumulated sample by sample.
//
// The update's H only selects/differences rows of P (velocity, the last two clone poses), so
//   HP  = rows of P,  S = HP H^T + R_z (9 x 9) = C C^T,  W = C^-1 HP,
//   delta_x = W^T C^-1 r,   P+ = P - W^T W                    (symmetric by construction; the
// reference's (I - K H) P followed by (P + P^T)/2 is the same matrix up to rounding).
#include "kernels.h"
#include "increment.cuh"

namespace ob {

namespace {

constexpr int ZCHUNK = 64;

// in-place lower Cholesky of a 9 x 9 SPD matrix (row-major), single thread
__device__ void chol9(double* A) {
  for (int j = 0; j < 9; ++j) {
    double d = A[j * 9 + j];
    for (int k = 0; k < j; ++k) d -= A[j * 9 + k] * A[j * 9 + k];
    d = sqrt(d);
    A[j * 9 + j] = d;
    for (int i = j + 1; i < 9; ++i) {
      double s = A[i * 9 + j];
      for (int k = 0; k < j; ++k) s -= A[i * 9 + k] * A[j * 9 + k];
      A[i * 9 + j] = s / d;
    }
  }
}

}  // namespace

__global__ void __launch_bounds__(256) k_zupt(ZuptArgs a) {
  __shared__ double HP[9][ORCVIO_LEG + 6 * ORCVIO_MAX_OBS + 2];
  __shared__ double dxs[ORCVIO_LEG + 6 * ORCVIO_MAX_OBS + 2];
  __shared__ double M[81], Pm[81], bvec[9], S[81], rz[9], yz[9];
  __shared__ double stage[ZCHUNK][31];      // per-sample A_i (27), e_i (3), dt
  __shared__ double cd[2];
  __shared__ int s_do;
  const int fi = blockIdx.x;
  const int N = a.N[fi];
  const int mode = a.mode[fi];
  const int tid = threadIdx.x;
  if (tid == 0) {
    a.decision[fi] = 0;
    a.info[2 * fi] = 0.0;
    a.info[2 * fi + 1] = 0.0;
  }
  if (N < 2 || mode == 0) return;
  const int L = ORCVIO_LEG, D = L + 6 * N, ldp = a.ldp;
  double* P = a.P + (size_t)fi * a.p_stride;
  double* imu = a.imu + (size_t)fi * IM_STRIDE;
  double* clones = a.clones + (size_t)fi * a.clone_stride;

  if (mode == 2) {
    // ------------------------------------------------ checkZUPTIMU
    const int s0 = a.samp_off[fi], cnt = a.samp_off[fi + 1] - s0;
    if (cnt < 2) return;                                  // :3133-3136
    const int n = cnt - 1;
    const double sw2 = 1.6968e-04 * 1.6968e-04, sa2 = 2.0000e-3 * 2.0000e-3;   // :3139-3146
    const double sigma_wb = 1.9393e-05, sigma_ab = 3.0000e-03;
    // stage per-sample rows A_i = [G_i | 0 | wRi] (3 x 9), e_i = -wRi acc - g and the weight dt / sigma_a^2
    // in chunks; thread (p, q) then sums its entry of M (and b, c) over the samples in order.
    double accum = 0.0, dt_sum = 0.0;
    for (int c0 = 0; c0 < n; c0 += ZCHUNK) {
      const int nc = min(ZCHUNK, n - c0);
      if (tid < nc) {
        const PropSample& sm = a.samples[s0 + c0 + tid];
        const double dt = a.samples[s0 + c0 + tid + 1].t - sm.t;
        double acc[3], Ra[3], G[9], sk[9];
        for (int k = 0; k < 3; ++k) acc[k] = sm.a[k] - imu[IM_BA + k];
        m3_vec(imu + IM_R, acc, Ra);
        if (a.flags & FL_LEFT) {
          m3_skew(Ra, G);
        } else {
          m3_skew(acc, sk);
          m3_mul(imu + IM_R, sk, G);
        }
        double* row = stage[tid];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            row[9 * r + c] = G[3 * r + c];
            row[9 * r + 3 + c] = 0.0;
            row[9 * r + 6 + c] = imu[IM_R + 3 * r + c];
          }
        row[27] = -Ra[0];
        row[28] = -Ra[1];
        row[29] = -Ra[2] + 9.81;                 // - IMUState::gravity, gravity = (0, 0, -9.81)
        row[30] = dt;
      }
      __syncthreads();
      if (tid < 91) {
        for (int i = 0; i < nc; ++i) {
          const double* row = stage[i];
          const double wa = row[30] / sa2;
          if (tid < 81) {
            const int p = tid / 9, q = tid % 9;
            accum += ((row[p] * row[q] + row[9 + p] * row[9 + q]) + row[18 + p] * row[18 + q]) * wa;
            if (p == q && p >= 3 && p < 6) accum += row[30] / sw2;   // gyro rows: H = [0 I 0], res = 0
          } else if (tid < 90) {
            const int p = tid - 81;
            accum += ((row[p] * row[27] + row[9 + p] * row[28]) + row[18 + p] * row[29]) * wa;
          } else {
            accum += ((row[27] * row[27] + row[28] * row[28]) + row[29] * row[29]) * wa;
            dt_sum += row[30];
          }
        }
      }
      __syncthreads();
    }
    if (tid < 81) {
      M[tid] = accum;
      // P_marg over [theta, b_g, b_a] (:3236-3256)
      const int idx[9] = {0, 1, 2, 9, 10, 11, 12, 13, 14};
      Pm[tid] = P[(size_t)idx[tid / 9] * ldp + idx[tid % 9]];
    } else if (tid < 90) {
      bvec[tid - 81] = accum;
    } else if (tid == 90) {
      cd[0] = accum;
      cd[1] = dt_sum;
    }
    __syncthreads();
    if (tid == 0) {
      const double c = cd[0], dtot = cd[1];
      for (int k = 3; k < 6; ++k) Pm[9 * k + k] += dtot * sigma_wb;     // Q_bias, :3229-3231
      for (int k = 6; k < 9; ++k) Pm[9 * k + k] += dtot * sigma_ab;
      chol9(Pm);                                           // P_m = L L^T (lower)
      double u[9], T1[81], Wm[81];
      for (int p = 0; p < 9; ++p) {                        // u = L^T b
        double s = 0.0;
        for (int k = p; k < 9; ++k) s += Pm[9 * k + p] * bvec[k];
        u[p] = s;
      }
      for (int p = 0; p < 9; ++p)                          // T1 = M L
        for (int q = 0; q < 9; ++q) {
          double s = 0.0;
          for (int k = q; k < 9; ++k) s += M[9 * p + k] * Pm[9 * k + q];
          T1[9 * p + q] = s;
        }
      for (int p = 0; p < 9; ++p)                          // Wm = I + L^T T1
        for (int q = 0; q < 9; ++q) {
          double s = (p == q) ? 1.0 : 0.0;
          for (int k = p; k < 9; ++k) s += Pm[9 * k + p] * T1[9 * k + q];
          Wm[9 * p + q] = s;
        }
      for (int p = 0; p < 9; ++p)                          // symmetrise rounding
        for (int q = 0; q < p; ++q) {
          const double v = 0.5 * (Wm[9 * p + q] + Wm[9 * q + p]);
          Wm[9 * p + q] = Wm[9 * q + p] = v;
        }
      chol9(Wm);
      double quad = 0.0;
      for (int p = 0; p < 9; ++p) {                        // z = C^-1 u, quad = z^T z
        double s = u[p];
        for (int k = 0; k < p; ++k) s -= Wm[9 * p + k] * u[k];
        u[p] = s / Wm[9 * p + p];
        quad += u[p] * u[p];
      }
      const double chi2 = c - quad;
      const double* v = imu + IM_V;
      const double vn = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
      a.info[2 * fi] = chi2;
      a.info[2 * fi + 1] = vn;
      s_do = !(chi2 > a.chi2_check[fi] || vn > 0.25);      // :3289-3300
    }
    __syncthreads();
    if (!s_do) return;
  }

  // -------------------------------------------------- measurementUpdate_ZUPT_vpq
  const int pc = L + 6 * N - 3, pp = L + 6 * N - 9, qc = L + 6 * N - 6, qp = L + 6 * N - 12;
  for (int e = tid; e < 9 * D; e += blockDim.x) {
    const int r = e / D, j = e % D, k = r % 3;
    double v;
    if (r < 3) v = P[(size_t)(3 + k) * ldp + j];
    else if (r < 6) v = P[(size_t)(pc + k) * ldp + j] - P[(size_t)(pp + k) * ldp + j];
    else v = -0.5 * P[(size_t)(qc + k) * ldp + j] + 0.5 * P[(size_t)(qp + k) * ldp + j];
    HP[r][j] = v;
  }
  __syncthreads();
  if (tid < 81) {
    const int r = tid / 9, c = tid % 9, k = c % 3;
    double v;
    if (c < 3) v = HP[r][3 + k];
    else if (c < 6) v = HP[r][pc + k] - HP[r][pp + k];
    else v = -0.5 * HP[r][qc + k] + 0.5 * HP[r][qp + k];
    if (r == c) v += (r < 3) ? a.noise_v : (r < 6) ? a.noise_p : a.noise_q;
    S[tid] = v;
  }
  if (tid == 96) {
    const double* cc = clones + (size_t)(N - 1) * CL_STRIDE;
    const double* cp = clones + (size_t)(N - 2) * CL_STRIDE;
    for (int k = 0; k < 3; ++k) {
      rz[k] = -imu[IM_V + k];
      rz[3 + k] = -(cc[CL_P + k] - cp[CL_P + k]);
    }
    double q1[4], q0[4];
    R_to_quat_xyzw(cc + CL_R, q1);
    R_to_quat_xyzw(cp + CL_R, q0);
    // Eigen product q_curr * conj(q_prev), components (w, x, y, z)
    const double aw = q1[3], ax = q1[0], ay = q1[1], az = q1[2];
    const double bw = q0[3], bx = -q0[0], by = -q0[1], bz = -q0[2];
    rz[6] = aw * bx + ax * bw + ay * bz - az * by;
    rz[7] = aw * by + ay * bw + az * bx - ax * bz;
    rz[8] = aw * bz + az * bw + ax * by - ay * bx;
  }
  __syncthreads();
  if (tid == 0) {
    for (int p = 0; p < 9; ++p)
      for (int q = 0; q < p; ++q) {
        const double v = 0.5 * (S[9 * p + q] + S[9 * q + p]);
        S[9 * p + q] = S[9 * q + p] = v;
      }
    chol9(S);
    for (int p = 0; p < 9; ++p) {
      double s = rz[p];
      for (int k = 0; k < p; ++k) s -= S[9 * p + k] * yz[k];
      yz[p] = s / S[9 * p + p];
    }
  }
  __syncthreads();
  for (int j = tid; j < D; j += blockDim.x) {             // W = C^-1 HP, column by column
    double w[9];
    double dx = 0.0;
#pragma unroll
    for (int p = 0; p < 9; ++p) {
      double s = HP[p][j];
#pragma unroll
      for (int k = 0; k < p; ++k) s -= S[9 * p + k] * w[k];
      w[p] = s / S[9 * p + p];
      dx += w[p] * yz[p];
    }
#pragma unroll
    for (int p = 0; p < 9; ++p) HP[p][j] = w[p];
    dxs[j] = dx;
    if (a.dx) a.dx[(size_t)fi * a.lddx + j] = dx;
  }
  __syncthreads();
  for (int e = tid; e < D * D; e += blockDim.x) {
    const int i = e / D, j = e % D;
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < 9; ++p) s += HP[p][i] * HP[p][j];
    P[(size_t)i * ldp + j] -= s;
  }
  cta_increment_state(dxs, imu, clones, N, a.flags, a.dx ? &a.dx[(size_t)fi * a.lddx + a.lddx - 1] : nullptr);
  if (tid == 0) a.decision[fi] = 1;
}

void launch_zupt(const ZuptArgs& a, cudaStream_t s) {
  k_zupt<<<a.n_filters, 256, 0, s>>>(a);
  check_launch("k_zupt");
}

}  // namespace ob
