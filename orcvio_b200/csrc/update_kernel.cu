// Stage 5 -- EKF gain, state increment and covariance update.
//
// Reference: OrcVIO::measurementUpdate_msckf / measurementUpdate_hybrid
// (src/orcvio.cpp:1685-1755, 1811-1907) and incrementState_IMUCam (:4468-4567):
//     S = H P H^T + sigma^2 I,  K^T = S.ldlt().solve(H P),  dx = K r,
//     P <- (I - K H) P,  P <- (P + P^T)/2.
// With H = [0 | R] (R = compressed 6N x 6N factor over the clone columns) this is
//     T = R P[c,:]            (n x D)        k_gemm<MODE_RP>
//     S = T[:,c] R^T + s^2 I  (n x n)        k_gemm<MODE_S>
//     S = L L^T, y = L^-1 r   (one CTA)      k_chol_solve
//     Y = L^-1 T              (n x D)        k_trsm
//     dx = Y^T y, state (+)= dx              k_apply_dx
//     P <- P - Y^T Y                          k_gemm<MODE_P>
// P - Y^T Y equals (I - K H) P in exact arithmetic and is symmetric by construction, which
// is what the reference's trailing (P + P^T)/2 enforces.  The reference formula is the
// simple form, NOT Joseph form, and so is this.
//
// All kernels are batched over the filters of a batch (blockIdx.z / blockIdx.y = filter) so
// that one launch serves a single filter (tiles provide the parallelism) and the
// Monte-Carlo batch (filters provide it).
#include "kernels.h"
#include "increment.cuh"

namespace ob {

constexpr int MODE_RP = 0, MODE_S = 1, MODE_P = 2;
constexpr int GT = 32;   // GEMM tile

template <int MODE>
__global__ void __launch_bounds__(256) k_gemm(UpdArgs a) {
  const int fi = blockIdx.z;
  const FilterWork fw = a.fw[fi];
  if (!fw.active) return;
  const int n = 6 * fw.N, D = fw.D, L = ORCVIO_LEG;
  int M, Nn, K;
  if (MODE == MODE_RP) { M = n; Nn = D; K = n; }
  else if (MODE == MODE_S) { M = n; Nn = n; K = n; }
  else { M = D; Nn = D; K = n; }
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  if (i0 >= M || j0 >= Nn) return;
  double* P = a.P + (size_t)fi * a.p_stride;
  const double* Rm = a.Rm + (size_t)fi * a.r_stride;
  double* T = a.T + (size_t)fi * a.t_stride;
  double* S = a.S + (size_t)fi * a.r_stride;
  const int ldp = a.ldp, ldr = a.ldr, ldt = a.ldt;
  __shared__ double As[GT][GT + 1];
  __shared__ double Bs[GT][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[2][2] = {{0, 0}, {0, 0}};
  int k_begin = 0;
  if (MODE == MODE_RP) k_begin = (i0 / GT) * GT;            // R upper triangular: k >= i
  if (MODE == MODE_S) k_begin = (j0 / GT) * GT;             // R[j][k] = 0 for k < j
  for (int k0 = k_begin; k0 < K; k0 += GT) {
    // load A tile (rows i0.., cols k0..) and B tile (rows k0.., cols j0..)
    for (int e = threadIdx.x; e < GT * GT; e += 256) {
      const int r_ = e / GT, c_ = e % GT;
      double av = 0.0, bv = 0.0;
      if (MODE == MODE_P) {
        // A = Y^T: stage as As[k][i], reading Y rows contiguously
        const int ak = k0 + r_, ai = i0 + c_;
        if (ak < K && ai < M) av = T[(size_t)ak * ldt + ai];
      } else {
        const int ai = i0 + r_, ak = k0 + c_;
        if (ai < M && ak < K) {
          if (MODE == MODE_RP) av = Rm[(size_t)ai * ldr + ak];
          else av = T[(size_t)ai * ldt + L + ak];
        }
      }
      const int bk = k0 + r_, bj = j0 + c_;
      if (MODE == MODE_S) {
        // B = R^T: stage as Bs[k][j], reading R rows contiguously (r_ -> j, c_ -> k)
        const int bj2 = j0 + r_, bk2 = k0 + c_;
        if (bj2 < Nn && bk2 < K) bv = Rm[(size_t)bj2 * ldr + bk2];
        Bs[c_][r_] = bv;
      } else {
        if (bk < K && bj < Nn) {
          if (MODE == MODE_RP) bv = P[(size_t)(L + bk) * ldp + bj];
          else bv = T[(size_t)bk * ldt + bj];
        }
        Bs[r_][c_] = bv;                      // Bs[k][j]
      }
      As[r_][c_] = av;                        // As[i][k]  (MODE_P: As[k][i])
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < GT; ++kk) {
      double a0, a1;
      if (MODE == MODE_P) { a0 = As[kk][2 * ty]; a1 = As[kk][2 * ty + 1]; }
      else { a0 = As[2 * ty][kk]; a1 = As[2 * ty + 1][kk]; }
      const double b0 = Bs[kk][2 * tx], b1 = Bs[kk][2 * tx + 1];
      acc[0][0] += a0 * b0; acc[0][1] += a0 * b1;
      acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
    }
    __syncthreads();
  }
  for (int u = 0; u < 2; ++u)
    for (int v = 0; v < 2; ++v) {
      const int i = i0 + 2 * ty + u, j = j0 + 2 * tx + v;
      if (i >= M || j >= Nn) continue;
      if (MODE == MODE_RP) T[(size_t)i * ldt + j] = acc[u][v];
      else if (MODE == MODE_S) S[(size_t)i * ldr + j] = acc[u][v] + (i == j ? a.sigma2 : 0.0);
      else P[(size_t)i * ldp + j] -= acc[u][v];
    }
}

// Blocked left-looking Cholesky of S (lower, in place) with r_thin carried along as an
// extra bottom row, so that the row comes out as y = L^-1 r_thin.  One CTA per filter.
constexpr int CB = 32;
__global__ void __launch_bounds__(256) k_chol_solve(UpdArgs a, int ncap) {
  extern __shared__ double sm[];
  const int fi = blockIdx.x;
  const FilterWork fw = a.fw[fi];
  if (!fw.active) return;
  const int n = 6 * fw.N;
  double* S = a.S + (size_t)fi * a.r_stride;
  double* yv = a.yv + (size_t)fi * a.ldr;
  const double* rth = a.rthin + (size_t)fi * a.ldr;
  const int ldr = a.ldr;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* pan = sm;                           // [(ncap+1)][CB+1]  current panel
  double* lk = pan + (size_t)(ncap + 1) * (CB + 1);   // [(ncap+1)][CB+1]  staged chunk of previous columns
  auto rowp = [&](int i) -> double* { return (i < n) ? (S + (size_t)i * ldr) : yv; };
  // y row starts as r_thin
  for (int j = tid; j < n; j += nt) yv[j] = rth[j];
  __syncthreads();
  for (int p0 = 0; p0 < n; p0 += CB) {
    const int nb = min(CB, n - p0);
    const int mrows = n + 1 - p0;             // panel rows p0 .. n (row n = y)
    for (int e = tid; e < mrows * nb; e += nt) {
      const int i = e / nb, j = e % nb;
      pan[i * (CB + 1) + j] = rowp(p0 + i)[p0 + j];
    }
    __syncthreads();
    // subtract contributions of the already factored columns [0, p0)
    for (int k0 = 0; k0 < p0; k0 += CB) {
      for (int e = tid; e < mrows * CB; e += nt) {
        const int i = e / CB, k = e % CB;
        lk[i * (CB + 1) + k] = rowp(p0 + i)[k0 + k];
      }
      __syncthreads();
      for (int e = tid; e < mrows * nb; e += nt) {
        const int i = e / nb, j = e % nb;
        if (i < nb && j > i) continue;        // upper part of the diagonal block unused
        double s = 0.0;
        const double* li = lk + i * (CB + 1);
        const double* lj = lk + j * (CB + 1);
#pragma unroll 8
        for (int k = 0; k < CB; ++k) s += li[k] * lj[k];
        pan[i * (CB + 1) + j] -= s;
      }
      __syncthreads();
    }
    // factor the diagonal block
    for (int k = 0; k < nb; ++k) {
      if (tid == 0) pan[k * (CB + 1) + k] = sqrt(pan[k * (CB + 1) + k]);
      __syncthreads();
      const double dk = pan[k * (CB + 1) + k];
      for (int i = k + 1 + tid; i < mrows; i += nt) pan[i * (CB + 1) + k] /= dk;
      __syncthreads();
      const int remc = nb - k - 1;
      for (int e = tid; e < (mrows - k - 1) * remc; e += nt) {
        const int i = k + 1 + e / remc, j = k + 1 + e % remc;
        if (i < nb && j > i) continue;
        pan[i * (CB + 1) + j] -= pan[i * (CB + 1) + k] * pan[j * (CB + 1) + k];
      }
      __syncthreads();
    }
    for (int e = tid; e < mrows * nb; e += nt) {
      const int i = e / nb, j = e % nb;
      if (i < nb && j > i) continue;
      rowp(p0 + i)[p0 + j] = pan[i * (CB + 1) + j];
    }
    __syncthreads();
  }
}

// Y = L^-1 T, in place in T.  One CTA per (32-column strip, filter).
__global__ void __launch_bounds__(256) k_trsm(UpdArgs a, int ncap) {
  extern __shared__ double sm[];
  const int fi = blockIdx.y;
  const FilterWork fw = a.fw[fi];
  if (!fw.active) return;
  const int n = 6 * fw.N, D = fw.D;
  const int c0 = blockIdx.x * CB;
  if (c0 >= D) return;
  const int nc = min(CB, D - c0);
  const double* S = a.S + (size_t)fi * a.r_stride;
  double* T = a.T + (size_t)fi * a.t_stride;
  const int ldr = a.ldr, ldt = a.ldt;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* Ys = sm;                               // [ncap][CB+1]
  double* Lt = Ys + (size_t)ncap * (CB + 1);     // [CB][CB+1]
  for (int e = tid; e < n * CB; e += nt) {
    const int i = e / CB, j = e % CB;
    Ys[i * (CB + 1) + j] = (j < nc) ? T[(size_t)i * ldt + c0 + j] : 0.0;
  }
  __syncthreads();
  for (int r0 = 0; r0 < n; r0 += CB) {
    const int nr = min(CB, n - r0);
    // Ys[r0.., :] -= L[r0.., 0:r0] * Ys[0:r0, :]
    for (int k0 = 0; k0 < r0; k0 += CB) {
      for (int e = tid; e < CB * CB; e += nt) {
        const int i = e / CB, k = e % CB;
        Lt[i * (CB + 1) + k] = (i < nr) ? S[(size_t)(r0 + i) * ldr + k0 + k] : 0.0;
      }
      __syncthreads();
      for (int e = tid; e < nr * CB; e += nt) {
        const int i = e / CB, j = e % CB;
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < CB; ++k) s += Lt[i * (CB + 1) + k] * Ys[(k0 + k) * (CB + 1) + j];
        Ys[(r0 + i) * (CB + 1) + j] -= s;
      }
      __syncthreads();
    }
    // diagonal block: forward substitution, one thread per column
    for (int e = tid; e < CB * CB; e += nt) {
      const int i = e / CB, k = e % CB;
      Lt[i * (CB + 1) + k] = (i < nr && k < nr) ? S[(size_t)(r0 + i) * ldr + r0 + k] : 0.0;
    }
    __syncthreads();
    if (tid < CB) {
      const int j = tid;
      for (int i = 0; i < nr; ++i) {
        double s = Ys[(r0 + i) * (CB + 1) + j];
        for (int k = 0; k < i; ++k) s -= Lt[i * (CB + 1) + k] * Ys[(r0 + k) * (CB + 1) + j];
        Ys[(r0 + i) * (CB + 1) + j] = s / Lt[i * (CB + 1) + i];
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < n * CB; e += nt) {
    const int i = e / CB, j = e % CB;
    if (j < nc) T[(size_t)i * ldt + c0 + j] = Ys[i * (CB + 1) + j];
  }
}

// dx = Y^T y; large-update guard; state and clone increments (incrementState_IMUCam).
__global__ void __launch_bounds__(256) k_apply_dx(UpdArgs a) {
  __shared__ double dxs[ORCVIO_LEG + 6 * ORCVIO_MAX_OBS];
  const int fi = blockIdx.x;
  const FilterWork fw = a.fw[fi];
  if (!fw.active) return;
  const int n = 6 * fw.N, D = fw.D;
  const double* T = a.T + (size_t)fi * a.t_stride;
  const double* yv = a.yv + (size_t)fi * a.ldr;
  double* imu = a.imu + (size_t)fi * IM_STRIDE;
  double* clones = a.clones + (size_t)fi * a.clone_stride;
  const int tid = threadIdx.x;
  for (int i = tid; i < D; i += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += T[(size_t)k * a.ldt + i] * yv[k];
    dxs[i] = s;
    if (a.dx) a.dx[(size_t)fi * a.lddx + i] = s;
  }
  __syncthreads();
  cta_increment_state(dxs, imu, clones, fw.N, a.flags, a.dx ? &a.dx[(size_t)fi * a.lddx + a.lddx - 1] : nullptr);
}

void launch_update_tail(const UpdArgs& a, int max_N, cudaStream_t s) {
  const int nmax = 6 * max_N, Dmax = ORCVIO_LEG + nmax;
  const int B = a.n_filters;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  size_t sm_trsm = ((size_t)nmax * (CB + 1) + CB * (CB + 1)) * sizeof(double);
  dim3 g3((Dmax + CB - 1) / CB, B);
  k_trsm<<<g3, 256, sm_trsm, s>>>(a, nmax);
  check_launch("k_trsm");
  k_apply_dx<<<B, 256, 0, s>>>(a);
  check_launch("k_apply_dx");
}

void launch_update(const UpdArgs& a, int max_N, cudaStream_t s, int* launches) {
  const int nmax = 6 * max_N, Dmax = ORCVIO_LEG + nmax;
  const int B = a.n_filters;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_chol_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  dim3 g0((Dmax + GT - 1) / GT, (nmax + GT - 1) / GT, B);
  k_gemm<MODE_RP><<<g0, 256, 0, s>>>(a);
  check_launch("k_gemm<RP>");
  dim3 g1((nmax + GT - 1) / GT, (nmax + GT - 1) / GT, B);
  k_gemm<MODE_S><<<g1, 256, 0, s>>>(a);
  check_launch("k_gemm<S>");
  size_t sm_chol = (size_t)2 * (nmax + 1) * (CB + 1) * sizeof(double);
  k_chol_solve<<<B, 256, sm_chol, s>>>(a, nmax);
  check_launch("k_chol_solve");
  size_t sm_trsm = ((size_t)nmax * (CB + 1) + CB * (CB + 1)) * sizeof(double);
  dim3 g3((Dmax + CB - 1) / CB, B);
  k_trsm<<<g3, 256, sm_trsm, s>>>(a, nmax);
  check_launch("k_trsm");
  k_apply_dx<<<B, 256, 0, s>>>(a);
  check_launch("k_apply_dx");
  dim3 g5((Dmax + GT - 1) / GT, (Dmax + GT - 1) / GT, B);
  k_gemm<MODE_P><<<g5, 256, 0, s>>>(a);
  check_launch("k_gemm<P>");
  if (launches) *launches += 6;
}

}  // namespace ob
