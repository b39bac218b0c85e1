// 3 x 3 singular value decomposition by one-sided Jacobi (Hestenes) rotations and the 3 x 3 determinant, shared by the
// Kabsch fit (kabsch_kernel.cu) and the Umeyama trajectory alignment (metrics_kernel.cu).
#pragma once
#include "kernels.h"

namespace ob {
namespace svd3 {

__device__ inline void svd3_hestenes(const double* C, double* U, double* S, double* V) {
  double A[9];                               // columns rotated in place: A = C V
  for (int i = 0; i < 9; ++i) A[i] = C[i];
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < 3; ++i) {
          al += A[3 * i + p] * A[3 * i + p];
          be += A[3 * i + q] * A[3 * i + q];
          ga += A[3 * i + p] * A[3 * i + q];
        }
        if (fabs(ga) <= 1e-300 || fabs(ga) <= 2.3e-16 * sqrt(al * be)) continue;
        off = fmax(off, fabs(ga) / sqrt(al * be));
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; ++i) {
          const double ap = A[3 * i + p], aq = A[3 * i + q];
          A[3 * i + p] = c * ap - s * aq;
          A[3 * i + q] = s * ap + c * aq;
          const double vp = V[3 * i + p], vq = V[3 * i + q];
          V[3 * i + p] = c * vp - s * vq;
          V[3 * i + q] = s * vp + c * vq;
        }
      }
    if (off == 0.0) break;
  }
  double nrm[3];
  for (int j = 0; j < 3; ++j) nrm[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  int ord[3] = {0, 1, 2};                    // singular values in descending order
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (nrm[ord[b]] > nrm[ord[a]]) { const int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
  double Vs[9], Us[9];
  for (int j = 0; j < 3; ++j) {
    S[j] = nrm[ord[j]];
    for (int i = 0; i < 3; ++i) {
      Vs[3 * i + j] = V[3 * i + ord[j]];
      Us[3 * i + j] = nrm[ord[j]] > 0 ? A[3 * i + ord[j]] / nrm[ord[j]] : 0.0;
    }
  }
  if (S[2] <= 1e-13 * S[0]) {                // rank 2: complete the left basis
    Us[2] = Us[3] * Us[7] - Us[6] * Us[4];
    Us[5] = Us[6] * Us[1] - Us[0] * Us[7];
    Us[8] = Us[0] * Us[4] - Us[3] * Us[1];
  }
  for (int i = 0; i < 9; ++i) { U[i] = Us[i]; V[i] = Vs[i]; }
}

__device__ inline double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

}  // namespace svd3
}  // namespace ob
