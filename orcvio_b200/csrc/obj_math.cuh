// Small dense helpers shared by the object kernels (obj_kernel.cu: stage-3 rows; objlm_kernel.cu: the LM normal
// equations): fixed-size products, odotOperator / circledCirc (include/orcvio/utils/se3_ops.hpp:229-240, 510-519),
// rigid inverse, project_image_df (:340-352).
#pragma once
#include "kernels.h"

namespace ob {
namespace objm {

template <int M, int K, int N>
__device__ __forceinline__ void mm(const double* A, const double* B, double* C) {   // C(MxN) = A(MxK) B(KxN)
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s += A[i * K + k] * B[k * N + j];
      C[i * N + j] = s;
    }
}
template <int M, int N>
__device__ __forceinline__ void tr(const double* A, double* At) {   // At(NxM) = A(MxN)^T
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) At[j * M + i] = A[i * N + j];
}
__device__ __forceinline__ void skew3(const double* w, double* S) { m3_skew(w, S); }

// odotOperator(x): 4x6 [x4 I, -skew(x123); 0]
__device__ __forceinline__ void odot(const double* x, double* O) {
  double S[9];
  skew3(x, S);
#pragma unroll
  for (int i = 0; i < 24; ++i) O[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    O[i * 6 + i] = x[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) O[i * 6 + 3 + j] = -S[3 * i + j];
  }
}
// circledCirc(x): 6x4, out[3:, :3] = -skew(x123), out[:3, 3] = x123
__device__ __forceinline__ void circ(const double* x, double* Cc) {
  double S[9];
  skew3(x, S);
#pragma unroll
  for (int i = 0; i < 24; ++i) Cc[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Cc[i * 4 + 3] = x[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) Cc[(3 + i) * 4 + j] = -S[3 * i + j];
  }
}
// inverse of a rigid 4x4 (Sophus SE3::inverse)
__device__ __forceinline__ void inv_rigid(const double* T, double* Ti) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) Ti[4 * i + j] = T[4 * j + i];
    Ti[4 * i + 3] = -((T[i] * T[3] + T[4 + i] * T[7]) + T[8 + i] * T[11]);
  }
  Ti[12] = Ti[13] = Ti[14] = 0.0;
  Ti[15] = 1.0;
}
__device__ __forceinline__ void dpi_of(const double* p, double* d) {   // project_image_df, 2x3
  const double z = p[2], zsq = z * z;
  d[0] = 1 / z; d[1] = 0.0; d[2] = -p[0] / zsq;
  d[3] = 0.0; d[4] = 1 / z; d[5] = -p[1] / zsq;
}

// Sophus SE3::log of a rigid transform -> [upsilon, omega]
HD void se3_log(const double* T, double* xi) {
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  const double tr_ = R[0] + R[4] + R[8];
  double q[4];   // x y z w  (rotationToQuaternion, math_utils.hpp:188-227)
  int k = 0;
  double best = R[0];
  if (R[4] > best) { best = R[4]; k = 1; }
  if (R[8] > best) { best = R[8]; k = 2; }
  if (tr_ > best) { best = tr_; k = 3; }
  if (k == 0) {
    q[0] = sqrt(1 + 2 * R[0] - tr_) / 2.0;
    q[1] = (R[1] + R[3]) / (4 * q[0]); q[2] = (R[2] + R[6]) / (4 * q[0]); q[3] = (R[7] - R[5]) / (4 * q[0]);
  } else if (k == 1) {
    q[1] = sqrt(1 + 2 * R[4] - tr_) / 2.0;
    q[0] = (R[1] + R[3]) / (4 * q[1]); q[2] = (R[5] + R[7]) / (4 * q[1]); q[3] = (R[2] - R[6]) / (4 * q[1]);
  } else if (k == 2) {
    q[2] = sqrt(1 + 2 * R[8] - tr_) / 2.0;
    q[0] = (R[2] + R[6]) / (4 * q[2]); q[1] = (R[5] + R[7]) / (4 * q[2]); q[3] = (R[3] - R[1]) / (4 * q[2]);
  } else {
    q[3] = sqrt(1 + tr_) / 2.0;
    q[0] = (R[7] - R[5]) / (4 * q[3]); q[1] = (R[2] - R[6]) / (4 * q[3]); q[2] = (R[3] - R[1]) / (4 * q[3]);
  }
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double qn = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= qn;
  const double n2 = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
  const double n = sqrt(n2), w = q[3];
  const double two_atan = (n2 < 1e-20) ? (2.0 / w - (2.0 / 3.0) * n2 / (w * w * w)) : (2.0 * atan2(n, w) / n);
  double om[3] = {two_atan * q[0], two_atan * q[1], two_atan * q[2]};
  const double th = sqrt((om[0] * om[0] + om[1] * om[1]) + om[2] * om[2]);
  double W[9], W2[9];
  m3_skew(om, W);
  m3_mul(W, W, W2);
  double c2;
  if (th < 1e-10) c2 = 1.0 / 12.0;
  else {
    const double half = 0.5 * th;
    c2 = (1 - th * cos(half) / (2 * sin(half))) / (th * th);
  }
  const double t[3] = {T[3], T[7], T[11]};
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int j = 0; j < 3; ++j) s += ((i == j ? 1.0 : 0.0) - 0.5 * W[3 * i + j] + c2 * W2[3 * i + j]) * t[j];
    xi[i] = s;
    xi[3 + i] = om[i];
  }
}


// Sophus SE3::exp([upsilon, omega]) -> row-major 4 x 4: R = exp(omega), t = V(omega) upsilon
HD void se3_exp(const double* xi, double* T) {
  double R[9], W[9], W2[9];
  so3_exp(xi + 3, R);
  const double th = v3_norm(xi + 3);
  m3_skew(xi + 3, W);
  m3_mul(W, W, W2);
  double V[9];
  if (th < 1e-10) {
    for (int i = 0; i < 9; ++i) V[i] = R[i];
  } else {
    const double a = (1 - cos(th)) / (th * th), b = (th - sin(th)) / (th * th * th);
    for (int i = 0; i < 9; ++i) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * W[i] + b * W2[i];
  }
  double t[3];
  m3_vec(V, xi, t);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
    T[4 * i + 3] = t[i];
  }
  T[12] = T[13] = T[14] = 0.0;
  T[15] = 1.0;
}

}  // namespace objm
}  // namespace ob
