// Small fixed-size FP64 helpers shared by host code and kernels.
// Row-major 3x3 matrices in double[9].  Formulas follow the reference's math utils
// (include/orcvio/utils/math_utils.hpp, se3_ops.hpp) and Sophus' SO3 exponential.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace ob {

HD void m3_mul(const double* A, const double* B, double* C) {  // C = A B
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
}
HD void m3_mulT(const double* A, const double* B, double* C) {  // C = A B^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = (A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1]) + A[3 * i + 2] * B[3 * j + 2];
}
HD void m3_Tmul(const double* A, const double* B, double* C) {  // C = A^T B
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = (A[i] * B[j] + A[3 + i] * B[3 + j]) + A[6 + i] * B[6 + j];
}
HD void m3_vec(const double* A, const double* v, double* o) {  // o = A v
  for (int i = 0; i < 3; ++i) o[i] = (A[3 * i] * v[0] + A[3 * i + 1] * v[1]) + A[3 * i + 2] * v[2];
}
// o = A^-1 v by cofactors (Eigen's fixed-size 3 x 3 `.inverse()`): the reference inverts camera rotations that are only
// orthonormal to the digits of the yaml extrinsics (src/orcvio.cpp:2707, 3638), so the transpose is not a substitute
HD void m3_inv_vec(const double* A, const double* v, double* o) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = (A[0] * c00 + A[1] * c01) + A[2] * c02;
  const double id = 1.0 / det;
  const double i00 = c00 * id, i01 = (A[2] * A[7] - A[1] * A[8]) * id, i02 = (A[1] * A[5] - A[2] * A[4]) * id;
  const double i10 = c01 * id, i11 = (A[0] * A[8] - A[2] * A[6]) * id, i12 = (A[2] * A[3] - A[0] * A[5]) * id;
  const double i20 = c02 * id, i21 = (A[1] * A[6] - A[0] * A[7]) * id, i22 = (A[0] * A[4] - A[1] * A[3]) * id;
  o[0] = (i00 * v[0] + i01 * v[1]) + i02 * v[2];
  o[1] = (i10 * v[0] + i11 * v[1]) + i12 * v[2];
  o[2] = (i20 * v[0] + i21 * v[1]) + i22 * v[2];
}
HD void m3_Tvec(const double* A, const double* v, double* o) {  // o = A^T v
  for (int i = 0; i < 3; ++i) o[i] = (A[i] * v[0] + A[3 + i] * v[1]) + A[6 + i] * v[2];
}
HD void m3_skew(const double* w, double* S) {  // math_utils.hpp:27-39
  S[0] = 0;      S[1] = -w[2];  S[2] = w[1];
  S[3] = w[2];   S[4] = 0;      S[5] = -w[0];
  S[6] = -w[1];  S[7] = w[0];   S[8] = 0;
}
HD void m3_eye(double* A) {
  for (int i = 0; i < 9; ++i) A[i] = 0;
  A[0] = A[4] = A[8] = 1;
}
HD double v3_norm(const double* v) { return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

// Eigen::Quaterniond(w,x,y,z).toRotationMatrix()
HD void quat_wxyz_to_R(double w, double x, double y, double z, double* R) {
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w;
  double txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Sophus::SO3d::exp(omega).matrix()  (reference call sites src/orcvio.cpp:919,4331,4497,4542)
HD void so3_exp(const double* om, double* R) {
  double th2 = (om[0] * om[0] + om[1] * om[1]) + om[2] * om[2];
  double th = sqrt(th2);
  double half = 0.5 * th;
  double imag, real;
  if (th < 1e-10) {
    double p4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * p4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * p4;
  } else {
    imag = sin(half) / th;
    real = cos(half);
  }
  quat_wxyz_to_R(real, imag * om[0], imag * om[1], imag * om[2], R);
}

// Jl_operator / Hl_operator, math_utils.hpp:230-270
HD void Jl_op(const double* g, double* J) {
  double n = v3_norm(g);
  m3_eye(J);
  if (n < 1.0e-5) return;
  double S[9], SS[9];
  m3_skew(g, S);
  m3_mul(S, S, SS);
  double a = (1 - cos(n)) / (n * n);
  double b = (n - sin(n)) / (n * n * n);
  for (int i = 0; i < 9; ++i) J[i] = J[i] + a * S[i] + b * SS[i];
}
HD void Hl_op(const double* g, double* H) {
  double n = v3_norm(g);
  for (int i = 0; i < 9; ++i) H[i] = 0;
  H[0] = H[4] = H[8] = 0.5;
  if (n < 1.0e-5) return;
  double S[9], SS[9];
  m3_skew(g, S);
  m3_mul(S, S, SS);
  double a = (n - sin(n)) / (n * n * n);
  double b = (2 * (cos(n) - 1) + n * n) / (2 * n * n * n * n);
  for (int i = 0; i < 9; ++i) H[i] = H[i] + a * S[i] + b * SS[i];
}

// quaternionToRotation ([x,y,z,w]), math_utils.hpp:164-177
HD void quat_xyzw_to_R(const double* q, double* R) {
  double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
  R[0] = 1 - 2 * (qy * qy + qz * qz); R[1] = 2 * (qx * qy - qw * qz);     R[2] = 2 * (qx * qz + qw * qy);
  R[3] = 2 * (qx * qy + qw * qz);     R[4] = 1 - 2 * (qx * qx + qz * qz); R[5] = 2 * (qy * qz - qw * qx);
  R[6] = 2 * (qx * qz - qw * qy);     R[7] = 2 * (qy * qz + qw * qx);     R[8] = 1 - 2 * (qx * qx + qy * qy);
}

// rotationToQuaternion ([x,y,z,w], w >= 0, normalised), math_utils.hpp:188-227
HD void R_to_quat_xyzw(const double* R, double* q) {
  double tr = R[0] + R[4] + R[8];
  double score[4] = {R[0], R[4], R[8], tr};
  int k = 0;
  for (int i = 1; i < 4; ++i)
    if (score[i] > score[k]) k = i;
  if (k == 0) {
    q[0] = sqrt(1 + 2 * R[0] - tr) / 2.0;
    q[1] = (R[1] + R[3]) / (4 * q[0]);
    q[2] = (R[2] + R[6]) / (4 * q[0]);
    q[3] = (R[7] - R[5]) / (4 * q[0]);
  } else if (k == 1) {
    q[1] = sqrt(1 + 2 * R[4] - tr) / 2.0;
    q[0] = (R[1] + R[3]) / (4 * q[1]);
    q[2] = (R[5] + R[7]) / (4 * q[1]);
    q[3] = (R[2] - R[6]) / (4 * q[1]);
  } else if (k == 2) {
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
  if (q[3] < 0) {
    for (int i = 0; i < 4; ++i) q[i] = -q[i];
  }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}

// Eigen::AngleAxisd(R).angle() via Eigen's Quaternion(Matrix3) (src/orcvio.cpp:2606-2607)
HD double angle_axis_angle(const double* R) {
  double t = R[0] + R[4] + R[8];
  double q[4] = {0, 0, 0, 0};  // w,x,y,z
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t;
    q[2] = (R[2] - R[6]) * t;
    q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[1 + j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[1 + k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  double n = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  return 2.0 * atan2(n, fabs(q[0]));
}

}  // namespace ob
