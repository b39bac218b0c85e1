// CPU oracle, C++ restatement -- TEST / BENCH INFRASTRUCTURE, not product code.
//
// A second, independent CPU restatement of the reference's frame update ("stack -> compress ->
// update"), used (a) as a second opinion for the NumPy oracle (tests/test_oracle_cpu.py checks
// that the two agree) and (b) as the timed CPU baseline of bench.py.  One function per reference
// function, DENSE where the reference is dense; citations are /root/reference paths:
//   Feature::checkMotion / triangulate_position   include/orcvio/feat/feature.hpp:353-396, 583-719
//   OrcVIO::measurementJacobian_msckf             src/orcvio.cpp:1071-1168
//   OrcVIO::featureJacobian_msckf                 src/orcvio.cpp:1171-1226
//   nullspace_project_inplace_svd                 include/orcvio/utils/math_utils.hpp:287-312
//   OrcVIO::gatingTestFeature                     src/orcvio.cpp:1953-1976
//   stacking + SPQR compression                   src/orcvio.cpp:2498-2556
//   OrcVIO::measurementUpdate_hybrid              src/orcvio.cpp:1766-1950 (empty EKF parts)
//   OrcVIO::incrementState_IMUCam                 src/orcvio.cpp:4468-4567
// Third-party numerics that are absent from /root/reference are restated by their published
// algorithms: Eigen LDLT (3x3, pivoted) for the LM step, Householder QR for JacobiSVD's left
// nullspace and for SuiteSparseQR (any orthogonal basis gives the same gate value and posterior,
// SURVEY 0 #1/#3), Cholesky for S.ldlt().solve, boost::math chi-square quantile by inverting the
// regularised incomplete gamma function.  The compression skips structural zeros (rows sorted by
// first clone, row-profile Householder) because SPQR does -- this favours the CPU baseline.
//
// parity: UNPINNED by the reference's own tests for these functions; pinned by
// agreement NumPy oracle == this file == CUDA path.  Build: oracle/Makefile
// (g++ -O2 -ffp-contract=off: the reference is built for generic x86-64, no FMA contraction).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

namespace {

constexpr int LEG = 22;
constexpr int FL_LARVIO = 1, FL_LEFT = 2, FL_DISCARD = 4;

// ------------------------------------------------------------------ small fixed-size helpers
inline void matT_mat(const double* A, const double* B, double* o) {   // A^T B
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = (A[i] * B[j] + A[3 + i] * B[3 + j]) + A[6 + i] * B[6 + j];
}
inline void matT_vec(const double* A, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = (A[i] * v[0] + A[3 + i] * v[1]) + A[6 + i] * v[2];
}
inline void mat_vec(const double* A, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = (A[3 * i] * v[0] + A[3 * i + 1] * v[1]) + A[3 * i + 2] * v[2];
}
inline void mat_mat(const double* A, const double* B, double* o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      o[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
}
inline void mat_matT(const double* A, const double* B, double* o) {   // A B^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      o[3 * i + j] = (A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1]) + A[3 * i + 2] * B[3 * j + 2];
}
inline void skew(const double* w, double* S) {
  S[0] = 0; S[1] = -w[2]; S[2] = w[1];
  S[3] = w[2]; S[4] = 0; S[5] = -w[0];
  S[6] = -w[1]; S[7] = w[0]; S[8] = 0;
}

// general dense C (m x n) = A (m x k) * B (k x n), row-major
void gemm(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; ++i) {
    double* c = C + (size_t)i * n;
    for (int j = 0; j < n; ++j) c[j] = 0.0;
    for (int p = 0; p < k; ++p) {
      const double a = A[(size_t)i * k + p];
      if (a == 0.0) continue;      // Eigen's dense product does not skip; skipping favours the CPU
      const double* b = B + (size_t)p * n;
      for (int j = 0; j < n; ++j) c[j] += a * b[j];
    }
  }
}

// ------------------------------------------------------------------ chi-square quantile
double gamma_p(double a, double x) {          // regularised lower incomplete gamma P(a, x)
  if (x <= 0) return 0.0;
  const double gln = std::lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 2000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - gln);
  }
  double b = x + 1.0 - a, c = 1.0 / 1e-300, d = 1.0 / b, h = d;
  for (int i = 1; i < 2000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < 1e-300) d = 1e-300;
    c = b + an / c;
    if (std::fabs(c) < 1e-300) c = 1e-300;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - gln) * h;
}

double chi2_quantile(double p, int dof) {     // boost::math::quantile(chi_squared(dof), p)
  const double a = 0.5 * dof;
  double lo = 0.0, hi = std::max(4.0 * dof, 50.0);
  while (gamma_p(a, 0.5 * hi) < p) hi *= 2.0;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (gamma_p(a, 0.5 * mid) < p) lo = mid;
    else hi = mid;
    if (hi - lo <= 1e-15 * hi) break;
  }
  double x = 0.5 * (lo + hi);
  for (int it = 0; it < 3; ++it) {            // Newton polish on the CDF
    const double f = gamma_p(a, 0.5 * x) - p;
    const double pdf = std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - std::lgamma(a)) * 0.5;
    if (pdf > 0) x -= f / pdf;
  }
  return x;
}

// ------------------------------------------------------------------ triangulation
struct TriCfg {
  double translation_threshold = 0.2, huber_epsilon = 0.01, estimation_precision = 5e-7,
         initial_damping = 1e-3, cost_threshold = 4.7673e-4, init_final_dist_threshold = 5.0;
  int outer_max = 10, inner_max = 10;
};

// Eigen::LDLT<Matrix3d>::compute / solve (largest-|diagonal| pivoting), feature.hpp:649
void ldlt3_solve(const double* M, const double* b, double* x) {
  const int n = 3;
  double a[3][3];
  int tr[3];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[i][j] = M[3 * i + j];
  for (int k = 0; k < n; ++k) {
    int big = k;
    double bigv = std::fabs(a[k][k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(a[i][i]) > bigv) { bigv = std::fabs(a[i][i]); big = i; }
    tr[k] = big;
    if (big != k) {
      for (int j = 0; j < k; ++j) std::swap(a[k][j], a[big][j]);
      for (int i = big + 1; i < n; ++i) std::swap(a[i][k], a[i][big]);
      std::swap(a[k][k], a[big][big]);
      for (int i = k + 1; i < big; ++i) std::swap(a[i][k], a[big][i]);
    }
    if (k > 0) {
      double temp[3];
      for (int j = 0; j < k; ++j) temp[j] = a[j][j] * a[k][j];
      double s = 0.0;
      for (int j = 0; j < k; ++j) s = s + a[k][j] * temp[j];
      a[k][k] = a[k][k] - s;
      for (int i = k + 1; i < n; ++i) {
        s = 0.0;
        for (int j = 0; j < k; ++j) s = s + a[i][j] * temp[j];
        a[i][k] = a[i][k] - s;
      }
    }
    const double akk = a[k][k];
    if (std::fabs(akk) > 0.0)
      for (int i = k + 1; i < n; ++i) a[i][k] = a[i][k] / akk;
  }
  for (int i = 0; i < n; ++i) x[i] = b[i];
  for (int k = 0; k < n; ++k)
    if (tr[k] != k) std::swap(x[k], x[tr[k]]);
  for (int i = 0; i < n; ++i)
    for (int r = i + 1; r < n; ++r) x[r] = x[r] - x[i] * a[r][i];
  const double tol = 2.2250738585072014e-308;
  for (int i = 0; i < n; ++i) x[i] = (std::fabs(a[i][i]) > tol) ? x[i] / a[i][i] : 0.0;
  for (int i = n - 2; i >= 0; --i) {
    double s = 0.0;
    for (int j = i + 1; j < n; ++j) s = s + a[j][i] * x[j];
    x[i] = x[i] - s;
  }
  for (int k = n - 1; k >= 0; --k)
    if (tr[k] != k) std::swap(x[k], x[tr[k]]);
}

inline void h_of(const double* R, const double* t, const double* x, double* h) {
  h[0] = ((R[0] * x[0] + R[1] * x[1]) + R[2] * 1.0) + x[2] * t[0];
  h[1] = ((R[3] * x[0] + R[4] * x[1]) + R[5] * 1.0) + x[2] * t[1];
  h[2] = ((R[6] * x[0] + R[7] * x[1]) + R[8] * 1.0) + x[2] * t[2];
}
inline double tri_cost(const double* R, const double* t, const double* x, const double* z) {  // :271-291
  double h[3];
  h_of(R, t, x, h);
  const double d0 = h[0] / h[2] - z[0], d1 = h[1] / h[2] - z[1];
  return d0 * d0 + d1 * d1;
}
inline void tri_jacobian(const double* R, const double* t, const double* x, const double* z, double eps,
                         double J[2][3], double* r, double* w) {                               // :293-329
  double h[3];
  h_of(R, t, x, h);
  const double W[3][3] = {{R[0], R[1], t[0]}, {R[3], R[4], t[1]}, {R[6], R[7], t[2]}};
  const double ih3 = 1 / h[2], c0 = h[0] / (h[2] * h[2]), c1 = h[1] / (h[2] * h[2]);
  for (int j = 0; j < 3; ++j) {
    J[0][j] = ih3 * W[0][j] - c0 * W[2][j];
    J[1][j] = ih3 * W[1][j] - c1 * W[2][j];
  }
  r[0] = h[0] / h[2] - z[0];
  r[1] = h[1] / h[2] - z[1];
  const double e = std::sqrt(r[0] * r[0] + r[1] * r[1]);
  *w = (e <= eps) ? 1.0 : std::sqrt(2.0 * eps / e);
}

bool check_motion(const double* R_first, const double* t_first, const double* t_last, const double* z_first,
                  double thr) {                                                                // :353-396
  double d[3] = {z_first[0], z_first[1], 1.0};
  const double n = std::sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
  d[0] /= n; d[1] /= n; d[2] /= n;
  double dw[3];
  mat_vec(R_first, d, dw);
  const double tr[3] = {t_last[0] - t_first[0], t_last[1] - t_first[1], t_last[2] - t_first[2]};
  const double par = (tr[0] * dw[0] + tr[1] * dw[1]) + tr[2] * dw[2];
  const double o[3] = {tr[0] - par * dw[0], tr[1] - par * dw[1], tr[2] - par * dw[2]};
  return std::sqrt((o[0] * o[0] + o[1] * o[1]) + o[2] * o[2]) > thr;
}

// Feature::triangulate_position, feature.hpp:583-719 (not previously initialised).
bool triangulate(int m, const double* const* cam_R, const double* const* cam_t, const double* meas,
                 const TriCfg& cfg, double* pos_w, int* n_outer, int* n_inner, double* final_cost) {
  std::vector<double> rel_R((size_t)9 * m), rel_t((size_t)3 * m);
  const double* Rl = cam_R[m - 1];
  const double* tl = cam_t[m - 1];
  for (int i = 0; i < m; ++i) {
    double tinv[3], rt[3];
    matT_vec(cam_R[i], cam_t[i], tinv);
    matT_mat(cam_R[i], Rl, &rel_R[9 * i]);
    matT_vec(cam_R[i], tl, rt);
    for (int k = 0; k < 3; ++k) rel_t[3 * i + k] = rt[k] + (-tinv[k]);
  }
  // generateInitialGuess (:331-351) from the last and the first view
  double init[3];
  {
    const double* R = &rel_R[0];
    const double* t = &rel_t[0];
    const double* z1 = meas + 2 * (m - 1);
    const double* z2 = meas;
    const double v[3] = {z1[0], z1[1], 1.0};
    double mm[3];
    mat_vec(R, v, mm);
    const double A0 = mm[0] - z2[0] * mm[2], A1 = mm[1] - z2[1] * mm[2];
    const double b0 = z2[0] * t[2] - t[0], b1 = z2[1] * t[2] - t[1];
    const double inv = 1.0 / (A0 * A0 + A1 * A1);
    const double depth = (inv * A0) * b0 + (inv * A1) * b1;
    init[0] = z1[0] * depth; init[1] = z1[1] * depth; init[2] = depth;
  }
  double sol[3] = {init[0] / init[2], init[1] / init[2], 1.0 / init[2]};
  double lam = cfg.initial_damping;
  int inner = 0, outer = 0, inner_total = 0;
  bool reduced = false;
  double delta_norm = 0.0, total_cost = 0.0;
  for (int i = 0; i < m; ++i) total_cost = total_cost + tri_cost(&rel_R[9 * i], &rel_t[3 * i], sol, meas + 2 * i);
  while (true) {
    double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    for (int i = 0; i < m; ++i) {
      double J[2][3], r[2], w;
      tri_jacobian(&rel_R[9 * i], &rel_t[3 * i], sol, meas + 2 * i, cfg.huber_epsilon, J, r, &w);
      if (w == 1) {
        for (int a_ = 0; a_ < 3; ++a_) {
          for (int c_ = 0; c_ < 3; ++c_) A[3 * a_ + c_] = A[3 * a_ + c_] + (J[0][a_] * J[0][c_] + J[1][a_] * J[1][c_]);
          b[a_] = b[a_] + (J[0][a_] * r[0] + J[1][a_] * r[1]);
        }
      } else {
        const double w2 = w * w;
        for (int a_ = 0; a_ < 3; ++a_) {
          for (int c_ = 0; c_ < 3; ++c_)
            A[3 * a_ + c_] = A[3 * a_ + c_] + ((w2 * J[0][a_]) * J[0][c_] + (w2 * J[1][a_]) * J[1][c_]);
          b[a_] = b[a_] + ((w2 * J[0][a_]) * r[0] + (w2 * J[1][a_]) * r[1]);
        }
      }
    }
    bool cont;
    do {
      double M[9];
      std::memcpy(M, A, sizeof(M));
      M[0] = A[0] + lam; M[4] = A[4] + lam; M[8] = A[8] + lam;
      double delta[3];
      ldlt3_solve(M, b, delta);
      const double ns[3] = {sol[0] - delta[0], sol[1] - delta[1], sol[2] - delta[2]};
      delta_norm = std::sqrt((delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2]);
      double new_cost = 0.0;
      for (int i = 0; i < m; ++i) new_cost = new_cost + tri_cost(&rel_R[9 * i], &rel_t[3 * i], ns, meas + 2 * i);
      ++inner_total;
      if (new_cost < total_cost) {
        reduced = true;
        sol[0] = ns[0]; sol[1] = ns[1]; sol[2] = ns[2];
        total_cost = new_cost;
        lam = lam / 10 > 1e-10 ? lam / 10 : 1e-10;
      } else {
        reduced = false;
        lam = lam * 10 < 1e12 ? lam * 10 : 1e12;
      }
      cont = (inner < cfg.inner_max) && !reduced;
      ++inner;
    } while (cont);
    inner = 0;
    const bool cont_outer = (outer < cfg.outer_max) && (delta_norm > cfg.estimation_precision);
    ++outer;
    if (!cont_outer) break;
  }
  const double fin[3] = {sol[0] / sol[2], sol[1] / sol[2], 1.0 / sol[2]};
  bool valid = true;
  for (int i = 0; i < m; ++i) {
    const double* R = &rel_R[9 * i];
    const double pz = ((R[6] * fin[0] + R[7] * fin[1]) + R[8] * fin[2]) + rel_t[3 * i + 2];
    if (pz <= 0) { valid = false; break; }
  }
  const double normalized_cost = total_cost / (double)(2 * m * m);
  const double d[3] = {fin[0] - init[0], fin[1] - init[1], fin[2] - init[2]};
  if (std::sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) > cfg.init_final_dist_threshold) valid = false;
  if (normalized_cost > cfg.cost_threshold) valid = false;
  double pw[3];
  mat_vec(Rl, fin, pw);
  for (int k = 0; k < 3; ++k) pos_w[k] = pw[k] + tl[k];
  if (n_outer) *n_outer = outer;
  if (n_inner) *n_inner = inner_total;
  if (final_cost) *final_cost = total_cost;
  return valid;
}

// ------------------------------------------------------------------ measurement Jacobian
// OrcVIO::measurementJacobian_msckf, src/orcvio.cpp:1071-1168.  H_x 2x6, H_f 2x3, r 2 (row-major).
void measurement_jacobian(int flags, const double* R_b2w, const double* t_b_w, const double* R_b2c,
                          const double* t_c_b, const double* p_w, const double* z, double* Hx, double* Hf,
                          double* r) {
  double R_w2c[9], Rt[3], t_c_w[3], d[3], p_c[3];
  mat_matT(R_b2c, R_b2w, R_w2c);
  mat_vec(R_b2w, t_c_b, Rt);
  for (int k = 0; k < 3; ++k) { t_c_w[k] = t_b_w[k] + Rt[k]; d[k] = p_w[k] - t_c_w[k]; }
  mat_vec(R_w2c, d, p_c);
  const double iz = 1 / p_c[2];
  const double dz[6] = {iz, 0, -p_c[0] / (p_c[2] * p_c[2]), 0, iz, -p_c[1] / (p_c[2] * p_c[2])};
  double dpc[18];      // 3 x 6
  double sign;
  if (flags & FL_LARVIO) {                       // :1147-1149
    double pbf[3] = {p_w[0] - t_b_w[0], p_w[1] - t_b_w[1], p_w[2] - t_b_w[2]}, Sk[9], A[9];
    skew(pbf, Sk);
    mat_mat(R_w2c, Sk, A);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { dpc[6 * i + j] = A[3 * i + j]; dpc[6 * i + 3 + j] = -R_w2c[3 * i + j]; }
    sign = 1.0;
  } else {
    // cTw = wTc.inverse() -- Eigen's general inverse of the 4x4 [R_w2c^T, t_c_w; 0 1]  (:1136, :1140)
    double Minv[9], tinv[3];
    {
      double M[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = R_w2c[3 * j + i];
      const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
      const double c10 = M[2] * M[7] - M[1] * M[8], c11 = M[0] * M[8] - M[2] * M[6], c12 = M[1] * M[6] - M[0] * M[7];
      const double c20 = M[1] * M[5] - M[2] * M[4], c21 = M[2] * M[3] - M[0] * M[5], c22 = M[0] * M[4] - M[1] * M[3];
      const double det = (M[0] * c00 + M[1] * c01) + M[2] * c02;
      Minv[0] = c00 / det; Minv[1] = c10 / det; Minv[2] = c20 / det;
      Minv[3] = c01 / det; Minv[4] = c11 / det; Minv[5] = c21 / det;
      Minv[6] = c02 / det; Minv[7] = c12 / det; Minv[8] = c22 / det;
      mat_vec(Minv, t_c_w, tinv);
      for (int k = 0; k < 3; ++k) tinv[k] = -tinv[k];
    }
    // get_cam_wrt_imu_se3_jacobian, se3_ops.hpp:531-552 (6x6, tangent order [trans, rot] x [theta, p])
    double J[36];
    std::memset(J, 0, sizeof(J));
    double X[12];      // temp * (cTw * odot(.)) or temp * odot(cTw * .): 3 x 6 = [x4 I | -[x]x]
    if (flags & FL_LEFT) {
      double Sp[9];
      skew(t_b_w, Sp);
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) J[6 * i + j] = Sp[3 * i + j];
        J[6 * (3 + i) + i] = 1.0;
        J[6 * i + 3 + i] = 1.0;
      }
      // cTw[0:3,:] * odot([p_w;1]) = [Minv * 1 | -Minv [p_w]x]  (the translation column meets a zero row)
      double Sk[9], MS[9];
      skew(p_w, Sk);
      mat_mat(Minv, Sk, MS);
      double O[18];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { O[6 * i + j] = Minv[3 * i + j]; O[6 * i + 3 + j] = -MS[3 * i + j]; }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += O[6 * i + k] * J[6 * k + j];
          dpc[6 * i + j] = s;
        }
      (void)X;
    } else {
      double St[9], RS[9];
      skew(t_c_b, St);
      mat_mat(R_b2c, St, RS);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          J[6 * i + j] = -RS[3 * i + j];
          J[6 * (3 + i) + j] = R_b2c[3 * i + j];
          J[6 * i + 3 + j] = R_w2c[3 * i + j];
        }
      double pci[3], Mp[3];
      mat_vec(Minv, p_w, Mp);
      for (int k = 0; k < 3; ++k) pci[k] = Mp[k] + tinv[k];
      double Sk[9], O[18];
      skew(pci, Sk);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { O[6 * i + j] = (i == j) ? 1.0 : 0.0; O[6 * i + 3 + j] = -Sk[3 * i + j]; }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += O[6 * i + k] * J[6 * k + j];
          dpc[6 * i + j] = s;
        }
    }
    sign = -1.0;
  }
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < 6; ++j)
      Hx[6 * i + j] = sign * ((dz[3 * i] * dpc[j] + dz[3 * i + 1] * dpc[6 + j]) + dz[3 * i + 2] * dpc[12 + j]);
    for (int j = 0; j < 3; ++j)
      Hf[3 * i + j] = (dz[3 * i] * R_w2c[j] + dz[3 * i + 1] * R_w2c[3 + j]) + dz[3 * i + 2] * R_w2c[6 + j];
  }
  r[0] = z[0] - p_c[0] / p_c[2];
  r[1] = z[1] - p_c[1] / p_c[2];
}

// ------------------------------------------------------------------ dense linear algebra
// Householder QR of A (m x n, row-major, ld) in place, applying Q^T to `ncarry` extra dense
// blocks is done by the caller through `apply`.  Returns the reflectors in A's lower part.
struct Reflector { double tau; };

// Left nullspace projection (math_utils.hpp:287-312): A = last rows-3 columns of the full
// orthogonal factor of H_f (2m x 3); H_x <- A^T H_x, r <- A^T r.  Dense over all D columns.
void nullspace_project(std::vector<double>& Hf, std::vector<double>& Hx, std::vector<double>& r, int rows, int D) {
  for (int k = 0; k < 3; ++k) {
    double sig = 0.0;
    for (int i = k + 1; i < rows; ++i) sig += Hf[3 * i + k] * Hf[3 * i + k];
    const double akk = Hf[3 * k + k];
    if (sig == 0.0) continue;
    const double mu = std::sqrt(akk * akk + sig);
    const double v0 = (akk <= 0.0) ? (akk - mu) : (-sig / (akk + mu));
    const double tau = 2.0 * v0 * v0 / (sig + v0 * v0);
    std::vector<double> v(rows, 0.0);
    v[k] = 1.0;
    for (int i = k + 1; i < rows; ++i) v[i] = Hf[3 * i + k] / v0;
    auto apply = [&](double* M, int ld, int ncols) {
      for (int j = 0; j < ncols; ++j) {
        double dot = 0.0;
        for (int i = k; i < rows; ++i) dot += v[i] * M[(size_t)i * ld + j];
        const double s = tau * dot;
        if (s == 0.0) continue;
        for (int i = k; i < rows; ++i) M[(size_t)i * ld + j] -= s * v[i];
      }
    };
    apply(Hf.data(), 3, 3);
    apply(Hx.data(), D, D);
    apply(r.data(), 1, 1);
  }
}

// Cholesky solve S x = b for SPD S (n x n, row-major); nrhs right-hand sides stored row-major n x nrhs.
bool chol_solve(std::vector<double>& S, int n, double* B, int nrhs) {
  for (int k = 0; k < n; ++k) {
    double dkk = S[(size_t)k * n + k];
    for (int p = 0; p < k; ++p) dkk -= S[(size_t)k * n + p] * S[(size_t)k * n + p];
    if (!(dkk > 0.0)) return false;
    dkk = std::sqrt(dkk);
    S[(size_t)k * n + k] = dkk;
    for (int i = k + 1; i < n; ++i) {
      double s = S[(size_t)i * n + k];
      const double* li = &S[(size_t)i * n];
      const double* lk = &S[(size_t)k * n];
      for (int p = 0; p < k; ++p) s -= li[p] * lk[p];
      S[(size_t)i * n + k] = s / dkk;
    }
  }
  // forward: L Y = B
  for (int i = 0; i < n; ++i) {
    double* bi = B + (size_t)i * nrhs;
    for (int p = 0; p < i; ++p) {
      const double l = S[(size_t)i * n + p];
      if (l == 0.0) continue;
      const double* bp = B + (size_t)p * nrhs;
      for (int j = 0; j < nrhs; ++j) bi[j] -= l * bp[j];
    }
    const double inv = 1.0 / S[(size_t)i * n + i];
    for (int j = 0; j < nrhs; ++j) bi[j] *= inv;
  }
  // backward: L^T X = Y
  for (int i = n - 1; i >= 0; --i) {
    double* bi = B + (size_t)i * nrhs;
    const double inv = 1.0 / S[(size_t)i * n + i];
    for (int j = 0; j < nrhs; ++j) bi[j] *= inv;
    for (int p = 0; p < i; ++p) {
      const double l = S[(size_t)i * n + p];
      if (l == 0.0) continue;
      double* bp = B + (size_t)p * nrhs;
      for (int j = 0; j < nrhs; ++j) bp[j] -= l * bi[j];
    }
  }
  return true;
}

// gatingTestFeature, src/orcvio.cpp:1953-1976: dense H (r x D) P (D x D) H^T like the reference.
double gating_gamma(const double* H, const double* r, int rows, int D, const double* P, double sigma2,
                    std::vector<double>& work) {
  work.resize((size_t)rows * D);
  gemm(H, P, work.data(), rows, D, D);
  std::vector<double> S((size_t)rows * rows);
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < rows; ++j) {
      double s = 0.0;
      const double* a = &work[(size_t)i * D];
      const double* b = H + (size_t)j * D;
      for (int k = 0; k < D; ++k) s += a[k] * b[k];
      S[(size_t)i * rows + j] = s + (i == j ? sigma2 : 0.0);
    }
  std::vector<double> x(r, r + rows);
  if (!chol_solve(S, rows, x.data(), 1)) return 1e300;
  double g = 0.0;
  for (int i = 0; i < rows; ++i) g += r[i] * x[i];
  return g;
}

// Sophus::SO3d::exp(omega).matrix() (quaternion form), call site src/orcvio.cpp:4497
void so3_exp(const double* w, double* R) {
  const double th2 = (w[0] * w[0] + w[1] * w[1]) + w[2] * w[2];
  const double th = std::sqrt(th2), half = 0.5 * th;
  double imag, real;
  if (th < 1e-10) {
    const double po4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * po4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * po4;
  } else {
    imag = std::sin(half) / th;
    real = std::cos(half);
  }
  const double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx,
               tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

struct FrameResult {
  int n_pass = 0;
  double seconds = 0.0;
};

// One frame: removeLostFeatures' stack -> compress -> update chain (src/orcvio.cpp:2498-2560).
int frame_update(const double* clone_R, const double* clone_p, int N, const double* R_b2c, const double* t_c_b,
                 const double* P_in, const int* feat_off, const int* obs_clone, const double* obs_z, int n_feat,
                 int flags, double sigma2, double chi2_p, const TriCfg& cfg, double* P_out, double* delta_x,
                 int* status, double* gamma_out, double* positions, double* clone_out) {
  const int D = LEG + 6 * N;
  std::vector<double> P((size_t)D * D);
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) P[(size_t)i * D + j] = P_in[(size_t)j * D + i];   // column-major in
  // camera poses of the clones (stateAugmentation :954-961)
  std::vector<double> camR((size_t)9 * N), camT((size_t)3 * N);
  for (int c = 0; c < N; ++c) {
    mat_matT(clone_R + 9 * c, R_b2c, &camR[9 * c]);
    double t[3];
    mat_vec(clone_R + 9 * c, t_c_b, t);
    for (int k = 0; k < 3; ++k) camT[3 * c + k] = clone_p[3 * c + k] + t[k];
  }
  std::vector<double> chi(64, 0.0);
  for (int dof = 1; dof < 64; ++dof) chi[dof] = chi2_quantile(chi2_p, dof);

  struct Block { int lo, hi; std::vector<double> H, r; };   // gated block over columns [lo, hi)
  std::vector<Block> blocks;
  std::vector<double> work;
  for (int f = 0; f < n_feat; ++f) {
    const int o0 = feat_off[f], m = feat_off[f + 1] - o0;
    status[f] = 0;
    gamma_out[f] = -1.0;
    if (positions) positions[3 * f] = positions[3 * f + 1] = positions[3 * f + 2] = 0.0;
    std::vector<const double*> Rs(m), ts(m);
    for (int k = 0; k < m; ++k) { Rs[k] = &camR[9 * obs_clone[o0 + k]]; ts[k] = &camT[3 * obs_clone[o0 + k]]; }
    if (!check_motion(Rs[0], ts[0], ts[m - 1], obs_z + 2 * o0, cfg.translation_threshold)) continue;
    double pw[3];
    if (!triangulate(m, Rs.data(), ts.data(), obs_z + 2 * o0, cfg, pw, nullptr, nullptr, nullptr)) continue;
    status[f] |= 1;
    if (positions) { positions[3 * f] = pw[0]; positions[3 * f + 1] = pw[1]; positions[3 * f + 2] = pw[2]; }
    // featureJacobian_msckf :1171-1226: dense 2m x D H_xj (MatrixXd::Zero, :1191)
    const int rows = 2 * m;
    std::vector<double> Hx((size_t)rows * D, 0.0), Hf((size_t)rows * 3), r(rows);
    int lo = D, hi = 0;
    for (int k = 0; k < m; ++k) {
      const int c = obs_clone[o0 + k];
      double hx[12], hf[6], ri[2];
      measurement_jacobian(flags, clone_R + 9 * c, clone_p + 3 * c, R_b2c, t_c_b, pw, obs_z + 2 * (o0 + k), hx, hf, ri);
      for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 6; ++j) Hx[(size_t)(2 * k + i) * D + LEG + 6 * c + j] = hx[6 * i + j];
        for (int j = 0; j < 3; ++j) Hf[3 * (2 * k + i) + j] = hf[3 * i + j];
        r[2 * k + i] = ri[i];
      }
      lo = std::min(lo, LEG + 6 * c);
      hi = std::max(hi, LEG + 6 * c + 6);
    }
    if (rows <= 3) continue;
    nullspace_project(Hf, Hx, r, rows, D);
    const int pr = rows - 3;
    const double* Hp = &Hx[(size_t)3 * D];
    const double* rp = &r[3];
    const double g = gating_gamma(Hp, rp, pr, D, P.data(), sigma2, work);
    gamma_out[f] = g;
    if (g < chi[pr]) {
      status[f] |= 2;
      Block b;
      b.lo = lo; b.hi = hi;
      b.H.assign(Hp, Hp + (size_t)pr * D);
      b.r.assign(rp, rp + pr);
      blocks.push_back(std::move(b));
    }
  }
  int M = 0;
  for (auto& b : blocks) M += (int)b.r.size();
  if (M == 0) {
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) P_out[(size_t)j * D + i] = P[(size_t)i * D + j];
    std::fill(delta_x, delta_x + D, 0.0);
    return 0;
  }
  // stacked H (M x D) with the residual as column D; rows sorted by first non-zero column
  std::vector<int> order(blocks.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return blocks[a].lo < blocks[b].lo; });
  const int ld = D + 1;
  std::vector<double> H((size_t)M * ld);
  std::vector<int> rlo(M);
  {
    int row = 0;
    for (int bi : order) {
      const Block& b = blocks[bi];
      const int pr = (int)b.r.size();
      for (int i = 0; i < pr; ++i, ++row) {
        std::memcpy(&H[(size_t)row * ld], &b.H[(size_t)i * D], D * sizeof(double));
        H[(size_t)row * ld + D] = b.r[i];
        rlo[row] = b.lo;
      }
    }
  }
  int n_thin = M;
  if (M > D) {
    // SPQR-like compression (:2532-2552): Householder QR that skips structural zeros.  At column k
    // only rows whose profile starts at or before k can be non-zero there; rows are sorted by
    // profile start so they form the contiguous range [k', e_k).
    int prow = 0;                    // next pivot row
    int colhi = 0;                   // right-most non-zero column of the active rows
    std::vector<int> rhi(M);
    {
      int row = 0;
      for (int bi : order) for (size_t i = 0; i < blocks[bi].r.size(); ++i) rhi[row++] = blocks[bi].hi;
    }
    int e = 0;
    std::vector<double> v(M);
    for (int k = 0; k < D && prow < M; ++k) {
      while (e < M && rlo[e] <= k) { colhi = std::max(colhi, rhi[e]); ++e; }
      if (e <= prow) continue;       // structurally empty column: no pivot row consumed
      double sig = 0.0;
      for (int i = prow + 1; i < e; ++i) sig += H[(size_t)i * ld + k] * H[(size_t)i * ld + k];
      const double akk = H[(size_t)prow * ld + k];
      if (sig == 0.0 && akk == 0.0) continue;
      if (sig > 0.0) {
        const double mu = std::sqrt(akk * akk + sig);
        const double v0 = (akk <= 0.0) ? (akk - mu) : (-sig / (akk + mu));
        const double tau = 2.0 * v0 * v0 / (sig + v0 * v0);
        v[prow] = 1.0;
        for (int i = prow + 1; i < e; ++i) v[i] = H[(size_t)i * ld + k] / v0;
        H[(size_t)prow * ld + k] = mu;
        for (int i = prow + 1; i < e; ++i) H[(size_t)i * ld + k] = 0.0;
        const int jend = std::max(colhi, k + 1);
        // columns k+1 .. colhi-1 and the residual column
        std::vector<double> dots(jend - k - 1 + 1, 0.0);
        for (int i = prow; i < e; ++i) {
          const double vi = v[i];
          const double* row = &H[(size_t)i * ld];
          for (int j = k + 1; j < jend; ++j) dots[j - k - 1] += vi * row[j];
          dots[jend - k - 1] += vi * row[D];
        }
        for (int i = prow; i < e; ++i) {
          const double tv = tau * v[i];
          double* row = &H[(size_t)i * ld];
          for (int j = k + 1; j < jend; ++j) row[j] -= tv * dots[j - k - 1];
          row[D] -= tv * dots[jend - k - 1];
        }
      }
      ++prow;
    }
    n_thin = std::min(prow, D);
  }
  // measurementUpdate_hybrid :1811-1907 with H_o = H_thin (n_thin x D)
  std::vector<double> Ht((size_t)n_thin * D), rt(n_thin);
  for (int i = 0; i < n_thin; ++i) {
    std::memcpy(&Ht[(size_t)i * D], &H[(size_t)i * ld], D * sizeof(double));
    rt[i] = H[(size_t)i * ld + D];
  }
  std::vector<double> HP((size_t)n_thin * D);
  gemm(Ht.data(), P.data(), HP.data(), n_thin, D, D);
  std::vector<double> S((size_t)n_thin * n_thin);
  for (int i = 0; i < n_thin; ++i)
    for (int j = 0; j < n_thin; ++j) {
      double s = 0.0;
      const double* a = &HP[(size_t)i * D];
      const double* b = &Ht[(size_t)j * D];
      for (int k = 0; k < D; ++k) s += a[k] * b[k];
      S[(size_t)i * n_thin + j] = s + (i == j ? sigma2 : 0.0);
    }
  // K^T = S^-1 (H P): n_thin x D
  std::vector<double> KT = HP;
  if (!chol_solve(S, n_thin, KT.data(), D)) return -1;
  std::vector<double> dx(D, 0.0);
  for (int i = 0; i < n_thin; ++i)
    for (int j = 0; j < D; ++j) dx[j] += KT[(size_t)i * D + j] * rt[i];
  // P <- (I - K H) P = P - K (H P);  P <- (P + P^T)/2
  std::vector<double> Pn = P;
  for (int k = 0; k < n_thin; ++k) {
    const double* kt = &KT[(size_t)k * D];
    const double* hp = &HP[(size_t)k * D];
    for (int i = 0; i < D; ++i) {
      const double a = kt[i];
      if (a == 0.0) continue;
      double* row = &Pn[(size_t)i * D];
      for (int j = 0; j < D; ++j) row[j] -= a * hp[j];
    }
  }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) P_out[(size_t)j * D + i] = (Pn[(size_t)i * D + j] + Pn[(size_t)j * D + i]) / 2.0;
  std::memcpy(delta_x, dx.data(), D * sizeof(double));
  // incrementState_IMUCam :4468-4567 (clone part)
  if (clone_out) {
    const double nv = std::sqrt((dx[3] * dx[3] + dx[4] * dx[4]) + dx[5] * dx[5]);
    const double np_ = std::sqrt((dx[6] * dx[6] + dx[7] * dx[7]) + dx[8] * dx[8]);
    const bool apply = !((nv > 1.0 || np_ > 1.5) && (flags & FL_DISCARD));
    const bool left = (flags & FL_LARVIO) || (flags & FL_LEFT);
    for (int c = 0; c < N; ++c) {
      double Rn[9];
      std::memcpy(Rn, clone_R + 9 * c, sizeof(Rn));
      double pn[3] = {clone_p[3 * c], clone_p[3 * c + 1], clone_p[3 * c + 2]};
      if (apply) {
        double Rt[9];
        so3_exp(&dx[LEG + 6 * c], Rt);
        if (left) mat_mat(Rt, clone_R + 9 * c, Rn);
        else mat_mat(clone_R + 9 * c, Rt, Rn);
        for (int k = 0; k < 3; ++k) pn[k] += dx[LEG + 6 * c + 3 + k];
      }
      std::memcpy(clone_out + 12 * c, Rn, sizeof(Rn));
      for (int k = 0; k < 3; ++k) clone_out[12 * c + 9 + k] = pn[k];
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int cpu_ref_frame_update(const double* clone_R, const double* clone_p, int N, const double* R_b2c,
                         const double* t_c_b, const double* P_in, const int* feat_off, const int* obs_clone,
                         const double* obs_z, int n_feat, int flags, double sigma2, double chi2_p,
                         double translation_threshold, double cost_threshold, double init_final_dist_threshold,
                         double* P_out, double* delta_x, int* status, double* gamma, double* positions,
                         double* clone_out) {
  TriCfg cfg;
  cfg.translation_threshold = translation_threshold;
  cfg.cost_threshold = cost_threshold;
  cfg.init_final_dist_threshold = init_final_dist_threshold;
  return frame_update(clone_R, clone_p, N, R_b2c, t_c_b, P_in, feat_off, obs_clone, obs_z, n_feat, flags, sigma2,
                      chi2_p, cfg, P_out, delta_x, status, gamma, positions, clone_out);
}

// Triangulation only (bit-exact pin against oracle/feature.py and the CUDA kernel).
int cpu_ref_triangulate(const double* cam_R, const double* cam_t, const int* feat_off, const int* obs_clone,
                        const double* obs_z, int n_feat, double translation_threshold, double cost_threshold,
                        double init_final_dist_threshold, double* out_pos, int* out_status, int* out_iters,
                        double* out_cost) {
  TriCfg cfg;
  cfg.translation_threshold = translation_threshold;
  cfg.cost_threshold = cost_threshold;
  cfg.init_final_dist_threshold = init_final_dist_threshold;
  for (int f = 0; f < n_feat; ++f) {
    const int o0 = feat_off[f], m = feat_off[f + 1] - o0;
    std::vector<const double*> Rs(m), ts(m);
    for (int k = 0; k < m; ++k) { Rs[k] = cam_R + 9 * obs_clone[o0 + k]; ts[k] = cam_t + 3 * obs_clone[o0 + k]; }
    int no = 0, ni = 0;
    double cost = 0;
    const bool ok = triangulate(m, Rs.data(), ts.data(), obs_z + 2 * o0, cfg, out_pos + 3 * f, &no, &ni, &cost);
    out_status[f] = ok ? 1 : 0;
    if (out_iters) { out_iters[2 * f] = no; out_iters[2 * f + 1] = ni; }
    if (out_cost) out_cost[f] = cost;
  }
  return 0;
}

double cpu_ref_chi2_quantile(double p, int dof) { return chi2_quantile(p, dof); }

}  // extern "C"
