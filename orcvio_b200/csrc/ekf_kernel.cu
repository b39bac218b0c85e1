// Hybrid EKF-SLAM feature rows, stage level -- the per-observation Jacobians of a
// 1-D inverse-depth feature anchored in a clone of the window, and the stacked, gated rows of the features
// that are already part of the state.
//
// Reference: OrcVIO::measurementJacobian_ekf_1didp (src/orcvio.cpp:1356-1478), featureJacobian_ekf
// (:1575-1651) and gatingTestFeature with dof 2 (:1953-1976), for feature_idp_dim == 1, use_schmidt == 0,
// if_FEJ == 0, estimate_td == 0 (every shipped yaml).  The filter-level integration of these rows (state
// augmentation with a feature block, delayed initialisation, anchor change) is not built yet: these entry
// points are the element-wise parity surface for H1 / H2, like orcvio_measurement_jacobians is for J1.
//
// Parallelisation: thread per observation (H1) / thread per feature (H2: 19 structurally non-zero columns,
// H P H^T from the corresponding 19 x 19 entries of P) -- a frame holds at most a few dozen such features.
#include <cstdio>
#include <vector>

#include "../../include/orcvio_b200.h"
#include "config.h"
#include "kernels.h"

namespace ob {

// J for one (observing clone k, anchor clone a) pair.  Hf 2, Ha 2x6, Hx 2x6, He 2x6 (row-major), r 2.
__device__ __forceinline__ void ekf_jacobian_1didp(const double* clk, const double* cla, const double* Rbc,
                                                   const double* tcb, double fx, double fy, double rho,
                                                   const double* pw, double zu, double zv, bool same,
                                                   double* Hf, double* Ha, double* Hx, double* He, double* r) {
  if (same) {      // the anchor frame's own observation carries no information (:1433-1441)
    for (int i = 0; i < 2; ++i) Hf[i] = 0.0, r[i] = 0.0;
    for (int i = 0; i < 12; ++i) Ha[i] = 0.0, Hx[i] = 0.0, He[i] = 0.0;
    return;
  }
  const double* Rk = clk + CL_R;           // body -> world
  const double* tk = clk + CL_P;
  const double* Ra = cla + CL_R;
  const double* ta = cla + CL_P;
  double Rw2ck[9], Rw2ca[9];
  m3_mulT(Rbc, Rk, Rw2ck);                 // R_b2c R_bk2w^T
  m3_mulT(Rbc, Ra, Rw2ca);
  double Rt[3];
  m3_vec(Rk, tcb, Rt);
  const double tck[3] = {tk[0] + Rt[0], tk[1] + Rt[1], tk[2] + Rt[2]};
  const double d[3] = {pw[0] - tck[0], pw[1] - tck[1], pw[2] - tck[2]};
  double pck[3];
  m3_vec(Rw2ck, d, pck);
  r[0] = zu - pck[0] / pck[2];
  r[1] = zv - pck[1] / pck[2];
  const double iz = 1 / pck[2];
  const double Jk[6] = {iz, 0, -pck[0] / (pck[2] * pck[2]), 0, iz, -pck[1] / (pck[2] * pck[2])};
  const double fan[3] = {fx, fy, 1.0};
  const double pca[3] = {fx / rho, fy / rho, 1.0 / rho};
  // J_d = R_w2ck R_w2ca^T f_an
  double t1[3], Jd[3];
  m3_Tvec(Rw2ca, fan, t1);
  m3_vec(Rw2ck, t1, Jd);
  const double Jrho = -1.0 / (rho * rho);
  for (int i = 0; i < 2; ++i) Hf[i] = ((Jk[3 * i] * Jd[0] + Jk[3 * i + 1] * Jd[1]) + Jk[3 * i + 2] * Jd[2]) * Jrho;
  const double pba[3] = {pw[0] - ta[0], pw[1] - ta[1], pw[2] - ta[2]};
  const double pbk[3] = {pw[0] - tk[0], pw[1] - tk[1], pw[2] - tk[2]};
  double S[9], A[9];
  // anchor pose: [-R_w2ck [p_baf]x | R_w2ck]
  m3_skew(pba, S);
  m3_mul(Rw2ck, S, A);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) {
      Ha[6 * i + j] = -((Jk[3 * i] * A[j] + Jk[3 * i + 1] * A[3 + j]) + Jk[3 * i + 2] * A[6 + j]);
      Ha[6 * i + 3 + j] = (Jk[3 * i] * Rw2ck[j] + Jk[3 * i + 1] * Rw2ck[3 + j]) + Jk[3 * i + 2] * Rw2ck[6 + j];
    }
  // pose of clone k: [R_w2ck [p_bkf]x | -R_w2ck]
  m3_skew(pbk, S);
  m3_mul(Rw2ck, S, A);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) {
      Hx[6 * i + j] = (Jk[3 * i] * A[j] + Jk[3 * i + 1] * A[3 + j]) + Jk[3 * i + 2] * A[6 + j];
      Hx[6 * i + 3 + j] = -((Jk[3 * i] * Rw2ck[j] + Jk[3 * i + 1] * Rw2ck[3 + j]) + Jk[3 * i + 2] * Rw2ck[6 + j]);
    }
  // extrinsics: [R_b2c (Skew(R_w2bk p_bkf - t_c_b) - R_w2bk R_w2ba^T Skew(R_b2c^T p_ca)) | R_b2c (R_w2bk R_w2ba^T - I)]
  double v[3], q[3], Rka[9], Sk1[9], Sk2[9], M[9], E1[9], E2[9];
  m3_Tvec(Rk, pbk, v);
  v[0] -= tcb[0]; v[1] -= tcb[1]; v[2] -= tcb[2];
  m3_skew(v, Sk1);
  m3_Tmul(Rk, Ra, Rka);                    // R_w2bk R_w2ba^T = R_bk2w^T R_ba2w
  m3_Tvec(Rbc, pca, q);
  m3_skew(q, Sk2);
  m3_mul(Rka, Sk2, M);
  for (int i = 0; i < 9; ++i) Sk1[i] -= M[i];
  m3_mul(Rbc, Sk1, E1);
  Rka[0] -= 1.0; Rka[4] -= 1.0; Rka[8] -= 1.0;
  m3_mul(Rbc, Rka, E2);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) {
      He[6 * i + j] = (Jk[3 * i] * E1[j] + Jk[3 * i + 1] * E1[3 + j]) + Jk[3 * i + 2] * E1[6 + j];
      He[6 * i + 3 + j] = (Jk[3 * i] * E2[j] + Jk[3 * i + 1] * E2[3 + j]) + Jk[3 * i + 2] * E2[6 + j];
    }
}

struct EkfArgs {
  const double* clones; const double* Rbc; const double* tcb;
  const int* anchor; const double* rho; const double* fan; const double* pos;   // per feature
  const int* obs_feat; const int* obs_clone; const double* obs_z; int n_obs;   // per observation
  double* Hf; double* Ha; double* Hx; double* He; double* r;
};

__global__ void __launch_bounds__(128) k_ekf_jac(EkfArgs a) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= a.n_obs) return;
  const int f = a.obs_feat[o], k = a.obs_clone[o], an = a.anchor[f];
  double Hf[2], Ha[12], Hx[12], He[12], r[2];
  ekf_jacobian_1didp(a.clones + (size_t)k * CL_STRIDE, a.clones + (size_t)an * CL_STRIDE, a.Rbc, a.tcb, a.fan[2 * f],
                     a.fan[2 * f + 1], a.rho[f], a.pos + 3 * (size_t)f, a.obs_z[2 * (size_t)o], a.obs_z[2 * (size_t)o + 1],
                     k == an, Hf, Ha, Hx, He, r);
  for (int i = 0; i < 2; ++i) { a.Hf[2 * (size_t)o + i] = Hf[i]; a.r[2 * (size_t)o + i] = r[i]; }
  for (int i = 0; i < 12; ++i) {
    a.Ha[12 * (size_t)o + i] = Ha[i];
    a.Hx[12 * (size_t)o + i] = Hx[i];
    a.He[12 * (size_t)o + i] = He[i];
  }
}

// featureJacobian_ekf + gate: feature f (state column L + 6N + f) observed by the newest clone N - 1.
struct EkfRowArgs {
  const double* clones; const double* Rbc; const double* tcb; int N; int n_feat;
  const int* anchor; const double* rho; const double* fan; const double* pos; const double* z;
  const double* P; int D;                 // column-major == row-major (symmetric)
  double sigma2, chi2;
  double* H; double* r; double* gamma; int* pass;
};

__global__ void __launch_bounds__(64) k_ekf_rows(EkfRowArgs a) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.n_feat) return;
  const int k = a.N - 1, an = a.anchor[f], D = a.D;
  double Hf[2], Ha[12], Hx[12], He[12], r[2];
  ekf_jacobian_1didp(a.clones + (size_t)k * CL_STRIDE, a.clones + (size_t)an * CL_STRIDE, a.Rbc, a.tcb, a.fan[2 * f],
                     a.fan[2 * f + 1], a.rho[f], a.pos + 3 * (size_t)f, a.z[2 * f], a.z[2 * f + 1], k == an, Hf, Ha, Hx,
                     He, r);
  double* H = a.H + (size_t)2 * f * D;
  for (int i = 0; i < 2 * D; ++i) H[i] = 0.0;
  const int cf = ORCVIO_LEG + 6 * a.N + f, ca = ORCVIO_LEG + 6 * an, ck = ORCVIO_LEG + 6 * k;
  for (int i = 0; i < 2; ++i) {
    H[(size_t)i * D + cf] = Hf[i];
    for (int j = 0; j < 6; ++j) H[(size_t)i * D + ca + j] = Ha[6 * i + j];
    for (int j = 0; j < 6; ++j) H[(size_t)i * D + ck + j] = Hx[6 * i + j];     // after H_a, like :1644-1645
    for (int j = 0; j < 6; ++j) H[(size_t)i * D + 15 + j] = He[6 * i + j];
  }
  a.r[2 * f] = r[0];
  a.r[2 * f + 1] = r[1];
  // S = H P H^T + sigma^2 I over the structurally non-zero columns
  int cols[19];
  int nc = 0;
  for (int j = 0; j < 6; ++j) cols[nc++] = 15 + j;
  for (int j = 0; j < 6; ++j) cols[nc++] = ca + j;
  if (ck != ca)
    for (int j = 0; j < 6; ++j) cols[nc++] = ck + j;
  cols[nc++] = cf;
  double s00 = 0.0, s01 = 0.0, s11 = 0.0;
  for (int x = 0; x < nc; ++x) {
    double t0 = 0.0, t1 = 0.0;
    for (int y = 0; y < nc; ++y) {
      const double p = a.P[(size_t)cols[x] * D + cols[y]];
      t0 += p * H[cols[y]];
      t1 += p * H[(size_t)D + cols[y]];
    }
    s00 += H[cols[x]] * t0;
    s01 += H[cols[x]] * t1;
    s11 += H[(size_t)D + cols[x]] * t1;
  }
  s00 += a.sigma2;
  s11 += a.sigma2;
  const double det = s00 * s11 - s01 * s01;
  const double g = (r[0] * (s11 * r[0] - s01 * r[1]) + r[1] * (s00 * r[1] - s01 * r[0])) / det;
  a.gamma[f] = g;
  a.pass[f] = g < a.chi2 ? 1 : 0;
}


// updateFeatureCov_1didp (src/orcvio.cpp:3611-3773): anchor change of one inverse-depth feature.  One CTA: the
// 1 x D Jacobian J (19 structurally non-zero entries) is built by thread 0, Pfleg = J P by all threads (one
// column each), then the feature's row / column of P is replaced (J P J^T on the diagonal).  P stays exactly
// symmetric (row and column are written from the same values), so the reference's (P + P^T)/2 is the identity.
struct ReanchorArgs {
  const double* clones; const double* Rbc; const double* tcb; int N;
  int feat_idx, old_idx, new_idx;
  double pw[3]; double rho_new;
  double* P; int D;
  double* J_out;                           // optional: the Jacobian row (D)
};

__global__ void __launch_bounds__(256) k_ekf_reanchor(ReanchorArgs a) {
  extern __shared__ double sm[];           // J (D), Pfleg (D)
  double* J = sm;
  double* Pf = sm + a.D;
  const int D = a.D, tid = threadIdx.x;
  for (int i = tid; i < D; i += blockDim.x) J[i] = 0.0;
  __syncthreads();
  const int c = ORCVIO_LEG + 6 * a.N + a.feat_idx;
  if (tid == 0) {
    const double* Ro = a.clones + (size_t)a.old_idx * CL_STRIDE + CL_R;
    const double* to = a.clones + (size_t)a.old_idx * CL_STRIDE + CL_P;
    const double* Rn = a.clones + (size_t)a.new_idx * CL_STRIDE + CL_R;
    const double* tn = a.clones + (size_t)a.new_idx * CL_STRIDE + CL_P;
    const double* Rbc = a.Rbc;
    const double* tcb = a.tcb;
    double Rc2w_o[9], Rc2w_n[9], Rt[3];
    m3_mulT(Ro, Rbc, Rc2w_o);              // R_b2w R_b2c^T
    m3_mulT(Rn, Rbc, Rc2w_n);
    m3_vec(Ro, tcb, Rt);
    const double d[3] = {a.pw[0] - (to[0] + Rt[0]), a.pw[1] - (to[1] + Rt[1]), a.pw[2] - (to[2] + Rt[2])};
    double po[3];
    m3_Tvec(Rc2w_o, d, po);                // R_c2w_old^-1 (p_w - t_c_w_old): the rotation's inverse is its transpose
    const double inv_old = 1 / po[2];
    const double fo[3] = {po[0] / po[2], po[1] / po[2], 1.0};
    const double pbo[3] = {a.pw[0] - to[0], a.pw[1] - to[1], a.pw[2] - to[2]};
    const double pbn[3] = {a.pw[0] - tn[0], a.pw[1] - tn[1], a.pw[2] - tn[2]};
    const double Jrd = -a.rho_new * a.rho_new;
    double v1[3], v2[3];
    m3_vec(Rc2w_o, fo, v1);
    m3_Tvec(Rc2w_n, v1, v2);               // R_w2c_new R_c2w_old f_old
    const double Jd = v2[2];
    // bottom rows of 3x3 products with R_w2c_new = Rc2w_n^T: row 2 of R_w2c_new is column 2 of Rc2w_n
    const double w2[3] = {Rc2w_n[2], Rc2w_n[5], Rc2w_n[8]};
    double S[9];
    double Jto[3], Jtn[3];
    m3_skew(pbo, S);
    for (int j = 0; j < 3; ++j) Jto[j] = -((w2[0] * S[j] + w2[1] * S[3 + j]) + w2[2] * S[6 + j]);
    m3_skew(pbn, S);
    for (int j = 0; j < 3; ++j) Jtn[j] = (w2[0] * S[j] + w2[1] * S[3 + j]) + w2[2] * S[6 + j];
    // extrinsics
    double u[3], q[3], Rno[9], Sk1[9], Sk2[9], M[9];
    m3_Tvec(Rn, pbn, u);
    u[0] -= tcb[0]; u[1] -= tcb[1]; u[2] -= tcb[2];
    m3_skew(u, Sk1);
    m3_Tmul(Rn, Ro, Rno);                  // R_w2b_new R_b2w_old
    m3_Tvec(Rbc, po, q);
    m3_skew(q, Sk2);
    m3_mul(Rno, Sk2, M);
    for (int i = 0; i < 9; ++i) Sk1[i] -= M[i];
    Rno[0] -= 1.0; Rno[4] -= 1.0; Rno[8] -= 1.0;
    double Jet[3], Jep[3];
    for (int j = 0; j < 3; ++j) {
      Jet[j] = (Rbc[6] * Sk1[j] + Rbc[7] * Sk1[3 + j]) + Rbc[8] * Sk1[6 + j];
      Jep[j] = (Rbc[6] * Rno[j] + Rbc[7] * Rno[3 + j]) + Rbc[8] * Rno[6 + j];
    }
    const double Jdro = -1 / (inv_old * inv_old);
    J[c] = Jrd * Jd * Jdro;
    const int co = ORCVIO_LEG + 6 * a.old_idx, cn = ORCVIO_LEG + 6 * a.new_idx;
    for (int j = 0; j < 3; ++j) { J[co + j] = Jrd * Jto[j]; J[co + 3 + j] = Jrd * w2[j]; }
    for (int j = 0; j < 3; ++j) { J[cn + j] = Jrd * Jtn[j]; J[cn + 3 + j] = Jrd * (-w2[j]); }   // after the old block
    for (int j = 0; j < 3; ++j) { J[15 + j] = Jrd * Jet[j]; J[18 + j] = Jrd * Jep[j]; }
  }
  __syncthreads();
  if (a.J_out)
    for (int i = tid; i < D; i += blockDim.x) a.J_out[i] = J[i];
  // Pfleg[j] = sum_i J[i] P[i][j]: J is zero outside 15..20, the two clone blocks and the feature column
  const int co = ORCVIO_LEG + 6 * a.old_idx, cn = ORCVIO_LEG + 6 * a.new_idx;
  for (int j = tid; j < D; j += blockDim.x) {
    double s = 0.0;
    for (int i = 15; i < 21; ++i) s += J[i] * a.P[(size_t)i * D + j];
    for (int i = co; i < co + 6; ++i) s += J[i] * a.P[(size_t)i * D + j];
    if (cn != co)
      for (int i = cn; i < cn + 6; ++i) s += J[i] * a.P[(size_t)i * D + j];
    s += J[c] * a.P[(size_t)c * D + j];
    Pf[j] = s;
  }
  __syncthreads();
  __shared__ double s_pff;
  if (tid == 0) {
    double s = 0.0;
    for (int j = 0; j < D; ++j) s += Pf[j] * J[j];
    s_pff = s;
  }
  __syncthreads();
  for (int j = tid; j < D; j += blockDim.x) {
    const double v = (j == c) ? s_pff : Pf[j];
    a.P[(size_t)c * D + j] = v;
    a.P[(size_t)j * D + c] = v;
  }
}

// rmLostFeaturesCov (src/orcvio.cpp:3776-3828): drop one row / column of P (D x D -> (D-1) x (D-1)).
__global__ void k_ekf_drop_state(const double* Pin, int D, int col, double* Pout) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int Dn = D - 1;
  if (e >= Dn * Dn) return;
  const int i = e / Dn, j = e - i * Dn;
  Pout[e] = Pin[(size_t)(i + (i >= col)) * D + (j + (j >= col))];
}

namespace {
struct Dev {
  void* p = nullptr;
  ~Dev() { if (p) cudaFree(p); }
  template <class T>
  bool put(const T* h, size_t n) {
    if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return false;
    return n == 0 || cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  template <class T>
  bool make(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) == cudaSuccess; }
  template <class T>
  T* as() { return (T*)p; }
};

std::vector<double> clone_records(const double* clone_R, const double* clone_p, int N) {
  std::vector<double> cl((size_t)N * CL_STRIDE, 0.0);
  for (int c = 0; c < N; ++c) {
    for (int k = 0; k < 9; ++k) cl[(size_t)c * CL_STRIDE + CL_R + k] = clone_R[9 * (size_t)c + k];
    for (int k = 0; k < 3; ++k) cl[(size_t)c * CL_STRIDE + CL_P + k] = clone_p[3 * (size_t)c + k];
  }
  return cl;
}
}  // namespace
}  // namespace ob

extern "C" int orcvio_ekf_measurement_jacobians(const double* clone_R, const double* clone_p, int n_clones,
                                                const double* R_b2c, const double* t_c_b, const int* anchor,
                                                const double* inv_depth, const double* f_an,
                                                const double* positions, const int* feat_off, const int* obs_clone,
                                                const double* obs_z, int n_feat, double* H_f, double* H_a,
                                                double* H_x, double* H_e, double* r) {
  using namespace ob;
  if (n_clones < 1 || n_feat < 0) return ORCVIO_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return ORCVIO_ERR_NO_DEVICE;
  const int n_obs = n_feat ? feat_off[n_feat] : 0;
  std::vector<int> obs_feat(n_obs);
  for (int f = 0; f < n_feat; ++f) {
    if (anchor[f] < 0 || anchor[f] >= n_clones || !(inv_depth[f] != 0.0)) return ORCVIO_ERR_ARG;
    for (int o = feat_off[f]; o < feat_off[f + 1]; ++o) {
      if (obs_clone[o] < 0 || obs_clone[o] >= n_clones) return ORCVIO_ERR_ARG;
      obs_feat[o] = f;
    }
  }
  if (n_obs == 0) return ORCVIO_OK;
  const std::vector<double> cl = clone_records(clone_R, clone_p, n_clones);
  Dev dcl, dR, dt, dan, drho, dfan, dpos, dof, doc, doz, dHf, dHa, dHx, dHe, dr;
  bool ok = dcl.put(cl.data(), cl.size()) && dR.put(R_b2c, 9) && dt.put(t_c_b, 3) && dan.put(anchor, n_feat) &&
            drho.put(inv_depth, n_feat) && dfan.put(f_an, 2 * (size_t)n_feat) && dpos.put(positions, 3 * (size_t)n_feat) &&
            dof.put(obs_feat.data(), n_obs) && doc.put(obs_clone, n_obs) && doz.put(obs_z, 2 * (size_t)n_obs) &&
            dHf.make<double>(2 * (size_t)n_obs) && dHa.make<double>(12 * (size_t)n_obs) &&
            dHx.make<double>(12 * (size_t)n_obs) && dHe.make<double>(12 * (size_t)n_obs) && dr.make<double>(2 * (size_t)n_obs);
  if (!ok) return ORCVIO_ERR_CUDA;
  EkfArgs a{dcl.as<double>(), dR.as<double>(), dt.as<double>(), dan.as<int>(), drho.as<double>(), dfan.as<double>(),
            dpos.as<double>(), dof.as<int>(), doc.as<int>(), doz.as<double>(), n_obs, dHf.as<double>(), dHa.as<double>(),
            dHx.as<double>(), dHe.as<double>(), dr.as<double>()};
  k_ekf_jac<<<(n_obs + 127) / 128, 128>>>(a);
  check_launch("k_ekf_jac");
  ok = cudaMemcpy(H_f, dHf.p, 2 * (size_t)n_obs * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
       cudaMemcpy(H_a, dHa.p, 12 * (size_t)n_obs * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
       cudaMemcpy(H_x, dHx.p, 12 * (size_t)n_obs * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
       cudaMemcpy(H_e, dHe.p, 12 * (size_t)n_obs * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
       cudaMemcpy(r, dr.p, 2 * (size_t)n_obs * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  return ok ? ORCVIO_OK : ORCVIO_ERR_CUDA;
}

extern "C" int orcvio_ekf_feature_rows(const double* clone_R, const double* clone_p, int n_clones, const double* R_b2c,
                                       const double* t_c_b, const int* anchor, const double* inv_depth,
                                       const double* f_an, const double* positions, const double* z_cur, int n_feat,
                                       const double* P, int D, double noise_var, double chi2_p, double* H, double* r,
                                       double* gamma, int* pass) {
  using namespace ob;
  if (n_clones < 1 || n_feat < 0 || D != ORCVIO_LEG + 6 * n_clones + n_feat) return ORCVIO_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return ORCVIO_ERR_NO_DEVICE;
  for (int f = 0; f < n_feat; ++f)
    if (anchor[f] < 0 || anchor[f] >= n_clones || !(inv_depth[f] != 0.0)) return ORCVIO_ERR_ARG;
  if (n_feat == 0) return ORCVIO_OK;
  const std::vector<double> cl = clone_records(clone_R, clone_p, n_clones);
  Dev dcl, dR, dt, dan, drho, dfan, dpos, dz, dP, dH, dr, dg, dp;
  bool ok = dcl.put(cl.data(), cl.size()) && dR.put(R_b2c, 9) && dt.put(t_c_b, 3) && dan.put(anchor, n_feat) &&
            drho.put(inv_depth, n_feat) && dfan.put(f_an, 2 * (size_t)n_feat) && dpos.put(positions, 3 * (size_t)n_feat) &&
            dz.put(z_cur, 2 * (size_t)n_feat) && dP.put(P, (size_t)D * D) && dH.make<double>((size_t)2 * n_feat * D) &&
            dr.make<double>(2 * (size_t)n_feat) && dg.make<double>(n_feat) && dp.make<int>(n_feat);
  if (!ok) return ORCVIO_ERR_CUDA;
  EkfRowArgs a{dcl.as<double>(), dR.as<double>(), dt.as<double>(), n_clones, n_feat, dan.as<int>(), drho.as<double>(),
               dfan.as<double>(), dpos.as<double>(), dz.as<double>(), dP.as<double>(), D, noise_var,
               chi2_quantile(chi2_p, 2), dH.as<double>(), dr.as<double>(), dg.as<double>(), dp.as<int>()};
  k_ekf_rows<<<(n_feat + 63) / 64, 64>>>(a);
  check_launch("k_ekf_rows");
  ok = cudaMemcpy(H, dH.p, (size_t)2 * n_feat * D * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
       cudaMemcpy(r, dr.p, 2 * (size_t)n_feat * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
       cudaMemcpy(gamma, dg.p, (size_t)n_feat * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
       cudaMemcpy(pass, dp.p, (size_t)n_feat * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
  return ok ? ORCVIO_OK : ORCVIO_ERR_CUDA;
}

extern "C" int orcvio_ekf_update_feature_cov(double* P, int D, const double* clone_R, const double* clone_p,
                                             int n_clones, const double* R_b2c, const double* t_c_b, int feat_idx,
                                             int old_idx, int new_idx, const double* p_w, double inv_depth_new,
                                             double* J_out) {
  using namespace ob;
  const int E = D - ORCVIO_LEG - 6 * n_clones;
  if (n_clones < 1 || E < 1 || feat_idx < 0 || feat_idx >= E || old_idx < 0 || old_idx >= n_clones || new_idx < 0 ||
      new_idx >= n_clones)
    return ORCVIO_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return ORCVIO_ERR_NO_DEVICE;
  const std::vector<double> cl = clone_records(clone_R, clone_p, n_clones);
  Dev dcl, dR, dt, dP, dJ;
  bool ok = dcl.put(cl.data(), cl.size()) && dR.put(R_b2c, 9) && dt.put(t_c_b, 3) && dP.put(P, (size_t)D * D) &&
            dJ.make<double>(D);
  if (!ok) return ORCVIO_ERR_CUDA;
  ReanchorArgs a{};
  a.clones = dcl.as<double>(); a.Rbc = dR.as<double>(); a.tcb = dt.as<double>(); a.N = n_clones;
  a.feat_idx = feat_idx; a.old_idx = old_idx; a.new_idx = new_idx;
  for (int k = 0; k < 3; ++k) a.pw[k] = p_w[k];
  a.rho_new = inv_depth_new;
  a.P = dP.as<double>(); a.D = D; a.J_out = dJ.as<double>();
  k_ekf_reanchor<<<1, 256, 2 * (size_t)D * sizeof(double)>>>(a);
  check_launch("k_ekf_reanchor");
  ok = cudaMemcpy(P, dP.p, (size_t)D * D * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
       (!J_out || cudaMemcpy(J_out, dJ.p, (size_t)D * 8, cudaMemcpyDeviceToHost) == cudaSuccess);
  return ok ? ORCVIO_OK : ORCVIO_ERR_CUDA;
}

extern "C" int orcvio_ekf_remove_feature_cov(const double* P, int D, int n_clones, int feat_idx, double* P_out) {
  using namespace ob;
  const int E = D - ORCVIO_LEG - 6 * n_clones;
  if (n_clones < 1 || E < 1 || feat_idx < 0 || feat_idx >= E) return ORCVIO_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return ORCVIO_ERR_NO_DEVICE;
  Dev dP, dO;
  const int Dn = D - 1;
  if (!dP.put(P, (size_t)D * D) || !dO.make<double>((size_t)Dn * Dn)) return ORCVIO_ERR_CUDA;
  k_ekf_drop_state<<<(Dn * Dn + 255) / 256, 256>>>(dP.as<double>(), D, ORCVIO_LEG + 6 * n_clones + feat_idx, dO.as<double>());
  check_launch("k_ekf_drop_state");
  return cudaMemcpy(P_out, dO.p, (size_t)Dn * Dn * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? ORCVIO_OK : ORCVIO_ERR_CUDA;
}
