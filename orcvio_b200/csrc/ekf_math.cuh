// Device arithmetic of the hybrid EKF-SLAM feature rows (1-D inverse depth), shared by the stage-level entry points
// (ekf_kernel.cu) and the in-filter kernels (hybrid_kernel.cu).
// Reference: OrcVIO::measurementJacobian_ekf_1didp (src/orcvio.cpp:1356-1478), updateFeatureCov_1didp (:3611-3699).
#pragma once
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

// updateFeatureCov_1didp (:3611-3699): d(rho_new) with respect to (rho_old, old anchor pose, new anchor pose,
// extrinsics) when a feature moves its anchor from clone `old` to clone `new`; pw = world position, rho_new = the
// inverse depth already re-expressed in the new anchor.  The new-anchor block is written after the old one (:3714-3717):
// callers pass J_old and J_new pointing into the same row, so old == new resolves like the reference.
__device__ __forceinline__ void ekf_reanchor_jacobian(const double* clo, const double* cln, const double* Rbc,
                                                      const double* tcb, const double* pw, double rho_new, double* J_f,
                                                      double* J_old, double* J_new, double* J_e) {
  const double* Ro = clo + CL_R;
  const double* to = clo + CL_P;
  const double* Rn = cln + CL_R;
  const double* tn = cln + CL_P;
  double Rc2w_o[9], Rc2w_n[9], Rt[3];
  m3_mulT(Ro, Rbc, Rc2w_o);              // R_b2w R_b2c^T
  m3_mulT(Rn, Rbc, Rc2w_n);
  m3_vec(Ro, tcb, Rt);
  const double d[3] = {pw[0] - (to[0] + Rt[0]), pw[1] - (to[1] + Rt[1]), pw[2] - (to[2] + Rt[2])};
  double po[3];
  m3_inv_vec(Rc2w_o, d, po);             // R_c2w_old.inverse() * (p_w - t_c_w_old), :3638
  const double inv_old = 1 / po[2];
  const double fo[3] = {po[0] / po[2], po[1] / po[2], 1.0};
  const double pbo[3] = {pw[0] - to[0], pw[1] - to[1], pw[2] - to[2]};
  const double pbn[3] = {pw[0] - tn[0], pw[1] - tn[1], pw[2] - tn[2]};
  const double Jrd = -rho_new * rho_new;
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
  *J_f = Jrd * Jd * Jdro;
  for (int j = 0; j < 3; ++j) { J_old[j] = Jrd * Jto[j]; J_old[3 + j] = Jrd * w2[j]; }
  for (int j = 0; j < 3; ++j) { J_new[j] = Jrd * Jtn[j]; J_new[3 + j] = Jrd * (-w2[j]); }   // after the old block
  if (J_e)
    for (int j = 0; j < 3; ++j) { J_e[j] = Jrd * Jet[j]; J_e[3 + j] = Jrd * Jep[j]; }
}

}  // namespace ob
