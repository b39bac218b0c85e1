// Hybrid MSCKF / EKF-SLAM mode inside the filter (max_features_in_one_grid > 0: config/euroc.yaml,
// config/kitti_odom.yaml) -- the device side of the EKF-SLAM branches of the reference's
//   removeLostFeatures        src/orcvio.cpp:2196-2579 (rows of the features of the state :2449-2495, new features
//                             :2343-2446 incl. the sparsification :2413-2443)
//   measurementUpdate_hybrid  :1766-1950 (feature increments :1843-1882, delayed initialisation :1823-1832, :1903-1941)
//   measurementUpdate_msckf   :1700-1737 (feature increments after the prune-phase update)
//   pruneImuStateBuffer       :2665-2773 (anchor change) with updateFeatureCov_1didp :3611-3773
//   rmLostFeaturesCov         :3776-3828
// for feature_idp_dim == 1, use_schmidt == 0 (every shipped yaml).  The arithmetic is the one of ekf_math.cuh, whose
// stage-level entry points (ekf_kernel.cu) are parity-tested element by element against oracle/hybrid.py.
//
// How the rows join the whitened-form update (info_kernel.cu): a frame has at most 30 features in the state, so their
// rows are kept DENSE -- [window columns 22 .. D-1 | residual], one CTA per filter builds them (k_hybrid_rows):
//   * feature of the state, observed by the newest clone: 2 rows (anchor block, newest block, its own column), gated
//     with dof 2 against the block-sparse H P H^T; a rejected feature leaves two zero rows (exact no-ops);
//   * new feature: its MSCKF rows only decide the gate (k_jac_gate, CAND_GATE_ONLY); the 2 (m - 1) rows of
//     featureJacobian_ekf_new have ONE feature column, so W = [V U] of the reference is block diagonal and the
//     sparsification is one Householder reflection per feature: the reflected row 0 is the initialisation row
//     (H_1 | h_2, r_1), the other 2 (m - 1) - 1 rows have lost their feature part and join the dense rows;
// k_hybrid_aform multiplies the dense rows by the prior factor (A = H L, appended behind the MSCKF tiles' rows), the
// update runs with n = 6 N + E window columns, and k_hybrid_post applies the feature increments and appends the
// new features to P (delayed initialisation) in place.
#include "kernels.h"
#include "ekf_math.cuh"

namespace ob {

namespace {

__device__ __forceinline__ void feature_world_position(const double* cl_anchor, double rho, double ox, double oy,
                                                       double* pw) {
  // p_w = R_c2w p_c + t_c_w with p_c = (obs_anchor / rho, 1 / rho)   (:1866-1877)
  const double pc[3] = {ox / rho, oy / rho, 1.0 / rho};
  double t[3];
  m3_vec(cl_anchor + CL_RC, pc, t);
  for (int i = 0; i < 3; ++i) pw[i] = t[i] + cl_anchor[CL_PC + i];
}

}  // namespace

// ---------------------------------------------------------------- dense rows of the EKF-SLAM features
__global__ void __launch_bounds__(128) k_hybrid_rows(HybArgs a) {
  __shared__ double hv[2 * ORCVIO_MAX_OBS], rv[2 * ORCVIO_MAX_OBS];
  __shared__ int orow[ORCVIO_MAX_OBS];
  __shared__ double s_tau, s_beta;
  __shared__ int s_k, s_cnt;
  const int fi = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const HybWork hw = a.hw[fi];
  if (tid == 0) a.n_new[fi] = 0;
  if (!hw.active) return;
  const int N = hw.N, E = hw.E, n = 6 * N + E, ldh = a.ldh;
  const double* clones = a.clones + (size_t)fi * a.clone_stride;
  const double* imu = a.imu + (size_t)fi * IM_STRIDE;
  const double* Rbc = imu + IM_RBC;
  const double* tcb = imu + IM_TCB;
  const double* P = a.P + (size_t)fi * a.p_stride;
  const int ldp = a.ldp;
  double* Hd = a.Hd + (size_t)hw.dense_off * ldh;
  for (int e = tid; e < hw.n_dense * ldh; e += nt) Hd[e] = 0.0;
  __syncthreads();
  // ---- features of the state: featureJacobian_ekf + gatingTestFeature(dof 2)   (:1575-1651, :2451-2461)
  for (int i = tid; i < E; i += nt) {
    const HybFeat ft = a.feats[hw.feat_begin + i];
    const int k = N - 1, an = ft.anchor;
    const double* fid = a.fidp + ((size_t)fi * a.fcap + ft.slot) * FI_STRIDE;
    const double* pw = a.fpos + ((size_t)fi * a.fcap + ft.slot) * FP_STRIDE;
    double Hf[2], Ha[12], Hx[12], He[12], r[2];
    ekf_jacobian_1didp(clones + (size_t)k * CL_STRIDE, clones + (size_t)an * CL_STRIDE, Rbc, tcb, fid[1], fid[2], fid[0],
                       pw, ft.zu, ft.zv, k == an, Hf, Ha, Hx, He, r);
    // the 2 x D block over its structurally non-zero columns; H_x is written after H_a like :1644-1645
    int cols[13];
    double h0[13], h1[13];
    int nc = 0;
    const int ca = ORCVIO_LEG + 6 * an, ck = ORCVIO_LEG + 6 * k, cf = ORCVIO_LEG + 6 * N + i;
    if (ck != ca)
      for (int j = 0; j < 6; ++j) { cols[nc] = ca + j; h0[nc] = Ha[j]; h1[nc] = Ha[6 + j]; ++nc; }
    for (int j = 0; j < 6; ++j) { cols[nc] = ck + j; h0[nc] = Hx[j]; h1[nc] = Hx[6 + j]; ++nc; }
    cols[nc] = cf; h0[nc] = Hf[0]; h1[nc] = Hf[1]; ++nc;
    // (the extrinsic columns 15..20 of H meet exactly zero rows of P: estimate_extrin = 0)
    double s00 = 0.0, s01 = 0.0, s11 = 0.0;
    for (int x = 0; x < nc; ++x) {
      double t0 = 0.0, t1 = 0.0;
      for (int y = 0; y < nc; ++y) {
        const double p = P[(size_t)cols[x] * ldp + cols[y]];
        t0 += p * h0[y];
        t1 += p * h1[y];
      }
      s00 += h0[x] * t0;
      s01 += h0[x] * t1;
      s11 += h1[x] * t1;
    }
    s00 += a.sigma2;
    s11 += a.sigma2;
    const double det = s00 * s11 - s01 * s01;
    const double g = (r[0] * (s11 * r[0] - s01 * r[1]) + r[1] * (s00 * r[1] - s01 * r[0])) / det;
    const int pass = g < a.chi2_dof2 ? 1 : 0;
    a.ekf_gamma[hw.feat_begin + i] = g;
    a.ekf_pass[hw.feat_begin + i] = pass;
    if (pass) {
      double* r0 = Hd + (size_t)(2 * i) * ldh;
      double* r1 = r0 + ldh;
      for (int x = 0; x < nc; ++x) {
        r0[cols[x] - ORCVIO_LEG] = h0[x];
        r1[cols[x] - ORCVIO_LEG] = h1[x];
      }
      r0[n] = r[0];
      r1[n] = r[1];
    }
  }
  // ---- new features, one after the other (the survivors take consecutive columns behind the old state)
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  double* M = a.scratch + (size_t)fi * 64 * ldh;
  for (int j = hw.new_begin; j < hw.new_end; ++j) {
    const HybNew nw = a.news[j];
    const bool ok = (a.status[nw.cand] & ST_GATE_PASS) != 0;
    if (tid == 0) a.new_ok[j] = ok ? 1 : 0;
    if (!ok) continue;                                   // uniform: rejected by the MSCKF gate (:2371-2411)
    const int an = nw.anchor;
    if (tid == 0) {
      int k = 0;
      for (int i = 0; i < nw.obs_m; ++i)
        if (a.obs_clone[nw.obs_off + i] != an) orow[k++] = nw.obs_off + i;   // the anchor frame's own observation is not used
      s_k = k;
    }
    __syncthreads();
    const int k = s_k, rows = 2 * k;
    for (int e = tid; e < rows * ldh; e += nt) M[e] = 0.0;
    __syncthreads();
    if (tid < k) {
      const int o = orow[tid], c = a.obs_clone[o];
      const double* fid = a.fidp + ((size_t)fi * a.fcap + nw.slot) * FI_STRIDE;
      const double* pw = a.fpos + ((size_t)fi * a.fcap + nw.slot) * FP_STRIDE;
      double Hf[2], Ha[12], Hx[12], He[12], r[2];
      ekf_jacobian_1didp(clones + (size_t)c * CL_STRIDE, clones + (size_t)an * CL_STRIDE, Rbc, tcb, fid[1], fid[2], fid[0],
                         pw, a.obs_z[2 * (size_t)o], a.obs_z[2 * (size_t)o + 1], false, Hf, Ha, Hx, He, r);
      for (int i = 0; i < 2; ++i) {
        double* row = M + (size_t)(2 * tid + i) * ldh;
        for (int q = 0; q < 6; ++q) row[6 * an + q] = Ha[6 * i + q];
        for (int q = 0; q < 6; ++q) row[6 * c + q] = Hx[6 * i + q];
        hv[2 * tid + i] = Hf[i];
        rv[2 * tid + i] = r[i];
      }
    }
    __syncthreads();
    if (tid == 0) {            // Householder vector of h (LAPACK dlarfg convention): (I - tau v v^T) h = beta e_0, v_0 = 1
      double sig = 0.0;
      for (int i = 1; i < rows; ++i) sig += hv[i] * hv[i];
      const double alpha = hv[0];
      double beta = alpha, tau = 0.0;
      if (sig > 0.0) {
        const double nrm = sqrt(alpha * alpha + sig);
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        const double sc = 1.0 / (alpha - beta);
        for (int i = 1; i < rows; ++i) hv[i] *= sc;
      }
      hv[0] = 1.0;
      s_tau = tau;
      s_beta = beta;
    }
    __syncthreads();
    const double tau = s_tau;
    const int sidx = s_cnt;                              // survivor index = its column behind the old state
    double* H1 = a.H1 + ((size_t)fi * a.new_cap + sidx) * ldh;
    for (int c = tid; c <= 6 * N; c += nt) {             // column 6 N stands for the residual
      const bool is_r = (c == 6 * N);
      double dot = 0.0;
      for (int i = 0; i < rows; ++i) dot += hv[i] * (is_r ? rv[i] : M[(size_t)i * ldh + c]);
      dot *= tau;
      for (int i = 0; i < rows; ++i) {
        const double v = (is_r ? rv[i] : M[(size_t)i * ldh + c]) - dot * hv[i];
        if (i == 0) {
          if (is_r) a.r1[(size_t)fi * a.new_cap + sidx] = v;
          else H1[c] = v;
        } else {
          double* dst = Hd + (size_t)(2 * E + nw.row_off + i - 1) * ldh;
          dst[is_r ? n : c] = v;
        }
      }
    }
    for (int c = 6 * N + tid; c < ldh; c += nt) H1[c] = 0.0;   // no dependence on the features of the state
    if (tid == 0) {
      a.h2[(size_t)fi * a.new_cap + sidx] = s_beta;
      s_cnt = sidx + 1;
    }
    __syncthreads();
  }
  if (tid == 0) a.n_new[fi] = s_cnt;
}

void launch_hybrid_rows(const HybArgs& a, cudaStream_t s) {
  k_hybrid_rows<<<a.n_filters, 128, 0, s>>>(a);
  check_launch("k_hybrid_rows");
}

// ---------------------------------------------------------------- A = [r | H L] for the dense rows
// grid (dense rows, filters); L[k][j] = FT[j][22 + k] (zero for k < j).  Rows go behind the MSCKF tiles' rows of
// the filter; the columns right of n are zeroed as far as k_syrk reads them.
__global__ void __launch_bounds__(128) k_hybrid_aform(HybArgs a, const double* FT_all, size_t t_stride, int ldt,
                                                      double* Amat, int lda, const FilterWork* fw_all) {
  const int fi = blockIdx.y, row = blockIdx.x;
  const HybWork hw = a.hw[fi];
  if (!hw.active || row >= hw.n_dense) return;
  const FilterWork fw = fw_all[fi];
  const int n = 6 * hw.N + hw.E;
  const double* h = a.Hd + (size_t)(hw.dense_off + row) * a.ldh;
  const double* FT = FT_all + (size_t)fi * t_stride;
  double* out = Amat + (size_t)(fw.arow0 + hw.arow_dense + row) * lda;
  for (int j = threadIdx.x; j < lda; j += blockDim.x) {
    if (j == 0) { out[0] = h[n]; continue; }
    const int jl = j - 1;
    if (jl >= n) { out[j] = 0.0; continue; }
    const double* src = FT + (size_t)jl * ldt + ORCVIO_LEG;
    double s0 = 0.0, s1 = 0.0;
    int k = jl;
    for (; k + 1 < n; k += 2) { s0 += h[k] * src[k]; s1 += h[k + 1] * src[k + 1]; }
    if (k < n) s0 += h[k] * src[k];
    out[j] = s0 + s1;
  }
}

void launch_hybrid_aform(const HybArgs& a, const double* FT, size_t t_stride, int ldt, double* Amat, int lda,
                         const FilterWork* fw, int max_dense, cudaStream_t s) {
  if (max_dense <= 0) return;
  dim3 g(max_dense, a.n_filters);
  k_hybrid_aform<<<g, 128, 0, s>>>(a, FT, t_stride, ldt, Amat, lda, fw);
  check_launch("k_hybrid_aform");
}

// ---------------------------------------------------------------- feature increments, delayed initialisation
// One CTA per filter, after the legacy update (P+, dx in place).  WITH_INIT: the survivors of this frame's new
// candidates are appended to P:  HH = H_1 / h_2,  P[new, :D] = -HH P,  P22 = HH P HH^T + s^2 / h_2^2 (symmetrised),
// dx_new = -HH dx + r_1 / h_2   (src/orcvio.cpp:1823-1832, 1903-1941).
template <bool WITH_INIT>
__global__ void __launch_bounds__(256) k_hybrid_post(HybArgs a) {
  extern __shared__ double sm[];                         // hh (ldh) per survivor processed at a time + nHHP row (ldp)
  const int fi = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const HybWork hw = a.hw[fi];
  if (!hw.active) return;
  const int N = hw.N, E = hw.E, D = ORCVIO_LEG + 6 * N + E, n = 6 * N + E, ldp = a.ldp;
  const double* clones = a.clones + (size_t)fi * a.clone_stride;
  const double* dx = a.dx + (size_t)fi * a.lddx;
  double* P = a.P + (size_t)fi * a.p_stride;
  // ---- features of the state: invDepth += its component of dx, world position from the updated anchor (:1843-1882)
  for (int i = tid; i < E; i += nt) {
    const HybFeat ft = a.feats[hw.feat_begin + i];
    double* fid = a.fidp + ((size_t)fi * a.fcap + ft.slot) * FI_STRIDE;
    double* pw = a.fpos + ((size_t)fi * a.fcap + ft.slot) * FP_STRIDE;
    fid[0] += dx[ORCVIO_LEG + 6 * N + i];
    feature_world_position(clones + (size_t)ft.anchor * CL_STRIDE, fid[0], fid[1], fid[2], pw);
  }
  if (!WITH_INIT) return;
  const int F = a.n_new[fi];
  if (F == 0) return;
  double* hh = sm;                                       // ldh
  double* nrow = sm + a.ldh;                             // ldp: the new row -HH_s P
  __shared__ double s_dxn;
  // survivors in order; the HybNew record of survivor s is the s-th accepted candidate
  int jn = hw.new_begin;
  for (int s = 0; s < F; ++s) {
    while (!a.new_ok[jn]) ++jn;                          // uniform
    const HybNew nw = a.news[jn];
    ++jn;
    const double* H1 = a.H1 + ((size_t)fi * a.new_cap + s) * a.ldh;
    const double ih = 1.0 / a.h2[(size_t)fi * a.new_cap + s];
    for (int c = tid; c < n; c += nt) hh[c] = H1[c] * ih;
    __syncthreads();
    // -HH_s P over the old state (H_1 has no IMU columns: rows 22.. of P only)
    for (int c = tid; c < D; c += nt) {
      double s0 = 0.0, s1 = 0.0;
      int k = 0;
      for (; k + 1 < n; k += 2) {
        s0 += hh[k] * P[(size_t)(ORCVIO_LEG + k) * ldp + c];
        s1 += hh[k + 1] * P[(size_t)(ORCVIO_LEG + k + 1) * ldp + c];
      }
      if (k < n) s0 += hh[k] * P[(size_t)(ORCVIO_LEG + k) * ldp + c];
      nrow[c] = -(s0 + s1);
    }
    if (tid == 0) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += hh[k] * dx[ORCVIO_LEG + k];
      s_dxn = -acc + a.r1[(size_t)fi * a.new_cap + s] * ih;
    }
    __syncthreads();
    const int rnew = D + s;
    for (int c = tid; c < D; c += nt) {
      P[(size_t)rnew * ldp + c] = nrow[c];
      P[(size_t)c * ldp + rnew] = nrow[c];
    }
    // P22[s][t], t <= s:  -nHHP_s . HH_t  (+ sigma^2 / h_2^2 on the diagonal), mean of the two orders like the
    // reference's symmetrisation
    for (int t = tid; t <= s; t += nt) {
      const double* H1t = a.H1 + ((size_t)fi * a.new_cap + t) * a.ldh;
      const double iht = 1.0 / a.h2[(size_t)fi * a.new_cap + t];
      double v0 = 0.0, v1 = 0.0;
      for (int k = 0; k < n; ++k) {
        v0 -= nrow[ORCVIO_LEG + k] * (H1t[k] * iht);                                  // -(nHHP_s) HH_t^T
        v1 -= P[(size_t)(D + t) * ldp + ORCVIO_LEG + k] * hh[k];                       // -(nHHP_t) HH_s^T
      }
      double v = (t == s) ? v0 : 0.5 * (v0 + v1);
      if (t == s) v += a.sigma2 * ih * ih;
      P[(size_t)rnew * ldp + D + t] = v;
      P[(size_t)(D + t) * ldp + rnew] = v;
    }
    if (tid == 0) {
      double* fid = a.fidp + ((size_t)fi * a.fcap + nw.slot) * FI_STRIDE;
      double* pw = a.fpos + ((size_t)fi * a.fcap + nw.slot) * FP_STRIDE;
      fid[0] += s_dxn;
      feature_world_position(clones + (size_t)nw.anchor * CL_STRIDE, fid[0], fid[1], fid[2], pw);
    }
    __syncthreads();
  }
}

void launch_hybrid_post(const HybArgs& a, cudaStream_t s) {
  const size_t smem = ((size_t)a.ldh + a.ldp) * sizeof(double);
  k_hybrid_post<true><<<a.n_filters, 256, smem, s>>>(a);
  check_launch("k_hybrid_post");
}

void launch_hybrid_feature_increment(const HybArgs& a, cudaStream_t s) {
  k_hybrid_post<false><<<a.n_filters, 256, 0, s>>>(a);
  check_launch("k_hybrid_feature_increment");
}

// ---------------------------------------------------------------- initializeInvParamPosition: the idp record
// feature.hpp:536-547: invDepth = 1 / final_position.z, obs_anchor = final_position.xy * invDepth
__global__ void k_hybrid_commit(const CommitRec* recs, int n, const double* final_pos, const double* spec_pos,
                                double* fpos, long long* fgen, double* fidp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CommitRec rc = recs[i];
  const double* fp = final_pos + 3 * (size_t)rc.cand;
  double* fid = fidp + (size_t)rc.slot * FI_STRIDE;     // slot already includes the filter's offset
  const double rho = 1 / fp[2];
  fid[0] = rho;
  fid[1] = fp[0] * rho;
  fid[2] = fp[1] * rho;
  fid[3] = 0.0;
  for (int k = 0; k < 3; ++k) fpos[(size_t)rc.slot * FP_STRIDE + k] = spec_pos[(size_t)rc.slot * FP_STRIDE + k];
  fgen[rc.slot] = rc.gen;
}

void launch_hybrid_commit(const CommitRec* recs, int n, const double* final_pos, const double* spec_pos, double* fpos,
                          long long* fgen, double* fidp, cudaStream_t s) {
  if (n <= 0) return;
  k_hybrid_commit<<<(n + 127) / 128, 128, 0, s>>>(recs, n, final_pos, spec_pos, fpos, fgen, fidp);
  check_launch("k_hybrid_commit");
}

// ---------------------------------------------------------------- P <- P[keep, keep]   (rmLostFeaturesCov :3776-3828)
__global__ void __launch_bounds__(256) k_compact_cov(double* P_all, size_t p_stride, int ldp, const int* newidx_all,
                                                     const int* D_old) {
  extern __shared__ double rowbuf[];
  const int fi = blockIdx.x;
  const int D = D_old[fi];
  if (D <= 0) return;
  double* P = P_all + (size_t)fi * p_stride;
  const int* newidx = newidx_all + (size_t)fi * ldp;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = 0; i < D; ++i) {
    const int ni = newidx[i];
    if (ni < 0) continue;                                // uniform across the CTA
    for (int j = tid; j < D; j += nt) rowbuf[j] = P[(size_t)i * ldp + j];
    __syncthreads();
    for (int j = tid; j < D; j += nt)
      if (newidx[j] >= 0) P[(size_t)ni * ldp + newidx[j]] = rowbuf[j];   // ni <= i, newidx[j] <= j: in place is safe
    __syncthreads();
  }
}

void launch_compact_cov(double* P, size_t p_stride, int ldp, const int* newidx, const int* D_old, int n_filters,
                        cudaStream_t s) {
  k_compact_cov<<<n_filters, 256, (size_t)ldp * sizeof(double), s>>>(P, p_stride, ldp, newidx, D_old);
  check_launch("k_compact_cov");
}

// ---------------------------------------------------------------- anchor change   (:2665-2773, :3611-3773)
// One CTA per filter, its records one after the other (every change sees the covariance left by the previous one,
// like the reference's loop over the map server).
__global__ void __launch_bounds__(256) k_reanchor(const ReanchorRec* recs, const int* rec_off, const HybWork* hw_all,
                                                  const double* clones_all, size_t clone_stride, const double* imu_all,
                                                  const double* fpos, double* fidp, int fcap, double* P_all,
                                                  size_t p_stride, int ldp) {
  extern __shared__ double sm[];                         // J (ldp), Pfleg (ldp)
  const int fi = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const HybWork hw = hw_all[fi];
  const int D = ORCVIO_LEG + 6 * hw.N + hw.E;
  double* J = sm;
  double* Pf = sm + ldp;
  __shared__ double s_pff;
  const double* clones = clones_all + (size_t)fi * clone_stride;
  const double* imu = imu_all + (size_t)fi * IM_STRIDE;
  double* P = P_all + (size_t)fi * p_stride;
  for (int q = rec_off[fi]; q < rec_off[fi + 1]; ++q) {
    const ReanchorRec rc = recs[q];
    const double* pw = fpos + ((size_t)fi * fcap + rc.slot) * FP_STRIDE;
    double* fid = fidp + ((size_t)fi * fcap + rc.slot) * FI_STRIDE;
    const double* cln = clones + (size_t)rc.new_idx * CL_STRIDE;
    for (int i = tid; i < D; i += nt) J[i] = 0.0;
    __syncthreads();
    if (tid == 0) {
      // new inverse depth (and, for a feature of the state, the corrected anchor observation): p_new = R_c2w_new^-1 (p_w - t)
      const double d[3] = {pw[0] - cln[CL_PC], pw[1] - cln[CL_PC + 1], pw[2] - cln[CL_PC + 2]};
      double pn[3];
      m3_inv_vec(cln + CL_RC, d, pn);                    // R_c2w_new.inverse() (:2707)
      const double rho = 1 / pn[2];
      fid[0] = rho;
      if (rc.col >= 0) {
        fid[1] = pn[0] / pn[2];
        fid[2] = pn[1] / pn[2];
        const int c = ORCVIO_LEG + 6 * hw.N + rc.col;
        ekf_reanchor_jacobian(clones + (size_t)rc.old_idx * CL_STRIDE, cln, imu + IM_RBC, imu + IM_TCB, pw, rho, J + c,
                              J + ORCVIO_LEG + 6 * rc.old_idx, J + ORCVIO_LEG + 6 * rc.new_idx, nullptr);
      } else {
        fid[1] = rc.zu;                                  // :2762-2764: the stored observation, not corrected
        fid[2] = rc.zv;
      }
    }
    __syncthreads();
    if (rc.col < 0) continue;                            // uniform
    const int c = ORCVIO_LEG + 6 * hw.N + rc.col;
    const int co = ORCVIO_LEG + 6 * rc.old_idx, cn = ORCVIO_LEG + 6 * rc.new_idx;
    // Pfleg = J P: J is zero outside the two clone blocks and the feature column (its extrinsic part meets zero rows)
    for (int j = tid; j < D; j += nt) {
      double s = 0.0;
      for (int i = co; i < co + 6; ++i) s += J[i] * P[(size_t)i * ldp + j];
      if (cn != co)
        for (int i = cn; i < cn + 6; ++i) s += J[i] * P[(size_t)i * ldp + j];
      s += J[c] * P[(size_t)c * ldp + j];
      Pf[j] = s;
    }
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int j = 0; j < D; ++j) s += Pf[j] * J[j];
      s_pff = s;
    }
    __syncthreads();
    for (int j = tid; j < D; j += nt) {
      const double v = (j == c) ? s_pff : Pf[j];
      P[(size_t)c * ldp + j] = v;
      P[(size_t)j * ldp + c] = v;
    }
    __syncthreads();
  }
}

void launch_reanchor(const ReanchorRec* recs, const int* rec_off, int n_filters, const HybWork* hw,
                     const double* clones, size_t clone_stride, const double* imu, const double* fpos, double* fidp,
                     int fcap, double* P, size_t p_stride, int ldp, cudaStream_t s) {
  k_reanchor<<<n_filters, 256, 2 * (size_t)ldp * sizeof(double), s>>>(recs, rec_off, hw, clones, clone_stride, imu, fpos,
                                                                     fidp, fcap, P, p_stride, ldp);
  check_launch("k_reanchor");
}

// ---------------------------------------------------------------- host mirror of the EKF features
__global__ void k_hybrid_gather(const int* filt, const int* slots, int n, const double* fpos, const double* fidp,
                                int fcap, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t s = (size_t)filt[i] * fcap + slots[i];
  for (int k = 0; k < 3; ++k) out[6 * (size_t)i + k] = fpos[s * FP_STRIDE + k];
  for (int k = 0; k < 3; ++k) out[6 * (size_t)i + 3 + k] = fidp[s * FI_STRIDE + k];
}

void launch_hybrid_gather(const int* filt, const int* slots, int n, const double* fpos, const double* fidp, int fcap,
                          double* out, cudaStream_t s) {
  if (n <= 0) return;
  k_hybrid_gather<<<(n + 127) / 128, 128, 0, s>>>(filt, slots, n, fpos, fidp, fcap, out);
  check_launch("k_hybrid_gather");
}

}  // namespace ob
