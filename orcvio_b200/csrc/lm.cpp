// See lm.h.  Host code: the unknown has at most a few dozen dimensions; the model is evaluated on the device.
#include "lm.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace ob {

namespace {

constexpr double EPS = std::numeric_limits<double>::epsilon();
constexpr double DWARF = std::numeric_limits<double>::min();

double norm2(const std::vector<double>& v) {
  double s = 0.0;
  for (double x : v) s += x * x;
  return std::sqrt(s);
}

// R (n x n upper, row-major, rows >= rank zero) with R^T R = P^T (J^T J) P, perm[j] = original column at position j
int pivoted_cholesky(const double* JtJ, int n, std::vector<double>& R, std::vector<int>& perm) {
  std::vector<double> A(JtJ, JtJ + (size_t)n * n);
  perm.resize(n);
  for (int j = 0; j < n; ++j) perm[j] = j;
  R.assign((size_t)n * n, 0.0);
  double maxpiv = 0.0;
  int rank = n;
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int j = k + 1; j < n; ++j)
      if (A[(size_t)j * n + j] > A[(size_t)p * n + p]) p = j;
    if (p != k) {                        // symmetric swap of rows / columns k and p, and of the computed part of R
      for (int i = 0; i < n; ++i) std::swap(A[(size_t)i * n + k], A[(size_t)i * n + p]);
      for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
      for (int i = 0; i < k; ++i) std::swap(R[(size_t)i * n + k], R[(size_t)i * n + p]);
      std::swap(perm[k], perm[p]);
    }
    const double d = A[(size_t)k * n + k];
    if (k == 0) maxpiv = d;
    // (ColPivHouseholderQR::rank(): |r_kk| > eps * min(m, n) * |r_00|, on the squares here)
    if (!(d > 0.0) || std::sqrt(d) <= std::sqrt(maxpiv) * EPS * n) { rank = k; break; }
    const double r = std::sqrt(d);
    R[(size_t)k * n + k] = r;
    for (int j = k + 1; j < n; ++j) R[(size_t)k * n + j] = A[(size_t)k * n + j] / r;
    for (int i = k + 1; i < n; ++i)
      for (int j = i; j < n; ++j) {
        A[(size_t)i * n + j] -= R[(size_t)k * n + i] * R[(size_t)k * n + j];
        A[(size_t)j * n + i] = A[(size_t)i * n + j];
      }
  }
  return rank;
}

// lmqrsolv (LMqrsolv.h:22-103).  s: n x n row-major, upper triangle = R (its lower part is scratch)
void qrsolv(std::vector<double>& s, int n, const std::vector<int>& perm, const std::vector<double>& diag,
            const std::vector<double>& qtb, std::vector<double>& x, std::vector<double>& sdiag) {
  auto S = [&](int i, int j) -> double& { return s[(size_t)i * n + j]; };
  std::vector<double> save(n), wa(qtb);
  for (int i = 0; i < n; ++i) save[i] = S(i, i);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) S(i, j) = S(j, i);
  sdiag.assign(n, 0.0);
  for (int j = 0; j < n; ++j) {
    const int l = perm[j];
    if (diag[l] == 0.0) break;
    for (int k = j; k < n; ++k) sdiag[k] = 0.0;
    sdiag[j] = diag[l];
    double qtbpj = 0.0;
    for (int k = j; k < n; ++k) {
      // JacobiRotation::makeGivens(-s(k,k), sdiag[k]) (real case)
      const double p = -S(k, k), q = sdiag[k];
      double c, sn;
      if (q == 0.0) { c = p < 0 ? -1.0 : 1.0; sn = 0.0; }
      else if (p == 0.0) { c = 0.0; sn = q < 0 ? 1.0 : -1.0; }
      else if (std::fabs(p) > std::fabs(q)) {
        const double t = q / p;
        double u = std::sqrt(1.0 + t * t);
        if (p < 0) u = -u;
        c = 1.0 / u;
        sn = -t * c;
      } else {
        const double t = p / q;
        double u = std::sqrt(1.0 + t * t);
        if (q < 0) u = -u;
        sn = -1.0 / u;
        c = -t * sn;
      }
      S(k, k) = c * S(k, k) + sn * sdiag[k];
      const double temp = c * wa[k] + sn * qtbpj;
      qtbpj = -sn * wa[k] + c * qtbpj;
      wa[k] = temp;
      for (int i = k + 1; i < n; ++i) {
        const double t2 = c * S(i, k) + sn * sdiag[i];
        sdiag[i] = -sn * S(i, k) + c * sdiag[i];
        S(i, k) = t2;
      }
    }
  }
  int nsing = 0;
  while (nsing < n && sdiag[nsing] != 0.0) ++nsing;
  for (int i = nsing; i < n; ++i) wa[i] = 0.0;
  for (int i = nsing - 1; i >= 0; --i) {          // s(:nsing,:nsing)^T upper = the modified factor stored below the diagonal
    double acc = wa[i];
    for (int j = i + 1; j < nsing; ++j) acc -= S(j, i) * wa[j];
    wa[i] = acc / S(i, i);
  }
  for (int i = 0; i < n; ++i) { sdiag[i] = S(i, i); S(i, i) = save[i]; }
  x.assign(n, 0.0);
  for (int j = 0; j < n; ++j) x[perm[j]] = wa[j];
}

// lmpar2 (LMpar.h:20-158)
void lmpar(const std::vector<double>& R, int n, const std::vector<int>& perm, int rank, const std::vector<double>& diag,
           const std::vector<double>& qtb, double delta, double& par, std::vector<double>& x) {
  std::vector<double> s(R), wa1(qtb), wa2(n), sdiag;
  auto S = [&](int i, int j) -> double& { return s[(size_t)i * n + j]; };
  for (int i = rank; i < n; ++i) wa1[i] = 0.0;
  for (int i = rank - 1; i >= 0; --i) {
    double acc = qtb[i];
    for (int j = i + 1; j < rank; ++j) acc -= S(i, j) * wa1[j];
    wa1[i] = acc / S(i, i);
  }
  x.assign(n, 0.0);
  for (int j = 0; j < n; ++j) x[perm[j]] = wa1[j];
  for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = norm2(wa2);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) { par = 0.0; return; }
  double parl = 0.0;
  if (rank == n) {
    for (int j = 0; j < n; ++j) wa1[j] = diag[perm[j]] * wa2[perm[j]] / dxnorm;
    for (int i = 0; i < n; ++i) {                  // R^T (lower) forward solve
      double acc = wa1[i];
      for (int j = 0; j < i; ++j) acc -= S(j, i) * wa1[j];
      wa1[i] = acc / S(i, i);
    }
    const double temp = norm2(wa1);
    parl = fp / delta / temp / temp;
  }
  for (int j = 0; j < n; ++j) {
    double acc = 0.0;
    for (int i = 0; i <= j; ++i) acc += S(i, j) * qtb[i];
    wa1[j] = acc / diag[perm[j]];
  }
  const double gnorm = norm2(wa1);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = DWARF / std::min(delta, 0.1);
  par = std::max(par, parl);
  par = std::min(par, paru);
  if (par == 0.0) par = gnorm / dxnorm;
  int iter = 0;
  for (;;) {
    ++iter;
    if (par == 0.0) par = std::max(DWARF, 0.001 * paru);
    std::vector<double> d(n);
    for (int j = 0; j < n; ++j) d[j] = std::sqrt(par) * diag[j];
    qrsolv(s, n, perm, d, qtb, x, sdiag);
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = norm2(wa2);
    double temp = fp;
    fp = dxnorm - delta;
    if (std::fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    for (int j = 0; j < n; ++j) wa1[j] = diag[perm[j]] * (wa2[perm[j]] / dxnorm);
    for (int j = 0; j < n; ++j) {
      wa1[j] /= sdiag[j];
      temp = wa1[j];
      for (int i = j + 1; i < n; ++i) wa1[i] -= S(i, j) * temp;
    }
    temp = norm2(wa1);
    const double parc = fp / delta / temp / temp;
    if (fp > 0.0) parl = std::max(parl, par);
    if (fp < 0.0) paru = std::min(paru, par);
    par = std::max(parl, par + parc);
  }
}

}  // namespace

void LmSolver::start(int n_, const std::vector<double>& x0, const LmOptions& opt_, PlusFn plus_, NormFn norm_) {
  n = n_;
  opt = opt_;
  plus = std::move(plus_);
  scaled_norm = std::move(norm_);
  x = x0;
  x_req = x0;
  res = LmResult();
  res.status = LM_RUNNING;
  phase_ = FIRST;
  iter_ = 1;
  par_ = delta_ = xnorm_ = gnorm_ = 0.0;
  JtJ_.assign((size_t)n * n, 0.0);
  Jtf_.assign(n, 0.0);
  qtf_.assign(n, 0.0);
  wa2_.assign(n, 0.0);
  diag_.assign(n, 1.0);
  if (n <= 0 || opt.ftol < 0 || opt.xtol < 0 || opt.gtol < 0 || opt.maxfev <= 0 || opt.factor <= 0) finish(LM_IMPROPER_INPUT);
}

void LmSolver::finish(int status) {
  res.status = status;
  res.iterations = iter_;
  res.fnorm = fnorm_;
  phase_ = DONE;
}

// head of minimizeOneStep (LMonestep.h:22-93) on the normal equations held for x; false = finished
bool LmSolver::outer_begin() {
  ++res.njev;
  for (int j = 0; j < n; ++j) wa2_[j] = std::sqrt(std::max(JtJ_[(size_t)j * n + j], 0.0));
  rank_ = pivoted_cholesky(JtJ_.data(), n, R_, perm_);
  if (iter_ == 1) {
    for (int j = 0; j < n; ++j) diag_[j] = wa2_[j] == 0.0 ? 1.0 : wa2_[j];
    xnorm_ = scaled_norm(diag_.data(), x);
    delta_ = opt.factor * xnorm_;
    if (delta_ == 0.0) delta_ = opt.factor;
  }
  // (Q^T f)_{1..n} = R^-T P^T J^T f over the leading `rank` rows
  for (int i = 0; i < n; ++i) qtf_[i] = 0.0;
  for (int i = 0; i < rank_; ++i) {
    double acc = Jtf_[perm_[i]];
    for (int j = 0; j < i; ++j) acc -= R_[(size_t)j * n + i] * qtf_[j];
    qtf_[i] = acc / R_[(size_t)i * n + i];
  }
  gnorm_ = 0.0;
  if (fnorm_ != 0.0)
    for (int j = 0; j < n; ++j)
      if (wa2_[perm_[j]] != 0.0) {
        double acc = 0.0;
        for (int i = 0; i <= j; ++i) acc += R_[(size_t)i * n + j] * (qtf_[i] / fnorm_);
        gnorm_ = std::max(gnorm_, std::fabs(acc / wa2_[perm_[j]]));
      }
  if (gnorm_ <= opt.gtol) { finish(LM_COSINUS_TOO_SMALL); return false; }
  for (int j = 0; j < n; ++j) diag_[j] = std::max(diag_[j], wa2_[j]);
  return true;
}

// the trust-region step of the inner loop (LMonestep.h:95-113): fills x_req with the trial point
void LmSolver::propose() {
  lmpar(R_, n, perm_, rank_, diag_, qtf_, delta_, par_, wa1_);
  for (double& v : wa1_) v = -v;
  plus(x, wa1_.data(), x_req);
  double p2 = 0.0;
  for (int j = 0; j < n; ++j) p2 += diag_[j] * wa1_[j] * diag_[j] * wa1_[j];
  pnorm_ = std::sqrt(p2);
  if (iter_ == 1) delta_ = std::min(delta_, pnorm_);
  phase_ = TRY;
}

bool LmSolver::feed(double fnorm_new, const double* JtJ_new, const double* Jtf_new) {
  if (phase_ == DONE) return false;
  if (!(fnorm_new >= 0.0) || !std::isfinite(fnorm_new)) { finish(LM_USER_ASKED); return false; }
  ++res.nfev;
  if (phase_ == FIRST) {
    fnorm_ = fnorm_new;
    std::copy(JtJ_new, JtJ_new + (size_t)n * n, JtJ_.begin());
    std::copy(Jtf_new, Jtf_new + n, Jtf_.begin());
    if (!outer_begin()) return false;
    propose();
    return true;
  }
  // rest of the inner loop (LMonestep.h:115-203) with the evaluation at the trial point
  const double fnorm1 = fnorm_new;
  double actred = -1.0;
  if (0.1 * fnorm1 < fnorm_) actred = 1.0 - (fnorm1 / fnorm_) * (fnorm1 / fnorm_);
  double w3 = 0.0;                                   // |R P^T p|
  for (int i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int j = i; j < n; ++j) acc += R_[(size_t)i * n + j] * wa1_[perm_[j]];
    w3 += acc * acc;
  }
  const double temp1 = w3 / (fnorm_ * fnorm_);
  const double temp2 = par_ * pnorm_ * pnorm_ / (fnorm_ * fnorm_);
  const double prered = temp1 + temp2 / 0.5;
  const double dirder = -(temp1 + temp2);
  const double ratio = prered != 0.0 ? actred / prered : 0.0;
  if (ratio <= 0.25) {
    double temp = actred >= 0.0 ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
    if (0.1 * fnorm1 >= fnorm_ || temp < 0.1) temp = 0.1;
    delta_ = temp * std::min(delta_, pnorm_ / 0.1);
    par_ /= temp;
  } else if (!(par_ != 0.0 && ratio < 0.75)) {
    delta_ = pnorm_ / 0.5;
    par_ = 0.5 * par_;
  }
  const bool accepted = ratio >= 1e-4;
  if (accepted) {
    x = x_req;
    std::copy(JtJ_new, JtJ_new + (size_t)n * n, JtJ_.begin());
    std::copy(Jtf_new, Jtf_new + n, Jtf_.begin());
    xnorm_ = scaled_norm(diag_.data(), x);
    fnorm_ = fnorm1;
    ++iter_;
  }
  const bool small = std::fabs(actred) <= opt.ftol && prered <= opt.ftol && 0.5 * ratio <= 1.0;
  int status = LM_RUNNING;
  if (small && delta_ <= opt.xtol * xnorm_) status = LM_REL_ERROR_AND_REDUCTION_TOO_SMALL;
  else if (small) status = LM_REL_REDUCTION_TOO_SMALL;
  else if (delta_ <= opt.xtol * xnorm_) status = LM_REL_ERROR_TOO_SMALL;
  else if (res.nfev >= opt.maxfev) status = LM_TOO_MANY_FEV;
  else if (std::fabs(actred) <= EPS && prered <= EPS && 0.5 * ratio <= 1.0) status = LM_FTOL_TOO_SMALL;
  else if (delta_ <= EPS * xnorm_) status = LM_XTOL_TOO_SMALL;
  else if (gnorm_ <= EPS) status = LM_GTOL_TOO_SMALL;
  if (status != LM_RUNNING) { finish(status); return false; }
  if (accepted && !outer_begin()) return false;
  propose();
  return true;
}

LmResult lm_minimize(int n, const EvalFn& eval, std::vector<double>& x, const LmOptions& opt, PlusFn plus, NormFn norm) {
  LmSolver s;
  s.start(n, x, opt, std::move(plus), std::move(norm));
  std::vector<double> JtJ((size_t)std::max(n, 1) * std::max(n, 1)), Jtf(std::max(n, 1));
  while (s.running()) {
    const double fn = eval(s.request(), JtJ.data(), Jtf.data());
    s.feed(fn, JtJ.data(), Jtf.data());
  }
  x = s.x;
  return s.res;
}

}  // namespace ob
