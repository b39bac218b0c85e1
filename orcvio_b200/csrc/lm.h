// Levenberg-Marquardt driver of the object optimiser: the control flow of the reference's vendored
// MINPACK port (include/orcvio/utils/EigenLevenbergMarquardt/: LevenbergMarquardt.h:300-395, LMonestep.h:22-206,
// LMpar.h:20-158, LMqrsolv.h:22-103) over a manifold-valued unknown (`plus`, `scaled_norm`).
//
// What differs from the reference:
//  * the model enters through its NORMAL EQUATIONS (J^T J, J^T f, |f|) -- they are what the device reduces a 600 x 45
//    Jacobian to, 17 KB per evaluation instead of the Jacobian itself -- and the triangular factor the trust-region step
//    needs comes from a diagonally pivoted Cholesky of J^T J (the same pivot order as the column-pivoted QR of J:
//    largest remaining column norm) with (Q^T f)_{1..n} = R^-T P^T J^T f;
//  * it is a resumable state machine: `request()` is the point to evaluate next, `feed()` consumes (|f|, J^T J, J^T f)
//    AT THAT POINT (the normal equations ride along with every trial evaluation: on the device they cost the same
//    launch, and an accepted trial point is the next linearisation point), so a batch of objects advances in lock-step
//    with ONE kernel launch per round.
// Everything else (scaling, lmpar, the Givens elimination of sqrt(par) D, gain ratio, step-bound update, the stopping
// tests and their order, the nfev / njev counts) follows the reference; its own known-answer tests
// (src/tests/test_levenberg_marquardt.cpp) are reproduced through orcvio_lm_known_answer.
#pragma once
#include <functional>
#include <vector>

namespace ob {

enum LmStatus {   // LevenbergMarquardtSpace::Status
  LM_NOT_STARTED = -2, LM_RUNNING = -1, LM_IMPROPER_INPUT = 0, LM_REL_REDUCTION_TOO_SMALL = 1, LM_REL_ERROR_TOO_SMALL = 2,
  LM_REL_ERROR_AND_REDUCTION_TOO_SMALL = 3, LM_COSINUS_TOO_SMALL = 4, LM_TOO_MANY_FEV = 5, LM_FTOL_TOO_SMALL = 6,
  LM_XTOL_TOO_SMALL = 7, LM_GTOL_TOO_SMALL = 8, LM_USER_ASKED = 9
};

struct LmOptions {
  double ftol = 1.4901161193847656e-08, xtol = 1.4901161193847656e-08, gtol = 0.0, factor = 100.0;
  int maxfev = 400;
};

struct LmResult { int status = LM_NOT_STARTED, nfev = 0, njev = 0, iterations = 0; double fnorm = 0.0; };

// x (any parametrisation) (+) dx (n tangent entries) -> out;   scaled_norm(diag, x)
using PlusFn = std::function<void(const std::vector<double>&, const double* dx, std::vector<double>&)>;
using NormFn = std::function<double(const double* diag, const std::vector<double>&)>;
// |f(x)| (negative / non-finite = evaluation failed), J^T J (n x n row-major) and J^T f at x
using EvalFn = std::function<double(const std::vector<double>& x, double* JtJ, double* Jtf)>;

class LmSolver {
 public:
  void start(int n, const std::vector<double>& x0, const LmOptions& opt, PlusFn plus, NormFn scaled_norm);
  bool running() const { return phase_ != DONE; }
  const std::vector<double>& request() const { return x_req; }
  bool feed(double fnorm, const double* JtJ, const double* Jtf);   // false once finished
  std::vector<double> x;      // current (finally: optimal) point
  LmResult res;

 private:
  enum Phase { FIRST, TRY, DONE };
  bool outer_begin();
  void propose();
  void finish(int status);
  int n = 0, rank_ = 0, iter_ = 1;
  Phase phase_ = DONE;
  LmOptions opt;
  PlusFn plus;
  NormFn scaled_norm;
  std::vector<double> x_req, JtJ_, Jtf_, R_, qtf_, wa2_, diag_, wa1_;
  std::vector<int> perm_;
  double par_ = 0, delta_ = 0, xnorm_ = 0, fnorm_ = 0, gnorm_ = 0, pnorm_ = 0;
};

// blocking form over a host callback
LmResult lm_minimize(int n, const EvalFn& eval, std::vector<double>& x, const LmOptions& opt, PlusFn plus, NormFn norm);

}  // namespace ob
