// EM for new individuals with FREE hyper-parameters — the repaired free-hp branch of train_new_gp_EM (MC.R:142-186).
//
// What the reference intends (its code cannot run: `t` is undefined and the objective adds db$Timestamp):
//   repeat up to n_loop_max = 25 times                                                            MC.R:143-147
//     E step  tau_k  ∝ pi_k N(y*; m_k(t*), Psi_theta(t*) + C_k(t*, t*))          update_tau_star_k_EM, CF.R:764-786
//     M step  theta' = argmin_theta  - sum_k tau_k log N(y*; m_k(t*), Psi_theta(t*) + C_k(t*, t*))   MC.R:151-166
//             by L-BFGS-B WITHOUT an analytic gradient (opm(hp, LL_GP, ...), MC.R:166: optim's central differences,
//             ndeps = 1e-3)
//     stop when 0 < sum |theta' - theta| < 1e-3                                                   MC.R:175-181
//   return theta' of the last M step and tau_k of the last E step                                  MC.R:192
//
// This file is HOST code only: the driver runs B independent EM loops in lock step and asks a callback for the K
// log-densities of many (individual, theta) points at once — 7 points per active individual and round (the centre and
// the six central-difference neighbours). On the GPU path the callback is one launch of new_indiv_logl_kernel
// (predict.cu); the CPU tests drive the same code with the oracle's log-density (tests/test_new_indiv_em.py), so the
// control flow is covered without a GPU. Optimiser: lbfgs.h (as for the M-step of training).
#include <cmath>
#include <vector>

#include "../../include/mcb.h"
#include "lbfgs.h"

namespace {

constexpr double kNdeps = 1e-3;   // optim's default finite-difference step
constexpr int kPts = 7;           // centre, +-e1, +-e2, +-e3

enum Stage { ST_ESTEP = 0, ST_MSTEP = 1, ST_DONE = 2 };

struct Indiv {
  int stage = ST_ESTEP;
  int loops = 0;
  double cur[3];
  mcb::Lbfgs opt;
};

// objective and central-difference gradient of one individual from the K log-densities at its 7 points
void objective(const double* logl /*[7][K]*/, const double* tau, int K, double& f, double g[3]) {
  double v[kPts];
  for (int q = 0; q < kPts; ++q) {
    double s = 0.0;
    for (int k = 0; k < K; ++k)
      if (tau[k] != 0.0) s -= tau[k] * logl[q * K + k];   // a cluster of weight exactly 0 does not enter (0 * -Inf)
    v[q] = s;
  }
  f = v[0];
  for (int j = 0; j < 3; ++j) g[j] = (v[1 + 2 * j] - v[2 + 2 * j]) / (2.0 * kNdeps);
}

}  // namespace

extern "C" int mcb_new_indiv_em(int B, int K, const double* theta0, int theta0_per_indiv, const double* pi_k,
                                int n_loop_max, mcb_logl_fn fn, void* ctx, double* out_theta, double* out_tau,
                                int* out_loops) {
  if (B < 1 || K < 1 || !theta0 || !pi_k || !fn || !out_theta || !out_tau || n_loop_max < 1) return MCB_EINVAL;
  std::vector<Indiv> st(B);
  for (int b = 0; b < B; ++b) {
    const double* t0 = theta0 + (theta0_per_indiv ? 3 * (size_t)b : 0);
    for (int j = 0; j < 3; ++j) { st[b].cur[j] = t0[j]; out_theta[3 * (size_t)b + j] = t0[j]; }
    for (int k = 0; k < K; ++k) out_tau[(size_t)b * K + k] = NAN;
    if (out_loops) out_loops[b] = 0;
  }
  std::vector<int> active, pt_ind;
  std::vector<double> pt_theta, logl;
  for (;;) {
    active.clear();
    for (int b = 0; b < B; ++b)
      if (st[b].stage != ST_DONE) active.push_back(b);
    if (active.empty()) break;
    const int npts = kPts * (int)active.size();
    pt_ind.resize(npts);
    pt_theta.resize(3 * (size_t)npts);
    logl.resize((size_t)npts * K);
    for (size_t a = 0; a < active.size(); ++a) {
      const Indiv& s = st[active[a]];
      const double* x = s.stage == ST_ESTEP ? s.cur : s.opt.x;
      for (int q = 0; q < kPts; ++q) {
        const size_t p = a * kPts + q;
        pt_ind[p] = active[a];
        for (int j = 0; j < 3; ++j) pt_theta[3 * p + j] = x[j];
        if (q > 0) pt_theta[3 * p + (q - 1) / 2] += (q & 1) ? kNdeps : -kNdeps;
      }
    }
    const int rc = fn(ctx, npts, pt_ind.data(), pt_theta.data(), logl.data());
    if (rc != MCB_OK) return rc;
    for (size_t a = 0; a < active.size(); ++a) {
      const int b = active[a];
      Indiv& s = st[b];
      const double* L = logl.data() + a * kPts * K;
      double* tau = out_tau + (size_t)b * K;
      double f, g[3];
      int status;
      if (s.stage == ST_ESTEP) {
        // E step at the centre point (CF.R:781-785: max-shift, then pi_k, then normalise)
        double mx = -INFINITY;
        bool fin = true;
        for (int k = 0; k < K; ++k) { fin = fin && std::isfinite(L[k]); mx = std::fmax(mx, L[k]); }
        if (!fin) {   // the covariance at the current theta is not positive definite: nothing valid to return
          for (int k = 0; k < K; ++k) tau[k] = NAN;
          if (out_loops) out_loops[b] = -1;
          s.stage = ST_DONE;
          continue;
        }
        double den = 0.0;
        for (int k = 0; k < K; ++k) { tau[k] = pi_k[k] * std::exp(L[k] - mx); den += tau[k]; }
        for (int k = 0; k < K; ++k) tau[k] /= den;
        mcb::lb_init(s.opt, 3, s.cur, 100, 1e7);
        objective(L, tau, K, f, g);
        status = mcb::lb_advance(s.opt, f, g);   // the first evaluation of the M step re-uses this round's points
        s.stage = ST_MSTEP;
      } else {
        objective(L, tau, K, f, g);
        status = mcb::lb_advance(s.opt, f, g);
      }
      if (status == mcb::LB_EVAL) continue;
      // the M step has ended (MC.R:166-181)
      s.loops++;
      if (out_loops) out_loops[b] = s.loops;
      if (status == mcb::LB_NONFINITE) {   // MC.R:167-172: values of the last valid iteration
        s.stage = ST_DONE;
        continue;
      }
      double eps = 0.0;
      for (int j = 0; j < 3; ++j) {
        eps += std::fabs(s.opt.x[j] - s.cur[j]);
        out_theta[3 * (size_t)b + j] = s.opt.x[j];
      }
      if ((eps > 0.0 && eps < 1e-3) || s.loops >= n_loop_max) {
        s.stage = ST_DONE;
        continue;
      }
      for (int j = 0; j < 3; ++j) s.cur[j] = s.opt.x[j];
      s.stage = ST_ESTEP;
    }
  }
  return MCB_OK;
}
