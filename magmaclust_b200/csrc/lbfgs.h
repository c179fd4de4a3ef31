// Limited-memory BFGS with a More'-Thuente line search, as a reverse-communication state machine
// for tiny problems (D <= 3 hyper-parameters), usable from host code and from device code.
//
// Replaces (for the built-in M-step) the third-party optimiser the reference calls at
// Computing_functions.R:647-648,656-657,666-668,675-677:
//     optimr::opm(par, fn, gr, method = "L-BFGS-B", control = list(kkt = F))  ->  stats::optim
// with R's defaults lmm = 5, factr = 1e7, pgtol = 0, maxit = 100 and NO bounds. Without bounds
// L-BFGS-B's Cauchy-point / subspace steps reduce to the plain L-BFGS direction d = -H g with
// H0 = (s'y / y'y) I, followed by the MINPACK-2 line search (ftol 1e-3, gtol 0.9, xtol 0.1, first
// step min(1/|d|, stpmx), at most 20 evaluations per search, memory reset on a failed search).
// This file restates those published algorithms (Byrd-Lu-Nocedal-Zhu 1995; More'-Thuente 1994);
// it is not a transcription of lbfgsb.c. Optimiser trajectories are "parity unpinned" (SURVEY §8c).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MCB_HD __host__ __device__ __forceinline__
// the larger building blocks stay out of line in device code: the persistent M-step kernel is instruction-fetch bound and
// this update runs on one thread between two evaluations (same arithmetic either way)
#define MCB_HD_BIG static __host__ __device__ __noinline__
#else
#define MCB_HD inline
#define MCB_HD_BIG static inline
#endif

// History of the device-side form: with a run-time D the short loops over D and over the stored pairs unrolled with remainder
// code (4,505 of the 9,272 SASS instructions of the persistent M-step kernel, profiles/r01f_ptxas.txt); -DMCB_LBFGS_ROLLED
// kept them rolled (9,272 -> 6,632 instructions, theta_i phase 4.44 -> 4.29 ms) at the price of local-memory temporaries.
// Since the end of round 2 the kernels run lb_advance_t<3> (bottom of this file: D and the memory depth are compile-time
// constants); LB_LOOP only shapes lb_advance_any, the run-time-D form the host keeps for other sizes and for the twin test.
#if defined(MCB_LBFGS_ROLLED) && defined(__CUDA_ARCH__)
#define LB_LOOP _Pragma("unroll 1")
#else
#define LB_LOOP
#endif

// full unrolling of the compile-time-D loops (device passes only: the host compiler does not know the pragma)
#ifdef __CUDA_ARCH__
#define LB_UNROLL _Pragma("unroll")
#else
#define LB_UNROLL
#endif

namespace mcb {

constexpr int LB_MAXD = 3;
constexpr int LB_MEM = 5;

struct LineSearch {
  double finit, ginit, gtest, width, width1;
  double stx, fx, gx, sty, fy, gy, stmin, stmax;
  double stpmax;
  int brackt, stage;
};

// one safeguarded cubic/quadratic interpolation step (More'-Thuente section 4)
MCB_HD_BIG void mt_step(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                    double fp, double dp, int& brackt, double stpmin, double stpmax) {
  double sgnd = dp * (dx / fabs(dx));
  double stpf, stpc, stpq, theta, s, gamma, p, q, r;
  if (fp > fx) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else stpc = (stp > stx) ? stpmax : stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpf = stp + r * (sty - stp);
    } else {
      stpf = (stp > stx) ? stpmax : stpmin;
    }
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

enum LsStatus { LS_CONTINUE = 0, LS_CONVERGED = 1, LS_WARNING = 2, LS_ERROR = 3 };

MCB_HD void ls_start(LineSearch& ls, double f, double g, double stp, double stpmax) {
  const double ftol = 1e-3;
  ls.brackt = 0; ls.stage = 1;
  ls.finit = f; ls.ginit = g; ls.gtest = ftol * g;
  ls.stpmax = stpmax;
  ls.width = stpmax; ls.width1 = 2.0 * ls.width;
  ls.stx = 0; ls.fx = f; ls.gx = g;
  ls.sty = 0; ls.fy = f; ls.gy = g;
  ls.stmin = 0; ls.stmax = stp + 4.0 * stp;
}

// one line-search iteration: (f, g) evaluated at step `stp`; updates `stp` to the next trial
MCB_HD LsStatus ls_iterate(LineSearch& ls, double f, double g, double& stp) {
  const double gtol = 0.9, xtol = 0.1, stpmin = 0.0;
  const double stpmax = ls.stpmax;
  double ftest = ls.finit + stp * ls.gtest;
  if (ls.stage == 1 && f <= ftest && g >= 0.0) ls.stage = 2;
  LsStatus st = LS_CONTINUE;
  if (ls.brackt && (stp <= ls.stmin || stp >= ls.stmax)) st = LS_WARNING;
  if (ls.brackt && ls.stmax - ls.stmin <= xtol * ls.stmax) st = LS_WARNING;
  if (stp == stpmax && f <= ftest && g <= ls.gtest) st = LS_WARNING;
  if (stp == stpmin && (f > ftest || g >= ls.gtest)) st = LS_WARNING;
  if (f <= ftest && fabs(g) <= gtol * (-ls.ginit)) st = LS_CONVERGED;
  if (st != LS_CONTINUE) return st;
  if (ls.stage == 1 && f <= ls.fx && f > ftest) {
    double fm = f - stp * ls.gtest, fxm = ls.fx - ls.stx * ls.gtest, fym = ls.fy - ls.sty * ls.gtest;
    double gm = g - ls.gtest, gxm = ls.gx - ls.gtest, gym = ls.gy - ls.gtest;
    mt_step(ls.stx, fxm, gxm, ls.sty, fym, gym, stp, fm, gm, ls.brackt, ls.stmin, ls.stmax);
    ls.fx = fxm + ls.stx * ls.gtest; ls.fy = fym + ls.sty * ls.gtest;
    ls.gx = gxm + ls.gtest; ls.gy = gym + ls.gtest;
  } else {
    mt_step(ls.stx, ls.fx, ls.gx, ls.sty, ls.fy, ls.gy, stp, f, g, ls.brackt, ls.stmin, ls.stmax);
  }
  if (ls.brackt) {
    if (fabs(ls.sty - ls.stx) >= 0.66 * ls.width1) stp = ls.stx + 0.5 * (ls.sty - ls.stx);
    ls.width1 = ls.width;
    ls.width = fabs(ls.sty - ls.stx);
    ls.stmin = fmin(ls.stx, ls.sty);
    ls.stmax = fmax(ls.stx, ls.sty);
  } else {
    ls.stmin = stp + 1.1 * (stp - ls.stx);
    ls.stmax = stp + 4.0 * (stp - ls.stx);
  }
  stp = fmax(stp, stpmin);
  stp = fmin(stp, stpmax);
  if ((ls.brackt && (stp <= ls.stmin || stp >= ls.stmax)) ||
      (ls.brackt && ls.stmax - ls.stmin <= xtol * ls.stmax))
    stp = ls.stx;
  return LS_CONTINUE;
}

enum LbStatus {
  LB_EVAL = 0,        // evaluate f,g at x and call lb_advance again
  LB_CONVERGED = 1,   // relative reduction of f <= factr*epsmch  (optim: convergence 0)
  LB_MAXIT = 2,       // iteration limit reached                   (optim: convergence 1)
  LB_ABNORMAL = 3,    // line search failed with empty memory       (optim: convergence 52)
  LB_ZEROGRAD = 4,    // projected gradient exactly zero (pgtol = 0)
  LB_NONFINITE = 5    // f or g not finite at the starting point
};

struct Lbfgs {
  int D, maxit;
  double factr;
  // current accepted iterate and the line-search base point
  double x[LB_MAXD], g[LB_MAXD], f;
  double xb[LB_MAXD], gb[LB_MAXD], fb;
  double d[LB_MAXD];
  double S[LB_MEM][LB_MAXD], Y[LB_MEM][LB_MAXD], rho[LB_MEM];
  int col, head;  // number of stored pairs, index of the oldest
  int iter, nfev, nls, status;
  double stp, gd, gdold, dnorm;
  LineSearch ls;
  int started;
};

MCB_HD void lb_init(Lbfgs& o, int D, const double* x0, int maxit = 100, double factr = 1e7) {
  o.D = D; o.maxit = maxit; o.factr = factr;
  LB_LOOP for (int j = 0; j < D; ++j) o.x[j] = x0[j];
  o.col = 0; o.head = 0; o.iter = 0; o.nfev = 0; o.nls = 0; o.status = LB_EVAL; o.started = 0;
  o.stp = 0; o.f = 0;
}

MCB_HD_BIG void lb_direction(Lbfgs& o) {
  // two-loop recursion over the stored pairs, newest first
  double q[LB_MAXD], alpha[LB_MEM];
  const int D = o.D;
  LB_LOOP for (int j = 0; j < D; ++j) q[j] = o.g[j];
  LB_LOOP for (int c = o.col - 1; c >= 0; --c) {
    int p = (o.head + c) % LB_MEM;
    double a = 0;
    LB_LOOP for (int j = 0; j < D; ++j) a += o.S[p][j] * q[j];
    a *= o.rho[p];
    alpha[c] = a;
    LB_LOOP for (int j = 0; j < D; ++j) q[j] -= a * o.Y[p][j];
  }
  if (o.col > 0) {
    int p = (o.head + o.col - 1) % LB_MEM;
    double yy = 0;
    LB_LOOP for (int j = 0; j < D; ++j) yy += o.Y[p][j] * o.Y[p][j];
    double h0 = 1.0 / (o.rho[p] * yy);  // s'y / y'y
    LB_LOOP for (int j = 0; j < D; ++j) q[j] *= h0;
  }
  LB_LOOP for (int c = 0; c < o.col; ++c) {
    int p = (o.head + c) % LB_MEM;
    double b = 0;
    LB_LOOP for (int j = 0; j < D; ++j) b += o.Y[p][j] * q[j];
    b *= o.rho[p];
    LB_LOOP for (int j = 0; j < D; ++j) q[j] += (alpha[c] - b) * o.S[p][j];
  }
  LB_LOOP for (int j = 0; j < D; ++j) o.d[j] = -q[j];
}

MCB_HD_BIG void lb_begin_search(Lbfgs& o) {
  const int D = o.D;
  for (;;) {
    lb_direction(o);
    double gd = 0, dn = 0;
    LB_LOOP for (int j = 0; j < D; ++j) { gd += o.g[j] * o.d[j]; dn += o.d[j] * o.d[j]; }
    if (gd >= 0.0) {
      // not a descent direction: drop the memory and retry with steepest descent
      if (o.col == 0) { o.status = LB_ABNORMAL; return; }
      o.col = 0; o.head = 0;
      continue;
    }
    o.gd = gd; o.gdold = gd; o.dnorm = sqrt(dn);
    break;
  }
  const double stpmx = 1e10;
  o.stp = (o.iter == 0) ? fmin(1.0 / o.dnorm, stpmx) : 1.0;
  LB_LOOP for (int j = 0; j < D; ++j) { o.xb[j] = o.x[j]; o.gb[j] = o.g[j]; }
  o.fb = o.f;
  ls_start(o.ls, o.f, o.gd, o.stp, stpmx);
  o.nls = 0;
  LB_LOOP for (int j = 0; j < D; ++j) o.x[j] = o.xb[j] + o.stp * o.d[j];
  o.status = LB_EVAL;
}

// Feed the value and gradient at the current o.x. Returns the new status; while it is LB_EVAL the
// caller evaluates at o.x again. On termination o.x/o.f hold the best accepted point.
// (any D <= LB_MAXD, loops over D rolled in device code; lb_advance below dispatches to the compile-time-D version)
MCB_HD int lb_advance_any(Lbfgs& o, double f, const double* g) {
  const int D = o.D;
  const double epsmch = 2.220446049250313e-16;
  o.nfev++;
  if (!o.started) {
    o.started = 1;
    o.f = f;
    bool fin = isfinite(f);
    double pg = 0;
    LB_LOOP for (int j = 0; j < D; ++j) { o.g[j] = g[j]; fin = fin && isfinite(g[j]); pg = fmax(pg, fabs(g[j])); }
    if (!fin) { o.status = LB_NONFINITE; return o.status; }
    if (pg <= 0.0) { o.status = LB_ZEROGRAD; return o.status; }
    lb_begin_search(o);
    return o.status;
  }
  // inside a line search: o.x = xb + stp d
  double gd = 0;
  LB_LOOP for (int j = 0; j < D; ++j) gd += g[j] * o.d[j];
  o.nls++;
  LsStatus st;
  bool fin = isfinite(f) && isfinite(gd);
  if (!fin) {
    // treat a non-finite trial as a failed search step: bisect towards the base point
    st = (o.nls >= 20) ? LS_ERROR : LS_CONTINUE;
    o.stp *= 0.5;
    o.ls.stpmax = o.stp;
  } else {
    st = ls_iterate(o.ls, f, gd, o.stp);
  }
  if (st == LS_CONTINUE && o.nls < 20) {
    LB_LOOP for (int j = 0; j < D; ++j) o.x[j] = o.xb[j] + o.stp * o.d[j];
    o.status = LB_EVAL;
    return o.status;
  }
  if (st != LS_CONVERGED && st != LS_WARNING) {
    // line search failed (20 evaluations or error): restore the base point, reset memory
    LB_LOOP for (int j = 0; j < D; ++j) { o.x[j] = o.xb[j]; o.g[j] = o.gb[j]; }
    o.f = o.fb;
    if (o.col == 0) { o.status = LB_ABNORMAL; return o.status; }
    o.col = 0; o.head = 0;
    lb_begin_search(o);
    return o.status;
  }
  // accept the step
  o.iter++;
  double s[LB_MAXD], y[LB_MAXD];
  double sy = 0, yy = 0;
  LB_LOOP for (int j = 0; j < D; ++j) {
    s[j] = o.x[j] - o.xb[j];
    y[j] = g[j] - o.gb[j];
    sy += s[j] * y[j]; yy += y[j] * y[j];
  }
  double fold = o.fb;
  o.f = f;
  double pg = 0;
  LB_LOOP for (int j = 0; j < D; ++j) { o.g[j] = g[j]; pg = fmax(pg, fabs(g[j])); }
  if (pg <= 0.0) { o.status = LB_ZEROGRAD; return o.status; }
  double ddum = fmax(fmax(fabs(fold), fabs(f)), 1.0);
  if (fold - f <= epsmch * o.factr * ddum) { o.status = LB_CONVERGED; return o.status; }
  if (o.iter >= o.maxit) { o.status = LB_MAXIT; return o.status; }
  // curvature pair; skipped when s'y <= eps * (-g_old's)  (L-BFGS-B's update test)
  double ddum2 = -o.gdold * o.stp;
  if (sy > epsmch * ddum2 && yy > 0.0) {
    int p;
    if (o.col < LB_MEM) { p = (o.head + o.col) % LB_MEM; o.col++; }
    else { p = o.head; o.head = (o.head + 1) % LB_MEM; }
    LB_LOOP for (int j = 0; j < D; ++j) { o.S[p][j] = s[j]; o.Y[p][j] = y[j]; }
    o.rho[p] = 1.0 / sy;
  }
  lb_begin_search(o);
  return o.status;
}

// ---- the same state machine with D known at compile time ----------------------------------------------------------
// In device code the update above runs on ONE thread between two evaluations while the rest of the CTA waits: 7.2 k cycles
// alone on an SM, 13.7 k next to three other CTAs (profiles/r01e_eval_phases.txt) = 16 % of an evaluation. With a run-time D
// the short loops over D either unroll with remainder code (4,505 SASS instructions) or stay rolled, and then q[], s[], y[],
// alpha[] are indexed dynamically and live in LOCAL memory (184 bytes of stack in the M-step kernel): every step of the
// two-loop recursion waits for a local-memory round trip. Here D and the memory depth are constants: all loops unroll
// exactly, every temporary is a register, the ring index needs no modulo. Same operations in the same order as above
// (tests/test_host_logic.py runs both side by side and compares the trajectories bit for bit on the host).
template <int DT>
MCB_HD void lb_direction_t(Lbfgs& o) {
  double q[DT], alpha[LB_MEM];
  const int col = o.col, head = o.head;
LB_UNROLL
  for (int j = 0; j < DT; ++j) q[j] = o.g[j];
LB_UNROLL
  for (int c = LB_MEM - 1; c >= 0; --c) {
    alpha[c] = 0.0;
    if (c < col) {
      int p = head + c;
      if (p >= LB_MEM) p -= LB_MEM;
      double a = 0;
LB_UNROLL
      for (int j = 0; j < DT; ++j) a += o.S[p][j] * q[j];
      a *= o.rho[p];
      alpha[c] = a;
LB_UNROLL
      for (int j = 0; j < DT; ++j) q[j] -= a * o.Y[p][j];
    }
  }
  if (col > 0) {
    int p = head + col - 1;
    if (p >= LB_MEM) p -= LB_MEM;
    double yy = 0;
LB_UNROLL
    for (int j = 0; j < DT; ++j) yy += o.Y[p][j] * o.Y[p][j];
    const double h0 = 1.0 / (o.rho[p] * yy);  // s'y / y'y
LB_UNROLL
    for (int j = 0; j < DT; ++j) q[j] *= h0;
  }
LB_UNROLL
  for (int c = 0; c < LB_MEM; ++c) {
    if (c < col) {
      int p = head + c;
      if (p >= LB_MEM) p -= LB_MEM;
      double b = 0;
LB_UNROLL
      for (int j = 0; j < DT; ++j) b += o.Y[p][j] * q[j];
      b *= o.rho[p];
LB_UNROLL
      for (int j = 0; j < DT; ++j) q[j] += (alpha[c] - b) * o.S[p][j];
    }
  }
LB_UNROLL
  for (int j = 0; j < DT; ++j) o.d[j] = -q[j];
}

template <int DT>
MCB_HD void lb_begin_search_t(Lbfgs& o) {
  for (;;) {
    lb_direction_t<DT>(o);
    double gd = 0, dn = 0;
LB_UNROLL
    for (int j = 0; j < DT; ++j) { gd += o.g[j] * o.d[j]; dn += o.d[j] * o.d[j]; }
    if (gd >= 0.0) {
      if (o.col == 0) { o.status = LB_ABNORMAL; return; }
      o.col = 0; o.head = 0;
      continue;
    }
    o.gd = gd; o.gdold = gd; o.dnorm = sqrt(dn);
    break;
  }
  const double stpmx = 1e10;
  o.stp = (o.iter == 0) ? fmin(1.0 / o.dnorm, stpmx) : 1.0;
LB_UNROLL
  for (int j = 0; j < DT; ++j) { o.xb[j] = o.x[j]; o.gb[j] = o.g[j]; }
  o.fb = o.f;
  ls_start(o.ls, o.f, o.gd, o.stp, stpmx);
  o.nls = 0;
LB_UNROLL
  for (int j = 0; j < DT; ++j) o.x[j] = o.xb[j] + o.stp * o.d[j];
  o.status = LB_EVAL;
}

template <int DT>
MCB_HD int lb_advance_t(Lbfgs& o, double f, const double* g) {
  const double epsmch = 2.220446049250313e-16;
  o.nfev++;
  if (!o.started) {
    o.started = 1;
    o.f = f;
    bool fin = isfinite(f);
    double pg = 0;
LB_UNROLL
    for (int j = 0; j < DT; ++j) { o.g[j] = g[j]; fin = fin && isfinite(g[j]); pg = fmax(pg, fabs(g[j])); }
    if (!fin) { o.status = LB_NONFINITE; return o.status; }
    if (pg <= 0.0) { o.status = LB_ZEROGRAD; return o.status; }
    lb_begin_search_t<DT>(o);
    return o.status;
  }
  double gd = 0;
LB_UNROLL
  for (int j = 0; j < DT; ++j) gd += g[j] * o.d[j];
  o.nls++;
  LsStatus st;
  const bool fin = isfinite(f) && isfinite(gd);
  if (!fin) {
    st = (o.nls >= 20) ? LS_ERROR : LS_CONTINUE;
    o.stp *= 0.5;
    o.ls.stpmax = o.stp;
  } else {
    st = ls_iterate(o.ls, f, gd, o.stp);
  }
  if (st == LS_CONTINUE && o.nls < 20) {
LB_UNROLL
    for (int j = 0; j < DT; ++j) o.x[j] = o.xb[j] + o.stp * o.d[j];
    o.status = LB_EVAL;
    return o.status;
  }
  if (st != LS_CONVERGED && st != LS_WARNING) {
LB_UNROLL
    for (int j = 0; j < DT; ++j) { o.x[j] = o.xb[j]; o.g[j] = o.gb[j]; }
    o.f = o.fb;
    if (o.col == 0) { o.status = LB_ABNORMAL; return o.status; }
    o.col = 0; o.head = 0;
    lb_begin_search_t<DT>(o);
    return o.status;
  }
  o.iter++;
  double s[DT], y[DT];
  double sy = 0, yy = 0;
LB_UNROLL
  for (int j = 0; j < DT; ++j) {
    s[j] = o.x[j] - o.xb[j];
    y[j] = g[j] - o.gb[j];
    sy += s[j] * y[j]; yy += y[j] * y[j];
  }
  const double fold = o.fb;
  o.f = f;
  double pg = 0;
LB_UNROLL
  for (int j = 0; j < DT; ++j) { o.g[j] = g[j]; pg = fmax(pg, fabs(g[j])); }
  if (pg <= 0.0) { o.status = LB_ZEROGRAD; return o.status; }
  const double ddum = fmax(fmax(fabs(fold), fabs(f)), 1.0);
  if (fold - f <= epsmch * o.factr * ddum) { o.status = LB_CONVERGED; return o.status; }
  if (o.iter >= o.maxit) { o.status = LB_MAXIT; return o.status; }
  const double ddum2 = -o.gdold * o.stp;
  if (sy > epsmch * ddum2 && yy > 0.0) {
    int p;
    if (o.col < LB_MEM) { p = o.head + o.col; if (p >= LB_MEM) p -= LB_MEM; o.col++; }
    else { p = o.head; o.head = (o.head + 1 == LB_MEM) ? 0 : o.head + 1; }
LB_UNROLL
    for (int j = 0; j < DT; ++j) { o.S[p][j] = s[j]; o.Y[p][j] = y[j]; }
    o.rho[p] = 1.0 / sy;
  }
  lb_begin_search_t<DT>(o);
  return o.status;
}

// Every device-side caller optimises three hyper-parameters; the host also runs D = 2 (per-cluster theta_k with a fixed noise).
MCB_HD int lb_advance(Lbfgs& o, double f, const double* g) {
#if defined(__CUDA_ARCH__) && !defined(MCB_LBFGS_ANYD)
  return lb_advance_t<3>(o, f, g);   // D == 3 is the only size the kernels set up (lb_init(.., 3, ..))
#else
  if (o.D == 3) return lb_advance_t<3>(o, f, g);
  if (o.D == 2) return lb_advance_t<2>(o, f, g);
  return lb_advance_any(o, f, g);
#endif
}

}  // namespace mcb
