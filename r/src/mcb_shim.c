/* mcb_shim.c — thin .Call marshalling between R and libmagmaclust_b200.so (include/mcb.h).
 *
 * R and its headers are not installed in the build image nor on the GPU box. The file is compiled in CI against a
 * stand-in for R's C API (tests/r_stub/include: ~100 lines implementing exactly the entry points used here) and
 * driven by tests/r_stub/drive_shim.c: tests/test_r_shim.py (CPU: every mcbR_* symbol links, error path; GPU: e_step_VEM,
 * one loop body of training_VEM and the callbacks through the shim against the ctypes path). It contains no arithmetic: every entry
 * point copies pointers/lengths out of SEXPs, calls one mcb_* function and wraps the result.
 *
 * Build (where R exists):  R CMD SHLIB mcb_shim.c -I../../include -L../../magmaclust_b200/lib -lmagmaclust_b200
 */
#include <R.h>
#include <Rinternals.h>

#include "mcb.h"

static void handle_finalizer(SEXP ptr) {
  mcb_handle* h = (mcb_handle*)R_ExternalPtrAddr(ptr);
  if (h) {
    mcb_destroy(h);
    R_ClearExternalPtr(ptr);
  }
}

static mcb_handle* get_handle(SEXP ptr) {
  mcb_handle* h = (mcb_handle*)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("magmaclust_b200: handle was destroyed");
  return h;
}

/* Rf_error long-jumps: nothing allocated with malloc may be live when it is raised. All buffers below are R
 * vectors (PROTECTed), so the only thing to release before raising is the PROTECT stack. */
static void check(mcb_handle* h, int rc, int nprotect) {
  if (rc == MCB_OK) return;
  const char* msg = mcb_last_error(h);
  UNPROTECT(nprotect);
  Rf_error("libmagmaclust_b200 error %d: %s", rc, msg ? msg : "");
}

SEXP mcbR_create(SEXP device) {
  mcb_handle* h = NULL;
  int rc = mcb_create(&h, Rf_asInteger(device));
  if (rc != MCB_OK) Rf_error("mcb_create failed (%d): %s", rc, mcb_last_error(NULL));
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* offsets, grid_idx: integer (0-based); y, grid: double */
SEXP mcbR_set_data(SEXP hp, SEXP offsets, SEXP grid_idx, SEXP y, SEXP grid) {
  mcb_handle* h = get_handle(hp);
  int M = Rf_length(offsets) - 1, N = Rf_length(grid);
  long long P = Rf_xlength(y);
  check(h, mcb_set_data(h, M, N, P, INTEGER(offsets), INTEGER(grid_idx), REAL(y), REAL(grid), M), 0);
  return R_NilValue;
}

/* theta_k: 3 x K, theta_i: 3 x M (column-major = [K][3] / [M][3] row-major), tau: K x M (= [M][K]) */
SEXP mcbR_e_step_VEM(SEXP hp, SEXP theta_k, SEXP theta_i, SEXP pi_k, SEXP m_k, SEXP old_tau, SEXP N_) {
  mcb_handle* h = get_handle(hp);
  int K = Rf_length(pi_k), M = Rf_length(theta_i) / 3, N = Rf_asInteger(N_);
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, N, K));      /* column k = mean of cluster k */
  SEXP cov = PROTECT(Rf_alloc3DArray(REALSXP, N, N, K));   /* symmetric: row/column major agree */
  SEXP tau = PROTECT(Rf_allocMatrix(REALSXP, K, M));
  SEXP logL = PROTECT(Rf_allocMatrix(REALSXP, K, M));
  check(h, mcb_e_step_VEM(h, K, REAL(theta_k), REAL(theta_i), REAL(pi_k), REAL(m_k), Rf_length(m_k), REAL(old_tau),
                          REAL(mean), REAL(cov), REAL(tau), REAL(logL)), 4);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, mean);
  SET_VECTOR_ELT(out, 1, cov);
  SET_VECTOR_ELT(out, 2, tau);
  SET_VECTOR_ELT(out, 3, logL);
  UNPROTECT(5);
  return out;
}

/* optimiser callbacks: fn and gr of optimr::opm(..., method = "L-BFGS-B") become one .Call each */
SEXP mcbR_eval_indiv_common(SEXP hp, SEXP theta, SEXP want_grad) {
  mcb_handle* h = get_handle(hp);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 4));
  double* o = REAL(out);
  check(h, mcb_eval_indiv_common(h, REAL(theta), o, Rf_asLogical(want_grad) ? o + 1 : NULL), 1);
  UNPROTECT(1);
  return out;
}

SEXP mcbR_eval_indiv(SEXP hp, SEXP theta /* 3 x M */, SEXP want_grad) {
  mcb_handle* h = get_handle(hp);
  int M = Rf_length(theta) / 3;
  SEXP f = PROTECT(Rf_allocVector(REALSXP, M));
  SEXP g = PROTECT(Rf_allocMatrix(REALSXP, 3, M));
  check(h, mcb_eval_indiv(h, REAL(theta), REAL(f), Rf_asLogical(want_grad) ? REAL(g) : NULL), 2);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, f);
  SET_VECTOR_ELT(out, 1, g);
  UNPROTECT(3);
  return out;
}

SEXP mcbR_eval_mean(SEXP hp, SEXP k /* 0-based, -1 = common */, SEXP theta, SEXP pen_diag, SEXP want_grad) {
  mcb_handle* h = get_handle(hp);
  int nhp = Rf_length(theta), kk = Rf_asInteger(k);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 1 + nhp));
  double* o = REAL(out);
  double* g = Rf_asLogical(want_grad) ? o + 1 : NULL;
  int rc = kk < 0 ? mcb_eval_mean_common(h, REAL(theta), nhp, Rf_asReal(pen_diag), o, g)
                  : mcb_eval_mean(h, kk, REAL(theta), nhp, Rf_asReal(pen_diag), o, g);
  check(h, rc, 1);
  UNPROTECT(1);
  return out;
}

SEXP mcbR_elbo(SEXP hp, SEXP theta_k, SEXP theta_i) {
  mcb_handle* h = get_handle(hp);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 2));
  int resident = Rf_isNull(theta_k);
  check(h, mcb_elbo(h, resident ? NULL : REAL(theta_k), resident ? NULL : REAL(theta_i), REAL(out)), 1);
  UNPROTECT(1);
  return out;
}

/* built-in M-step (GPU-resident L-BFGS); returns list(theta_k 3 x K, theta_i 3 x M, pi_k, stats) */
SEXP mcbR_m_step(SEXP hp, SEXP common_hp_k, SEXP common_hp_i, SEXP K_, SEXP M_) {
  mcb_handle* h = get_handle(hp);
  int K = Rf_asInteger(K_), M = Rf_asInteger(M_);
  SEXP tk = PROTECT(Rf_allocMatrix(REALSXP, 3, K));
  SEXP ti = PROTECT(Rf_allocMatrix(REALSXP, 3, M));
  SEXP pi = PROTECT(Rf_allocVector(REALSXP, K));
  SEXP st = PROTECT(Rf_allocVector(INTSXP, 4));
  check(h, mcb_m_step(h, Rf_asLogical(common_hp_k), Rf_asLogical(common_hp_i), REAL(tk), REAL(ti), REAL(pi), INTEGER(st)), 4);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, tk);
  SET_VECTOR_ELT(out, 1, ti);
  SET_VECTOR_ELT(out, 2, pi);
  SET_VECTOR_ELT(out, 3, st);
  UNPROTECT(5);
  return out;
}

SEXP mcbR_set_state(SEXP hp, SEXP theta_k, SEXP theta_i, SEXP pi_k, SEXP m_k, SEXP tau) {
  mcb_handle* h = get_handle(hp);
  check(h, mcb_set_state(h, Rf_length(pi_k), REAL(theta_k), REAL(theta_i), REAL(pi_k), REAL(m_k), Rf_length(m_k), REAL(tau)), 0);
  return R_NilValue;
}

SEXP mcbR_vem_iteration(SEXP hp, SEXP common_hp_k, SEXP common_hp_i) {
  mcb_handle* h = get_handle(hp);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 3));
  check(h, mcb_vem_iteration(h, Rf_asLogical(common_hp_k), Rf_asLogical(common_hp_i), REAL(out)), 1);
  UNPROTECT(1);
  return out;
}

/* the same iteration with `param` of its E-step in R arrays: the library copies them out on a second stream while the
   M-step runs (mcb.h: mcb_vem_iteration_host). K, M, N as set by mcbR_set_data / mcbR_set_state */
SEXP mcbR_vem_iteration_host(SEXP hp, SEXP common_hp_k, SEXP common_hp_i, SEXP K_, SEXP M_, SEXP N_) {
  mcb_handle* h = get_handle(hp);
  int K = Rf_asInteger(K_), M = Rf_asInteger(M_), N = Rf_asInteger(N_);
  SEXP stats = PROTECT(Rf_allocVector(REALSXP, 3));
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, N, K));
  SEXP cov = PROTECT(Rf_alloc3DArray(REALSXP, N, N, K));
  SEXP tau = PROTECT(Rf_allocMatrix(REALSXP, K, M));
  check(h, mcb_vem_iteration_host(h, Rf_asLogical(common_hp_k), Rf_asLogical(common_hp_i), REAL(stats), REAL(mean),
                                  REAL(cov), REAL(tau)), 4);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, stats);
  SET_VECTOR_ELT(out, 1, mean);
  SET_VECTOR_ELT(out, 2, cov);
  SET_VECTOR_ELT(out, 3, tau);
  UNPROTECT(5);
  return out;
}

SEXP mcbR_posterior_mu_k(SEXP hp, SEXP timestamps, SEXP tau, SEXP K_) {
  mcb_handle* h = get_handle(hp);
  int T = Rf_length(timestamps), K = Rf_asInteger(K_);
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, T, K));
  SEXP cov = PROTECT(Rf_alloc3DArray(REALSXP, T, T, K));
  check(h, mcb_posterior_mu_k(h, T, REAL(timestamps), Rf_isNull(tau) ? NULL : REAL(tau), REAL(mean), REAL(cov)), 2);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, mean);
  SET_VECTOR_ELT(out, 1, cov);
  UNPROTECT(3);
  return out;
}

SEXP mcbR_predict(SEXP hp, SEXP offsets, SEXP obs_idx, SEXP y, SEXP theta_new, SEXP pi_k, SEXP pred_idx, SEXP K_) {
  mcb_handle* h = get_handle(hp);
  int B = Rf_length(offsets) - 1, K = Rf_asInteger(K_), Tp = Rf_length(pred_idx);
  SEXP tau = PROTECT(Rf_allocMatrix(REALSXP, K, B));
  SEXP mean = PROTECT(Rf_alloc3DArray(REALSXP, Tp, K, B));
  SEXP var = PROTECT(Rf_alloc3DArray(REALSXP, Tp, K, B));
  /* no targets (update_tau_star_k_EM, CF.R:764-786): responsibilities only */
  check(h, mcb_predict(h, B, INTEGER(offsets), INTEGER(obs_idx), REAL(y), REAL(theta_new),
                       Rf_isNull(pi_k) ? NULL : REAL(pi_k), Tp, INTEGER(pred_idx), REAL(tau), Tp > 0 ? REAL(mean) : NULL,
                       Tp > 0 ? REAL(var) : NULL), 3);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, tau);
  SET_VECTOR_ELT(out, 1, mean);
  SET_VECTOR_ELT(out, 2, var);
  UNPROTECT(4);
  return out;
}

/* kern_to_cov / kern_to_inv, vector branch (CF.R:68-86, 130-149): x sorted, theta = c(a, b); n x n matrix (symmetric, so the
 * row-major result needs no transpose). want_inv: the inverse, with attr(, "logdet") = log det of the covariance. */
/* train_new_gp_EM with hp_i = NULL (MC.R:142-186, repaired) for B new individuals in CSR; theta0 of length 3 */
SEXP mcbR_train_new_indiv(SEXP hp, SEXP offsets, SEXP obs_idx, SEXP y, SEXP theta0, SEXP pi_k, SEXP n_loop_max) {
  mcb_handle* h = get_handle(hp);
  int B = Rf_length(offsets) - 1, K = Rf_length(pi_k);
  SEXP theta = PROTECT(Rf_allocMatrix(REALSXP, 3, B));   /* column b = theta_new of individual b */
  SEXP tau = PROTECT(Rf_allocMatrix(REALSXP, K, B));
  SEXP loops = PROTECT(Rf_allocVector(INTSXP, B));
  check(h, mcb_train_new_indiv(h, B, INTEGER(offsets), INTEGER(obs_idx), REAL(y), REAL(theta0), 0, REAL(pi_k),
                               Rf_asInteger(n_loop_max), REAL(theta), REAL(tau), INTEGER(loops)), 3);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, theta);
  SET_VECTOR_ELT(out, 1, tau);
  SET_VECTOR_ELT(out, 2, loops);
  UNPROTECT(4);
  return out;
}

SEXP mcbR_kern(SEXP hp, SEXP x, SEXP theta, SEXP sigma, SEXP want_inv) {
  mcb_handle* h = get_handle(hp);
  int n = Rf_length(x);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  if (Rf_asLogical(want_inv)) {
    double logdet = 0.0;
    check(h, mcb_kern_to_inv(h, n, REAL(x), REAL(theta), Rf_asReal(sigma), REAL(out), &logdet), 1);
    SEXP ld = PROTECT(Rf_ScalarReal(logdet));
    Rf_setAttrib(out, Rf_install("logdet"), ld);
    UNPROTECT(1);
  } else {
    check(h, mcb_kern_to_cov(h, n, REAL(x), REAL(theta), Rf_asReal(sigma), REAL(out)), 1);
  }
  UNPROTECT(1);
  return out;
}

/* dmvnorm(x, mu, inv_Sigma, log) CF.R:10-35 */
SEXP mcbR_dmvnorm(SEXP hp, SEXP x, SEXP mu, SEXP inv_Sigma, SEXP want_log) {
  mcb_handle* h = get_handle(hp);
  double out = 0.0;
  check(h, mcb_dmvnorm(h, Rf_length(mu), REAL(x), REAL(mu), REAL(inv_Sigma), Rf_asLogical(want_log), &out), 0);
  return Rf_ScalarReal(out);
}

/* update_tau_i_k_VEM (CF.R:733-762): responsibilities from given means (N x K) and covariances (N x N x K) */
SEXP mcbR_update_tau(SEXP hp, SEXP mean, SEXP cov, SEXP K_, SEXP M_) {
  mcb_handle* h = get_handle(hp);
  int K = Rf_asInteger(K_), M = Rf_asInteger(M_);
  SEXP tau = PROTECT(Rf_allocMatrix(REALSXP, K, M));
  check(h, mcb_set_posterior(h, REAL(mean), REAL(cov), NULL), 1);
  check(h, mcb_get_tau_new(h, REAL(tau)), 1);
  UNPROTECT(1);
  return tau;
}


/* ---- round 2: the rest of the reference's interface -------------------------------------------------------------------- */

/* Content hash of a column (numeric, integer or character): decides on the R side whether the resident copy of `db` is
 * still the caller's data (FNV-1a, 64 bit, returned as 16 hex digits; R has no 64-bit integer type). */
SEXP mcbR_hash(SEXP x) {
  unsigned long long hsh = 1469598103934665603ULL;
  const unsigned char* p = NULL;
  size_t nb = 0;
  if (TYPEOF(x) == REALSXP) { p = (const unsigned char*)REAL(x); nb = sizeof(double) * (size_t)Rf_xlength(x); }
  else if (TYPEOF(x) == INTSXP || TYPEOF(x) == LGLSXP) { p = (const unsigned char*)INTEGER(x); nb = sizeof(int) * (size_t)Rf_xlength(x); }
  if (p) {
    for (size_t i = 0; i < nb; ++i) { hsh ^= p[i]; hsh *= 1099511628211ULL; }
  } else if (TYPEOF(x) == STRSXP) {
    for (R_xlen_t j = 0; j < Rf_xlength(x); ++j) {
      const char* c = CHAR(STRING_ELT(x, j));
      for (; *c; ++c) { hsh ^= (unsigned char)*c; hsh *= 1099511628211ULL; }
      hsh ^= 0xffu; hsh *= 1099511628211ULL;   /* element separator */
    }
  } else {
    Rf_error("mcbR_hash: unsupported column type");
  }
  char buf[17];
  snprintf(buf, sizeof buf, "%016llx", hsh);
  return Rf_mkString(buf);
}

/* mu_k_param / list_mu_param given explicitly (CF.R:218, 280, 327, 631): mean N x K, cov N x N x K, tau K x M or NULL */
SEXP mcbR_set_posterior(SEXP hp, SEXP mean, SEXP cov, SEXP tau) {
  mcb_handle* h = get_handle(hp);
  check(h, mcb_set_posterior(h, REAL(mean), REAL(cov), Rf_isNull(tau) ? NULL : REAL(tau)), 0);
  return R_NilValue;
}

/* list_mu of pred_gp_clust / mean_k, cov_k of update_tau_star_k_EM given explicitly: mean T x K, cov T x T x K */
SEXP mcbR_set_pred_posterior(SEXP hp, SEXP timestamps, SEXP mean, SEXP cov, SEXP pi_k) {
  mcb_handle* h = get_handle(hp);
  int T = Rf_length(timestamps), K = Rf_length(mean) / (T > 0 ? T : 1);
  check(h, mcb_set_pred_posterior(h, K, T, REAL(timestamps), REAL(mean), REAL(cov), Rf_isNull(pi_k) ? NULL : REAL(pi_k)), 0);
  return R_NilValue;
}

/* logL_GP_mod / gr_GP_mod (Z = 1) and the common_hp_k forms on explicit arguments: t[n], y n x Z, mean (Z or n x Z values),
 * new_cov n x n x Z, theta (2 or 3), pen_diag; returns c(f, g...) */
SEXP mcbR_gp_mod(SEXP hp, SEXP t, SEXP y, SEXP mean, SEXP new_cov, SEXP theta, SEXP pen_diag, SEXP want_grad) {
  mcb_handle* h = get_handle(hp);
  int n = Rf_length(t), Z = Rf_length(y) / (n > 0 ? n : 1), nhp = Rf_length(theta);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 1 + nhp));
  double* o = REAL(out);
  check(h, mcb_gp_mod(h, n, Z, REAL(t), REAL(y), REAL(mean), Rf_length(mean), REAL(new_cov), REAL(theta), nhp,
                      Rf_asReal(pen_diag), o, Rf_asLogical(want_grad) ? o + 1 : NULL), 1);
  UNPROTECT(1);
  return out;
}

/* hp = new_hp of the last M-step (training_VEM returns it even when it was not accepted, MC.R:95) */
SEXP mcbR_get_new_hp(SEXP hp, SEXP K_, SEXP M_) {
  mcb_handle* h = get_handle(hp);
  int K = Rf_asInteger(K_), M = Rf_asInteger(M_);
  SEXP tk = PROTECT(Rf_allocMatrix(REALSXP, 3, K));
  SEXP ti = PROTECT(Rf_allocMatrix(REALSXP, 3, M));
  SEXP pi = PROTECT(Rf_allocVector(REALSXP, K));
  check(h, mcb_get_new_hp(h, REAL(tk), REAL(ti), REAL(pi)), 3);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, tk);
  SET_VECTOR_ELT(out, 1, ti);
  SET_VECTOR_ELT(out, 2, pi);
  UNPROTECT(4);
  return out;
}

/* terms of BIC (CF.R:354-432) at the resident hp / hyper-posterior */
SEXP mcbR_bic_terms(SEXP hp) {
  mcb_handle* h = get_handle(hp);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 3));
  check(h, mcb_bic_terms(h, REAL(out)), 1);
  UNPROTECT(1);
  return out;
}

SEXP mcbR_version(void) { return Rf_ScalarInteger(mcb_version()); }
