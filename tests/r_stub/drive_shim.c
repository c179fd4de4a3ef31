/* Drives r/src/mcb_shim.c the way R's .Call would, against the stand-in R API of tests/r_stub/include (test infrastructure).
 *
 *   drive_shim abi                      no GPU needed: version, hash, and the error path of mcbR_create when there is no device
 *   drive_shim estep <in.bin> <out.bin> GPU: e_step_VEM, one loop body of training_VEM and two callbacks through the shim
 *
 * in.bin : int32 M, N, K, P | int32 offsets[M+1] | int32 grid_idx[P] | f64 y[P] | f64 grid[N] | f64 theta_k[K][3] |
 *          f64 theta_i[M][3] | f64 pi[K] | f64 m_k[K] | f64 tau[M][K]
 * out.bin: f64 mean[K][N] | cov[K][N][N] | tau[M][K] | eval_indiv_common (4) | eval_mean common (4) | vem stats (3) |
 *          vem mean[K][N] | new theta_k[K][3] | elbo (2) | bic terms (3) | posterior_mu_k mean[K][N] on the training grid |
 *          tau_star[K] of individual 0 (no targets) | pred mean[K][N] | pred var[K][N] | gp_mod (4) on cluster 0 of the E-step |
 *          the same prediction after set_pred_posterior with the arrays just read: tau[K] | free-hp EM: theta[3], tau[K]
 */
#include <R.h>
#include <Rinternals.h>

jmp_buf mcb_stub_error_jmp;
char mcb_stub_error_msg[1024];
int mcb_stub_protect_depth = 0;

SEXP mcbR_create(SEXP), mcbR_version(void), mcbR_hash(SEXP);
SEXP mcbR_set_data(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_e_step_VEM(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_eval_indiv_common(SEXP, SEXP, SEXP);
SEXP mcbR_eval_mean(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_set_state(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_vem_iteration_host(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_get_new_hp(SEXP, SEXP, SEXP);
SEXP mcbR_elbo(SEXP, SEXP, SEXP), mcbR_bic_terms(SEXP);
SEXP mcbR_posterior_mu_k(SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_predict(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_gp_mod(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_set_pred_posterior(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mcbR_train_new_indiv(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

static SEXP reals(const double* p, long n) { SEXP s = Rf_allocVector(REALSXP, n); memcpy(REAL(s), p, sizeof(double) * n); return s; }
static SEXP ints(const int* p, long n) { SEXP s = Rf_allocVector(INTSXP, n); memcpy(INTEGER(s), p, sizeof(int) * n); return s; }

static int abi(void) {
  if (Rf_asInteger(mcbR_version()) <= 0) return 2;
  double v[3] = {1.0, 2.5, -3.0};
  SEXP h1 = mcbR_hash(reals(v, 3));
  v[2] = 3.0;
  SEXP h2 = mcbR_hash(reals(v, 3));
  if (strlen(CHAR(STRING_ELT(h1, 0))) != 16 || !strcmp(CHAR(STRING_ELT(h1, 0)), CHAR(STRING_ELT(h2, 0)))) return 3;
  printf("hash %s %s\n", CHAR(STRING_ELT(h1, 0)), CHAR(STRING_ELT(h2, 0)));
  if (setjmp(mcb_stub_error_jmp) == 0) {
    SEXP h = mcbR_create(Rf_ScalarInteger(0));
    printf("create ok (a CUDA device is visible) %p\n", R_ExternalPtrAddr(h));
    if (h->finalizer) h->finalizer(h);
  } else {
    printf("create raised: %s\n", mcb_stub_error_msg);     /* the Rf_error path: message from mcb_last_error(NULL) */
    if (!strstr(mcb_stub_error_msg, "mcb_create failed")) return 4;
  }
  if (mcb_stub_protect_depth != 0) return 5;
  printf("abi ok\n");
  return 0;
}

static int estep(const char* in, const char* outp) {
  FILE* f = fopen(in, "rb");
  if (!f) return 10;
  int hd[4];
  if (fread(hd, sizeof(int), 4, f) != 4) return 11;
  const int M = hd[0], N = hd[1], K = hd[2], P = hd[3];
  int* off = malloc(sizeof(int) * (M + 1));
  int* gi = malloc(sizeof(int) * P);
  double* y = malloc(sizeof(double) * P);
  double* grid = malloc(sizeof(double) * N);
  double* tk = malloc(sizeof(double) * 3 * K);
  double* ti = malloc(sizeof(double) * 3 * M);
  double* pi = malloc(sizeof(double) * K);
  double* mk = malloc(sizeof(double) * K);
  double* tau = malloc(sizeof(double) * M * K);
  size_t ok = fread(off, sizeof(int), M + 1, f) + fread(gi, sizeof(int), P, f) + fread(y, sizeof(double), P, f) +
              fread(grid, sizeof(double), N, f) + fread(tk, sizeof(double), 3 * K, f) + fread(ti, sizeof(double), 3 * M, f) +
              fread(pi, sizeof(double), K, f) + fread(mk, sizeof(double), K, f) + fread(tau, sizeof(double), (size_t)M * K, f);
  fclose(f);
  if (ok != (size_t)(M + 1 + P + P + N + 3 * K + 3 * M + K + K + M * K)) return 12;
  if (setjmp(mcb_stub_error_jmp) != 0) {
    fprintf(stderr, "Rf_error: %s\n", mcb_stub_error_msg);
    return 13;
  }
  SEXP h = mcbR_create(Rf_ScalarInteger(0));
  mcbR_set_data(h, ints(off, M + 1), ints(gi, P), reals(y, P), reals(grid, N));
  /* R matrices are column-major: theta_k 3 x K has the same bytes as [K][3]; tau K x M the same as [M][K] */
  SEXP res = mcbR_e_step_VEM(h, reals(tk, 3 * K), reals(ti, 3 * M), reals(pi, K), reals(mk, K), reals(tau, (long)M * K),
                             Rf_ScalarInteger(N));
  FILE* o = fopen(outp, "wb");
  if (!o) return 14;
  fwrite(REAL(VECTOR_ELT(res, 0)), sizeof(double), (size_t)K * N, o);          /* mean: N x K column-major = [K][N] */
  fwrite(REAL(VECTOR_ELT(res, 1)), sizeof(double), (size_t)K * N * N, o);
  fwrite(REAL(VECTOR_ELT(res, 2)), sizeof(double), (size_t)M * K, o);
  double th3[3] = {0.9, 0.7, 0.3};
  SEXP ei = mcbR_eval_indiv_common(h, reals(th3, 3), Rf_ScalarLogical(1));
  fwrite(REAL(ei), sizeof(double), 4, o);
  SEXP em = mcbR_eval_mean(h, Rf_ScalarInteger(-1), reals(th3, 3), Rf_ScalarReal(0.0), Rf_ScalarLogical(1));
  fwrite(REAL(em), sizeof(double), 4, o);
  /* one loop body of training_VEM (MC.R:44-93) from the same state */
  mcbR_set_state(h, reals(tk, 3 * K), reals(ti, 3 * M), reals(pi, K), reals(mk, K), reals(tau, (long)M * K));
  SEXP it = mcbR_vem_iteration_host(h, Rf_ScalarLogical(1), Rf_ScalarLogical(1), Rf_ScalarInteger(K), Rf_ScalarInteger(M),
                                    Rf_ScalarInteger(N));
  fwrite(REAL(VECTOR_ELT(it, 0)), sizeof(double), 3, o);
  fwrite(REAL(VECTOR_ELT(it, 1)), sizeof(double), (size_t)K * N, o);
  SEXP nh = mcbR_get_new_hp(h, Rf_ScalarInteger(K), Rf_ScalarInteger(M));
  fwrite(REAL(VECTOR_ELT(nh, 0)), sizeof(double), 3 * (size_t)K, o);
  /* logL_monitoring_VEM and BIC terms at the resident state (CF.R:327-352, 354-432) */
  fwrite(REAL(mcbR_elbo(h, R_NilValue, R_NilValue)), sizeof(double), 2, o);
  fwrite(REAL(mcbR_bic_terms(h)), sizeof(double), 3, o);
  /* posterior_mu_k -> update_tau_star_k_EM -> pred_gp_clust for individual 0 as the new individual (MC.R:196-274) */
  SEXP pm = mcbR_posterior_mu_k(h, reals(grid, N), R_NilValue, Rf_ScalarInteger(K));
  fwrite(REAL(VECTOR_ELT(pm, 0)), sizeof(double), (size_t)K * N, o);
  const int n0 = off[1];
  int o2[2] = {0, n0};
  int* all = malloc(sizeof(int) * N);
  for (int j = 0; j < N; ++j) all[j] = j;
  SEXP t0 = mcbR_predict(h, ints(o2, 2), ints(gi, n0), reals(y, n0), reals(th3, 3), R_NilValue, ints(all, 0), Rf_ScalarInteger(K));
  fwrite(REAL(VECTOR_ELT(t0, 0)), sizeof(double), K, o);
  SEXP p1 = mcbR_predict(h, ints(o2, 2), ints(gi, n0), reals(y, n0), reals(th3, 3), R_NilValue, ints(all, N), Rf_ScalarInteger(K));
  /* Tp x K x B column-major = [B][K][Tp] */
  fwrite(REAL(VECTOR_ELT(p1, 1)), sizeof(double), (size_t)K * N, o);
  fwrite(REAL(VECTOR_ELT(p1, 2)), sizeof(double), (size_t)K * N, o);
  /* logL_GP_mod / gr_GP_mod on explicit arguments: mean process 0 of the E-step, prior mean m_k[0] */
  SEXP gm = mcbR_gp_mod(h, reals(grid, N), reals(REAL(VECTOR_ELT(res, 0)), N), reals(mk, 1), reals(REAL(VECTOR_ELT(res, 1)), (long)N * N),
                        reals(th3, 3), Rf_ScalarReal(0.0), Rf_ScalarLogical(1));
  fwrite(REAL(gm), sizeof(double), 4, o);
  /* a list_mu that does not come from this handle's posterior_mu_k (MC.R:305): upload, then the same responsibilities */
  mcbR_set_pred_posterior(h, reals(grid, N), VECTOR_ELT(pm, 0), VECTOR_ELT(pm, 1), reals(pi, K));
  SEXP t1 = mcbR_predict(h, ints(o2, 2), ints(gi, n0), reals(y, n0), reals(th3, 3), reals(pi, K), ints(all, 0), Rf_ScalarInteger(K));
  fwrite(REAL(VECTOR_ELT(t1, 0)), sizeof(double), K, o);
  /* train_new_gp_EM with hp_i = NULL (MC.R:142-186) */
  SEXP tn = mcbR_train_new_indiv(h, ints(o2, 2), ints(gi, n0), reals(y, n0), reals(th3, 3), reals(pi, K), Rf_ScalarInteger(25));
  fwrite(REAL(VECTOR_ELT(tn, 0)), sizeof(double), 3, o);
  fwrite(REAL(VECTOR_ELT(tn, 1)), sizeof(double), K, o);
  free(all);
  fclose(o);
  if (h->finalizer) h->finalizer(h);
  if (mcb_stub_protect_depth != 0) return 15;
  printf("estep ok\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "abi")) return abi();
  if (argc >= 4 && !strcmp(argv[1], "estep")) return estep(argv[2], argv[3]);
  fprintf(stderr, "usage: drive_shim abi | estep <in.bin> <out.bin>\n");
  return 1;
}
