/* mcb.h — C ABI of libmagmaclust_b200.so: the B200 (sm_100a) implementation of MAGMAclust's
 * variational-EM hot path. Plain C, plain pointers and sizes; no torch / R types cross this boundary.
 *
 * The reference (ArthurLeroy/MAGMAclust) has no native interface at all: its "operator API" is a set of R
 * closures in a flat namespace. Each entry point below names the R function(s) it replaces
 * (CF.R = Computing_functions.R, MC.R = MAGMAclust.R, line numbers into /root/reference). INTEGRATION.md
 * shows the `.Call` shim and the ctypes binding over exactly these symbols.
 *
 * Conventions
 *   - All reals are FP64. All matrices are dense, row-major, and symmetric where the maths says so (so
 *     R's column-major view is identical).
 *   - The training set is canonicalised ONCE into a CSR-like layout (replaces the tibble `db` with columns
 *     ID, Timestamp, Output and the string-named matrices of CF.R:83,146):
 *         grid[N]        sorted unique timestamps over all individuals  (all_t, CF.R:601)
 *         offsets[M+1]   rows of individual i are [offsets[i], offsets[i+1])
 *         grid_idx[P]    position of each row's timestamp in grid[]      (match(Timestamp, all_t) - 1)
 *         y[P]           outputs
 *     Rows of one individual must be sorted by timestamp and unique (the reference silently assumes it,
 *     SURVEY.md §8.0 Q2).
 *   - Individuals are positional (order of unique(db$ID)); clusters are positional (order of names(m_k)).
 *   - theta_i is [M][3] = (log variance, log lengthscale^2, noise sd), theta_k is [K][3] with theta_k[.][2]
 *     the jitter sd ("pen_diag", CF.R:662); tau is [M][K] (row i = individual).
 *   - Return value: 0 ok; <0 invalid argument / state; >0 CUDA, NCCL or numerical failure. Nothing throws
 *     or exits across the ABI; mcb_last_error() gives the text. Where the reference silently switches to
 *     MASS::ginv on a singular matrix (CF.R:142,177,612), this library reports MCB_ENOTPD instead.
 *   - There is no CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef MCB_H
#define MCB_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcb_handle mcb_handle;

enum {
  MCB_OK = 0,
  MCB_EINVAL = -1,   /* bad argument */
  MCB_ESTATE = -2,   /* call sequence error (e.g. E-step before data upload) */
  MCB_ECUDA = 1,     /* CUDA runtime error */
  MCB_ENOTPD = 2,    /* a covariance matrix was not positive definite (non-positive pivot) */
  MCB_ENCCL = 3,     /* NCCL error / NCCL not loadable */
  MCB_ENOMEM = 4
};

/* ---- lifetime ------------------------------------------------------------------------------------- */
int mcb_version(void);
int mcb_device_count(void);                       /* 0 when no CUDA device is usable */
int mcb_create(mcb_handle** out, int device);     /* binds the handle to one GPU, creates its streams */
void mcb_destroy(mcb_handle* h);
const char* mcb_last_error(const mcb_handle* h);  /* NULL handle -> last creation error */

/* ---- multi-GPU: individuals are sharded over ranks; grid accumulators are all-reduced (NCCL) -------- */
/* One process per GPU. Rank 0 calls mcb_comm_unique_id and ships the 128 bytes to the other ranks by any
 * out-of-band channel; every rank then calls mcb_comm_init. After that, every mcb_* call that contains a
 * reduction (E-step accumulators, common-hp objectives, ELBO, pi_k) is collective over the ranks. */
int mcb_comm_unique_id(unsigned char out_id[128]);
int mcb_comm_init(mcb_handle* h, int nranks, int rank, const unsigned char id[128]);

/* ---- data ----------------------------------------------------------------------------------------- */
/* Upload the (local shard of the) training set. M_total = number of individuals over all ranks (== M when
 * single-GPU); it is the denominator of pi_k = mean_i tau_ik (CF.R:682). The call invalidates the state (call
 * mcb_set_state next). A dataset of the same shape (M, N, P, longest individual, M_total) as the previous one re-uses
 * the device buffers and the captured optimiser graphs, so a loop that re-uploads every iteration stays cheap. */
int mcb_set_data(mcb_handle* h, int M, int N, long long P, const int* offsets, const int* grid_idx,
                 const double* y, const double* grid, long long M_total);

/* ---- model state (device resident between calls) --------------------------------------------------- */
/* m_k: prior means of the mean processes, m_k_len == K (scalar per cluster, CF.R:718) or K*N. */
int mcb_set_state(mcb_handle* h, int K, const double* theta_k, const double* theta_i, const double* pi_k,
                  const double* m_k, int m_k_len, const double* tau);
int mcb_get_hp(mcb_handle* h, double* theta_k, double* theta_i, double* pi_k);
int mcb_get_tau(mcb_handle* h, double* tau);             /* the tau the next E-step will use as "old" */

/* ---- E-step: e_step_VEM (CF.R:589-629) = kern_to_inv (CF.R:123-191) + update_inv_VEM (CF.R:690-707)
 *      + update_mean_VEM (CF.R:709-731) + solve (CF.R:612) + update_tau_i_k_VEM (CF.R:733-762) --------- */
/* Runs on the resident state. Results stay on the device (hyper-posterior mean[K][N], cov[K][N][N], new
 * tau[M][K], log-likelihood matrix mat_logL[M][K]) and are fetched with the getters below. */
int mcb_e_step(mcb_handle* h);
int mcb_get_mean(mcb_handle* h, double* mean /* [K][N] */);
int mcb_get_cov(mcb_handle* h, int k, double* cov /* [N][N] */);
int mcb_get_tau_new(mcb_handle* h, double* tau /* [M][K] */);
int mcb_get_mat_logL(mcb_handle* h, double* logL /* [M][K] */);
int mcb_get_logdet_cov(mcb_handle* h, double* logdet /* [K] : log det C_k (finite; see MC.R:53 / Q8) */);
/* tau <- tau_new (MC.R:90) */
int mcb_accept_tau(mcb_handle* h);

/* Host-buffer convenience = one call per reference call; what the `.Call` shim binds for e_step_VEM.
 * Any output pointer may be NULL. cov is [K][N][N]. */
int mcb_e_step_VEM(mcb_handle* h, int K, const double* theta_k, const double* theta_i, const double* pi_k,
                   const double* m_k, int m_k_len, const double* old_tau, double* out_mean, double* out_cov,
                   double* out_tau, double* out_mat_logL);

/* Install an explicit hyper-posterior — the `mu_k_param` / `list_mu_param` / `new_cov` arguments of the
 * reference's callbacks (CF.R:218, 280, 327, 631) — instead of the one left by the last E-step:
 * mean[K][N], cov[K][N][N], tau[M][K] (host). Evaluated against the resident hp.
 * tau == NULL: the responsibilities are computed from (mean, cov, resident hp and pi_k) instead, i.e.
 * update_tau_i_k_VEM (CF.R:733-762); read them with mcb_get_tau_new, the K x M log-likelihoods with mcb_get_mat_logL. */
int mcb_set_posterior(mcb_handle* h, const double* mean, const double* cov, const double* tau);

/* ---- M-step objectives and gradients (the optimiser callbacks of CF.R:647-677) ----------------------- */
/* All evaluate against the hyper-posterior (mean, cov, tau_new) left on the device by the last E-step.
 * Individuals: F_i(theta) of logL_clust_multi_GP (CF.R:218-248) and its gradient gr_clust_multi_GP
 * (CF.R:477-508). theta is [M][3] (one row per individual) -> f[M], g[M][3]; g may be NULL. */
int mcb_eval_indiv(mcb_handle* h, const double* theta /* [M][3] */, double* f, double* g);
/* Shared theta_i: logL_clust_multi_GP_common_hp_i (CF.R:280-325) / gr_..._common_hp_i (CF.R:539-586):
 * sum over ALL individuals (all ranks) at one theta[3] -> f[1], g[3]. */
int mcb_eval_indiv_common(mcb_handle* h, const double* theta /* [3] */, double* f, double* g);
/* Mean processes: logL_GP_mod (CF.R:194-216) / gr_GP_mod (CF.R:448-475) for cluster k with nhp = 2
 * (theta = (a, b), sigma = pen_diag) or nhp = 3 (theta = (a, b, sigma)); g has nhp entries. */
int mcb_eval_mean(mcb_handle* h, int k, const double* theta, int nhp, double pen_diag, double* f, double* g);
/* Shared theta_k: logL_GP_mod_common_hp_k (CF.R:250-278) / gr_GP_mod_common_hp_k (CF.R:510-537). */
int mcb_eval_mean_common(mcb_handle* h, const double* theta, int nhp, double pen_diag, double* f, double* g);

/* ---- ELBO monitor: logL_monitoring_VEM (CF.R:327-352) + 0.5*(K*N + sum_k log det C_k) (MC.R:51-53) ----- */
/* Evaluated at the given hp against the resident hyper-posterior. out[0] = logL_monitoring_VEM value,
 * out[1] = the MC.R:51-53 total (finite log-determinants). theta_* NULL -> use the resident hp. */
int mcb_elbo(mcb_handle* h, const double* theta_k, const double* theta_i, double out[2]);

/* ---- BIC (CF.R:354-432, clust = TRUE; "next" row of SURVEY 8f) --------------------------------------- */
/* Terms of the reference's BIC at the resident hp against the resident hyper-posterior (mcb_set_state +
 * mcb_set_posterior install both): out[0] = sum_i sum_k tau_ik (l_ik + log(pi_k / tau_ik)) (CF.R:366-399),
 * out[1] = sum_k [dmvnorm(mean_k; m_1, K_k) - 0.5 sum(K_k^-1 o C_k)] (CF.R:409-414), out[2] = sum_k log det C_k
 * (CF.R:415). The host adds 0.5 K (N + N log 2 pi) and the penalty (counts of distinct hp values). */
int mcb_bic_terms(mcb_handle* h, double out[3]);

/* ---- built-in M-step: m_step_VEM (CF.R:631-687) with the L-BFGS of csrc/lbfgs.h -------------------- */
/* Optimises theta_i (one batched on-device L-BFGS per individual, or one shared), then theta_k, then
 * pi_k = mean_i tau_ik. Writes the NEW hp to new_* (host, may be NULL) and keeps them as "candidate" on the
 * device; mcb_accept_hp() makes them current (MC.R:85,89). stats[0..3] = #objective evaluations for
 * theta_i (max over individuals), theta_k, and the two optimisers' iteration counts. */
int mcb_m_step(mcb_handle* h, int common_hp_k, int common_hp_i, double* new_theta_k, double* new_theta_i,
               double* new_pi_k, int* stats);
int mcb_accept_hp(mcb_handle* h);
/* the last M-step's candidate (training_VEM returns hp = new_hp even when it was not accepted, MC.R:95) */
int mcb_get_new_hp(mcb_handle* h, double* theta_k, double* theta_i, double* pi_k);

/* ---- one VEM iteration = the loop body of training_VEM (MC.R:44-93) ---------------------------------- */
/* E-step, ELBO(hp), M-step, eps = (L(new_hp) - L(hp)) / |L(new_hp)|, acceptance rule of MC.R:83-90.
 * out[0] = ELBO total appended to logL_iterations (MC.R:77), out[1] = eps, out[2] = 0 go on, 1 converged (MC.R:83-87),
 * 2 the M-step's candidate hp are unusable (a covariance at them is not positive definite; eps = -inf): the reference's
 * M-step failure (MC.R:62-67) -- stop WITHOUT appending out[0], convergence stays FALSE, the last valid hp are kept. */
int mcb_vem_iteration(mcb_handle* h, int common_hp_k, int common_hp_i, double out[3]);
/* Host-buffer form of the same call (what the `.Call` shim binds for the loop body of training_VEM): additionally delivers
 * `param` of this iteration's E-step (MC.R:48) to caller-owned host buffers: out_mean[K][N], out_cov[K][N][N],
 * out_tau[M][K] (any may be NULL). The copies start as soon as the E-step has finished, on their own stream, and run
 * behind the M-step; everything has landed when the call returns. */
int mcb_vem_iteration_host(mcb_handle* h, int common_hp_k, int common_hp_i, double out[3], double* out_mean,
                           double* out_cov, double* out_tau);

/* ---- prediction -------------------------------------------------------------------------------------- */
/* posterior_mu_k (MC.R:196-233): hyper-posterior on a prediction grid `timestamps[T]` (sorted, must
 * contain every training timestamp) with theta_k[.][2] := 0.1 (MC.R:206), the resident hp and a frozen
 * tau[M][K] (NULL: param$tau_i_k of the last E-step, MC.R:208). The result stays resident as the
 * "prediction posterior" (mean_p[K][T], cov_p[K][T][T]); out_mean / out_cov may be NULL. */
int mcb_posterior_mu_k(mcb_handle* h, int T, const double* timestamps, const double* tau, double* out_mean,
                       double* out_cov);
/* Install an explicit prediction posterior instead: the `list_mu` argument of pred_gp_clust (MC.R:235) and the `mean_k`,
 * `cov_k` arguments of update_tau_star_k_EM (CF.R:764) when they do NOT come from this handle's last mcb_posterior_mu_k
 * (full_algo_clust(mu_k = ...), MC.R:305; a model read from disk). timestamps[T] sorted and unique, mean[K][T],
 * cov[K][T][T]; pi_k[K] = default mixing proportions of mcb_predict / mcb_train_new_indiv (NULL: 1/K). Needs no training
 * set on the handle. */
int mcb_set_pred_posterior(mcb_handle* h, int K, int T, const double* timestamps, const double* mean, const double* cov,
                           const double* pi_k);
/* Batched new individuals against the prediction posterior:
 *   update_tau_star_k_EM (CF.R:764-786) -> out_tau[B][K]
 *   pred_gp_clust (MC.R:235-274)        -> out_mean[B][K][Tp], out_var[B][K][Tp]
 * new individuals in CSR (offsets[B+1], obs_idx = positions in the prediction grid, y), shared theta_new[3]
 * (train_new_gp_EM fixed-hp branch, MC.R:187-191), targets pred_idx[Tp] = positions in the prediction grid.
 * pi_k[K]: mixing proportions (NULL: mean_i tau_ik of the tau frozen in mcb_posterior_mu_k, MC.R:141).
 * out_mean/out_var may both be NULL (tau only). Any number of observations per new individual (above 64, or above the shared-memory budget, a global-memory version of the kernel runs). */
int mcb_predict(mcb_handle* h, int B, const int* offsets, const int* obs_idx, const double* y,
                const double* theta_new, const double* pi_k, int Tp, const int* pred_idx, double* out_tau,
                double* out_mean, double* out_var);

/* New individuals with FREE hyper-parameters: the free-hp branch of train_new_gp_EM (MC.R:142-186), repaired (upstream it
 * cannot run: undefined `t`, stray `db$Timestamp -` in the objective). Batched over B new individuals against
 * the prediction posterior of mcb_posterior_mu_k.
 *
 * mcb_new_indiv_logl: the building block. For npts evaluation points (pt_ind[p] = new individual, pt_theta[p][3] = its
 * trial hyper-parameters) out_logl[p][K] = log N(y*; m_k(t*), Psi_theta(t*) + C_k(t*, t*)) — dmvnorm(..., log = T) of
 * CF.R:776-780 and MC.R:156-161. NaN where that covariance is not numerically positive definite.
 *
 * mcb_train_new_indiv: the EM loop (at most n_loop_max passes, MC.R:143: 25). E step: update_tau_star_k_EM at the current
 * theta; M step: L-BFGS on - sum_k tau_k log N(...) with optim's central-difference gradient (no `gr` at MC.R:166,
 * ndeps = 1e-3); stop when 0 < sum |theta' - theta| < 1e-3 (MC.R:175-181). theta0[3], or [B][3] if theta0_per_indiv;
 * pi_k NULL: as in mcb_predict. out_theta[B][3] = theta' of the last M step, out_tau[B][K] = tau_k of the last E step
 * (MC.R:192), out_loops[B] (may be NULL) = passes done, -1 if the start point already gave a non-positive-definite
 * covariance (out_tau NaN then).
 *
 * mcb_new_indiv_em is the same loop as host code over a caller-supplied log-density (what mcb_train_new_indiv runs with the
 * GPU kernel behind `fn`): fn(ctx, npts, pt_ind, pt_theta, out_logl) fills out_logl[npts][K] and returns MCB_OK. */
typedef int (*mcb_logl_fn)(void* ctx, int npts, const int* pt_ind, const double* pt_theta, double* out_logl);
int mcb_new_indiv_logl(mcb_handle* h, int B, const int* offsets, const int* obs_idx, const double* y, int npts,
                       const int* pt_ind, const double* pt_theta, double* out_logl);
int mcb_train_new_indiv(mcb_handle* h, int B, const int* offsets, const int* obs_idx, const double* y,
                        const double* theta0, int theta0_per_indiv, const double* pi_k, int n_loop_max,
                        double* out_theta, double* out_tau, int* out_loops);
int mcb_new_indiv_em(int B, int K, const double* theta0, int theta0_per_indiv, const double* pi_k, int n_loop_max,
                     mcb_logl_fn fn, void* ctx, double* out_theta, double* out_tau, int* out_loops);

/* ---- initialisation -------------------------------------------------------------------------------------- */
/* ini_kmeans (MAGMAclust.R:562-598): k-means of the feature rows feat[M][D] (one per individual: the outputs on a common
 * grid, or (min, mean, max), MAGMAclust.R:569-580) from `nstart` starts, start s taking the rows start_rows[s][0..K) as
 * initial centres (kmeans(centers = k) draws k distinct rows). Lloyd iterations (at most max_iter assignment passes) to a
 * fixed point, all starts side by side on the device; the start with the smallest total within-cluster sum of squares wins
 * (the first one on ties). out_labels[M] in [0, K); out_centers[K][D], out_cost, out_iters may be NULL. Every floating-point
 * operation has a prescribed order (csrc/kmeans.cu), so the result is reproducible bit for bit (stats::kmeans itself is
 * Hartigan-Wong on R's RNG: unpinned). Needs no training set on the handle. */
int mcb_kmeans(mcb_handle* h, int M, int D, const double* feat, int K, int nstart, const int* start_rows, int max_iter,
               int* out_labels, double* out_centers, double* out_cost, int* out_iters);

/* ---- stand-alone building blocks (the vector branches the reference also calls directly, e.g. MC.R:261-264) ------------ */
/* kern_to_cov(x, kern, theta, sigma), vector branch CF.R:68-86: SE covariance of the SORTED timestamps x[n] (the reference
 * sorts them, CF.R:76; pass them sorted), theta2 = (a, b): exp(a - exp(-b)/2 d^2) off the diagonal, exp(a) + sigma^2 on it.
 * out: n x n row-major. */
int mcb_kern_to_cov(mcb_handle* h, int n, const double* x, const double* theta2, double sigma, double* out);
/* kern_to_inv(x, kern, theta, sigma), vector branch CF.R:130-149: inverse of the matrix above (+ log det of the covariance in
 * *logdet when not NULL). MCB_ENOTPD where the reference would fall over to MASS::ginv. */
int mcb_kern_to_inv(mcb_handle* h, int n, const double* x, const double* theta2, double sigma, double* out, double* logdet);
/* dmvnorm(x, mu, inv_Sigma, log) CF.R:10-35: Gaussian (log-)density from a symmetric positive definite PRECISION matrix
 * inv_Sigma (n x n row-major): -(n log 2pi - log det inv_Sigma + z' inv_Sigma z) / 2, z = x - mu. */
int mcb_dmvnorm(mcb_handle* h, int n, const double* x, const double* mu, const double* inv_Sigma, int want_log, double* out);

/* logL_GP_mod / gr_GP_mod (CF.R:194-216, 448-475; Z = 1) and logL_GP_mod_common_hp_k / gr_GP_mod_common_hp_k (CF.R:250-278,
 * 510-537; Z clusters sharing one theta) on EXPLICIT arguments (hp, db, mean, kern, new_cov, pen_diag) instead of the
 * resident hyper-posterior: t[n] sorted timestamps (db$Timestamp), y[Z][n] (db$Output per cluster), mean with mean_len == Z
 * (scalars, CF.R:206) or Z*n entries, new_cov[Z][n][n]; theta[nhp], nhp == 3: sigma = theta[2], nhp == 2: sigma = pen_diag
 * (CF.R:207). f[1]; g[nhp] may be NULL. The resident forms above (mcb_eval_mean, mcb_eval_mean_common) are the fast path
 * of the M-step; this one serves callers that pass other data (as update_tau_i_k_VEM does, CF.R:746). */
int mcb_gp_mod(mcb_handle* h, int n, int Z, const double* t, const double* y, const double* mean, int mean_len,
               const double* new_cov, const double* theta, int nhp, double pen_diag, double* f, double* g);

/* ---- measurement helpers ----------------------------------------------------------------------------- */
/* Device time (ms, CUDA events on the handle's stream) of the kernels launched by the last compute call,
 * per phase: [0] per-individual E-step kernel, [1] dense hyper-posterior, [2] tau kernel, [3] M-step
 * individuals, [4] M-step mean processes, [5] ELBO, [6] prediction, [7] collectives. Also the number of
 * kernel launches of the last call. */
int mcb_last_timings(mcb_handle* h, float ms[8], long long* launches);
/* Counters of the last M-step: out[0] max #evaluations per individual optimiser, out[1] #evaluations of the
 * theta_k optimiser(s), out[2], out[3] the iteration counts, out[4] total individual objective+gradient
 * evaluations on this rank, out[5] number of local individuals. */
int mcb_last_mstep_stats(mcb_handle* h, long long out[6]);
/* Per-individual outcome of the last M-step with individual-specific theta_i (one optimisation per individual, CF.R:653-659):
 * nfev[M] objective + gradient evaluations, niter[M] iterations, status[M] = 1 converged (optim: 0), 2 iteration limit
 * (optim: 1), 3 abnormal line search (optim: 52), 4 zero gradient, 5 objective not finite at the start point (the start
 * point is returned). Any pointer may be NULL. */
int mcb_last_mstep_indiv(mcb_handle* h, int* nfev, int* niter, int* status);
/* Page-lock / release a host buffer (cudaHostRegister) so that the copies of the host-buffer entry points run
 * from pinned memory. */
int mcb_pin_host(void* ptr, size_t bytes);
int mcb_unpin_host(void* ptr);
/* Page-locked host memory (cudaHostAlloc) for buffers that cross PCIe on every call, e.g. the results of mcb_predict. */
int mcb_host_alloc(void** out, size_t bytes);
int mcb_host_free(void* ptr);
/* CUDA-event stopwatch on the handle's stream (the stream every kernel of this library is launched on):
 * start synchronises and records; stop records, synchronises and returns the elapsed device time. */
int mcb_timer_start(mcb_handle* h);
int mcb_timer_stop(mcb_handle* h, float* ms);
/* Evict the working set from L2 between timed steps: writes a 256 MiB scratch buffer on the stream. */
int mcb_flush_l2(mcb_handle* h);
/* FP64 roofline denominators measured on this GPU: out[0] DFMA TFLOP/s, out[1] DMMA m8n8k4 TFLOP/s. */
int mcb_measure_fp64_peaks(mcb_handle* h, double out[2]);

#ifdef __cplusplus
}
#endif
#endif /* MCB_H */
