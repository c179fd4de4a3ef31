// C ABI of libmagmaclust_b200.so (see include/mcb.h): handle lifetime, data/state upload, and the host
// orchestration of the E-step (CF.R:589-629), the M-step objective/gradient evaluators (CF.R:194-325,
// 448-586), the ELBO monitor (CF.R:327-352, MC.R:51-53), the built-in M-step (CF.R:631-687) and one
// iteration of training_VEM's loop (MC.R:44-93).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "handle.h"
#include "lbfgs.h"

using namespace mcb;

static std::string g_create_err;

#define H_CUDA(expr) MCB_CUDA(h, expr)
static inline cudaError_t sync_copy(mcb_handle* h, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, h->st);
  return e != cudaSuccess ? e : cudaStreamSynchronize(h->st);
}
#define H_CHECK(cond, code, text) \
  do {                            \
    if (!(cond)) {                \
      h->err = (text);            \
      return (code);              \
    }                             \
  } while (0)

// ---- timing spans ------------------------------------------------------------------------------------
static void spans_reset(mcb_handle* h) {
  h->spans.clear();
  h->ev_next = 0;
  h->launches = 0;
}
static int ev_get(mcb_handle* h) {
  if (h->ev_next == (int)h->evpool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->evpool.push_back(e);
  }
  return h->ev_next++;
}
struct PhaseScope {
  mcb_handle* h;
  int phase, e0;
  cudaStream_t s;
  PhaseScope(mcb_handle* h_, int ph, cudaStream_t s_ = nullptr) : h(h_), phase(ph), s(s_ ? s_ : h_->st) {
    e0 = ev_get(h);
    cudaEventRecord(h->evpool[e0], s);
  }
  ~PhaseScope() {
    int e1 = ev_get(h);
    cudaEventRecord(h->evpool[e1], s);
    h->spans.push_back({phase, e0, e1});
  }
};

// ---- lifetime ----------------------------------------------------------------------------------------
extern "C" int mcb_version(void) { return 100; }

extern "C" int mcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int mcb_create(mcb_handle** out, int device) {
  if (!out) return MCB_EINVAL;
  *out = nullptr;
  int n = mcb_device_count();
  if (n <= 0) {
    g_create_err = "no CUDA device available (this library has no CPU fallback)";
    return MCB_ECUDA;
  }
  if (device < 0 || device >= n) {
    g_create_err = "device index out of range";
    return MCB_EINVAL;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return MCB_ECUDA;
  }
  mcb_handle* h = new mcb_handle();
  h->device = device;
  if ((e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = [&] { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);
                 return cudaStreamCreateWithPriority(&h->st_hi, cudaStreamNonBlocking, hi); }()) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming)) != cudaSuccess ||
      (e = h->mstep_ctl.ensure(mcb::kMstepCtlWords)) != cudaSuccess ||
      (e = cudaMallocHost(&h->hbuf, sizeof(double) * mcb_handle::kHbuf)) != cudaSuccess ||
      (e = h->info.ensure(4)) != cudaSuccess || (e = h->scal.ensure(64)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    delete h;
    return MCB_ECUDA;
  }
  cudaMemset(h->info.p, 0, sizeof(int) * 4);
  cudaDeviceSynchronize();   // default-stream fill vs the handle's non-blocking streams
  *out = h;
  return MCB_OK;
}

extern "C" void mcb_destroy(mcb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->st);
  comm_destroy(h);
  for (auto& kv : h->eval_graphs) cudaGraphExecDestroy(kv.second.exec);
  if (h->mopt_exec) cudaGraphExecDestroy(h->mopt_exec);
  if (h->st_cap) cudaStreamDestroy(h->st_cap);
  if (h->st_hi) cudaStreamDestroy(h->st_hi);
  if (h->st_copy) { cudaStreamDestroy(h->st_copy); cudaEventDestroy(h->ev_copy0); cudaEventDestroy(h->ev_copy1); }
  h->mopt_state.release();
  for (auto e : h->evpool) cudaEventDestroy(e);
  mcb::DevBuf<double>* dbl[] = {&h->t, &h->y, &h->grid, &h->theta_i, &h->theta_k, &h->pi, &h->mk, &h->tau,
                                &h->theta_i_new, &h->theta_k_new, &h->pi_new, &h->tau_new, &h->logL, &h->Winv,
                                &h->pvec, &h->yWy, &h->logdet_i, &h->c1, &h->c2, &h->Fi, &h->Fnew, &h->kscr, &h->fi_tmp, &h->gi_tmp,
                                &h->theta_g, &h->Kinv, &h->cov, &h->wk, &h->mean, &h->rvec, &h->pz, &h->logdetK,
                                &h->logdetA, &h->eK, &h->eD1, &h->eD2, &h->eCsum, &h->eX, &h->etheta, &h->ered,
                                &h->elogdet, &h->ws.Pbuf, &h->ws.Cold, &h->ws.red, &h->ws.slabV, &h->ws.slabP, &h->ws.slabL,
                                &h->ws_aux.Pbuf, &h->ws_aux.Cold, &h->ws_aux.red, &h->ws_aux.slabV, &h->ws_aux.slabP, &h->ws_aux.slabL, &h->scal, &h->pgrid, &h->pKinv,
                                &h->pcov, &h->pwk, &h->pmean, &h->plogdet, &h->gscr};
  for (auto* b : dbl) b->release();
  mcb::DevBuf<int>* ints[] = {&h->off, &h->gidx, &h->nfev_i, &h->niter_i, &h->status_i, &h->gmap_d, &h->egmap,
                              &h->info, &h->pmapidx, &h->gscr_i};
  for (auto* b : ints) b->release();
  h->woff.release();
  h->ws.bar.release();
  h->ws_aux.bar.release();
  h->flush.release();
  if (h->tm0) { cudaEventDestroy(h->tm0); cudaEventDestroy(h->tm1); }
  if (h->hbuf) cudaFreeHost(h->hbuf);
  if (h->st2) { cudaStreamSynchronize(h->st2); cudaStreamDestroy(h->st2); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  h->mstep_ctl.release();
  h->order_i.release();
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
}

extern "C" const char* mcb_last_error(const mcb_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

// ---- data --------------------------------------------------------------------------------------------
extern "C" int mcb_set_data(mcb_handle* h, int M, int N, long long P, const int* offsets, const int* grid_idx,
                            const double* y, const double* grid, long long M_total) {
  if (!h) return MCB_EINVAL;
  H_CHECK(M > 0 && N > 0 && P > 0 && offsets && grid_idx && y && grid, MCB_EINVAL, "mcb_set_data: bad argument");
  H_CHECK(offsets[0] == 0 && offsets[M] == P, MCB_EINVAL, "mcb_set_data: offsets must span [0, P]");
  H_CHECK(M_total >= M, MCB_EINVAL, "mcb_set_data: M_total < M");
  for (int j = 1; j < N; ++j) H_CHECK(grid[j] > grid[j - 1], MCB_EINVAL, "mcb_set_data: grid must be sorted and unique");
  std::vector<long long> woff(M + 1, 0);
  std::vector<double> t((size_t)P);
  int nmax = 0, nmax_all = 0;
  std::vector<int> big_ids;
  std::vector<long long> big_wsoff(1, 0);
  for (int i = 0; i < M; ++i) {
    const int n = offsets[i + 1] - offsets[i];
    H_CHECK(n >= 1, MCB_EINVAL, "mcb_set_data: empty individual");
    nmax_all = std::max(nmax_all, n);
    if (n <= mcb::kSmallMax) {
      nmax = std::max(nmax, n);
    } else {
      // no size limit in the reference (CF.R:153-190): above the tile buckets of the shared-memory kernels an individual
      // takes the global-memory kernels of indiv_big.cu
      big_ids.push_back(i);
      big_wsoff.push_back(big_wsoff.back() + (long long)(n + 2) * (n + 2) + (long long)n * n);
    }
    woff[i + 1] = woff[i] + (long long)n * n;
    for (int a = offsets[i]; a < offsets[i + 1]; ++a) {
      H_CHECK(grid_idx[a] >= 0 && grid_idx[a] < N, MCB_EINVAL, "mcb_set_data: grid_idx out of range");
      H_CHECK(a == offsets[i] || grid_idx[a] > grid_idx[a - 1], MCB_EINVAL,
              "mcb_set_data: rows of an individual must be sorted by timestamp and unique");
      t[a] = grid[grid_idx[a]];
    }
  }
  if (nmax == 0) nmax = 1;   // every individual is on the global-memory path: the shared-memory kernels are not launched
  H_CHECK(eval_smem_bytes(nmax) <= 227 * 1024, MCB_EINVAL, "mcb_set_data: shared-memory budget of the evaluators exceeded");
  H_CUDA(cudaSetDevice(h->device));
  // a dataset of another shape forces the state buffers to be re-sized and the captured graphs (which hold M, N, P and
  // the buffer addresses) to be rebuilt; re-uploading a dataset of the SAME shape (the host-buffer loop) keeps both
  const bool same_shape = M == h->M && N == h->N && P == h->P && nmax == h->nmax && M_total == h->M_total &&
                          big_ids == h->h_big_ids && nmax_all == h->nmax_all;
  if (!same_shape) h->K = 0;
  h->M = M; h->N = N; h->ld = (N + 15) & ~15; h->P = P; h->nmax = nmax; h->M_total = M_total; h->W2 = woff[M];
  h->nmax_all = nmax_all; h->nbig = (int)big_ids.size(); h->h_big_ids = big_ids;
  if (h->nbig > 0) {
    const size_t nb = big_ids.size();
    H_CUDA(h->big_ids.ensure(nb)); H_CUDA(h->big_wsoff.ensure(nb + 1)); H_CUDA(h->big_ws.ensure((size_t)big_wsoff.back()));
    H_CUDA(h->big_active.ensure(nb)); H_CUDA(h->big_theta.ensure(3 * nb)); H_CUDA(h->big_f.ensure(nb)); H_CUDA(h->big_g.ensure(3 * nb));
    H_CUDA(cudaMemcpyAsync(h->big_ids.p, big_ids.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->big_wsoff.p, big_wsoff.data(), sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, h->st));
  }
  H_CUDA(h->off.ensure(M + 1)); H_CUDA(h->gidx.ensure(P)); H_CUDA(h->woff.ensure(M + 1));
  H_CUDA(h->t.ensure(P)); H_CUDA(h->y.ensure(P)); H_CUDA(h->grid.ensure(N));
  H_CUDA(cudaMemcpyAsync(h->off.p, offsets, sizeof(int) * (M + 1), cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->gidx.p, grid_idx, sizeof(int) * P, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->woff.p, woff.data(), sizeof(long long) * (M + 1), cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->t.p, t.data(), sizeof(double) * P, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->y.p, y, sizeof(double) * P, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->grid.p, grid, sizeof(double) * N, cudaMemcpyHostToDevice, h->st));
  H_CUDA(h->Winv.ensure(h->W2)); H_CUDA(h->c2.ensure(h->W2));
  H_CUDA(h->pvec.ensure(P)); H_CUDA(h->c1.ensure(P));
  H_CUDA(h->yWy.ensure(M)); H_CUDA(h->logdet_i.ensure(M)); H_CUDA(h->Fi.ensure(M));
  H_CUDA(h->fi_tmp.ensure(M)); H_CUDA(h->gi_tmp.ensure(3 * (size_t)M));
  H_CUDA(h->nfev_i.ensure(M)); H_CUDA(h->niter_i.ensure(M)); H_CUDA(h->status_i.ensure(M)); H_CUDA(h->Fnew.ensure(M));
  H_CUDA(h->theta_i.ensure(3 * (size_t)M)); H_CUDA(h->theta_i_new.ensure(3 * (size_t)M));
  H_CUDA(cudaStreamSynchronize(h->st));  // host vectors go out of scope
  h->has_data = true; h->has_state = false; h->has_post = false; h->has_cand = false; h->has_pred = false;
  // the evaluation counts of the last M-step only ORDER the work queue of the next one (never its results): they stay a
  // useful prediction for a re-uploaded dataset of the same shape
  if (!same_shape) h->nfev_valid = false;
  return MCB_OK;
}

// ---- state -------------------------------------------------------------------------------------------
static void drop_graphs(mcb_handle* h) {
  for (auto& kv : h->eval_graphs) cudaGraphExecDestroy(kv.second.exec);
  h->eval_graphs.clear();
  if (h->mopt_exec) cudaGraphExecDestroy(h->mopt_exec);
  h->mopt_exec = nullptr;
  if (h->iopt_exec) cudaGraphExecDestroy(h->iopt_exec);
  h->iopt_exec = nullptr;
}

static int alloc_state(mcb_handle* h, int K) {
  drop_graphs(h);
  const size_t KN = (size_t)K * h->ld, KNN = (size_t)K * h->N * h->ld;
  H_CUDA(h->theta_k.ensure(3 * K)); H_CUDA(h->theta_k_new.ensure(3 * K));
  H_CUDA(h->pi.ensure(K)); H_CUDA(h->pi_new.ensure(K));
  H_CUDA(h->mk.ensure(KN)); H_CUDA(h->wk.ensure(KN)); H_CUDA(h->mean.ensure(KN));
  H_CUDA(h->rvec.ensure(KN)); H_CUDA(h->pz.ensure(KN));
  H_CUDA(h->tau.ensure((size_t)h->M * K)); H_CUDA(h->tau_new.ensure((size_t)h->M * K));
  H_CUDA(h->logL.ensure((size_t)h->M * K));
  H_CUDA(h->Kinv.ensure(KNN)); H_CUDA(h->cov.ensure(KNN));
  H_CUDA(h->logdetK.ensure(K)); H_CUDA(h->logdetA.ensure(K));
  H_CUDA(h->gmap_d.ensure(K)); H_CUDA(h->theta_g.ensure(3 * K));
  H_CUDA(h->eK.ensure(KNN)); H_CUDA(h->eD1.ensure(KNN)); H_CUDA(h->eD2.ensure(KNN));
  H_CUDA(h->eCsum.ensure(KNN)); H_CUDA(h->eX.ensure(KNN));
  H_CUDA(h->etheta.ensure(3 * K)); H_CUDA(h->ered.ensure(16 * (size_t)K)); H_CUDA(h->elogdet.ensure(K));
  H_CUDA(h->egmap.ensure(K));
  H_CUDA(spd_inverse_reserve(h->ws, h->N, K));
  // the dense kernels rely on zero padding of every row up to ld: clear whatever an earlier shape left behind
  mcb::DevBuf<double>* dense[] = {&h->Kinv, &h->cov, &h->eK, &h->eD1, &h->eD2, &h->eCsum, &h->eX,
                                  &h->mk, &h->wk, &h->mean, &h->rvec, &h->pz};
  for (auto* b : dense) H_CUDA(cudaMemsetAsync(b->p, 0, sizeof(double) * b->cap, h->st));
  return MCB_OK;
}

extern "C" int mcb_set_state(mcb_handle* h, int K, const double* theta_k, const double* theta_i,
                             const double* pi_k, const double* m_k, int m_k_len, const double* tau) {
  if (!h) return MCB_EINVAL;
  H_CHECK(h->has_data, MCB_ESTATE, "mcb_set_state: call mcb_set_data first");
  H_CHECK(K >= 1 && theta_k && theta_i && pi_k && m_k && tau, MCB_EINVAL, "mcb_set_state: bad argument");
  H_CHECK(m_k_len == K || m_k_len == K * h->N, MCB_EINVAL, "mcb_set_state: m_k_len must be K or K*N");
  H_CUDA(cudaSetDevice(h->device));
  if (K != h->K) {
    MCB_TRY(alloc_state(h, K));
    h->K = K;
  }
  const int N = h->N, ld = h->ld, M = h->M;
  std::vector<double> mk((size_t)K * ld, 0.0);
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < N; ++j) mk[(size_t)k * ld + j] = (m_k_len == K) ? m_k[k] : m_k[(size_t)k * N + j];
  H_CUDA(cudaMemcpyAsync(h->theta_k.p, theta_k, sizeof(double) * 3 * K, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->theta_i.p, theta_i, sizeof(double) * 3 * M, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->pi.p, pi_k, sizeof(double) * K, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->mk.p, mk.data(), sizeof(double) * K * ld, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->tau.p, tau, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  h->h_theta_k.assign(theta_k, theta_k + 3 * K);
  h->h_pi.assign(pi_k, pi_k + K);
  h->has_state = true; h->has_post = false; h->has_cand = false;
  return MCB_OK;
}

extern "C" int mcb_get_hp(mcb_handle* h, double* theta_k, double* theta_i, double* pi_k) {
  if (!h) return MCB_EINVAL;
  H_CHECK(h->has_state, MCB_ESTATE, "mcb_get_hp: no state");
  H_CUDA(cudaSetDevice(h->device));
  if (theta_k) H_CUDA(cudaMemcpyAsync(theta_k, h->theta_k.p, sizeof(double) * 3 * h->K, cudaMemcpyDeviceToHost, h->st));
  if (theta_i) H_CUDA(cudaMemcpyAsync(theta_i, h->theta_i.p, sizeof(double) * 3 * h->M, cudaMemcpyDeviceToHost, h->st));
  if (pi_k) H_CUDA(cudaMemcpyAsync(pi_k, h->pi.p, sizeof(double) * h->K, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

extern "C" int mcb_get_new_hp(mcb_handle* h, double* theta_k, double* theta_i, double* pi_k) {
  if (!h) return MCB_EINVAL;
  H_CHECK(h->has_cand, MCB_ESTATE, "mcb_get_new_hp: no M-step candidate");
  H_CUDA(cudaSetDevice(h->device));
  if (theta_k) memcpy(theta_k, h->h_theta_k_new.data(), sizeof(double) * 3 * h->K);
  if (pi_k) memcpy(pi_k, h->h_pi_new.data(), sizeof(double) * h->K);
  if (theta_i) {
    H_CUDA(cudaMemcpyAsync(theta_i, h->theta_i_new.p, sizeof(double) * 3 * h->M, cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
  }
  return MCB_OK;
}

extern "C" int mcb_get_tau(mcb_handle* h, double* tau) {
  if (!h || !tau) return MCB_EINVAL;
  H_CHECK(h->has_state, MCB_ESTATE, "mcb_get_tau: no state");
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaMemcpyAsync(tau, h->tau.p, sizeof(double) * (size_t)h->M * h->K, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

// ---- helpers -----------------------------------------------------------------------------------------
// distinct rows of theta[K][3] -> gmap, returns G and fills groups[G][3]
static int group_theta(const double* theta, int K, std::vector<int>& gmap, std::vector<double>& groups) {
  gmap.assign(K, -1);
  groups.clear();
  int G = 0;
  for (int k = 0; k < K; ++k) {
    for (int g = 0; g < G; ++g)
      if (groups[3 * g] == theta[3 * k] && groups[3 * g + 1] == theta[3 * k + 1] && groups[3 * g + 2] == theta[3 * k + 2]) {
        gmap[k] = g;
        break;
      }
    if (gmap[k] < 0) {
      gmap[k] = G++;
      groups.insert(groups.end(), theta + 3 * k, theta + 3 * k + 3);
    }
  }
  return G;
}

// Multi-rank runs: the not-positive-definite flag is raised by kernels that work on rank-local individuals, so one rank could
// return MCB_ENOTPD alone while its peers wait in their next collective. The flag therefore travels as one more double in
// the all-reduce that precedes the check (info_to_double before it, info_from_double after it), or through its own 8-byte
// all-reduce where there is none (synced = false): every rank then takes the same branch.
__global__ void info_to_double_kernel(const int* info, double* d) { *d = (*info != 0) ? 1.0 : 0.0; }
__global__ void info_from_double_kernel(const double* d, int* info) { if (*d != 0.0) *info = 1; }
static void info_to_double(mcb_handle* h, double* d) {
  if (h->nranks > 1) info_to_double_kernel<<<1, 1, 0, h->st>>>(h->info.p, d);
}
static void info_from_double(mcb_handle* h, const double* d) {
  if (h->nranks > 1) info_from_double_kernel<<<1, 1, 0, h->st>>>(d, h->info.p);
}

static int check_info(mcb_handle* h, const char* where, bool synced = false) {
  if (h->nranks > 1 && !synced) {
    double* d = h->scal.p + 56;
    info_to_double(h, d);
    MCB_TRY(comm_allreduce_sum(h, d, 1));
    info_from_double(h, d);
  }
  int flag = 0;
  H_CUDA(cudaMemcpyAsync(&flag, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  if (flag) {
    cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st);
    h->err = std::string(where) + ": a covariance matrix is not positive definite (non-positive pivot); "
             "the reference would silently fall back to MASS::ginv here";
    return MCB_ENOTPD;
  }
  return MCB_OK;
}

// MCB_DEBUG=1: check the not-positive-definite flag after every stage (adds host syncs)
static int dbg_stage(mcb_handle* h, const char* stage) {
  static const bool on = getenv("MCB_DEBUG") != nullptr;
  return on ? check_info(h, stage) : MCB_OK;
}

// ---- E-step ------------------------------------------------------------------------------------------
static int e_step_impl(mcb_handle* h) {
  const int M = h->M, N = h->N, ld = h->ld, K = h->K;
  const long long sN = (long long)N * ld;
  std::vector<double> groups;
  h->G = group_theta(h->h_theta_k.data(), K, h->gmap, groups);
  const int G = h->G;
  H_CUDA(cudaMemcpyAsync(h->gmap_d.p, h->gmap.data(), sizeof(int) * K, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->theta_g.p, groups.data(), sizeof(double) * 3 * G, cudaMemcpyHostToDevice, h->st));
  {
    PhaseScope ps(h, PH_DENSE);
    // K_k^-1 for each distinct theta_k (CF.R:604)
    se_build(h->st, h->Kinv.p, nullptr, nullptr, N, ld, sN, h->grid.p, h->theta_g.p, 3, G, &h->launches);
    H_CUDA(spd_inverse(h->st, h->Kinv.p, N, ld, sN, G, h->logdetK.p, h->info.p, h->ws, &h->launches));
    MCB_TRY(dbg_stage(h, "e_step: solve(K_k)"));
    // A_k starts as K_k^-1 (rank 0 only when sharded: the all-reduce adds it exactly once), w_k as K_k^-1 m_k
    copy_mapped(h->st, N, ld, sN, h->Kinv.p, h->gmap_d.p, h->cov.p, K, h->rank == 0 ? 1.0 : 0.0, &h->launches);
    if (h->rank == 0) symv_mapped(h->st, N, ld, sN, h->Kinv.p, h->gmap_d.p, h->mk.p, h->wk.p, K, &h->launches);
    else H_CUDA(cudaMemsetAsync(h->wk.p, 0, sizeof(double) * K * ld, h->st));
  }
  {
    PhaseScope ps(h, PH_INDIV);
    H_CUDA(launch_factor_scatter(h->st, h->idata(), M, h->nmax, h->theta_i.p, h->tau.p, K, h->cov.p, ld, sN,
                                 h->wk.p, h->iwork(), h->info.p));
    ++h->launches;
    MCB_TRY(dbg_stage(h, "e_step: solve(Psi_i)"));
  }
  // Small grids: all-reduce A_k, w_k and replicate the K solves on every rank (latency bound, cheaper than a
  // second exchange). Large grids (N > 768): cluster k is OWNED by rank k % nranks: its A_k is reduced onto the
  // owner, inverted there only, and C_k / log|A_k| are broadcast back (one cluster per GPU at K = nranks).
  const bool owned = h->nranks > 1 && !spd_inverse_slab_ok(N, ld);
  if (h->nranks > 1) {
    PhaseScope ps(h, PH_COMM);
    if (owned) MCB_TRY(comm_reduce_to_owners(h, h->cov.p, (size_t)sN, sN, K));
    else MCB_TRY(comm_allreduce_sum(h, h->cov.p, (size_t)K * sN));
    MCB_TRY(comm_allreduce_sum(h, h->wk.p, (size_t)K * ld));
  }
  {
    PhaseScope ps(h, PH_DENSE);
    // C_k = A_k^-1 (CF.R:612), m_k = C_k w_k (CF.R:620)
    if (owned) {
      H_CUDA(cudaMemsetAsync(h->logdetA.p, 0, sizeof(double) * K, h->st));
      for (int k = h->rank; k < K; k += h->nranks)
        H_CUDA(spd_inverse(h->st, h->cov.p + (size_t)k * sN, N, ld, sN, 1, h->logdetA.p + k, h->info.p, h->ws,
                           &h->launches));
    } else {
      H_CUDA(spd_inverse(h->st, h->cov.p, N, ld, sN, K, h->logdetA.p, h->info.p, h->ws, &h->launches));
    }
  }
  if (owned) {
    PhaseScope ps(h, PH_COMM);
    MCB_TRY(comm_bcast_from_owners(h, h->cov.p, (size_t)sN, sN, K));
    MCB_TRY(comm_bcast_from_owners(h, h->logdetA.p, 1, 1, K));
  }
  {
    PhaseScope ps(h, PH_DENSE);
    symv(h->st, N, h->cov.p, ld, sN, h->wk.p, ld, h->mean.p, ld, K, &h->launches);
    vec_sub(h->st, K * ld, h->mean.p, h->mk.p, h->rvec.p, &h->launches);
  }
  {
    PhaseScope ps(h, PH_TAU);
    H_CUDA(launch_tau(h->st, h->idata(), M, h->nmax, K, h->mean.p, h->cov.p, ld, sN, h->pi.p, h->iwork(),
                      h->tau_new.p, h->logL.p, 0));
    ++h->launches;
  }
  H_CUDA(cudaMemcpyAsync(h->hbuf, h->logdetK.p, sizeof(double) * G, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaMemcpyAsync(h->hbuf + K, h->logdetA.p, sizeof(double) * K, cudaMemcpyDeviceToHost, h->st));
  MCB_TRY(check_info(h, "e_step_VEM"));
  h->h_logdetK.assign(h->hbuf, h->hbuf + G);
  h->h_logdetA.assign(h->hbuf + K, h->hbuf + 2 * K);
  h->has_post = true;
  h->csum_valid = false;
  return MCB_OK;
}

extern "C" int mcb_e_step(mcb_handle* h) {
  if (!h) return MCB_EINVAL;
  H_CHECK(h->has_state, MCB_ESTATE, "mcb_e_step: call mcb_set_data and mcb_set_state first");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  return e_step_impl(h);
}

static int copy_rows(mcb_handle* h, double* dst, const double* src, int rows, int N, int ld) {
  H_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * N, src, sizeof(double) * ld, sizeof(double) * N, rows,
                           cudaMemcpyDeviceToHost, h->st));
  return MCB_OK;
}

extern "C" int mcb_get_mean(mcb_handle* h, double* mean) {
  if (!h || !mean) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_get_mean: no E-step result");
  H_CUDA(cudaSetDevice(h->device));
  MCB_TRY(copy_rows(h, mean, h->mean.p, h->K, h->N, h->ld));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

extern "C" int mcb_get_cov(mcb_handle* h, int k, double* cov) {
  if (!h || !cov) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_get_cov: no E-step result");
  H_CHECK(k >= 0 && k < h->K, MCB_EINVAL, "mcb_get_cov: k out of range");
  H_CUDA(cudaSetDevice(h->device));
  MCB_TRY(copy_rows(h, cov, h->cov.p + (size_t)k * h->N * h->ld, h->N, h->N, h->ld));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

extern "C" int mcb_get_tau_new(mcb_handle* h, double* tau) {
  if (!h || !tau) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_get_tau_new: no E-step result");
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaMemcpyAsync(tau, h->tau_new.p, sizeof(double) * (size_t)h->M * h->K, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

extern "C" int mcb_get_mat_logL(mcb_handle* h, double* logL) {
  if (!h || !logL) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_get_mat_logL: no E-step result");
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaMemcpyAsync(logL, h->logL.p, sizeof(double) * (size_t)h->M * h->K, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

extern "C" int mcb_get_logdet_cov(mcb_handle* h, double* logdet) {
  if (!h || !logdet) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_get_logdet_cov: no E-step result");
  for (int k = 0; k < h->K; ++k) logdet[k] = -h->h_logdetA[k];
  return MCB_OK;
}

extern "C" int mcb_accept_tau(mcb_handle* h) {
  if (!h) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_accept_tau: no E-step result");
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaMemcpyAsync(h->tau.p, h->tau_new.p, sizeof(double) * (size_t)h->M * h->K, cudaMemcpyDeviceToDevice, h->st));
  return MCB_OK;
}

extern "C" int mcb_e_step_VEM(mcb_handle* h, int K, const double* theta_k, const double* theta_i,
                              const double* pi_k, const double* m_k, int m_k_len, const double* old_tau,
                              double* out_mean, double* out_cov, double* out_tau, double* out_mat_logL) {
  MCB_TRY(mcb_set_state(h, K, theta_k, theta_i, pi_k, m_k, m_k_len, old_tau));
  MCB_TRY(mcb_e_step(h));
  const int N = h->N, ld = h->ld, M = h->M;
  if (out_mean) MCB_TRY(copy_rows(h, out_mean, h->mean.p, K, N, ld));
  if (out_cov)
    for (int k = 0; k < K; ++k)
      MCB_TRY(copy_rows(h, out_cov + (size_t)k * N * N, h->cov.p + (size_t)k * N * ld, N, N, ld));
  if (out_tau) H_CUDA(cudaMemcpyAsync(out_tau, h->tau_new.p, sizeof(double) * (size_t)M * K, cudaMemcpyDeviceToHost, h->st));
  if (out_mat_logL) H_CUDA(cudaMemcpyAsync(out_mat_logL, h->logL.p, sizeof(double) * (size_t)M * K, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

// Install an explicit hyper-posterior (the `mu_k_param` / `list_mu_param` argument of the reference's
// callbacks) instead of the one left by mcb_e_step: mean[K][N], cov[K][N][N], tau[M][K] from host memory.
extern "C" int mcb_set_posterior(mcb_handle* h, const double* mean, const double* cov, const double* tau) {
  if (!h || !mean || !cov) return MCB_EINVAL;
  H_CHECK(h->has_state, MCB_ESTATE, "mcb_set_posterior: call mcb_set_data and mcb_set_state first");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  const int M = h->M, N = h->N, ld = h->ld, K = h->K;
  const long long sN = (long long)N * ld;
  H_CUDA(cudaMemcpy2DAsync(h->mean.p, sizeof(double) * ld, mean, sizeof(double) * N, sizeof(double) * N, K,
                           cudaMemcpyHostToDevice, h->st));
  for (int k = 0; k < K; ++k)
    H_CUDA(cudaMemcpy2DAsync(h->cov.p + k * sN, sizeof(double) * ld, cov + (size_t)k * N * N, sizeof(double) * N,
                             sizeof(double) * N, N, cudaMemcpyHostToDevice, h->st));
  if (tau) H_CUDA(cudaMemcpyAsync(h->tau_new.p, tau, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice, h->st));
  // Psi_i^-1 at the resident theta_i (no scatter: K = 0), then corr1/corr2/F_i with the supplied tau, or, without one,
  // with the responsibilities these means and covariances imply (update_tau_i_k_VEM, CF.R:733-762)
  H_CUDA(launch_factor_scatter(h->st, h->idata(), M, h->nmax, h->theta_i.p, h->tau.p, 0, h->cov.p, ld, sN, h->wk.p,
                               h->iwork(), h->info.p));
  vec_sub(h->st, K * ld, h->mean.p, h->mk.p, h->rvec.p, &h->launches);
  H_CUDA(launch_tau(h->st, h->idata(), M, h->nmax, K, h->mean.p, h->cov.p, ld, sN, h->pi.p, h->iwork(),
                    h->tau_new.p, h->logL.p, tau ? 1 : 0));
  h->launches += 2;
  // K_k^-1 and log|K_k| at the resident theta_k, log|C_k| of the supplied covariances (for the ELBO)
  std::vector<double> groups;
  h->G = group_theta(h->h_theta_k.data(), K, h->gmap, groups);
  H_CUDA(cudaMemcpyAsync(h->gmap_d.p, h->gmap.data(), sizeof(int) * K, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->theta_g.p, groups.data(), sizeof(double) * 3 * h->G, cudaMemcpyHostToDevice, h->st));
  se_build(h->st, h->Kinv.p, nullptr, nullptr, N, ld, sN, h->grid.p, h->theta_g.p, 3, h->G, &h->launches);
  H_CUDA(spd_inverse(h->st, h->Kinv.p, N, ld, sN, h->G, h->logdetK.p, h->info.p, h->ws, &h->launches));
  H_CUDA(cudaMemcpyAsync(h->eX.p, h->cov.p, sizeof(double) * K * sN, cudaMemcpyDeviceToDevice, h->st));
  H_CUDA(spd_inverse(h->st, h->eX.p, N, ld, sN, K, h->logdetA.p, h->info.p, h->ws, &h->launches));
  H_CUDA(cudaMemcpyAsync(h->hbuf, h->logdetK.p, sizeof(double) * h->G, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaMemcpyAsync(h->hbuf + K, h->logdetA.p, sizeof(double) * K, cudaMemcpyDeviceToHost, h->st));
  MCB_TRY(check_info(h, "mcb_set_posterior"));
  h->h_logdetK.assign(h->hbuf, h->hbuf + h->G);
  h->h_logdetA.resize(K);
  for (int k = 0; k < K; ++k) h->h_logdetA[k] = -h->hbuf[K + k];  // stored as log|A_k| = -log|C_k|
  h->has_post = true;
  h->csum_valid = false;
  return MCB_OK;
}

// ---- mean-process objective / gradient -------------------------------------------------------------
// Evaluates G groups of clusters, group g with hyper-parameters theta_h[g] = (a, b, sigma) shared by the
// clusters z with gmap_h[z] == g (gmap -1: cluster not evaluated). f_out[g] and g_out[g][3] (d/da, d/db,
// d/dsigma) follow logL_GP_mod(_common_hp_k) / gr_GP_mod(_common_hp_k) (CF.R:194-216, 250-278, 448-475,
// 510-537). use_resident: take K^-1 and log|K| from the last E-step instead of building them.
// Stream work of one evaluation (no host synchronisation): reads theta / gmap from the pinned staging area,
// leaves the reductions, log|K| and the not-PD flag in the pinned result area. Captured into a CUDA graph.
static int eval_mean_enqueue(mcb_handle* h, int G, bool want_grad, bool use_resident, bool refresh_csum) {
  const int N = h->N, ld = h->ld, K = h->K;
  const long long sN = (long long)N * ld;
  double* stage = h->hbuf + 1024;
  int* stage_i = (int*)(h->hbuf + 2048);
  H_CUDA(cudaMemcpyAsync(h->etheta.p, stage, sizeof(double) * 3 * G, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->egmap.p, stage_i, sizeof(int) * K, cudaMemcpyHostToDevice, h->st));
  const double* Kinv;
  const double* logdetK;
  if (use_resident) {
    Kinv = h->Kinv.p;
    logdetK = h->logdetK.p;
  } else {
    se_build(h->st, h->eK.p, want_grad ? h->eD1.p : nullptr, want_grad ? h->eD2.p : nullptr, N, ld, sN, h->grid.p,
             h->etheta.p, 3, G, &h->launches);
    H_CUDA(spd_inverse(h->st, h->eK.p, N, ld, sN, G, h->elogdet.p, h->info.p, h->ws, &h->launches));
    Kinv = h->eK.p;
    logdetK = h->elogdet.p;
  }
  // sum of the hyper-posterior covariances of each group: theta independent, refreshed only when the hyper-posterior
  // or the grouping changed since the last evaluation
  if (refresh_csum) sum_groups(h->st, N, ld, sN, h->cov.p, h->egmap.p, K, h->eCsum.p, G, &h->launches);
  symv_mapped(h->st, N, ld, sN, Kinv, h->egmap.p, h->rvec.p, h->pz.p, K, &h->launches);
  double* redg = h->ered.p;          // [G][4]
  double* redz = redg + 4 * K;       // [K][4]
  double* redt = redz + 4 * K;       // [G][4] (3 used)
  H_CUDA(cudaMemsetAsync(h->ered.p, 0, sizeof(double) * 12 * K, h->st));
  group_reduce(h->st, N, ld, sN, Kinv, h->eCsum.p, sN, want_grad ? h->eD1.p : nullptr,
               want_grad ? h->eD2.p : nullptr, redg, G, &h->launches);
  cluster_reduce(h->st, N, ld, sN, h->rvec.p, h->pz.p, want_grad ? h->eD1.p : nullptr,
                 want_grad ? h->eD2.p : nullptr, h->egmap.p, redz, K, &h->launches);
  if (want_grad) {
    // X = K^-1 Csum ; G = X K^-1 only through its traces against dK/da, dK/db and I
    gemm_nt(h->st, N, N, N, Kinv, ld, sN, h->eCsum.p, ld, sN, h->eX.p, ld, sN, 1.0, 0.0, G, &h->launches);
    gemm_nt_trace(h->st, N, h->eX.p, ld, sN, Kinv, ld, sN, h->eD1.p, h->eD2.p, ld, sN, redt, 4, G, &h->launches);
  }
  double* hr = h->hbuf + 4096;
  H_CUDA(cudaMemcpyAsync(hr, h->ered.p, sizeof(double) * 12 * K, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaMemcpyAsync(hr + 12 * K, logdetK, sizeof(double) * G, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaMemcpyAsync(h->hbuf + 8000, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  return MCB_OK;
}

// Evaluates G groups of clusters, group g with hyper-parameters theta_h[g] = (a, b, sigma) shared by the
// clusters z with gmap_h[z] == g (gmap -1: cluster not evaluated). f_out[g] and g_out[g][3] (d/da, d/db,
// d/dsigma) follow logL_GP_mod(_common_hp_k) / gr_GP_mod(_common_hp_k) (CF.R:194-216, 250-278, 448-475,
// 510-537). use_resident: take K^-1 and log|K| from the last E-step instead of building them.
// The ~40 kernels of one evaluation are replayed from a CUDA graph (one per shape), so the optimiser loop pays
// one graph launch + one synchronisation per objective evaluation instead of ~40 launch calls.
static int eval_mean_groups(mcb_handle* h, int G, const double* theta_h, const int* gmap_h, bool want_grad,
                            bool use_resident, double* f_out, double* g_out) {
  const int N = h->N, K = h->K;
  double* stage = h->hbuf + 1024;
  int* stage_i = (int*)(h->hbuf + 2048);
  memcpy(stage, theta_h, sizeof(double) * 3 * G);
  memcpy(stage_i, gmap_h, sizeof(int) * K);
  static const bool no_graph = getenv("MCB_NO_GRAPH") != nullptr;
  const std::vector<int> gm(gmap_h, gmap_h + K);
  const bool refresh_csum = !h->csum_valid || gm != h->csum_gmap;
  h->csum_gmap = gm;
  h->csum_valid = true;
  if (no_graph) {
    MCB_TRY(eval_mean_enqueue(h, G, want_grad, use_resident, refresh_csum));
  } else {
    const int key = ((((h->ws.max_ctas * 64 + G) * 2 + (want_grad ? 1 : 0)) * 2 + (use_resident ? 1 : 0)) * 2 + (refresh_csum ? 1 : 0));
    auto it = h->eval_graphs.find(key);
    if (it == h->eval_graphs.end()) {
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      const long long before = h->launches;
      H_CUDA(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
      const int rc = eval_mean_enqueue(h, G, want_grad, use_resident, refresh_csum);
      cudaError_t ce = cudaStreamEndCapture(h->st, &graph);
      if (rc != MCB_OK || ce != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        if (rc == MCB_OK) { h->err = std::string("graph capture failed: ") + cudaGetErrorString(ce); return MCB_ECUDA; }
        return rc;
      }
      H_CUDA(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      mcb_handle::EvalGraph eg{exec, h->launches - before};
      h->launches = before;
      it = h->eval_graphs.emplace(key, eg).first;
    }
    H_CUDA(cudaGraphLaunch(it->second.exec, h->st));
    h->launches += it->second.kernels;
  }
  H_CUDA(cudaStreamSynchronize(h->st));
  if (*(int*)(h->hbuf + 8000)) {
    cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st);
    h->err = "mean-process objective: a covariance matrix is not positive definite (non-positive pivot); "
             "the reference would silently fall back to MASS::ginv here";
    return MCB_ENOTPD;
  }
  const double* hr = h->hbuf + 4096;
  const double* hg = hr;
  const double* hz = hr + 4 * K;
  const double* ht = hr + 8 * K;
  const double* hl = hr + 12 * K;
  for (int g = 0; g < G; ++g) {
    int ng = 0;
    double q = 0, pp = 0, d1 = 0, d2 = 0;
    for (int z = 0; z < K; ++z)
      if (gmap_h[z] == g) {
        ++ng;
        q += hz[4 * z]; pp += hz[4 * z + 1]; d1 += hz[4 * z + 2]; d2 += hz[4 * z + 3];
      }
    f_out[g] = 0.5 * ng * ((double)N * kLog2Pi + hl[g]) + 0.5 * q + 0.5 * hg[4 * g];
    if (want_grad && g_out) {
      const double sg = theta_h[3 * g + 2];
      g_out[3 * g + 0] = -0.5 * (d1 + ht[4 * g + 0] - ng * hg[4 * g + 1]);
      g_out[3 * g + 1] = -0.5 * (d2 + ht[4 * g + 1] - ng * hg[4 * g + 2]);
      g_out[3 * g + 2] = -sg * (pp + ht[4 * g + 2] - ng * hg[4 * g + 3]);
    }
  }
  return MCB_OK;
}

extern "C" int mcb_eval_mean(mcb_handle* h, int k, const double* theta, int nhp, double pen_diag, double* f,
                             double* g) {
  if (!h || !theta || !f) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_eval_mean: run the E-step first");
  H_CHECK(k >= 0 && k < h->K && (nhp == 2 || nhp == 3), MCB_EINVAL, "mcb_eval_mean: bad argument");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  PhaseScope ps(h, PH_MSTEP_K);
  double th[3] = {theta[0], theta[1], nhp == 3 ? theta[2] : pen_diag};
  std::vector<int> gm(h->K, -1);
  gm[k] = 0;
  double fo[1], go[3];
  MCB_TRY(eval_mean_groups(h, 1, th, gm.data(), g != nullptr, false, fo, go));
  *f = fo[0];
  if (g) for (int j = 0; j < nhp; ++j) g[j] = go[j];
  return MCB_OK;
}

extern "C" int mcb_eval_mean_common(mcb_handle* h, const double* theta, int nhp, double pen_diag, double* f,
                                    double* g) {
  if (!h || !theta || !f) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_eval_mean_common: run the E-step first");
  H_CHECK(nhp == 2 || nhp == 3, MCB_EINVAL, "mcb_eval_mean_common: nhp must be 2 or 3");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  PhaseScope ps(h, PH_MSTEP_K);
  double th[3] = {theta[0], theta[1], nhp == 3 ? theta[2] : pen_diag};
  std::vector<int> gm(h->K, 0);
  double fo[1], go[3];
  MCB_TRY(eval_mean_groups(h, 1, th, gm.data(), g != nullptr, false, fo, go));
  *f = fo[0];
  if (g) for (int j = 0; j < nhp; ++j) g[j] = go[j];
  return MCB_OK;
}

// ---- individual objective / gradient ---------------------------------------------------------------
extern "C" int mcb_eval_indiv(mcb_handle* h, const double* theta, double* f, double* g) {
  if (!h || !theta || !f) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_eval_indiv: run the E-step first");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  const int M = h->M;
  H_CUDA(cudaMemcpyAsync(h->theta_i_new.p, theta, sizeof(double) * 3 * M, cudaMemcpyHostToDevice, h->st));
  {
    PhaseScope ps(h, PH_MSTEP_I);
    H_CUDA(launch_indiv_eval(h->st, h->idata(), M, h->nmax, h->iwork(), h->theta_i_new.p, 3, g != nullptr,
                             h->fi_tmp.p, h->gi_tmp.p, nullptr, h->info.p));
    ++h->launches;
  }
  H_CUDA(cudaMemcpyAsync(f, h->fi_tmp.p, sizeof(double) * M, cudaMemcpyDeviceToHost, h->st));
  if (g) H_CUDA(cudaMemcpyAsync(g, h->gi_tmp.p, sizeof(double) * 3 * M, cudaMemcpyDeviceToHost, h->st));
  h->has_cand = false; h->fnew_valid = false; h->fsum_valid = false;   // theta_i_new holds trial points now
  return check_info(h, "logL_clust_multi_GP");
}

// sum over all individuals (all ranks) of (F_i, dF_i) at one shared theta; result in out4 (host)
static int eval_indiv_common(mcb_handle* h, const double* theta3, bool want_grad, double* out4) {
  double* stage = h->hbuf + 512;
  memcpy(stage, theta3, sizeof(double) * 3);
  double* dth = h->scal.p + 8;
  double* dsum = h->scal.p + 16;
  H_CUDA(cudaMemcpyAsync(dth, stage, sizeof(double) * 3, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemsetAsync(dsum, 0, sizeof(double) * 4, h->st));
  H_CUDA(launch_indiv_eval(h->st, h->idata(), h->M, h->nmax, h->iwork(), dth, 0, want_grad, nullptr, nullptr, dsum,
                           h->info.p));
  ++h->launches;
  MCB_TRY(comm_allreduce_sum(h, dsum, 4));
  H_CUDA(cudaMemcpyAsync(h->hbuf + 520, dsum, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->st));
  if (h->nranks > 1) {
    // A non-positive pivot makes that individual's F_i NaN (log of the pivot), so the all-reduced sum is not finite on
    // EVERY rank: all ranks take the same branch without exchanging the local flag (which is cleared)
    H_CUDA(cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    if (!std::isfinite(h->hbuf[520])) {
      h->err = "logL_clust_multi_GP_common_hp_i: a covariance matrix is not positive definite";
      return MCB_ENOTPD;
    }
  } else {
    MCB_TRY(check_info(h, "logL_clust_multi_GP_common_hp_i"));
  }
  memcpy(out4, h->hbuf + 520, sizeof(double) * 4);
  return MCB_OK;
}

extern "C" int mcb_eval_indiv_common(mcb_handle* h, const double* theta, double* f, double* g) {
  if (!h || !theta || !f) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_eval_indiv_common: run the E-step first");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  PhaseScope ps(h, PH_MSTEP_I);
  double out[4];
  MCB_TRY(eval_indiv_common(h, theta, g != nullptr, out));
  *f = out[0];
  if (g) { g[0] = out[1]; g[1] = out[2]; g[2] = out[3]; }
  return MCB_OK;
}

// ---- ELBO --------------------------------------------------------------------------------------------
// logL_monitoring_VEM at (theta_k, theta_i) host arrays, or at the resident hp when both are NULL.
static int elbo_impl(mcb_handle* h, const double* theta_k, const double* theta_i_dev, double out[2]) {
  PhaseScope ps(h, PH_ELBO);
  const int K = h->K, N = h->N;
  double sum_k = 0.0;
  {
    std::vector<int> gmap;
    std::vector<double> groups;
    const bool resident = (theta_k == nullptr);
    const int G = group_theta(resident ? h->h_theta_k.data() : theta_k, K, gmap, groups);
    std::vector<double> f(G);
    MCB_TRY(eval_mean_groups(h, G, groups.data(), gmap.data(), false, resident, f.data(), nullptr));
    for (int g = 0; g < G; ++g) sum_k += f[g];
  }
  double sum_i;
  if (theta_i_dev == h->theta_i_new.p && h->fsum_valid) {
    // shared theta_i: the optimiser's objective at the point it returned IS sum_i F_i(theta_i_new), already summed over ranks
    sum_i = h->fsum_new;
  } else {
    double* dsum = h->scal.p + 24;
    H_CUDA(cudaMemsetAsync(dsum, 0, sizeof(double), h->st));
    if (theta_i_dev == nullptr) {
      vec_sum(h->st, h->M, h->Fi.p, dsum, &h->launches);
    } else if (theta_i_dev == h->theta_i_new.p && h->fnew_valid) {
      // the per-individual optimiser already evaluated F_i at the point it returned
      vec_sum(h->st, h->M, h->Fnew.p, dsum, &h->launches);
    } else {
      H_CUDA(launch_indiv_eval(h->st, h->idata(), h->M, h->nmax, h->iwork(), theta_i_dev, 3, 0, nullptr, nullptr, dsum,
                               h->info.p));
      ++h->launches;
    }
    MCB_TRY(comm_allreduce_sum(h, dsum, 1));
    H_CUDA(cudaMemcpyAsync(h->hbuf + 600, dsum, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (h->nranks > 1) {
      // as in eval_indiv_common: a non-positive pivot poisons the all-reduced sum on every rank
      H_CUDA(cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st));
      H_CUDA(cudaStreamSynchronize(h->st));
    } else {
      MCB_TRY(check_info(h, "logL_monitoring_VEM"));
    }
    sum_i = h->hbuf[600];
  }
  if (!std::isfinite(sum_i)) { h->err = "logL_monitoring_VEM: covariance not positive definite"; return MCB_ENOTPD; }
  out[0] = -sum_k - sum_i;
  double ld = 0.0;
  for (int k = 0; k < K; ++k) ld += -h->h_logdetA[k];
  out[1] = out[0] + 0.5 * ((double)K * N + ld);
  return MCB_OK;
}

extern "C" int mcb_elbo(mcb_handle* h, const double* theta_k, const double* theta_i, double out[2]) {
  if (!h || !out) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_elbo: run the E-step first");
  H_CHECK((theta_k == nullptr) == (theta_i == nullptr), MCB_EINVAL, "mcb_elbo: pass both theta_k and theta_i or neither");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  const double* ti = nullptr;
  if (theta_i) {
    H_CUDA(cudaMemcpyAsync(h->gi_tmp.p, theta_i, sizeof(double) * 3 * h->M, cudaMemcpyHostToDevice, h->st));
    ti = h->gi_tmp.p;
  }
  return elbo_impl(h, theta_k, ti, out);
}

// ---- BIC (CF.R:354-432, clust = TRUE) -----------------------------------------------------------------
// Terms of BIC at the resident hp against the resident hyper-posterior (install both with mcb_set_state +
// mcb_set_posterior, which also evaluates mat_logL at those hp):
//   out[0] = sum_i sum_k tau_ik (l_ik + log(pi_k / tau_ik))                         (LL_i, CF.R:366-399)
//   out[1] = sum_k [ dmvnorm(mean_k; m_1, K_k) - 0.5 sum(K_k^-1 o C_k) ]            (CF.R:414; every cluster is centred
//            on the FIRST prior mean m_k[[1]], CF.R:409)
//   out[2] = sum_k log det C_k          (maotai::pdeterminant of a full-rank SPD matrix, CF.R:415)
// The caller adds 0.5 K (N + N log 2 pi) and subtracts the penalty, which only counts distinct hp values.
extern "C" int mcb_bic_terms(mcb_handle* h, double out[3]) {
  if (!h || !out) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_bic_terms: no hyper-posterior (run the E-step or mcb_set_posterior)");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  const int K = h->K, ld = h->ld;
  double* d = h->scal.p + 48;
  H_CUDA(cudaMemsetAsync(d, 0, sizeof(double), h->st));
  bic_indiv(h->st, h->M, K, h->tau_new.p, h->logL.p, h->pi.p, d, &h->launches);
  MCB_TRY(comm_allreduce_sum(h, d, 1));
  H_CUDA(cudaMemcpyAsync(h->hbuf + 760, d, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  // cluster part with r_k = mean_k - m_1 (first prior mean for every cluster)
  for (int k = 0; k < K; ++k)
    H_CUDA(cudaMemcpyAsync(h->pz.p + (size_t)k * ld, h->mk.p, sizeof(double) * ld, cudaMemcpyDeviceToDevice, h->st));
  vec_sub(h->st, K * ld, h->mean.p, h->pz.p, h->rvec.p, &h->launches);
  std::vector<int> gmap;
  std::vector<double> groups;
  const int G = group_theta(h->h_theta_k.data(), K, gmap, groups);
  std::vector<double> f(G);
  const int rc = eval_mean_groups(h, G, groups.data(), gmap.data(), false, true, f.data(), nullptr);
  // restore r_k = mean_k - m_k for the other entry points
  vec_sub(h->st, K * ld, h->mean.p, h->mk.p, h->rvec.p, &h->launches);
  H_CUDA(cudaStreamSynchronize(h->st));
  if (rc != MCB_OK) return rc;
  out[0] = h->hbuf[760];
  out[1] = 0.0;
  for (int g = 0; g < G; ++g) out[1] -= f[g];
  out[2] = 0.0;
  for (int k = 0; k < K; ++k) out[2] += -h->h_logdetA[k];
  return MCB_OK;
}

// ---- device-driven optimisation of a shared theta_k -------------------------------------------------------
// The optimiser loop of CF.R:666-668 evaluates the objective ~25 times in sequence; driven from the host every
// evaluation pays a graph launch, a stream synchronisation and two small copies (~50 us against ~220 us of kernels).
// Here the loop itself lives on the device: a WHILE conditional graph node whose body is one evaluation followed by
// a one-warp kernel that forms f and the gradient from the reductions, advances the same L-BFGS state machine the
// host path uses (lbfgs.h) and decides whether to iterate again. The host launches one graph and waits once.
constexpr int kMoptWords = (int)((sizeof(Lbfgs) + 7) / 8);

// MCB_TRACE_K=1 (debug): globaltimer stamps between the nodes of the loop body, printed per segment after the M-step
static __device__ unsigned long long g_ktrace[8192];
static __device__ unsigned g_ktrace_n;
__global__ void ktrace_stamp_kernel(int tag) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const unsigned i = atomicAdd(&g_ktrace_n, 1u);
  if (i < 4096) { g_ktrace[2 * i] = t; g_ktrace[2 * i + 1] = (unsigned long long)tag; }
}
static bool ktrace_on() {
  static const bool on = getenv("MCB_TRACE_K") != nullptr;
  return on;
}
static void ktrace_dump() {
  static unsigned long long hostbuf[8192];
  unsigned n = 0;
  cudaMemcpyFromSymbol(&n, g_ktrace_n, sizeof n);
  cudaMemcpyFromSymbol(hostbuf, g_ktrace, sizeof hostbuf);
  if (n > 4096) n = 4096;
  double sum[32] = {0};
  int cnt[32] = {0};
  for (unsigned i = 0; i + 1 < n; ++i) {
    const int tag = (int)hostbuf[2 * i + 1];
    if (tag < 0 || tag >= 32) continue;
    sum[tag] += (double)(hostbuf[2 * i + 2] - hostbuf[2 * i]) * 1e-3;
    ++cnt[tag];
  }
  fprintf(stderr, "[ktrace] segment after stamp: mean us (count)\n");
  for (int t = 0; t < 32; ++t)
    if (cnt[t]) fprintf(stderr, "[ktrace]   %2d: %8.2f (%d)\n", t, sum[t] / cnt[t], cnt[t]);
  const unsigned zero = 0;
  cudaMemcpyToSymbol(g_ktrace_n, &zero, sizeof zero);
}
#define KTRACE(s, tag) do { if (ktrace_on()) ktrace_stamp_kernel<<<1, 1, 0, s>>>(tag); } while (0)

__global__ void mean_opt_init_kernel(const double* __restrict__ theta_k, double* state, double* etheta, int* egmap,
                                     int K) {
  if (threadIdx.x == 0) {
    Lbfgs* o = reinterpret_cast<Lbfgs*>(state);
    lb_init(*o, 3, theta_k);
    etheta[0] = theta_k[0]; etheta[1] = theta_k[1]; etheta[2] = theta_k[2];
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) egmap[k] = 0;
}

// result: [0..2] theta, [3] nfev, [4] iterations, [5] evaluations that met a non-positive pivot
__global__ void mean_opt_advance_kernel(double* state, const double* __restrict__ ered, const double* __restrict__ logdet,
                                        int* info, double* etheta, int N, int K, int fused,
                                        cudaGraphConditionalHandle handle, double* result) {
  __shared__ double words[kMoptWords];
  for (int w = threadIdx.x; w < kMoptWords; w += blockDim.x) words[w] = state[w];
  __syncthreads();
  if (threadIdx.x == 0) {
    Lbfgs& o = *reinterpret_cast<Lbfgs*>(words);
    // unfused layout: [G = 1][4] group sums, [K][4] per cluster r'p, p'p, p'D1 p, p'D2 p, [1][4] traces of G;
    // fused layout (MeanFuse): the same three groups of four, the cluster sums already added up
    const double* rg = ered;            // sum Kinv.Csum, tr(Kinv D1), tr(Kinv D2), tr(Kinv)
    const double* rz = ered + (fused ? 4 : 4 * K);
    const double* rt = ered + (fused ? 8 : 8 * K);   // tr(G D1), tr(G D2), tr(G)
    double q = 0, pp = 0, d1 = 0, d2 = 0;
    for (int z = 0; z < (fused ? 1 : K); ++z) { q += rz[4 * z]; pp += rz[4 * z + 1]; d1 += rz[4 * z + 2]; d2 += rz[4 * z + 3]; }
    double f = 0.5 * K * ((double)N * kLog2Pi + logdet[0]) + 0.5 * q + 0.5 * rg[0];
    double g[3];
    const double sg = o.x[2];
    g[0] = -0.5 * (d1 + rt[0] - K * rg[1]);
    g[1] = -0.5 * (d2 + rt[1] - K * rg[2]);
    g[2] = -sg * (pp + rt[2] - K * rg[3]);
    if (*info) {  // trial point with a non-PD covariance: infinite objective, the line search backs off
      f = INFINITY; g[0] = g[1] = g[2] = 0.0;
      *info = 0;
      result[5] += 1.0;
    }
    const int st = lb_advance(o, f, g);
    etheta[0] = o.x[0]; etheta[1] = o.x[1]; etheta[2] = o.x[2];
    result[0] = o.x[0]; result[1] = o.x[1]; result[2] = o.x[2];
    result[3] = (double)o.nfev; result[4] = (double)o.iter;
    cudaGraphSetConditional(handle, st == LB_EVAL ? 1u : 0u);
  }
  __syncthreads();
  for (int w = threadIdx.x; w < kMoptWords; w += blockDim.x) state[w] = words[w];
}

static int build_mean_opt_graph(mcb_handle* h) {
  const int N = h->N, ld = h->ld, K = h->K;
  const long long sN = (long long)N * ld;
  H_CUDA(h->mopt_state.ensure(kMoptWords + 16));
  if (!h->st_cap) H_CUDA(cudaStreamCreateWithFlags(&h->st_cap, cudaStreamNonBlocking));
  double* state = h->mopt_state.p;
  int* ginfo = h->info.p + 1;   // the loop's own not-positive-definite flag: it may run next to kernels that use info[0]
  double* result = state + kMoptWords;
  cudaGraph_t graph = nullptr;
  const long long before = h->launches;
  H_CUDA(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
  auto fail = [&](const char* what, cudaError_t ce) {
    cudaGraph_t g2 = nullptr;
    cudaStreamEndCapture(h->st, &g2);
    if (g2) cudaGraphDestroy(g2);
    h->launches = before;
    h->err = std::string("theta_k loop graph: ") + what + ": " + cudaGetErrorString(ce);
    return MCB_ECUDA;
  };
  cudaError_t ce;
  if ((ce = cudaMemsetAsync(result, 0, sizeof(double) * 16, h->st)) != cudaSuccess) return fail("memset", ce);
  if ((ce = cudaMemsetAsync(ginfo, 0, sizeof(int), h->st)) != cudaSuccess) return fail("memset", ce);
  mean_opt_init_kernel<<<1, 32, 0, h->st>>>(h->theta_k.p, state, h->etheta.p, h->egmap.p, K);
  ++h->launches;
  sum_groups(h->st, N, ld, sN, h->cov.p, h->egmap.p, K, h->eCsum.p, 1, &h->launches);
  h->mopt_pre = h->launches - before;
  // the WHILE node, inserted at the current capture frontier
  cudaStreamCaptureStatus cst;
  const cudaGraphNode_t* deps = nullptr;
  size_t ndeps = 0;
  if ((ce = cudaStreamGetCaptureInfo(h->st, &cst, nullptr, &graph, &deps, &ndeps)) != cudaSuccess) return fail("capture info", ce);
  cudaGraphConditionalHandle handle;
  if ((ce = cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault)) != cudaSuccess) return fail("handle", ce);
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t loop;
  if ((ce = cudaGraphAddNode(&loop, graph, deps, ndeps, &np)) != cudaSuccess) return fail("conditional node", ce);
  cudaGraph_t body = np.conditional.phGraph_out[0];
  // body: one evaluation with gradient (same enqueue order as eval_mean_enqueue, theta from device memory)
  {
    const long long b0 = h->launches;
    if ((ce = cudaStreamBeginCaptureToGraph(h->st_cap, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal)) != cudaSuccess)
      return fail("body capture", ce);
    cudaStream_t s = h->st_cap;
    static const bool unfused = getenv("MCB_MEAN_UNFUSED") != nullptr;
    const bool fused = !unfused && K <= 8 && spd_inverse_slab_ok(N, ld);
    cudaError_t ie = cudaSuccess;
    if (fused) {
      // four kernels per evaluation: build (+ clears the accumulators), slab inverse, two GEMM stages carrying the
      // reductions. ered: [0..3] group sums, [4..7] cluster sums (all clusters together), [8..10] traces
      if ((ie = spd_inverse_slab_reserve(h->ws, N, 1)) == cudaSuccess) {
        ZeroJob zj{h->ered.p, 12 * K, h->elogdet.p, 1, h->pz.p, K * ld, slab_sync_words(h->ws), 2};
        KTRACE(s, 0);
        se_build(s, h->eK.p, h->eD1.p, h->eD2.p, N, ld, sN, h->grid.p, h->etheta.p, 3, 1, &h->launches, &zj);
        KTRACE(s, 1);
        ie = spd_inverse_slab(s, h->eK.p, N, ld, sN, 1, h->elogdet.p, ginfo, h->ws, &h->launches, true);
        MeanFuse mf{K, ld, h->rvec.p, h->pz.p, h->ered.p};
        KTRACE(s, 2);
        mean_stage1(s, N, h->eK.p, h->eCsum.p, h->eX.p, ld, h->eD1.p, h->eD2.p, mf, &h->launches);
        KTRACE(s, 3);
        mean_stage2(s, N, h->eX.p, h->eK.p, ld, h->eD1.p, h->eD2.p, mf, &h->launches);
        KTRACE(s, 7);
      }
    } else {
      KTRACE(s, 0);
      se_build(s, h->eK.p, h->eD1.p, h->eD2.p, N, ld, sN, h->grid.p, h->etheta.p, 3, 1, &h->launches);
      KTRACE(s, 1);
      ie = spd_inverse(s, h->eK.p, N, ld, sN, 1, h->elogdet.p, ginfo, h->ws, &h->launches);
      KTRACE(s, 2);
      symv_mapped(s, N, ld, sN, h->eK.p, h->egmap.p, h->rvec.p, h->pz.p, K, &h->launches);
      double* redg = h->ered.p;
      double* redz = redg + 4 * K;
      double* redt = redz + 4 * K;
      cudaMemsetAsync(h->ered.p, 0, sizeof(double) * 12 * K, s);
      KTRACE(s, 3);
      group_reduce(s, N, ld, sN, h->eK.p, h->eCsum.p, sN, h->eD1.p, h->eD2.p, redg, 1, &h->launches);
      KTRACE(s, 4);
      cluster_reduce(s, N, ld, sN, h->rvec.p, h->pz.p, h->eD1.p, h->eD2.p, h->egmap.p, redz, K, &h->launches);
      KTRACE(s, 5);
      gemm_nt(s, N, N, N, h->eK.p, ld, sN, h->eCsum.p, ld, sN, h->eX.p, ld, sN, 1.0, 0.0, 1, &h->launches);
      KTRACE(s, 6);
      gemm_nt_trace(s, N, h->eX.p, ld, sN, h->eK.p, ld, sN, h->eD1.p, h->eD2.p, ld, sN, redt, 4, 1, &h->launches);
      KTRACE(s, 7);
    }
    mean_opt_advance_kernel<<<1, 32, 0, s>>>(state, h->ered.p, h->elogdet.p, ginfo, h->etheta.p, N, K, fused ? 1 : 0,
                                             handle, result);
    ++h->launches;
    cudaError_t ee = cudaStreamEndCapture(s, nullptr);
    h->mopt_body = h->launches - b0;
    if (ie != cudaSuccess) return fail("inverse in body", ie);
    if (ee != cudaSuccess) return fail("body end capture", ee);
  }
  if ((ce = cudaStreamUpdateCaptureDependencies(h->st, &loop, 1, cudaStreamSetCaptureDependencies)) != cudaSuccess)
    return fail("dependencies", ce);
  if ((ce = cudaMemcpyAsync(h->hbuf + 4096, result, sizeof(double) * 8, cudaMemcpyDeviceToHost, h->st)) != cudaSuccess)
    return fail("result copy", ce);
  h->launches = before;
  H_CUDA(cudaStreamEndCapture(h->st, &graph));
  cudaError_t inst = cudaGraphInstantiate(&h->mopt_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (inst != cudaSuccess) {
    h->mopt_exec = nullptr;
    h->err = std::string("theta_k loop graph: instantiate: ") + cudaGetErrorString(inst);
    return MCB_ECUDA;
  }
  return MCB_OK;
}

// ---- device-driven shared-theta_i optimisation --------------------------------------------------------------------------
// The optimiser of m_step_VEM's common_hp_i branch (CF.R:645-650: ONE opm() over sum_i F_i) as a WHILE graph, like the theta_k
// loop above: body = clear the 4 sums, the batched objective + gradient kernel over the local individuals (theta from device
// memory), the peer-memory all-reduce of the sums when the individuals are sharded over GPUs, one L-BFGS update on one thread.
// No host round trip per evaluation: at M = 100,000 on 8 GPUs the host turn-around (copy of 4 doubles, stream synchronisation,
// next launches) was 75 of the 374 us per evaluation.
// result: [0..2] theta, [3] nfev, [4] iterations, [5] status, [6] objective at theta, [7] evaluations with a non-finite sum
__global__ void indiv_opt_init_kernel(const double* __restrict__ x0, double* state, double* itheta) {
  if (threadIdx.x == 0) {
    Lbfgs* o = reinterpret_cast<Lbfgs*>(state);
    lb_init(*o, 3, x0);
    itheta[0] = x0[0]; itheta[1] = x0[1]; itheta[2] = x0[2];
  }
}

__global__ void indiv_opt_advance_kernel(double* state, const double* __restrict__ sum4, int* info, double* itheta,
                                         cudaGraphConditionalHandle handle, double* result) {
  __shared__ double words[kMoptWords];
  for (int w = threadIdx.x; w < kMoptWords; w += blockDim.x) words[w] = state[w];
  __syncthreads();
  if (threadIdx.x == 0) {
    Lbfgs& o = *reinterpret_cast<Lbfgs*>(words);
    double f = sum4[0];
    double g[3] = {sum4[1], sum4[2], sum4[3]};
    // a trial point where some Psi_i is not positive definite leaves NaN in that F_i, hence in the (all-reduced) sum on every
    // rank: an infinite objective for the line search (the reference's solve() would fail over to ginv, CF.R:142)
    if (!isfinite(f)) { f = INFINITY; g[0] = g[1] = g[2] = 0.0; result[7] += 1.0; }
    *info = 0;
    const int st = lb_advance(o, f, g);
    itheta[0] = o.x[0]; itheta[1] = o.x[1]; itheta[2] = o.x[2];
    result[0] = o.x[0]; result[1] = o.x[1]; result[2] = o.x[2];
    result[3] = (double)o.nfev; result[4] = (double)o.iter; result[5] = (double)st; result[6] = o.f;
    cudaGraphSetConditional(handle, st == LB_EVAL ? 1u : 0u);
  }
  __syncthreads();
  for (int w = threadIdx.x; w < kMoptWords; w += blockDim.x) state[w] = words[w];
}

__global__ void fill_theta_kernel(double* __restrict__ theta, int M, const double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) { theta[3 * (size_t)i] = x[0]; theta[3 * (size_t)i + 1] = x[1]; theta[3 * (size_t)i + 2] = x[2]; }
}

static int build_indiv_opt_graph(mcb_handle* h) {
  H_CUDA(h->iopt_state.ensure(kMoptWords + 16 + 8));
  if (!h->st_cap) H_CUDA(cudaStreamCreateWithFlags(&h->st_cap, cudaStreamNonBlocking));
  double* state = h->iopt_state.p;
  double* result = state + kMoptWords;
  double* itheta = result + 16;
  double* dsum = h->scal.p + 16;
  const double* x0 = h->scal.p + 32;
  cudaGraph_t graph = nullptr;
  const long long before = h->launches;
  H_CUDA(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
  auto fail = [&](const char* what, cudaError_t ce) {
    cudaGraph_t g2 = nullptr;
    cudaStreamEndCapture(h->st, &g2);
    if (g2) cudaGraphDestroy(g2);
    h->launches = before;
    h->err = std::string("theta_i loop graph: ") + what + ": " + cudaGetErrorString(ce);
    cudaGetLastError();
    return MCB_ECUDA;
  };
  cudaError_t ce;
  if ((ce = cudaMemsetAsync(result, 0, sizeof(double) * 16, h->st)) != cudaSuccess) return fail("memset", ce);
  indiv_opt_init_kernel<<<1, 32, 0, h->st>>>(x0, state, itheta);
  ++h->launches;
  h->iopt_pre = h->launches - before;
  cudaStreamCaptureStatus cst;
  const cudaGraphNode_t* deps = nullptr;
  size_t ndeps = 0;
  if ((ce = cudaStreamGetCaptureInfo(h->st, &cst, nullptr, &graph, &deps, &ndeps)) != cudaSuccess) return fail("capture info", ce);
  cudaGraphConditionalHandle handle;
  if ((ce = cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault)) != cudaSuccess) return fail("handle", ce);
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t loop;
  if ((ce = cudaGraphAddNode(&loop, graph, deps, ndeps, &np)) != cudaSuccess) return fail("conditional node", ce);
  cudaGraph_t body = np.conditional.phGraph_out[0];
  {
    const long long b0 = h->launches;
    if ((ce = cudaStreamBeginCaptureToGraph(h->st_cap, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal)) != cudaSuccess)
      return fail("body capture", ce);
    cudaStream_t s = h->st_cap;
    cudaError_t ie = cudaMemsetAsync(dsum, 0, sizeof(double) * 4, s);
    if (ie == cudaSuccess)
      ie = launch_indiv_eval(s, h->idata(), h->M, h->nmax, h->iwork(), itheta, 0, 1, nullptr, nullptr, dsum, h->info.p);
    ++h->launches;
    bool ok = comm_allreduce_small_on(h, s, dsum, 4);
    if (h->nranks > 1) ++h->launches;
    indiv_opt_advance_kernel<<<1, 32, 0, s>>>(state, dsum, h->info.p, itheta, handle, result);
    ++h->launches;
    cudaError_t ee = cudaStreamEndCapture(s, nullptr);
    h->iopt_body = h->launches - b0;
    if (ie != cudaSuccess) return fail("evaluation in body", ie);
    if (!ok) return fail("peer all-reduce in body", cudaErrorNotSupported);
    if (ee != cudaSuccess) return fail("body end capture", ee);
  }
  if ((ce = cudaStreamUpdateCaptureDependencies(h->st, &loop, 1, cudaStreamSetCaptureDependencies)) != cudaSuccess)
    return fail("dependencies", ce);
  fill_theta_kernel<<<(h->M + 255) / 256, 256, 0, h->st>>>(h->theta_i_new.p, h->M, result);
  if ((ce = cudaMemcpyAsync(h->hbuf + 9000, result, sizeof(double) * 8, cudaMemcpyDeviceToHost, h->st)) != cudaSuccess)
    return fail("result copy", ce);
  h->launches = before;
  H_CUDA(cudaStreamEndCapture(h->st, &graph));
  cudaError_t inst = cudaGraphInstantiate(&h->iopt_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (inst != cudaSuccess) {
    h->iopt_exec = nullptr;
    h->err = std::string("theta_i loop graph: instantiate: ") + cudaGetErrorString(inst);
    cudaGetLastError();
    return MCB_ECUDA;
  }
  return MCB_OK;
}

// Runs the whole optimisation; out[0..2] = theta, nfev / iterations returned through the pointers. Enqueue only when
// !wait (the caller synchronises h->st later and then calls with collect = true).
static int mean_opt_launch(mcb_handle* h, cudaStream_t stream = nullptr) {
  if (h->mopt_exec && h->mopt_budget != h->ws.max_ctas) {  // captured with another SM budget
    cudaGraphExecDestroy(h->mopt_exec);
    h->mopt_exec = nullptr;
  }
  if (!h->mopt_exec) MCB_TRY(build_mean_opt_graph(h));
  h->mopt_budget = h->ws.max_ctas;
  H_CUDA(cudaGraphLaunch(h->mopt_exec, stream ? stream : h->st));
  h->csum_valid = true;
  h->csum_gmap.assign(h->K, 0);
  return MCB_OK;
}

static void mean_opt_collect(mcb_handle* h, double theta[3], int* nfev, int* iters) {
  if (ktrace_on()) ktrace_dump();
  const double* r = h->hbuf + 4096;
  theta[0] = r[0]; theta[1] = r[1]; theta[2] = r[2];
  *nfev = (int)r[3];
  *iters = (int)r[4];
  h->launches += h->mopt_pre + (long long)*nfev * h->mopt_body;
}

// pen_diag = mean_i theta_i[3] over all ranks (CF.R:662)
static int compute_pen_diag(mcb_handle* h, double* pen_diag) {
  double* d = h->scal.p + 40;
  theta3_sum(h->st, h->M, h->theta_i_new.p, d, &h->launches);
  MCB_TRY(comm_allreduce_sum(h, d, 1));
  H_CUDA(cudaMemcpyAsync(h->hbuf + 700, d, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  *pen_diag = h->hbuf[700] / (double)h->M_total;
  return MCB_OK;
}

// The per-individual optimiser loop of m_step_VEM (CF.R:653-659) for the individuals of the global-memory path: the same
// L-BFGS state machine as the in-kernel one (lbfgs.h), advanced in lock step on the host, one batched objective + gradient
// launch over the still-active individuals per round (big_eval_kernel, indexed by position in big_ids).
static int mstep_big_individuals(mcb_handle* h) {
  const int nb = h->nbig;
  std::vector<Lbfgs> opt(nb);
  std::vector<int> st(nb, LB_EVAL), active(nb, 1);
  std::vector<double> th0(3 * (size_t)h->M), x(3 * (size_t)nb), f(nb), g(3 * (size_t)nb);
  H_CUDA(sync_copy(h, th0.data(), h->theta_i.p, sizeof(double) * 3 * h->M, cudaMemcpyDeviceToHost));
  for (int j = 0; j < nb; ++j) lb_init(opt[j], 3, &th0[3 * (size_t)h->h_big_ids[j]], 100);
  int left = nb;
  while (left > 0) {
    for (int j = 0; j < nb; ++j) for (int q = 0; q < 3; ++q) x[3 * (size_t)j + q] = opt[j].x[q];
    H_CUDA(cudaMemcpyAsync(h->big_theta.p, x.data(), sizeof(double) * 3 * nb, cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->big_active.p, active.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st));
    H_CUDA(launch_big_eval(h->st, h->idata(), h->iwork(), h->big_theta.p, 3, 1, h->big_active.p, 1, h->big_f.p, h->big_g.p,
                           nullptr, h->info.p));
    ++h->launches;
    H_CUDA(cudaMemcpyAsync(f.data(), h->big_f.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaMemcpyAsync(g.data(), h->big_g.p, sizeof(double) * 3 * nb, cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    for (int j = 0; j < nb; ++j) {
      if (!active[j]) continue;
      // a trial point where Psi_i is not positive definite (f = NaN) is a failed line-search step, as in the kernel
      st[j] = lb_advance(opt[j], f[j], &g[3 * (size_t)j]);
      if (st[j] != LB_EVAL) { active[j] = 0; --left; }
    }
  }
  H_CUDA(cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st));
  for (int j = 0; j < nb; ++j) {
    const int i = h->h_big_ids[j];
    const int iv[3] = {opt[j].nfev, opt[j].iter, st[j]};
    H_CUDA(cudaMemcpyAsync(h->theta_i_new.p + 3 * (size_t)i, opt[j].x, sizeof(double) * 3, cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->nfev_i.p + i, iv, sizeof(int), cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->niter_i.p + i, iv + 1, sizeof(int), cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->status_i.p + i, iv + 2, sizeof(int), cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));   // pageable sources on the stack
  }
  return MCB_OK;
}

// ---- M-step ------------------------------------------------------------------------------------------
static int m_step_impl(mcb_handle* h, int common_hp_k, int common_hp_i, int* stats) {
  const int M = h->M, K = h->K;
  int nfev_i = 0, nit_i = 0, nfev_k = 0, nit_k = 0;
  int fw = 0;
  h->fnew_valid = false;
  h->fsum_valid = false;
  // With individual-specific theta_i and a shared theta_k the two optimisations are independent (pen_diag only
  // replaces the third entry of theta_k afterwards, CF.R:666-668): run them concurrently.
  static const bool no_overlap = getenv("MCB_NO_OVERLAP") != nullptr;
  // (individuals on the global-memory path are optimised by a host loop after the kernel: no overlap then)
  const bool overlap = !no_overlap && !common_hp_i && common_hp_k && spd_inverse_slab_ok(h->N, h->ld) && h->nbig == 0;
  // Shared theta_i (one host-driven optimiser over sum_i F_i, one batched kernel per evaluation) and shared theta_k:
  // the two optimisations are independent here too. The device-side theta_k loop goes to a high-priority stream, so
  // its kernels (a few CTAs each) are placed as soon as CTAs of the batched theta_i evaluation retire.
  static const bool host_loop_k = getenv("MCB_HOST_THETA_K") != nullptr || getenv("MCB_NO_GRAPH") != nullptr;
  const bool overlap_c = !no_overlap && common_hp_i && common_hp_k && !host_loop_k && spd_inverse_slab_ok(h->N, h->ld);
  if (overlap_c) {
    H_CUDA(cudaEventRecord(h->ev_fork, h->st));
    H_CUDA(cudaStreamWaitEvent(h->st_hi, h->ev_fork, 0));
    {
      PhaseScope ps(h, PH_MSTEP_K, h->st_hi);
      h->ws.max_ctas = 0;
      MCB_TRY(mean_opt_launch(h, h->st_hi));
    }
    H_CUDA(cudaEventRecord(h->ev_join, h->st_hi));
  }
  // Work-queue order of the per-individual optimisations: longest first, predicted by the evaluation counts of
  // the previous M-step on the same individuals (shortens the tail of the persistent kernel).
  const int* order = nullptr;
  if (!common_hp_i && h->nfev_valid && M <= 65536) {
    std::vector<int> cnt(M), ord(M);
    H_CUDA(sync_copy(h, cnt.data(), h->nfev_i.p, sizeof(int) * M, cudaMemcpyDeviceToHost));
    for (int i = 0; i < M; ++i) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return cnt[x] > cnt[y]; });
    H_CUDA(h->order_i.ensure(M));
    H_CUDA(sync_copy(h, h->order_i.p, ord.data(), sizeof(int) * M, cudaMemcpyHostToDevice));
    order = h->order_i.p;
  }
  // --- individuals (CF.R:645-660)
  {
    PhaseScope ps(h, PH_MSTEP_I);
    if (common_hp_i) {
      // one optimiser on sum_i F_i, starting from theta_i[[1]] (CF.R:647)
      double x0[3];
      H_CUDA(cudaMemcpyAsync(h->hbuf, h->theta_i.p, sizeof(double) * 3, cudaMemcpyDeviceToHost, h->st));
      H_CUDA(cudaStreamSynchronize(h->st));
      if (h->nranks > 1) {
        // every rank must start from rank 0's first individual
        double* d = h->scal.p + 32;
        if (h->rank == 0) H_CUDA(cudaMemcpyAsync(d, h->theta_i.p, sizeof(double) * 3, cudaMemcpyDeviceToDevice, h->st));
        else H_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * 3, h->st));
        MCB_TRY(comm_allreduce_sum(h, d, 3));
        H_CUDA(cudaMemcpyAsync(h->hbuf, d, sizeof(double) * 3, cudaMemcpyDeviceToHost, h->st));
        H_CUDA(cudaStreamSynchronize(h->st));
      }
      memcpy(x0, h->hbuf, sizeof x0);
      // Device-side loop (WHILE graph) unless switched off, unavailable (several ranks without peer mailboxes) or it failed
      // to build once; the host loop below is the same state machine with one launch + one round trip per evaluation.
      static const bool host_loop_i = getenv("MCB_HOST_THETA_I") != nullptr || getenv("MCB_NO_GRAPH") != nullptr;
      bool done = false;
      if (!host_loop_i && !h->iopt_failed && (h->nranks == 1 || h->mbox_peers != nullptr)) {
        if (h->nranks == 1) H_CUDA(cudaMemcpyAsync(h->scal.p + 32, h->theta_i.p, sizeof(double) * 3, cudaMemcpyDeviceToDevice, h->st));
        if (!h->iopt_exec && build_indiv_opt_graph(h) != MCB_OK) {
          h->iopt_failed = true;   // (every rank builds the same graph from the same state: all fail or none)
        } else {
          H_CUDA(cudaGraphLaunch(h->iopt_exec, h->st));
          H_CUDA(cudaStreamSynchronize(h->st));
          const double* r = h->hbuf + 9000;
          nfev_i = (int)r[3]; nit_i = (int)r[4];
          h->launches += h->iopt_pre + (long long)nfev_i * h->iopt_body + 1;
          if ((int)r[5] != LB_NONFINITE && std::isfinite(r[6])) { h->fsum_new = r[6]; h->fsum_valid = true; }
          done = true;
        }
      }
      if (!done) {
      Lbfgs opt;
      lb_init(opt, 3, x0);
      int st = LB_EVAL;
      double out[4];
      while (st == LB_EVAL) {
        // a trial point where some Psi_i is not positive definite is an infinite objective for the line search
        // (the reference's solve() would fail over to ginv and carry on, CF.R:142)
        const int rc = eval_indiv_common(h, opt.x, true, out);
        if (rc == MCB_ENOTPD) { out[0] = INFINITY; out[1] = out[2] = out[3] = 0.0; }
        else if (rc != MCB_OK) return rc;
        st = lb_advance(opt, out[0], out + 1);
      }
      nfev_i = opt.nfev; nit_i = opt.iter;
      if (st != LB_NONFINITE && std::isfinite(opt.f)) { h->fsum_new = opt.f; h->fsum_valid = true; }
      std::vector<double> rep(3 * (size_t)M);
      for (int i = 0; i < M; ++i) { rep[3 * i] = opt.x[0]; rep[3 * i + 1] = opt.x[1]; rep[3 * i + 2] = opt.x[2]; }
      H_CUDA(cudaMemcpyAsync(h->theta_i_new.p, rep.data(), sizeof(double) * 3 * M, cudaMemcpyHostToDevice, h->st));
      H_CUDA(cudaStreamSynchronize(h->st));
      }
    } else if (!overlap) {
      H_CUDA(launch_indiv_mstep(h->st, h->idata(), M, h->nmax, h->iwork(), h->theta_i.p, h->theta_i_new.p, 100,
                                h->nfev_i.p, h->niter_i.p, h->status_i.p, h->mstep_ctl.p, 0, order, h->Fnew.p, &fw));
      h->fnew_valid = fw != 0;
      h->nfev_valid = true;
      ++h->launches;
      if (h->nbig > 0) {
        MCB_TRY(mstep_big_individuals(h));
        h->fnew_valid = false;   // F_i at the new theta_i is recomputed by the ELBO for everybody
      }
    }
  }
  if (overlap) {
    // theta_i optimisations on the second stream, leaving `reserve` SMs to the theta_k evaluations below
    H_CUDA(cudaEventRecord(h->ev_fork, h->st));
    H_CUDA(cudaStreamWaitEvent(h->st2, h->ev_fork, 0));
    {
      PhaseScope ps(h, PH_MSTEP_I, h->st2);
      // SMs left to the theta_k chain: one per CTA of the slab inverse (they must be co-resident), i.e. three (else two)
      // column parts per 32-row slab when that stays below ~a quarter of the GPU; the GEMMs / reductions between two
      // inverses share the same SMs
      const int nblk = (h->N + kNB - 1) / kNB;
      static const int extra = getenv("MCB_RESERVE_EXTRA") ? atoi(getenv("MCB_RESERVE_EXTRA")) : -1;
      // (re-tuned after the theta_i kernel got faster than the theta_k chain, gpurun_out/r02s3_resv*.log: three column
      // parts = 39 SMs and 8 one-CTA SMs at N = 401, M = 1000: 5.45 -> 5.25 ms per iteration)
      const int reserve = extra >= 0 ? std::min(nblk + extra, 96)
                                     : (3 * nblk <= 40 ? 3 * nblk : (2 * nblk <= 40 ? 2 * nblk : nblk));
      H_CUDA(launch_indiv_mstep(h->st2, h->idata(), M, h->nmax, h->iwork(), h->theta_i.p, h->theta_i_new.p, 100,
                                h->nfev_i.p, h->niter_i.p, h->status_i.p, h->mstep_ctl.p, reserve, order, h->Fnew.p, &fw));
      h->fnew_valid = fw != 0;
      h->ws.max_ctas = reserve;   // what the slab inverse may count on while that kernel runs
      h->nfev_valid = true;
      ++h->launches;
    }
    H_CUDA(cudaEventRecord(h->ev_join, h->st2));
  }
  // --- mean processes (CF.R:664-680)
  h->h_theta_k_new.assign(3 * K, 0.0);
  double pen_diag = 0.0;
  if (!overlap) MCB_TRY(compute_pen_diag(h, &pen_diag));
  {
    PhaseScope ps(h, PH_MSTEP_K);
    if (common_hp_k) {
      // NB reference behaviour: the start vector theta_k[[1]] has 3 entries, so sigma is optimised too and
      // then discarded: opm(...)[1, 1:2] keeps (a, b) and pen_diag becomes the third entry (CF.R:666-668)
      Lbfgs opt;
      static const bool host_loop = getenv("MCB_HOST_THETA_K") != nullptr || getenv("MCB_NO_GRAPH") != nullptr;
      if (overlap_c) {
        // launched before the theta_i optimiser (above): join
        H_CUDA(cudaStreamWaitEvent(h->st, h->ev_join, 0));
        H_CUDA(cudaStreamSynchronize(h->st));
        mean_opt_collect(h, opt.x, &nfev_k, &nit_k);
      } else if (!host_loop) {
        // the loop runs on the device (WHILE graph node); one launch, one wait
        MCB_TRY(mean_opt_launch(h));
        H_CUDA(cudaStreamSynchronize(h->st));
        mean_opt_collect(h, opt.x, &nfev_k, &nit_k);
      } else {
        lb_init(opt, 3, h->h_theta_k.data());
        std::vector<int> gm(K, 0);
        int st = LB_EVAL;
        double f, g[3];
        while (st == LB_EVAL) {
          const int rc = eval_mean_groups(h, 1, opt.x, gm.data(), true, false, &f, g);
          if (rc == MCB_ENOTPD) { f = INFINITY; g[0] = g[1] = g[2] = 0.0; }  // line search backs off (see above)
          else if (rc != MCB_OK) return rc;
          st = lb_advance(opt, f, g);
        }
        nfev_k = opt.nfev; nit_k = opt.iter;
      }
      if (overlap) {
        // join the theta_i stream, then pen_diag = mean_i theta_i[3]
        H_CUDA(cudaStreamWaitEvent(h->st, h->ev_join, 0));
        h->ws.max_ctas = 0;
        MCB_TRY(compute_pen_diag(h, &pen_diag));
      }
      for (int k = 0; k < K; ++k) {
        h->h_theta_k_new[3 * k] = opt.x[0];
        h->h_theta_k_new[3 * k + 1] = opt.x[1];
        h->h_theta_k_new[3 * k + 2] = pen_diag;
      }
    } else {
      // K optimisers over (a, b) with sigma = pen_diag, advanced in lock step: one batched launch sequence
      // evaluates every unfinished cluster (CF.R:673-679)
      std::vector<Lbfgs> opt(K);
      std::vector<int> st(K, LB_EVAL), gm(K), active;
      for (int k = 0; k < K; ++k) lb_init(opt[k], 2, h->h_theta_k.data() + 3 * k);
      std::vector<double> th(3 * K), f(K), g(3 * K);
      for (;;) {
        active.clear();
        std::fill(gm.begin(), gm.end(), -1);
        for (int k = 0; k < K; ++k)
          if (st[k] == LB_EVAL) {
            gm[k] = (int)active.size();
            th[3 * active.size()] = opt[k].x[0];
            th[3 * active.size() + 1] = opt[k].x[1];
            th[3 * active.size() + 2] = pen_diag;
            active.push_back(k);
          }
        if (active.empty()) break;
        const int rc = eval_mean_groups(h, (int)active.size(), th.data(), gm.data(), true, false, f.data(), g.data());
        if (rc == MCB_ENOTPD) {
          // the flag is shared by the batch: every active line search backs off this round
          std::fill(f.begin(), f.end(), INFINITY);
          std::fill(g.begin(), g.end(), 0.0);
        } else if (rc != MCB_OK) {
          return rc;
        }
        for (size_t a = 0; a < active.size(); ++a) {
          const int k = active[a];
          st[k] = lb_advance(opt[k], f[a], g.data() + 3 * a);
        }
      }
      for (int k = 0; k < K; ++k) {
        h->h_theta_k_new[3 * k] = opt[k].x[0];
        h->h_theta_k_new[3 * k + 1] = opt[k].x[1];
        h->h_theta_k_new[3 * k + 2] = pen_diag;
        nfev_k = std::max(nfev_k, opt[k].nfev);
        nit_k = std::max(nit_k, opt[k].iter);
      }
    }
  }
  memcpy(h->hbuf + 800, h->h_theta_k_new.data(), sizeof(double) * 3 * K);
  H_CUDA(cudaMemcpyAsync(h->theta_k_new.p, h->hbuf + 800, sizeof(double) * 3 * K, cudaMemcpyHostToDevice, h->st));
  // pi_k = mean_i tau_ik (CF.R:682)
  tau_colsum(h->st, M, K, h->tau_new.p, h->pi_new.p, &h->launches);
  MCB_TRY(comm_allreduce_sum(h, h->pi_new.p, K));
  H_CUDA(cudaMemcpyAsync(h->hbuf + 900, h->pi_new.p, sizeof(double) * K, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  h->h_pi_new.resize(K);
  for (int k = 0; k < K; ++k) h->h_pi_new[k] = h->hbuf[900 + k] / (double)h->M_total;
  memcpy(h->hbuf + 900, h->h_pi_new.data(), sizeof(double) * K);
  H_CUDA(cudaMemcpyAsync(h->pi_new.p, h->hbuf + 900, sizeof(double) * K, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  if (!common_hp_i && stats) {
    // max evaluations / iterations over individuals
    std::vector<int> tmp(2 * (size_t)M);
    H_CUDA(sync_copy(h, tmp.data(), h->nfev_i.p, sizeof(int) * M, cudaMemcpyDeviceToHost));
    H_CUDA(sync_copy(h, tmp.data() + M, h->niter_i.p, sizeof(int) * M, cudaMemcpyDeviceToHost));
    for (int i = 0; i < M; ++i) { nfev_i = std::max(nfev_i, tmp[i]); nit_i = std::max(nit_i, tmp[M + i]); }
  }
  if (stats) { stats[0] = nfev_i; stats[1] = nfev_k; stats[2] = nit_i; stats[3] = nit_k; }
  h->ms_stats[0] = nfev_i; h->ms_stats[1] = nfev_k; h->ms_stats[2] = nit_i; h->ms_stats[3] = nit_k;
  h->ms_stats[4] = common_hp_i ? (long long)nfev_i * h->M : -1;  // -1: per-individual counts live on the device
  h->has_cand = true;
  return MCB_OK;
}

extern "C" int mcb_m_step(mcb_handle* h, int common_hp_k, int common_hp_i, double* new_theta_k,
                          double* new_theta_i, double* new_pi_k, int* stats) {
  if (!h) return MCB_EINVAL;
  H_CHECK(h->has_post, MCB_ESTATE, "mcb_m_step: run the E-step first");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  MCB_TRY(m_step_impl(h, common_hp_k, common_hp_i, stats));
  if (new_theta_k) memcpy(new_theta_k, h->h_theta_k_new.data(), sizeof(double) * 3 * h->K);
  if (new_pi_k) memcpy(new_pi_k, h->h_pi_new.data(), sizeof(double) * h->K);
  if (new_theta_i) {
    H_CUDA(cudaMemcpyAsync(new_theta_i, h->theta_i_new.p, sizeof(double) * 3 * h->M, cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
  }
  return MCB_OK;
}

static int accept_hp_impl(mcb_handle* h) {
  H_CUDA(cudaMemcpyAsync(h->theta_i.p, h->theta_i_new.p, sizeof(double) * 3 * h->M, cudaMemcpyDeviceToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->theta_k.p, h->theta_k_new.p, sizeof(double) * 3 * h->K, cudaMemcpyDeviceToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(h->pi.p, h->pi_new.p, sizeof(double) * h->K, cudaMemcpyDeviceToDevice, h->st));
  h->h_theta_k = h->h_theta_k_new;
  h->h_pi = h->h_pi_new;
  return MCB_OK;
}

extern "C" int mcb_accept_hp(mcb_handle* h) {
  if (!h) return MCB_EINVAL;
  H_CHECK(h->has_cand, MCB_ESTATE, "mcb_accept_hp: no M-step candidate");
  H_CUDA(cudaSetDevice(h->device));
  return accept_hp_impl(h);
}

// ---- one iteration of training_VEM's loop (MC.R:44-93) --------------------------------------------
static int vem_iteration_impl(mcb_handle* h, int common_hp_k, int common_hp_i, double out[3], double* out_mean,
                              double* out_cov, double* out_tau);

extern "C" int mcb_vem_iteration(mcb_handle* h, int common_hp_k, int common_hp_i, double out[3]) {
  return vem_iteration_impl(h, common_hp_k, common_hp_i, out, nullptr, nullptr, nullptr);
}

extern "C" int mcb_vem_iteration_host(mcb_handle* h, int common_hp_k, int common_hp_i, double out[3], double* out_mean,
                                      double* out_cov, double* out_tau) {
  return vem_iteration_impl(h, common_hp_k, common_hp_i, out, out_mean, out_cov, out_tau);
}

static int vem_rest(mcb_handle* h, int common_hp_k, int common_hp_i, double out[3], bool copies) {
  double e_old[2], e_new[2];
  MCB_TRY(elbo_impl(h, nullptr, nullptr, e_old));           // MC.R:51-53; also the second term of MC.R:71
  MCB_TRY(m_step_impl(h, common_hp_k, common_hp_i, nullptr));
  {
    // MC.R:70. If the candidate hp make a covariance numerically singular (e.g. pen_diag = mean_i theta_i[3] ~ 0,
    // SURVEY Q12) the candidate is treated like the reference treats an M-step failure (MC.R:62-67): training
    // stops and the last valid hp are kept.
    const int rc = elbo_impl(h, h->h_theta_k_new.data(), h->theta_i_new.p, e_new);
    if (rc == MCB_ENOTPD) { e_new[0] = -INFINITY; e_new[1] = -INFINITY; }
    else if (rc != MCB_OK) return rc;
  }
  const double eps = std::isfinite(e_new[0]) ? (e_new[0] - e_old[0]) / fabs(e_new[0]) : -INFINITY;
  out[0] = e_old[1];
  out[1] = eps;
  if (!std::isfinite(eps)) {
    // the candidate hp are unusable: the reference leaves its loop BEFORE appending to logL_iterations, with convergence
    // still FALSE and the last valid hp (MC.R:62-67)
    out[2] = 2.0;
  } else if (eps < 1e-2) {
    if (eps > 0) MCB_TRY(accept_hp_impl(h));
    out[2] = 1.0;
  } else {
    MCB_TRY(accept_hp_impl(h));
    H_CUDA(cudaMemcpyAsync(h->tau.p, h->tau_new.p, sizeof(double) * (size_t)h->M * h->K, cudaMemcpyDeviceToDevice, h->st));
    out[2] = 0.0;
  }
  if (copies) H_CUDA(cudaStreamWaitEvent(h->st, h->ev_copy1, 0));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

static int vem_iteration_impl(mcb_handle* h, int common_hp_k, int common_hp_i, double out[3], double* out_mean,
                              double* out_cov, double* out_tau) {
  if (!h || !out) return MCB_EINVAL;
  H_CHECK(h->has_state, MCB_ESTATE, "mcb_vem_iteration: call mcb_set_data and mcb_set_state first");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  MCB_TRY(e_step_impl(h));
  const bool copies = out_mean || out_cov || out_tau;
  if (copies) {
    // param of this E-step to the caller's buffers, behind the M-step (mean, cov, tau_new are only read from here on)
    if (!h->st_copy) {
      H_CUDA(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking));
      H_CUDA(cudaEventCreateWithFlags(&h->ev_copy0, cudaEventDisableTiming));
      H_CUDA(cudaEventCreateWithFlags(&h->ev_copy1, cudaEventDisableTiming));
    }
    const int N = h->N, ld = h->ld, K = h->K;
    H_CUDA(cudaEventRecord(h->ev_copy0, h->st));
    H_CUDA(cudaStreamWaitEvent(h->st_copy, h->ev_copy0, 0));
    if (out_mean)
      H_CUDA(cudaMemcpy2DAsync(out_mean, sizeof(double) * N, h->mean.p, sizeof(double) * ld, sizeof(double) * N, K,
                               cudaMemcpyDeviceToHost, h->st_copy));
    if (out_cov)
      for (int k = 0; k < K; ++k)
        H_CUDA(cudaMemcpy2DAsync(out_cov + (size_t)k * N * N, sizeof(double) * N, h->cov.p + (size_t)k * N * ld,
                                 sizeof(double) * ld, sizeof(double) * N, N, cudaMemcpyDeviceToHost, h->st_copy));
    if (out_tau)
      H_CUDA(cudaMemcpyAsync(out_tau, h->tau_new.p, sizeof(double) * (size_t)h->M * K, cudaMemcpyDeviceToHost, h->st_copy));
    H_CUDA(cudaEventRecord(h->ev_copy1, h->st_copy));
  }
  const int rc = vem_rest(h, common_hp_k, common_hp_i, out, copies);
  if (copies) cudaStreamSynchronize(h->st_copy);   // never leave copies into caller memory in flight, error or not
  return rc;
}


// ---- stand-alone building blocks ---------------------------------------------------------------------------------
// scratch layout in h->kscr: [matrix n x ld][second matrix n x ld][x (ld)][theta 3][logdet 1][quad 1]
static int kern_scratch(mcb_handle* h, int n, int* ld_out, double** mat, double** mat2, double** vec, double** small) {
  const int ld = (n + 15) & ~15;
  const size_t need = 2 * (size_t)n * ld + 2 * (size_t)ld + 16;
  H_CUDA(h->kscr.ensure(need));
  *ld_out = ld;
  *mat = h->kscr.p;
  *mat2 = *mat + (size_t)n * ld;
  *vec = *mat2 + (size_t)n * ld;
  *small = *vec + 2 * (size_t)ld;
  return MCB_OK;
}

static int kern_build(mcb_handle* h, int n, const double* x, const double* theta2, double sigma, double** mat_out,
                      int* ld_out, double** small_out) {
  for (int a = 1; a < n; ++a) H_CHECK(x[a] >= x[a - 1], MCB_EINVAL, "kern_to_cov / kern_to_inv: timestamps must be sorted");
  int ld;
  double *mat, *mat2, *vec, *small;
  MCB_TRY(kern_scratch(h, n, &ld, &mat, &mat2, &vec, &small));
  const double th[3] = {theta2[0], theta2[1], sigma};
  H_CUDA(cudaMemcpyAsync(vec, x, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(small, th, sizeof th, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemsetAsync(mat, 0, sizeof(double) * (size_t)n * ld, h->st));   // zero pads (contract of the dense toolkit)
  se_build(h->st, mat, nullptr, nullptr, n, ld, (long long)n * ld, vec, small, 3, 1, &h->launches);
  *mat_out = mat;
  *ld_out = ld;
  *small_out = small;
  return MCB_OK;
}


extern "C" int mcb_kern_to_cov(mcb_handle* h, int n, const double* x, const double* theta2, double sigma, double* out) {
  if (!h) return MCB_EINVAL;
  H_CHECK(n >= 1 && x && theta2 && out, MCB_EINVAL, "mcb_kern_to_cov: bad argument");
  H_CUDA(cudaSetDevice(h->device));
  double *mat, *small;
  int ld;
  MCB_TRY(kern_build(h, n, x, theta2, sigma, &mat, &ld, &small));
  H_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * n, mat, sizeof(double) * ld, sizeof(double) * n, n,
                           cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  return MCB_OK;
}

extern "C" int mcb_kern_to_inv(mcb_handle* h, int n, const double* x, const double* theta2, double sigma, double* out,
                               double* logdet) {
  if (!h) return MCB_EINVAL;
  H_CHECK(n >= 1 && x && theta2 && out, MCB_EINVAL, "mcb_kern_to_inv: bad argument");
  H_CUDA(cudaSetDevice(h->device));
  double *mat, *small;
  int ld;
  MCB_TRY(kern_build(h, n, x, theta2, sigma, &mat, &ld, &small));
  H_CUDA(cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st));
  H_CUDA(spd_inverse(h->st, mat, n, ld, (long long)n * ld, 1, small + 4, h->info.p, h->ws_aux, &h->launches));
  H_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * n, mat, sizeof(double) * ld, sizeof(double) * n, n,
                           cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaMemcpyAsync(h->hbuf + 640, small + 4, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  MCB_TRY(check_info(h, "kern_to_inv"));
  if (logdet) *logdet = h->hbuf[640];
  return MCB_OK;
}

// quad[0] = z' A z with z = x - mu (one CTA; A n x ld)
__global__ void quad_form_kernel(int n, int ld, const double* __restrict__ A, const double* __restrict__ x,
                                 const double* __restrict__ mu, double* __restrict__ quad) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int a = e / n, b = e - a * n;
    acc += (x[a] - mu[a]) * A[(size_t)a * ld + b] * (x[b] - mu[b]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    quad[0] = s;
  }
}

extern "C" int mcb_dmvnorm(mcb_handle* h, int n, const double* x, const double* mu, const double* inv_Sigma, int want_log,
                           double* out) {
  if (!h) return MCB_EINVAL;
  H_CHECK(n >= 1 && x && mu && inv_Sigma && out, MCB_EINVAL, "mcb_dmvnorm: bad argument");
  H_CUDA(cudaSetDevice(h->device));
  int ld;
  double *mat, *mat2, *vec, *small;
  MCB_TRY(kern_scratch(h, n, &ld, &mat, &mat2, &vec, &small));
  H_CUDA(cudaMemsetAsync(mat, 0, sizeof(double) * 2 * (size_t)n * ld, h->st));
  H_CUDA(cudaMemcpy2DAsync(mat, sizeof(double) * ld, inv_Sigma, sizeof(double) * n, sizeof(double) * n, n,
                           cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpy2DAsync(mat2, sizeof(double) * ld, inv_Sigma, sizeof(double) * n, sizeof(double) * n, n,
                           cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(vec, x, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(vec + ld, mu, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
  quad_form_kernel<<<1, 256, 0, h->st>>>(n, ld, mat2, vec, vec + ld, small + 5);
  ++h->launches;
  // log det of the precision matrix from the pivots of its factorisation (the inverse itself is not used)
  H_CUDA(cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st));
  H_CUDA(spd_inverse(h->st, mat, n, ld, (long long)n * ld, 1, small + 4, h->info.p, h->ws_aux, &h->launches));
  H_CUDA(cudaMemcpyAsync(h->hbuf + 640, small + 4, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->st));
  MCB_TRY(check_info(h, "dmvnorm"));
  const double logdet_prec = h->hbuf[640], ssq = h->hbuf[641];
  const double ll = -0.5 * ((double)n * kLog2Pi - logdet_prec + ssq);
  *out = want_log ? ll : exp(ll);
  return MCB_OK;
}

// logL_GP_mod / gr_GP_mod (CF.R:194-216, 448-475; Z = 1) and logL_GP_mod_common_hp_k / gr_GP_mod_common_hp_k (CF.R:250-278,
// 510-537; Z clusters sharing one theta) on EXPLICIT arguments: the reference's callbacks take (hp, db, mean, kern, new_cov,
// pen_diag) and are also called on sub-problems that are not the resident hyper-posterior (update_tau_i_k_VEM evaluates
// logL_GP_mod per individual and cluster, CF.R:746). Same device code as the resident evaluators, own scratch buffers.
//   t[n] sorted timestamps, y[Z][n] outputs (db$Output per cluster), mean: Z values (scalar prior means) or [Z][n],
//   new_cov[Z][n][n]; theta[nhp] with sigma = theta[2] (nhp == 3) or pen_diag (nhp == 2); g (nhp entries) may be NULL.
extern "C" int mcb_gp_mod(mcb_handle* h, int n, int Z, const double* t, const double* y, const double* mean, int mean_len,
                          const double* new_cov, const double* theta, int nhp, double pen_diag, double* f, double* g) {
  if (!h) return MCB_EINVAL;
  H_CHECK(n >= 1 && Z >= 1 && t && y && mean && new_cov && theta && f, MCB_EINVAL, "mcb_gp_mod: bad argument");
  H_CHECK(nhp == 2 || nhp == 3, MCB_EINVAL, "mcb_gp_mod: nhp must be 2 or 3");
  H_CHECK(Z <= 256, MCB_EINVAL, "mcb_gp_mod: at most 256 clusters per call");
  H_CHECK(mean_len == Z || mean_len == Z * n, MCB_EINVAL, "mcb_gp_mod: mean must have Z or Z*n entries");
  for (int a = 1; a < n; ++a) H_CHECK(t[a] > t[a - 1], MCB_EINVAL, "mcb_gp_mod: timestamps must be sorted and unique");
  H_CUDA(cudaSetDevice(h->device));
  spans_reset(h);
  const int ld = (n + 15) & ~15;
  const long long sN = (long long)n * ld;
  // scratch: Kmat, D1, D2, Csum, X (n x ld each), cov[Z] (n x ld each), r[Z][ld], p[Z][ld], grid[ld], theta[3], logdet[1], red[12 Z]
  const size_t need = (5 + (size_t)Z) * sN + (2 * (size_t)Z + 1) * ld + 8 + 12 * (size_t)Z;
  H_CUDA(h->gscr.ensure(need));
  H_CUDA(h->gscr_i.ensure(Z));
  double* Kmat = h->gscr.p;
  double* D1 = Kmat + sN;
  double* D2 = D1 + sN;
  double* Csum = D2 + sN;
  double* X = Csum + sN;
  double* covz = X + sN;
  double* r = covz + (size_t)Z * sN;
  double* pz = r + (size_t)Z * ld;
  double* grid = pz + (size_t)Z * ld;
  double* th_d = grid + ld;
  double* logdet = th_d + 4;
  double* red = logdet + 4;
  H_CUDA(cudaMemsetAsync(h->gscr.p, 0, sizeof(double) * need, h->st));   // zero pads (contract of the dense toolkit)
  H_CUDA(cudaMemsetAsync(h->gscr_i.p, 0, sizeof(int) * Z, h->st));       // every cluster in group 0
  const double th[3] = {theta[0], theta[1], nhp == 3 ? theta[2] : pen_diag};
  std::vector<double> rh((size_t)Z * ld, 0.0);
  for (int z = 0; z < Z; ++z)
    for (int a = 0; a < n; ++a)
      rh[(size_t)z * ld + a] = y[(size_t)z * n + a] - (mean_len == Z ? mean[z] : mean[(size_t)z * n + a]);
  H_CUDA(cudaMemcpyAsync(grid, t, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(th_d, th, sizeof th, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpyAsync(r, rh.data(), sizeof(double) * Z * ld, cudaMemcpyHostToDevice, h->st));
  for (int z = 0; z < Z; ++z)
    H_CUDA(cudaMemcpy2DAsync(covz + z * sN, sizeof(double) * ld, new_cov + (size_t)z * n * n, sizeof(double) * n,
                             sizeof(double) * n, n, cudaMemcpyHostToDevice, h->st));
  const bool want_grad = g != nullptr;
  {
    PhaseScope ps(h, PH_MSTEP_K);
    se_build(h->st, Kmat, want_grad ? D1 : nullptr, want_grad ? D2 : nullptr, n, ld, sN, grid, th_d, 3, 1, &h->launches);
    H_CUDA(cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st));
    H_CUDA(spd_inverse(h->st, Kmat, n, ld, sN, 1, logdet, h->info.p, h->ws_aux, &h->launches));
    sum_groups(h->st, n, ld, sN, covz, h->gscr_i.p, Z, Csum, 1, &h->launches);
    symv_mapped(h->st, n, ld, sN, Kmat, h->gscr_i.p, r, pz, Z, &h->launches);
    double* redg = red;            // [1][4]
    double* redz = red + 4 * Z;    // [Z][4]
    double* redt = redz + 4 * Z;   // [1][4]
    group_reduce(h->st, n, ld, sN, Kmat, Csum, sN, want_grad ? D1 : nullptr, want_grad ? D2 : nullptr, redg, 1, &h->launches);
    cluster_reduce(h->st, n, ld, sN, r, pz, want_grad ? D1 : nullptr, want_grad ? D2 : nullptr, h->gscr_i.p, redz, Z,
                   &h->launches);
    if (want_grad) {
      gemm_nt(h->st, n, n, n, Kmat, ld, sN, Csum, ld, sN, X, ld, sN, 1.0, 0.0, 1, &h->launches);
      gemm_nt_trace(h->st, n, X, ld, sN, Kmat, ld, sN, D1, D2, ld, sN, redt, 4, 1, &h->launches);
    }
  }
  double* hr = h->hbuf + 4096;
  H_CUDA(cudaMemcpyAsync(hr, red, sizeof(double) * 12 * Z, cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaMemcpyAsync(hr + 12 * Z, logdet, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  MCB_TRY(check_info(h, "logL_GP_mod"));
  const double* hg = hr;
  const double* hz = hr + 4 * Z;
  const double* ht = hr + 8 * Z;
  double q = 0, pp = 0, d1 = 0, d2 = 0;
  for (int z = 0; z < Z; ++z) { q += hz[4 * z]; pp += hz[4 * z + 1]; d1 += hz[4 * z + 2]; d2 += hz[4 * z + 3]; }
  *f = 0.5 * Z * ((double)n * kLog2Pi + hr[12 * Z]) + 0.5 * q + 0.5 * hg[0];
  if (want_grad) {
    const double go[3] = {-0.5 * (d1 + ht[0] - Z * hg[1]), -0.5 * (d2 + ht[1] - Z * hg[2]), -th[2] * (pp + ht[2] - Z * hg[3])};
    for (int j = 0; j < nhp; ++j) g[j] = go[j];
  }
  return MCB_OK;
}

// ---- measurement -------------------------------------------------------------------------------------
// out[0..3] as mcb_m_step's stats; out[4] = total number of individual objective+gradient evaluations of the
// last M-step on this rank (sum over individuals), out[5] = number of individuals
extern "C" int mcb_last_mstep_stats(mcb_handle* h, long long out[6]) {
  if (!h || !out) return MCB_EINVAL;
  H_CHECK(h->has_cand, MCB_ESTATE, "mcb_last_mstep_stats: no M-step has run");
  H_CUDA(cudaSetDevice(h->device));
  for (int j = 0; j < 4; ++j) out[j] = h->ms_stats[j];
  if (h->ms_stats[4] < 0) {
    std::vector<int> tmp(h->M);
    H_CUDA(sync_copy(h, tmp.data(), h->nfev_i.p, sizeof(int) * h->M, cudaMemcpyDeviceToHost));
    long long sum = 0, mx = 0;
    for (int v : tmp) { sum += v; mx = std::max<long long>(mx, v); }
    h->ms_stats[4] = sum;
    h->ms_stats[0] = mx;
    out[0] = mx;
  }
  out[4] = h->ms_stats[4];
  out[5] = h->M;
  return MCB_OK;
}

// per-individual outcome of the last M-step with individual-specific theta_i (CF.R:653-659 runs one opm() per individual
// and keeps only its par; optim's counts and convergence code are what these are)
extern "C" int mcb_last_mstep_indiv(mcb_handle* h, int* nfev, int* niter, int* status) {
  if (!h) return MCB_EINVAL;
  H_CHECK(h->has_cand && h->nfev_valid, MCB_ESTATE,
          "mcb_last_mstep_indiv: the last M-step did not optimise theta_i per individual");
  H_CUDA(cudaSetDevice(h->device));
  if (nfev) H_CUDA(sync_copy(h, nfev, h->nfev_i.p, sizeof(int) * h->M, cudaMemcpyDeviceToHost));
  if (niter) H_CUDA(sync_copy(h, niter, h->niter_i.p, sizeof(int) * h->M, cudaMemcpyDeviceToHost));
  if (status) H_CUDA(sync_copy(h, status, h->status_i.p, sizeof(int) * h->M, cudaMemcpyDeviceToHost));
  return MCB_OK;
}

// page-locked host memory for buffers that cross PCIe on every call (results of mcb_predict: 1.6 GB at config C5)
extern "C" int mcb_host_alloc(void** out, size_t bytes) {
  if (!out || bytes == 0) return MCB_EINVAL;
  return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? MCB_OK : MCB_ENOMEM;
}

extern "C" int mcb_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? MCB_OK : MCB_ECUDA; }

extern "C" int mcb_pin_host(void* ptr, size_t bytes) {
  return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? MCB_OK : MCB_ECUDA;
}

extern "C" int mcb_unpin_host(void* ptr) { return cudaHostUnregister(ptr) == cudaSuccess ? MCB_OK : MCB_ECUDA; }

extern "C" int mcb_timer_start(mcb_handle* h) {
  if (!h) return MCB_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  if (!h->tm0) { H_CUDA(cudaEventCreate(&h->tm0)); H_CUDA(cudaEventCreate(&h->tm1)); }
  H_CUDA(cudaStreamSynchronize(h->st));
  H_CUDA(cudaEventRecord(h->tm0, h->st));
  return MCB_OK;
}

extern "C" int mcb_timer_stop(mcb_handle* h, float* ms) {
  if (!h || !ms) return MCB_EINVAL;
  H_CHECK(h->tm0 != nullptr, MCB_ESTATE, "mcb_timer_stop: timer not started");
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaEventRecord(h->tm1, h->st));
  H_CUDA(cudaEventSynchronize(h->tm1));
  H_CUDA(cudaEventElapsedTime(ms, h->tm0, h->tm1));
  return MCB_OK;
}

extern "C" int mcb_flush_l2(mcb_handle* h) {
  if (!h) return MCB_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)256 << 20;  // 256 MiB > 126 MB of L2
  H_CUDA(h->flush.ensure(n));
  H_CUDA(cudaMemsetAsync(h->flush.p, h->flush_byte++ & 0xff, n, h->st));
  return MCB_OK;
}

extern "C" int mcb_last_timings(mcb_handle* h, float ms[8], long long* launches) {
  if (!h || !ms) return MCB_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaStreamSynchronize(h->st));
  for (int i = 0; i < 8; ++i) ms[i] = 0.f;
  for (auto& sp : h->spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->evpool[sp.e0], h->evpool[sp.e1]) == cudaSuccess) ms[sp.phase] += t;
  }
  if (launches) *launches = h->launches;
  return MCB_OK;
}
