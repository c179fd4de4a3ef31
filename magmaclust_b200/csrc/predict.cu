// Prediction side of the path (sm_100a):
//   mcb_posterior_mu_k  posterior_mu_k (MC.R:196-233): the hyper-posterior block of the E-step on a prediction
//                       grid, theta_k[3] forced to 0.1 (MC.R:206), tau frozen
//   mcb_predict         update_tau_star_k_EM (CF.R:764-786) + pred_gp_clust (MC.R:235-274), batched over new
//                       individuals: one CTA per new individual, clusters in sequence
#include <algorithm>
#include <cmath>
#include <cstring>

#include "handle.h"

using namespace mcb;

#define H_CUDA(expr) MCB_CUDA(h, expr)
static inline cudaError_t sync_copy(mcb_handle* h, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, h->st);
  return e != cudaSuccess ? e : cudaStreamSynchronize(h->st);
}
static inline cudaError_t sync_copy2d(mcb_handle* h, void* dst, size_t dpitch, const void* src, size_t spitch,
                                      size_t width, size_t height, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, h->st);
  return e != cudaSuccess ? e : cudaStreamSynchronize(h->st);
}
#define H_CHECK(cond, code, text) \
  do {                            \
    if (!(cond)) {                \
      h->err = (text);            \
      return (code);              \
    }                             \
  } while (0)

namespace mcb {

constexpr int kTP = 256;
constexpr int kMaxObs = 64;  // observations per new individual supported by predict_kernel

__device__ __forceinline__ double wsum_p(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sweep of the bordered (n+1) x (n+1) matrix in shared memory with kTP threads (same algorithm as indiv.cu).
__device__ double sweep_p(double* A, int n, int ntot, int ld, int* bad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double logdet = 0.0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double d = A[k * ld + k];
    const double inv_d = 1.0 / d;
    logdet += log(d);
    if (!(d > 0.0) && threadIdx.x == 0) *bad = 1;
    const double* rk = A + k * ld;
    for (int a = warp; a < ntot; a += kTP / 32) {
      if (a == k) continue;
      double* ra = A + a * ld;
      const double ca = ra[k] * inv_d;
      for (int b = lane; b < ntot; b += 32)
        if (b != k) ra[b] -= ca * rk[b];
    }
    __syncthreads();
    for (int b = threadIdx.x; b < ntot; b += kTP) {
      if (b == k) continue;
      const double v = A[k * ld + b] * inv_d;
      A[k * ld + b] = v;
      A[b * ld + k] = v;
    }
    if (threadIdx.x == 0) A[k * ld + k] = -inv_d;
    __syncthreads();
  }
  return logdet;
}

// KC[k][i][j] = k_theta*(g_i, g_j) + C_k[i][j] on the prediction grid (diagonal exp(th1), no noise): the SE kernel of
// the new individuals depends only on the pair of grid points, so it is evaluated T^2 times here instead of
// B * K * T * n* times inside predict_kernel (MC.R:261-264 evaluates it per new individual and cluster).
__global__ void kc_build_kernel(int K, int T, int ldT, const double* __restrict__ pgrid, const double* __restrict__ pcov,
                                const double* __restrict__ theta, double* __restrict__ KC) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= T || j >= T) return;
  const double th1 = theta[0], th2 = theta[1];
  const double d = pgrid[i] - pgrid[j];
  const double e = (i == j) ? exp(th1) : exp(th1 - exp(-th2) / 2.0 * (d * d));
  const size_t sT = (size_t)T * ldT;
  for (int k = 0; k < K; ++k) KC[k * sT + (size_t)i * ldT + j] = e + pcov[k * sT + (size_t)i * ldT + j];
}

// BIG: more observations than the shared-memory version holds (n* > kMaxObs, or (n* + 1)^2 doubles above the shared-memory
// budget): the same layout lives in a global-memory workspace block per CTA (ws + blockIdx.x * wsstride) and the
// per-target quadratic form is done by one warp per target without a per-thread copy of the cross-covariance column.
// The reference has no size limit (MC.R:255-262).
template <bool BIG>
__global__ void __launch_bounds__(kTP) predict_kernel(int K, int T, int ldT, const double* __restrict__ pgrid,
                                                      const double* __restrict__ pmean,
                                                      const double* __restrict__ pcov, const int* __restrict__ off,
                                                      const int* __restrict__ oidx, const double* __restrict__ yv,
                                                      const double* __restrict__ theta, const double* __restrict__ pik,
                                                      int Tp, const int* __restrict__ pidx, int nmax,
                                                      double* __restrict__ out_tau, double* __restrict__ out_mean,
                                                      double* __restrict__ out_var, int* __restrict__ info,
                                                      double* ws, long long wsstride, int b0) {
  extern __shared__ double sm[];
  const int b = b0 + blockIdx.x;   // the batch is launched in chunks so that result copies overlap the next chunk
  const int o0 = off[b], n = off[b + 1] - o0;
  const int ntot = n + 1, ld = ntot | 1;
  double* A = BIG ? ws + (size_t)b * wsstride : sm;   // (nmax+1) x ld
  double* ts = A + (size_t)(nmax + 1) * ((nmax + 1) | 1);
  double* lk = ts + nmax;                       // K
  int* idx = (int*)(lk + K);
  const double sg = theta[2];
  for (int a = threadIdx.x; a < n; a += kTP) {
    idx[a] = oidx[o0 + a];
    ts[a] = pgrid[idx[a]];
  }
  __syncthreads();
  const long long sC = (long long)T * ldT;
  for (int k = 0; k < K; ++k) {
    const double* Ck = pcov + k * sC;
    const double* mk = pmean + (size_t)k * ldT;
    // S = Psi* + C_k[t*,t*], bordered with z = y* - m_k[t*]   (CF.R:776-777); Ck here is k_theta*(.,.) + C_k
    for (int e = threadIdx.x; e < ntot * ntot; e += kTP) {
      const int a = e / ntot, c = e - a * ntot;
      double v;
      if (a < n && c < n) {
        v = Ck[(size_t)idx[a] * ldT + idx[c]] + ((a == c) ? sg * sg : 0.0);
      } else if (a < n) {
        v = yv[o0 + a] - mk[idx[a]];
      } else if (c < n) {
        v = yv[o0 + c] - mk[idx[c]];
      } else {
        v = 0.0;
      }
      A[a * ld + c] = v;
    }
    const double logdet = sweep_p(A, n, ntot, ld, info);
    if (threadIdx.x == 0) lk[k] = -0.5 * ((double)n * kLog2Pi + logdet - A[n * ld + n]);  // dmvnorm, CF.R:780
    if (out_mean && BIG) {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      for (int j = warp; j < Tp; j += kTP / 32) {
        const int gj = pidx[j];
        double mean = 0.0, quad = 0.0;
        for (int a = lane; a < n; a += 32) {
          const double ga = Ck[(size_t)idx[a] * ldT + gj];
          mean += ga * A[a * ld + n];
          double row = 0.0;
          for (int c = 0; c < n; ++c) row -= A[a * ld + c] * Ck[(size_t)idx[c] * ldT + gj];   // A block = -S^-1
          quad += ga * row;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mean += __shfl_xor_sync(0xffffffffu, mean, o);
          quad += __shfl_xor_sync(0xffffffffu, quad, o);
        }
        if (lane == 0) {
          const size_t o = ((size_t)b * K + k) * Tp + j;
          out_mean[o] = mk[gj] + mean;
          out_var[o] = sg * sg + Ck[(size_t)gj * ldT + gj] - quad;
        }
      }
    } else if (out_mean) {
      // Mean_k = m_k[tp] + Gamma' S^-1 z ; Var_k = diag(Sigma(tp) + C_k[tp,tp] - Gamma' S^-1 Gamma)  (MC.R:261-268)
      for (int j = threadIdx.x; j < Tp; j += kTP) {
        const int gj = pidx[j];
        double gcol[kMaxObs];  // Gamma[:, j] = k(t*, tp_j) + C_k[t*, tp_j]   (MC.R:263)
        double mean = mk[gj];
        for (int a = 0; a < n; ++a) {
          const double ga = Ck[(size_t)idx[a] * ldT + gj];
          gcol[a] = ga;
          mean += ga * A[a * ld + n];
        }
        double quad = 0.0;
        for (int a = 0; a < n; ++a) {
          double row = 0.0;
          for (int c = 0; c < n; ++c) row -= A[a * ld + c] * gcol[c];  // A block = -S^-1
          quad += gcol[a] * row;
        }
        const size_t o = ((size_t)b * K + k) * Tp + j;
        out_mean[o] = mean;
        out_var[o] = sg * sg + Ck[(size_t)gj * ldT + gj] - quad;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double mx = lk[0];
    for (int k = 1; k < K; ++k) mx = fmax(mx, lk[k]);
    double den = 0.0;
    for (int k = 0; k < K; ++k) {
      lk[k] = pik[k] * exp(lk[k] - mx);
      den += lk[k];
    }
    for (int k = 0; k < K; ++k) out_tau[(size_t)b * K + k] = lk[k] / den;
  }
}

// log N(y*; m_k[t*], Psi_theta(t*) + C_k[t*,t*]) of new individuals at FREE hyper-parameters: one CTA per evaluation point
// (pt_ind[p] = new individual, pt_theta[p] = its trial theta), the objective terms of the repaired M-step of
// train_new_gp_EM (MC.R:151-166). The kernel matrix cannot be shared over the grid here (theta differs per point), so it is
// evaluated n*^2 times per cluster in place. A trial theta that makes a covariance numerically singular yields NaN for
// that term (the line search treats it as a failed step), not an error.
template <bool BIG>
__global__ void __launch_bounds__(kTP) new_indiv_logl_kernel(int K, int T, int ldT, const double* __restrict__ pgrid,
                                                             const double* __restrict__ pmean,
                                                             const double* __restrict__ pcov, const int* __restrict__ off,
                                                             const int* __restrict__ oidx, const double* __restrict__ yv,
                                                             const int* __restrict__ pt_ind,
                                                             const double* __restrict__ pt_theta, int nmax,
                                                             double* __restrict__ out, double* ws, long long wsstride) {
  extern __shared__ double sm[];
  __shared__ int bad;
  const int p = blockIdx.x;
  const int b = pt_ind[p];
  const int o0 = off[b], n = off[b + 1] - o0;
  const int ntot = n + 1, ld = ntot | 1;
  double* A = BIG ? ws + (size_t)blockIdx.x * wsstride : sm;   // (nmax+1) x ld
  double* ts = A + (size_t)(nmax + 1) * ((nmax + 1) | 1);
  int* idx = (int*)(ts + nmax);
  const double th1 = pt_theta[3 * p], th2 = pt_theta[3 * p + 1], sg = pt_theta[3 * p + 2];
  for (int a = threadIdx.x; a < n; a += kTP) {
    idx[a] = oidx[o0 + a];
    ts[a] = pgrid[idx[a]];
  }
  __syncthreads();
  const long long sC = (long long)T * ldT;
  for (int k = 0; k < K; ++k) {
    const double* Ck = pcov + k * sC;
    const double* mk = pmean + (size_t)k * ldT;
    if (threadIdx.x == 0) bad = 0;
    for (int e = threadIdx.x; e < ntot * ntot; e += kTP) {
      const int a = e / ntot, c = e - a * ntot;
      double v;
      if (a < n && c < n) {
        const double d = ts[a] - ts[c];
        const double kv = (a == c) ? exp(th1) + sg * sg : exp(th1 - exp(-th2) / 2.0 * (d * d));   // CF.R:43-50, 83-86
        v = kv + Ck[(size_t)idx[a] * ldT + idx[c]];
      } else if (a < n) {
        v = yv[o0 + a] - mk[idx[a]];
      } else if (c < n) {
        v = yv[o0 + c] - mk[idx[c]];
      } else {
        v = 0.0;
      }
      A[a * ld + c] = v;
    }
    const double logdet = sweep_p(A, n, ntot, ld, &bad);
    if (threadIdx.x == 0)
      out[(size_t)p * K + k] = bad ? nan("") : -0.5 * ((double)n * kLog2Pi + logdet - A[n * ld + n]);
    __syncthreads();
  }
}

}  // namespace mcb

extern "C" int mcb_posterior_mu_k(mcb_handle* h, int T, const double* timestamps, const double* tau,
                                  double* out_mean, double* out_cov) {
  if (!h || !timestamps) return MCB_EINVAL;
  H_CHECK(h->has_state, MCB_ESTATE, "mcb_posterior_mu_k: no trained state");
  H_CHECK(tau != nullptr || h->has_post, MCB_ESTATE, "mcb_posterior_mu_k: pass tau or run the E-step first");
  H_CHECK(T >= 1, MCB_EINVAL, "mcb_posterior_mu_k: T < 1");
  for (int j = 1; j < T; ++j) H_CHECK(timestamps[j] > timestamps[j - 1], MCB_EINVAL, "mcb_posterior_mu_k: timestamps must be sorted and unique");
  H_CUDA(cudaSetDevice(h->device));
  const int M = h->M, K = h->K, N = h->N;
  const int ldT = (T + 15) & ~15;
  const long long sT = (long long)T * ldT;
  // position of every training timestamp in the prediction grid (MC.R:323 builds t_pred as a superset)
  std::vector<double> grid(N);
  H_CUDA(sync_copy(h, grid.data(), h->grid.p, sizeof(double) * N, cudaMemcpyDeviceToHost));
  std::vector<int> gpos(N);
  for (int j = 0; j < N; ++j) {
    const double* it = std::lower_bound(timestamps, timestamps + T, grid[j]);
    H_CHECK(it != timestamps + T && *it == grid[j], MCB_EINVAL,
            "mcb_posterior_mu_k: the prediction grid must contain every training timestamp");
    gpos[j] = (int)(it - timestamps);
  }
  std::vector<int> gidx((size_t)h->P);
  H_CUDA(sync_copy(h, gidx.data(), h->gidx.p, sizeof(int) * h->P, cudaMemcpyDeviceToHost));
  for (auto& v : gidx) v = gpos[v];
  H_CUDA(h->pmapidx.ensure(h->P));
  H_CUDA(sync_copy(h, h->pmapidx.p, gidx.data(), sizeof(int) * h->P, cudaMemcpyHostToDevice));
  H_CUDA(h->pgrid.ensure(T));
  H_CUDA(sync_copy(h, h->pgrid.p, timestamps, sizeof(double) * T, cudaMemcpyHostToDevice));
  H_CUDA(h->pKinv.ensure((size_t)K * sT)); H_CUDA(h->pcov.ensure((size_t)K * sT));
  H_CUDA(h->pwk.ensure((size_t)K * ldT)); H_CUDA(h->pmean.ensure((size_t)K * ldT)); H_CUDA(h->plogdet.ensure(2 * K));
  H_CUDA(cudaMemsetAsync(h->pKinv.p, 0, sizeof(double) * h->pKinv.cap, h->st));  // zero row padding (ld % 16 == 0)
  H_CUDA(cudaMemsetAsync(h->pcov.p, 0, sizeof(double) * h->pcov.cap, h->st));
  const double* tau_d = h->tau_new.p;
  if (tau) {
    H_CUDA(sync_copy(h, h->logL.p, tau, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice));  // logL reused as staging
    tau_d = h->logL.p;
  }
  h->spans.clear(); h->ev_next = 0; h->launches = 0;
  // theta_k with the third entry forced to 0.1 (MC.R:206)
  std::vector<double> thk(h->h_theta_k);
  for (int k = 0; k < K; ++k) thk[3 * k + 2] = 0.1;
  std::vector<int> gmap(K, -1);
  std::vector<double> groups;
  int G = 0;
  for (int k = 0; k < K; ++k) {
    for (int g = 0; g < G; ++g)
      if (groups[3 * g] == thk[3 * k] && groups[3 * g + 1] == thk[3 * k + 1]) { gmap[k] = g; break; }
    if (gmap[k] < 0) { gmap[k] = G++; groups.insert(groups.end(), thk.begin() + 3 * k, thk.begin() + 3 * k + 3); }
  }
  H_CUDA(sync_copy(h, h->egmap.p, gmap.data(), sizeof(int) * K, cudaMemcpyHostToDevice));
  H_CUDA(sync_copy(h, h->etheta.p, groups.data(), sizeof(double) * 3 * G, cudaMemcpyHostToDevice));
  // prior means on the prediction grid: scalar priors replicate; vector priors need T == N
  std::vector<double> mk((size_t)K * h->ld), mkT((size_t)K * ldT, 0.0);
  H_CUDA(sync_copy(h, mk.data(), h->mk.p, sizeof(double) * K * h->ld, cudaMemcpyDeviceToHost));
  for (int k = 0; k < K; ++k) {
    bool constant = true;
    for (int j = 1; j < N; ++j) constant = constant && mk[(size_t)k * h->ld + j] == mk[(size_t)k * h->ld];
    H_CHECK(constant || T == N, MCB_EINVAL, "mcb_posterior_mu_k: vector prior means need T == N");
    for (int j = 0; j < T; ++j) mkT[(size_t)k * ldT + j] = constant ? mk[(size_t)k * h->ld] : mk[(size_t)k * h->ld + j];
  }
  double* mkT_d = h->pmean.p;  // staged here, overwritten by the result below
  H_CUDA(sync_copy(h, mkT_d, mkT.data(), sizeof(double) * K * ldT, cudaMemcpyHostToDevice));
  {
    se_build(h->st, h->pKinv.p, nullptr, nullptr, T, ldT, sT, h->pgrid.p, h->etheta.p, 3, G, &h->launches);
    H_CUDA(spd_inverse(h->st, h->pKinv.p, T, ldT, sT, G, h->plogdet.p, h->info.p, h->ws_aux, &h->launches));
    copy_mapped(h->st, T, ldT, sT, h->pKinv.p, h->egmap.p, h->pcov.p, K, h->rank == 0 ? 1.0 : 0.0, &h->launches);
    if (h->rank == 0) symv_mapped(h->st, T, ldT, sT, h->pKinv.p, h->egmap.p, mkT_d, h->pwk.p, K, &h->launches);
    else H_CUDA(cudaMemsetAsync(h->pwk.p, 0, sizeof(double) * K * ldT, h->st));
    IndivData dd = h->idata();
    dd.gidx = h->pmapidx.p;
    H_CUDA(launch_factor_scatter(h->st, dd, M, h->nmax, h->theta_i.p, tau_d, K, h->pcov.p, ldT, sT, h->pwk.p,
                                 h->iwork(), h->info.p));
    ++h->launches;
    MCB_TRY(comm_allreduce_sum(h, h->pcov.p, (size_t)K * sT));
    MCB_TRY(comm_allreduce_sum(h, h->pwk.p, (size_t)K * ldT));
    H_CUDA(spd_inverse(h->st, h->pcov.p, T, ldT, sT, K, h->plogdet.p + K, h->info.p, h->ws_aux, &h->launches));
    symv(h->st, T, h->pcov.p, ldT, sT, h->pwk.p, ldT, h->pmean.p, ldT, K, &h->launches);
  }
  int flag = 0;
  H_CUDA(cudaMemcpyAsync(&flag, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  if (flag) {
    cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st);
    h->err = "posterior_mu_k: a covariance matrix is not positive definite";
    return MCB_ENOTPD;
  }
  // pi_k = mean_i tau_ik of the frozen tau (MC.R:141), kept for mcb_predict
  tau_colsum(h->st, M, K, tau_d, h->plogdet.p, &h->launches);
  MCB_TRY(comm_allreduce_sum(h, h->plogdet.p, K));
  h->h_pi_pred.resize(K);
  H_CUDA(sync_copy(h, h->h_pi_pred.data(), h->plogdet.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
  for (auto& v : h->h_pi_pred) v /= (double)h->M_total;
  h->T = T; h->ldT = ldT; h->Kp = K; h->has_pred = true;
  if (out_mean)
    H_CUDA(sync_copy2d(h, out_mean, sizeof(double) * T, h->pmean.p, sizeof(double) * ldT, sizeof(double) * T, K, cudaMemcpyDeviceToHost));
  if (out_cov)
    for (int k = 0; k < K; ++k)
      H_CUDA(sync_copy2d(h, out_cov + (size_t)k * T * T, sizeof(double) * T, h->pcov.p + k * sT, sizeof(double) * ldT,
                          sizeof(double) * T, T, cudaMemcpyDeviceToHost));
  return MCB_OK;
}

// Install an explicit prediction posterior: the `list_mu` of pred_gp_clust (MC.R:235) / `mean_k`, `cov_k` of
// update_tau_star_k_EM (CF.R:764) when the caller supplies one (full_algo_clust(mu_k = ...), a model loaded from disk, a
// posterior computed by another handle) instead of the result of this handle's last mcb_posterior_mu_k.
extern "C" int mcb_set_pred_posterior(mcb_handle* h, int K, int T, const double* timestamps, const double* mean,
                                      const double* cov, const double* pi_k) {
  if (!h || !timestamps || !mean || !cov) return MCB_EINVAL;
  H_CHECK(K >= 1 && T >= 1, MCB_EINVAL, "mcb_set_pred_posterior: bad shape");
  for (int j = 1; j < T; ++j)
    H_CHECK(timestamps[j] > timestamps[j - 1], MCB_EINVAL, "mcb_set_pred_posterior: timestamps must be sorted and unique");
  H_CUDA(cudaSetDevice(h->device));
  const int ldT = (T + 15) & ~15;
  const long long sT = (long long)T * ldT;
  h->has_pred = false;
  H_CUDA(h->pgrid.ensure(T));
  H_CUDA(h->pKinv.ensure((size_t)K * sT)); H_CUDA(h->pcov.ensure((size_t)K * sT));
  H_CUDA(h->pwk.ensure((size_t)K * ldT)); H_CUDA(h->pmean.ensure((size_t)K * ldT)); H_CUDA(h->plogdet.ensure(2 * K));
  H_CUDA(cudaMemsetAsync(h->pcov.p, 0, sizeof(double) * h->pcov.cap, h->st));   // zero row padding (ld % 16 == 0)
  H_CUDA(cudaMemsetAsync(h->pmean.p, 0, sizeof(double) * h->pmean.cap, h->st));
  H_CUDA(cudaMemcpyAsync(h->pgrid.p, timestamps, sizeof(double) * T, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaMemcpy2DAsync(h->pmean.p, sizeof(double) * ldT, mean, sizeof(double) * T, sizeof(double) * T, K,
                           cudaMemcpyHostToDevice, h->st));
  for (int k = 0; k < K; ++k)
    H_CUDA(cudaMemcpy2DAsync(h->pcov.p + k * sT, sizeof(double) * ldT, cov + (size_t)k * T * T, sizeof(double) * T,
                             sizeof(double) * T, T, cudaMemcpyHostToDevice, h->st));
  H_CUDA(cudaStreamSynchronize(h->st));
  h->h_pi_pred.assign(K, 1.0 / K);
  if (pi_k) h->h_pi_pred.assign(pi_k, pi_k + K);
  h->T = T; h->ldT = ldT; h->Kp = K; h->has_pred = true;
  return MCB_OK;
}

extern "C" int mcb_predict(mcb_handle* h, int B, const int* offsets, const int* obs_idx, const double* y,
                           const double* theta_new, const double* pi_k, int Tp, const int* pred_idx,
                           double* out_tau, double* out_mean, double* out_var) {
  if (!h || !offsets || !obs_idx || !y || !theta_new || !out_tau) return MCB_EINVAL;
  H_CHECK(h->has_pred, MCB_ESTATE, "mcb_predict: call mcb_posterior_mu_k first");
  H_CHECK(B >= 1 && offsets[0] == 0, MCB_EINVAL, "mcb_predict: bad offsets");
  H_CHECK((out_mean == nullptr) == (out_var == nullptr), MCB_EINVAL, "mcb_predict: pass both out_mean and out_var or neither");
  H_CHECK(out_mean == nullptr || (Tp >= 1 && pred_idx), MCB_EINVAL, "mcb_predict: prediction targets missing");
  const int K = h->Kp, T = h->T;
  const long long Pn = offsets[B];
  int nmax = 0;
  for (int b = 0; b < B; ++b) {
    const int n = offsets[b + 1] - offsets[b];
    H_CHECK(n >= 1, MCB_EINVAL, "mcb_predict: empty new individual");
    nmax = std::max(nmax, n);
    for (int a = offsets[b]; a < offsets[b + 1]; ++a) {
      H_CHECK(obs_idx[a] >= 0 && obs_idx[a] < T, MCB_EINVAL, "mcb_predict: obs_idx out of range");
      H_CHECK(a == offsets[b] || obs_idx[a] > obs_idx[a - 1], MCB_EINVAL, "mcb_predict: observations must be sorted and unique");
    }
  }
  if (out_mean) for (int j = 0; j < Tp; ++j) H_CHECK(pred_idx[j] >= 0 && pred_idx[j] < T, MCB_EINVAL, "mcb_predict: pred_idx out of range");
  const size_t smem = sizeof(double) * ((size_t)(nmax + 1) * ((nmax + 1) | 1) + nmax + K) + sizeof(int) * nmax + 16;
  // above the shared-memory version's limits the same kernel runs on a global-memory workspace block per new individual
  const bool big = nmax > kMaxObs || smem > 227 * 1024;
  const long long wsstride = (long long)((smem + 7) / 8);
  mcb::DevBuf<double> d_ws;
  H_CUDA(cudaSetDevice(h->device));
  h->spans.clear(); h->ev_next = 0; h->launches = 0;
  mcb::DevBuf<int> d_off, d_idx, d_pidx;
  mcb::DevBuf<double> d_y, d_small, d_tau, d_mean, d_var;
  auto cleanup = [&]() {
    d_off.release(); d_idx.release(); d_pidx.release(); d_y.release(); d_small.release(); d_tau.release();
    d_mean.release(); d_var.release(); d_ws.release();
  };
  cudaError_t e = cudaSuccess;
  const size_t nout = out_mean ? (size_t)B * K * Tp : 0;
  if ((e = d_off.ensure(B + 1)) || (e = d_idx.ensure(Pn)) || (e = d_y.ensure(Pn)) || (e = d_small.ensure(3 + K)) ||
      (e = d_tau.ensure((size_t)B * K)) || (e = d_pidx.ensure(std::max(Tp, 1))) ||
      (nout && ((e = d_mean.ensure(nout)) || (e = d_var.ensure(nout)))) || (big && (e = d_ws.ensure((size_t)B * wsstride)))) {
    cleanup();
    h->err = std::string("mcb_predict: ") + cudaGetErrorString(e);
    return MCB_ENOMEM;
  }
  std::vector<double> small(theta_new, theta_new + 3);
  if (pi_k) small.insert(small.end(), pi_k, pi_k + K);
  else small.insert(small.end(), h->h_pi_pred.begin(), h->h_pi_pred.end());
  cudaMemcpyAsync(d_off.p, offsets, sizeof(int) * (B + 1), cudaMemcpyHostToDevice, h->st);
  cudaMemcpyAsync(d_idx.p, obs_idx, sizeof(int) * Pn, cudaMemcpyHostToDevice, h->st);
  cudaMemcpyAsync(d_y.p, y, sizeof(double) * Pn, cudaMemcpyHostToDevice, h->st);
  cudaMemcpyAsync(d_small.p, small.data(), sizeof(double) * (3 + K), cudaMemcpyHostToDevice, h->st);
  if (out_mean) cudaMemcpyAsync(d_pidx.p, pred_idx, sizeof(int) * Tp, cudaMemcpyHostToDevice, h->st);
  if (!big && smem > 48 * 1024) cudaFuncSetAttribute(predict_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  {
    int e0 = (int)h->evpool.size();
    cudaEvent_t a, b2;
    cudaEventCreate(&a); cudaEventCreate(&b2);
    h->evpool.push_back(a); h->evpool.push_back(b2);
    cudaEventRecord(a, h->st);
    kc_build_kernel<<<dim3((T + 31) / 32, (T + 7) / 8), dim3(32, 8), 0, h->st>>>(K, T, h->ldT, h->pgrid.p, h->pcov.p,
                                                                                 d_small.p, h->pKinv.p);
    ++h->launches;
    // Chunks of the batch: the means / variances of chunk c travel to the host (second stream) while chunk c + 1 is being
    // computed. 16 B K Tp bytes come back (1.6 GB at C5): with page-locked destinations (mcb_host_alloc / mcb_pin_host) the
    // call is bounded by that copy at PCIe rate; pageable destinations work but are staged by the driver.
    const int nch = out_mean ? std::max(1, std::min(16, B / 512)) : 1;
    if (out_mean && !h->st_copy) {
      cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&h->ev_copy0, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h->ev_copy1, cudaEventDisableTiming);
    }
    std::vector<cudaEvent_t> evc(nch, nullptr);
    for (int c = 0; c < nch; ++c) {
      const int bs = (int)((long long)B * c / nch), be = (int)((long long)B * (c + 1) / nch);
      if (big)
        predict_kernel<true><<<be - bs, kTP, 0, h->st>>>(K, T, h->ldT, h->pgrid.p, h->pmean.p, h->pKinv.p, d_off.p, d_idx.p,
                                                         d_y.p, d_small.p, d_small.p + 3, Tp, d_pidx.p, nmax, d_tau.p,
                                                         out_mean ? d_mean.p : nullptr, out_mean ? d_var.p : nullptr,
                                                         h->info.p, d_ws.p, wsstride, bs);
      else
        predict_kernel<false><<<be - bs, kTP, smem, h->st>>>(K, T, h->ldT, h->pgrid.p, h->pmean.p, h->pKinv.p, d_off.p, d_idx.p,
                                                             d_y.p, d_small.p, d_small.p + 3, Tp, d_pidx.p, nmax, d_tau.p,
                                                             out_mean ? d_mean.p : nullptr, out_mean ? d_var.p : nullptr,
                                                             h->info.p, nullptr, 0, bs);
      ++h->launches;
      if (out_mean) {
        cudaEventCreateWithFlags(&evc[c], cudaEventDisableTiming);
        cudaEventRecord(evc[c], h->st);
        cudaStreamWaitEvent(h->st_copy, evc[c], 0);
        const size_t o = (size_t)bs * K * Tp, cnt = (size_t)(be - bs) * K * Tp;
        cudaMemcpyAsync(out_mean + o, d_mean.p + o, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->st_copy);
        cudaMemcpyAsync(out_var + o, d_var.p + o, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->st_copy);
      }
    }
    cudaEventRecord(b2, h->st);
    h->spans.push_back({PH_PRED, e0, e0 + 1});
    cudaMemcpyAsync(out_tau, d_tau.p, sizeof(double) * (size_t)B * K, cudaMemcpyDeviceToHost, h->st);
    if (out_mean) cudaStreamSynchronize(h->st_copy);
    for (cudaEvent_t ev : evc) if (ev) cudaEventDestroy(ev);
  }
  int flag = 0;
  cudaMemcpyAsync(&flag, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, h->st);
  e = cudaStreamSynchronize(h->st);
  cleanup();
  if (e != cudaSuccess) {
    h->err = std::string("mcb_predict: ") + cudaGetErrorString(e);
    return MCB_ECUDA;
  }
  if (flag) {
    cudaMemsetAsync(h->info.p, 0, sizeof(int), h->st);
    h->err = "mcb_predict: Psi* + C_k[t*,t*] is not positive definite";
    return MCB_ENOTPD;
  }
  return MCB_OK;
}

// ---- new individuals with free hyper-parameters (repair of the free-hp branch of train_new_gp_EM, MC.R:142-186) ----
namespace {

// device copy of a batch of new individuals (CSR) + per-point scratch, alive for one EM run
struct NewIndivDev {
  mcb_handle* h;
  int B = 0, nmax = 0;
  size_t smem = 0;
  bool big = false;
  mcb::DevBuf<int> off, idx, pt_ind;
  mcb::DevBuf<double> y, pt_theta, out, ws;
  void release() { off.release(); idx.release(); pt_ind.release(); y.release(); pt_theta.release(); out.release(); ws.release(); }
};

int new_indiv_upload(mcb_handle* h, NewIndivDev& d, int B, const int* offsets, const int* obs_idx, const double* y,
                     const char* who) {
  const int T = h->T;
  H_CHECK(h->has_pred, MCB_ESTATE, std::string(who) + ": call mcb_posterior_mu_k first");
  H_CHECK(B >= 1 && offsets[0] == 0, MCB_EINVAL, std::string(who) + ": bad offsets");
  int nmax = 0;
  for (int b = 0; b < B; ++b) {
    const int n = offsets[b + 1] - offsets[b];
    H_CHECK(n >= 1, MCB_EINVAL, std::string(who) + ": empty new individual");
    nmax = std::max(nmax, n);
    for (int a = offsets[b]; a < offsets[b + 1]; ++a) {
      H_CHECK(obs_idx[a] >= 0 && obs_idx[a] < T, MCB_EINVAL, std::string(who) + ": obs_idx out of range");
      H_CHECK(a == offsets[b] || obs_idx[a] > obs_idx[a - 1], MCB_EINVAL,
              std::string(who) + ": observations must be sorted and unique");
    }
  }
  d.h = h; d.B = B; d.nmax = nmax;
  d.smem = sizeof(double) * ((size_t)(nmax + 1) * ((nmax + 1) | 1) + nmax) + sizeof(int) * nmax + 16;
  d.big = d.smem > 227 * 1024;   // beyond the shared-memory budget: global-memory workspace block per evaluation point
  H_CUDA(cudaSetDevice(h->device));
  const long long Pn = offsets[B];
  cudaError_t e = cudaSuccess;
  if ((e = d.off.ensure(B + 1)) || (e = d.idx.ensure(Pn)) || (e = d.y.ensure(Pn))) {
    d.release();
    h->err = std::string(who) + ": " + cudaGetErrorString(e);
    return MCB_ENOMEM;
  }
  cudaMemcpyAsync(d.off.p, offsets, sizeof(int) * (B + 1), cudaMemcpyHostToDevice, h->st);
  cudaMemcpyAsync(d.idx.p, obs_idx, sizeof(int) * Pn, cudaMemcpyHostToDevice, h->st);
  cudaMemcpyAsync(d.y.p, y, sizeof(double) * Pn, cudaMemcpyHostToDevice, h->st);
  if (!d.big && d.smem > 48 * 1024)
    cudaFuncSetAttribute(mcb::new_indiv_logl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem);
  if ((e = cudaStreamSynchronize(h->st)) != cudaSuccess) {
    d.release();
    h->err = std::string(who) + ": " + cudaGetErrorString(e);
    return MCB_ECUDA;
  }
  return MCB_OK;
}

// the mcb_logl_fn of the GPU path: npts evaluation points -> out[npts][K]
int new_indiv_eval(void* ctx, int npts, const int* pt_ind, const double* pt_theta, double* out) {
  NewIndivDev& d = *static_cast<NewIndivDev*>(ctx);
  mcb_handle* h = d.h;
  const int K = h->Kp;
  cudaError_t e = cudaSuccess;
  const long long wsstride = (long long)((d.smem + 7) / 8);
  if ((e = d.pt_ind.ensure(npts)) || (e = d.pt_theta.ensure(3 * (size_t)npts)) || (e = d.out.ensure((size_t)npts * K)) ||
      (d.big && (e = d.ws.ensure((size_t)npts * wsstride)))) {
    h->err = std::string("new-individual objective: ") + cudaGetErrorString(e);
    return MCB_ENOMEM;
  }
  cudaMemcpyAsync(d.pt_ind.p, pt_ind, sizeof(int) * npts, cudaMemcpyHostToDevice, h->st);
  cudaMemcpyAsync(d.pt_theta.p, pt_theta, sizeof(double) * 3 * npts, cudaMemcpyHostToDevice, h->st);
  if (d.big)
    mcb::new_indiv_logl_kernel<true><<<npts, mcb::kTP, 0, h->st>>>(K, h->T, h->ldT, h->pgrid.p, h->pmean.p, h->pcov.p, d.off.p,
                                                                   d.idx.p, d.y.p, d.pt_ind.p, d.pt_theta.p, d.nmax, d.out.p,
                                                                   d.ws.p, wsstride);
  else
    mcb::new_indiv_logl_kernel<false><<<npts, mcb::kTP, d.smem, h->st>>>(K, h->T, h->ldT, h->pgrid.p, h->pmean.p, h->pcov.p,
                                                                         d.off.p, d.idx.p, d.y.p, d.pt_ind.p, d.pt_theta.p,
                                                                         d.nmax, d.out.p, nullptr, 0);
  ++h->launches;
  cudaMemcpyAsync(out, d.out.p, sizeof(double) * (size_t)npts * K, cudaMemcpyDeviceToHost, h->st);
  if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(h->st)) != cudaSuccess) {
    h->err = std::string("new-individual objective: ") + cudaGetErrorString(e);
    return MCB_ECUDA;
  }
  return MCB_OK;
}

}  // namespace

extern "C" int mcb_new_indiv_logl(mcb_handle* h, int B, const int* offsets, const int* obs_idx, const double* y, int npts,
                                  const int* pt_ind, const double* pt_theta, double* out_logl) {
  if (!h || !offsets || !obs_idx || !y || !pt_ind || !pt_theta || !out_logl) return MCB_EINVAL;
  H_CHECK(npts >= 1, MCB_EINVAL, "mcb_new_indiv_logl: no evaluation point");
  for (int p = 0; p < npts; ++p) H_CHECK(pt_ind[p] >= 0 && pt_ind[p] < B, MCB_EINVAL, "mcb_new_indiv_logl: pt_ind out of range");
  NewIndivDev d;
  MCB_TRY(new_indiv_upload(h, d, B, offsets, obs_idx, y, "mcb_new_indiv_logl"));
  h->launches = 0;
  const int rc = new_indiv_eval(&d, npts, pt_ind, pt_theta, out_logl);
  d.release();
  return rc;
}

extern "C" int mcb_train_new_indiv(mcb_handle* h, int B, const int* offsets, const int* obs_idx, const double* y,
                                   const double* theta0, int theta0_per_indiv, const double* pi_k, int n_loop_max,
                                   double* out_theta, double* out_tau, int* out_loops) {
  if (!h || !offsets || !obs_idx || !y || !theta0 || !out_theta || !out_tau) return MCB_EINVAL;
  NewIndivDev d;
  h->err.clear();
  MCB_TRY(new_indiv_upload(h, d, B, offsets, obs_idx, y, "mcb_train_new_indiv"));
  h->launches = 0;
  const double* pi = pi_k ? pi_k : h->h_pi_pred.data();
  const int rc = mcb_new_indiv_em(B, h->Kp, theta0, theta0_per_indiv, pi, n_loop_max, new_indiv_eval, &d, out_theta,
                                  out_tau, out_loops);
  d.release();
  if (rc != MCB_OK && h->err.empty()) h->err = "mcb_train_new_indiv: the EM driver failed";
  return rc;
}
