// Per-individual FP64 kernels of the VEM path (sm_100a): one CTA per individual, the n_i x n_i problem
// resident in shared memory.
//
//   indiv_factor_scatter  kern_to_inv (CF.R:153-190) + update_inv_VEM (CF.R:690-707) + update_mean_VEM
//                         (CF.R:709-731): Psi_i^-1, Psi_i^-1 y_i, log|Psi_i|, scattered tau-weighted onto A_k, w_k
//   indiv_tau             update_tau_i_k_VEM (CF.R:733-762) via logL_GP_mod (CF.R:194-216) / dmvnorm
//                         (CF.R:10-35); also emits the theta-independent M-step inputs corr1/corr2
//                         (CF.R:236-245) and F_i at the current hp (logL_clust_multi_GP, CF.R:218-248)
//   indiv_eval            logL_clust_multi_GP + gr_clust_multi_GP (CF.R:218-248, 477-508) at a trial theta
//   indiv_mstep           the per-individual optimiser loop of m_step_VEM (CF.R:653-659), L-BFGS in-kernel
//
// The inverse is a symmetric sweep of the bordered matrix [[Psi, y, c1], [., 0]]: after sweeping the n
// pivots the block holds -Psi^-1, the borders hold Psi^-1 y and Psi^-1 c1, the corner -y'Psi^-1 y, and
// log|Psi| is the sum of the log pivots. Same flop count as Cholesky + inverse (n^3), but one uniform
// rank-1 update per step and no separate triangular phases.
#include "common.cuh"
#include "indiv.cuh"
#include "lbfgs.h"
#include "indiv_eval.cuh"
#include "indiv_eval2.cuh"
#include "indiv_eval3.cuh"

namespace mcb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum `NV` values over the CTA; result valid in all threads. red: smem [NV][kT/32]
template <int NV>
__device__ __forceinline__ void cta_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) v[q] = warp_sum(v[q]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) red[q * (kT / 32) + warp] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) s += red[q * (kT / 32) + w];
    v[q] = s;
  }
}

// Build the bordered SE covariance in shared memory (kern_to_cov semantics, CF.R:61-86):
// A[a][b] = exp(th1 - exp(-th2)/2 (ta-tb)^2) off-diagonal, exp(th1) + sigma^2 on the diagonal.
__device__ __forceinline__ void build_bordered(double* A, int ld, int n, int nbord, const double* t,
                                               const double* const* bord, double th1, double th2, double sg) {
  const double hl = exp(-th2) / 2.0;
  const double dg = exp(th1) + sg * sg;
  const int ntot = n + nbord;
  for (int e = threadIdx.x; e < ntot * ntot; e += kT) {
    const int a = e / ntot, b = e - a * ntot;
    double v;
    if (a < n && b < n) {
      const double d = t[a] - t[b];
      v = (a == b) ? dg : exp(th1 - hl * (d * d));
    } else if (a < n) {
      v = bord[b - n][a];
    } else if (b < n) {
      v = bord[a - n][b];
    } else {
      v = 0.0;
    }
    A[a * ld + b] = v;
  }
}

// Sweep pivots 0..n-1 of the symmetric ntot x ntot matrix in shared memory. Returns sum of log pivots
// (identical in every thread); *bad is set if a pivot is not positive.
__device__ double cta_sweep(double* A, int n, int ntot, int ld, int* bad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double logdet = 0.0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double d = A[k * ld + k];
    const double inv_d = 1.0 / d;
    logdet += log(d);
    if (!(d > 0.0) && threadIdx.x == 0) *bad = 1;
    const double* rk = A + k * ld;
    for (int a = warp; a < ntot; a += kT / 32) {
      if (a == k) continue;
      double* ra = A + a * ld;
      const double ca = ra[k] * inv_d;
      for (int b = lane; b < ntot; b += 32)
        if (b != k) ra[b] -= ca * rk[b];
    }
    __syncthreads();
    for (int b = threadIdx.x; b < ntot; b += kT) {
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

// -------------------------------------------------------------------------------------------------
// E-step, part 1
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) indiv_factor_scatter_kernel(IndivData dd, const double* __restrict__ theta_i,
                                                                  const double* __restrict__ tau, int K,
                                                                  double* __restrict__ acc, int ldN,
                                                                  long long strideA, double* __restrict__ wk,
                                                                  IndivWork w, int* __restrict__ info) {
  extern __shared__ double sm[];
  const int i = blockIdx.x;
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  if (n > dd.small_max) return;   // global-memory path (indiv_big.cu)
  const int ntot = n + 1, ld = ntot | 1;
  double* A = sm;
  double* t = A + (size_t)ntot * ld;
  double* y = t + n;
  int* idx = (int*)(y + n);
  for (int a = threadIdx.x; a < n; a += kT) {
    t[a] = dd.t[o0 + a];
    y[a] = dd.y[o0 + a];
    idx[a] = dd.gidx[o0 + a];
  }
  __syncthreads();
  const double th1 = theta_i[3 * i], th2 = theta_i[3 * i + 1], sg = theta_i[3 * i + 2];
  const double* bord[1] = {y};
  build_bordered(A, ld, n, 1, t, bord, th1, th2, sg);
  const double logdet = cta_sweep(A, n, ntot, ld, info);
  double* Wg = w.Winv + dd.woff[i];
  for (int e = threadIdx.x; e < n * n; e += kT) {
    const int a = e / n, b = e - a * n;
    Wg[e] = -A[a * ld + b];
  }
  for (int a = threadIdx.x; a < n; a += kT) w.pvec[o0 + a] = A[a * ld + n];
  if (threadIdx.x == 0) {
    w.yWy[i] = -A[n * ld + n];
    w.logdet[i] = logdet;
  }
  // scatter tau-weighted Psi^-1 and Psi^-1 y onto the union grid
  for (int k = 0; k < K; ++k) {
    const double tk = tau[(size_t)i * K + k];
    if (tk == 0.0) continue;  // exact zeros (one-hot initial tau) contribute nothing
    double* Ak = acc + k * strideA;
    for (int e = threadIdx.x; e < n * n; e += kT) {
      const int a = e / n, b = e - a * n;
      atomicAdd(Ak + (size_t)idx[a] * ldN + idx[b], -tk * A[a * ld + b]);
    }
    for (int a = threadIdx.x; a < n; a += kT) atomicAdd(wk + (size_t)k * ldN + idx[a], tk * A[a * ld + n]);
  }
}

// -------------------------------------------------------------------------------------------------
// E-step, part 2: responsibilities
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) indiv_tau_kernel(IndivData dd, int K, const double* __restrict__ mean,
                                                       const double* __restrict__ cov, int ldN, long long strideC,
                                                       const double* __restrict__ pi, IndivWork w,
                                                       double* __restrict__ tau_new, double* __restrict__ logL,
                                                       int fixed_tau) {
  extern __shared__ double sm[];
  __shared__ double red[3 * (kT / 32)];
  const int i = blockIdx.x;
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  if (n > dd.small_max) return;   // global-memory path (indiv_big.cu)
  double* W = sm;                 // n*n
  double* mu = W + (size_t)n * n;  // n
  double* p = mu + n;             // n
  double* lk = p + n;             // K : l_ik then tau_ik
  double* ak = lk + K;            // K : mu' W mu + sum W o C
  double* ek = ak + K;            // K : p' mu
  int* idx = (int*)(ek + K);
  const double* Wg = w.Winv + dd.woff[i];
  for (int e = threadIdx.x; e < n * n; e += kT) W[e] = Wg[e];
  for (int a = threadIdx.x; a < n; a += kT) {
    p[a] = w.pvec[o0 + a];
    idx[a] = dd.gidx[o0 + a];
  }
  __syncthreads();
  const double base = (double)n * kLog2Pi + w.logdet[i] + w.yWy[i];
  for (int k = 0; k < K; ++k) {
    const double* mk = mean + (size_t)k * ldN;
    const double* Ck = cov + k * strideC;
    for (int a = threadIdx.x; a < n; a += kT) mu[a] = mk[idx[a]];
    __syncthreads();
    double v[3] = {0.0, 0.0, 0.0};
    for (int e = threadIdx.x; e < n * n; e += kT) {
      const int a = e / n, b = e - a * n;
      const double wab = W[e];
      v[0] += wab * mu[a] * mu[b];
      v[1] += wab * Ck[(size_t)idx[a] * ldN + idx[b]];
    }
    for (int a = threadIdx.x; a < n; a += kT) v[2] += p[a] * mu[a];
    cta_sum<3>(v, red);
    if (threadIdx.x == 0) {
      // l_ik = -( 0.5 (n log 2pi + log|Psi| + z'Wz) + 0.5 sum(W o C_k) ),  z = y - mu   (CF.R:212-215, 746)
      lk[k] = -0.5 * (base - 2.0 * v[2] + v[0]) - 0.5 * v[1];
      ak[k] = v[0] + v[1];
      ek[k] = v[2];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double mx = lk[0];
    for (int k = 1; k < K; ++k) mx = fmax(mx, lk[k]);
    double den = 0.0;
    for (int k = 0; k < K; ++k) {
      logL[(size_t)i * K + k] = lk[k];
      const double u = pi[k] * exp(lk[k] - mx);  // CF.R:755-757: max over l only, pi multiplied after
      lk[k] = u;
      den += u;
    }
    double Fi = 0.5 * base;
    for (int k = 0; k < K; ++k) {
      // fixed_tau: the caller supplied mu_k_param$tau_i_k (M-step callbacks on an explicit posterior)
      const double tk = fixed_tau ? tau_new[(size_t)i * K + k] : lk[k] / den;
      lk[k] = tk;
      if (!fixed_tau) tau_new[(size_t)i * K + k] = tk;
      Fi += tk * (0.5 * ak[k] - ek[k]);
    }
    w.Fi[i] = Fi;
  }
  __syncthreads();
  // corr1 / corr2 with the NEW tau (they feed the M-step, CF.R:236-245)
  double* c2 = w.c2 + dd.woff[i];
  for (int e = threadIdx.x; e < n * n; e += kT) {
    const int a = e / n, b = e - a * n;
    double s = 0.0;
    for (int k = 0; k < K; ++k) {
      const double* mk = mean + (size_t)k * ldN;
      s += lk[k] * (mk[idx[a]] * mk[idx[b]] + cov[k * strideC + (size_t)idx[a] * ldN + idx[b]]);
    }
    c2[e] = s;
  }
  for (int a = threadIdx.x; a < n; a += kT) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += lk[k] * mean[(size_t)k * ldN + idx[a]];
    w.c1[o0 + a] = s;
  }
}

// Second version of the responsibilities kernel (round 2). The first one gathers C_k[idx, idx] twice from L2 (once for
// sum(W o C_k), once for corr2), both triangles, one cluster at a time with a CTA reduction and two barriers per cluster:
// ncu at C3 (profiles/r02_estep_ncu.txt, profiles/r02s2_tau2_ncu.txt): 5.6 long-scoreboard stall cycles per issue, 1.14 G L1 sectors per launch.
// Here every entry of the LOWER triangle of G_k = (m_k m_k' + C_k)[idx, idx] is gathered ONCE, for all clusters back to back
// (K independent L2 requests in flight per element and thread), reduced on the fly against W, and parked in shared memory;
// after ONE reduction over the CTA the new responsibilities are known and corr2 = sum_k tau_k G_k comes out of shared memory.
//   smem: G[K][L] (L = n (n + 1) / 2), mu[K][n], p[n], idx[n], small.   KB = compile-time bound on K (8).
constexpr int kTauKB = 8;
size_t tau2_smem_bytes(int nmax, int K) {
  const size_t L = (size_t)nmax * (nmax + 1) / 2;
  return sizeof(double) * (K * L + (size_t)K * nmax + nmax + 4 * kTauKB + 2 * kTauKB * (kT / 32) + 8) + sizeof(int) * (size_t)nmax + 16;
}

__global__ void __launch_bounds__(kT) indiv_tau2_kernel(IndivData dd, int K, const double* __restrict__ mean,
                                                        const double* __restrict__ cov, int ldN, long long strideC,
                                                        const double* __restrict__ pi, IndivWork w,
                                                        double* __restrict__ tau_new, double* __restrict__ logL,
                                                        int fixed_tau, int nmax) {
  extern __shared__ __align__(16) double sm[];
  const int i = blockIdx.x;
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  if (n > dd.small_max) return;   // global-memory path (indiv_big.cu)
  const int L = n * (n + 1) / 2;
  const size_t Lmax = (size_t)nmax * (nmax + 1) / 2;
  double* G = sm;                          // [K][Lmax]
  double* mu = G + (size_t)K * Lmax;       // [K][nmax]
  double* p = mu + (size_t)K * nmax;       // [nmax]
  double* lk = p + nmax;                   // [KB] tau_k
  double* red = lk + 4 * kTauKB;           // [2 KB][kT / 32]
  int* idx = (int*)(red + 2 * kTauKB * (kT / 32) + 8);
  const int tid = threadIdx.x;
  for (int a = tid; a < n; a += kT) {
    p[a] = w.pvec[o0 + a];
    idx[a] = dd.gidx[o0 + a];
  }
  __syncthreads();
  for (int e = tid; e < K * n; e += kT) {
    const int k = e / n, a = e - k * n;
    mu[(size_t)k * nmax + a] = mean[(size_t)k * ldN + idx[a]];
  }
  __syncthreads();
  const double* Wg = w.Winv + dd.woff[i];
  double acc[2 * kTauKB];                  // [k]: sum W o G_k (both triangles)   [KB + k]: p' mu_k
#pragma unroll
  for (int k = 0; k < 2 * kTauKB; ++k) acc[k] = 0.0;
  for (int e = tid; e < L; e += kT) {
    int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
    while ((a + 1) * (a + 2) / 2 <= e) ++a;
    while (a * (a + 1) / 2 > e) --a;
    const int b = e - a * (a + 1) / 2;
    const double wab = ((a == b) ? 1.0 : 2.0) * Wg[(size_t)a * n + b];
    const size_t gofs = (size_t)idx[a] * ldN + idx[b];
    double c[kTauKB];
#pragma unroll
    for (int k = 0; k < kTauKB; ++k) c[k] = (k < K) ? __ldg(cov + k * strideC + gofs) : 0.0;
#pragma unroll
    for (int k = 0; k < kTauKB; ++k) {
      if (k < K) {
        const double g = fma(mu[(size_t)k * nmax + a], mu[(size_t)k * nmax + b], c[k]);
        G[(size_t)k * Lmax + e] = g;
        acc[k] = fma(wab, g, acc[k]);
      }
    }
  }
  for (int a = tid; a < n; a += kT) {
#pragma unroll
    for (int k = 0; k < kTauKB; ++k)
      if (k < K) acc[kTauKB + k] = fma(p[a], mu[(size_t)k * nmax + a], acc[kTauKB + k]);
  }
  cta_sum_v<2 * kTauKB, kT>(acc, red);     // barriers inside: G is complete for every thread afterwards
  const double base = (double)n * kLog2Pi + w.logdet[i] + w.yWy[i];
  if (tid == 0) {
    double l[kTauKB];
    double mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kTauKB; ++k) {
      if (k < K) {
        // l_ik = -( 0.5 (n log 2pi + log|Psi| + z'Wz) + 0.5 sum(W o C_k) ),  z = y - mu   (CF.R:212-215, 746)
        l[k] = -0.5 * (base - 2.0 * acc[kTauKB + k] + acc[k]);
        logL[(size_t)i * K + k] = l[k];
        mx = fmax(mx, l[k]);
      }
    }
    double den = 0.0;
#pragma unroll
    for (int k = 0; k < kTauKB; ++k) {
      if (k < K) {
        l[k] = pi[k] * exp(l[k] - mx);      // CF.R:755-757: max over l only, pi multiplied after
        den += l[k];
      }
    }
    double Fi = 0.5 * base;
#pragma unroll
    for (int k = 0; k < kTauKB; ++k) {
      if (k < K) {
        // fixed_tau: the caller supplied mu_k_param$tau_i_k (M-step callbacks on an explicit posterior)
        const double tk = fixed_tau ? tau_new[(size_t)i * K + k] : l[k] / den;
        lk[k] = tk;
        if (!fixed_tau) tau_new[(size_t)i * K + k] = tk;
        Fi += tk * (0.5 * acc[k] - acc[kTauKB + k]);
      }
    }
    w.Fi[i] = Fi;
  }
  __syncthreads();
  // corr1 / corr2 with the NEW tau (they feed the M-step, CF.R:236-245)
  double tk[kTauKB];
#pragma unroll
  for (int k = 0; k < kTauKB; ++k) tk[k] = (k < K) ? lk[k] : 0.0;
  double* c2 = w.c2 + dd.woff[i];
  for (int e = tid; e < n * n; e += kT) {
    const int a = e / n, b = e - a * n;
    const int hi = a > b ? a : b, lo = a > b ? b : a;
    const int le = hi * (hi + 1) / 2 + lo;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kTauKB; ++k)
      if (k < K) s = fma(tk[k], G[(size_t)k * Lmax + le], s);
    c2[e] = s;
  }
  for (int a = tid; a < n; a += kT) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kTauKB; ++k)
      if (k < K) s = fma(tk[k], mu[(size_t)k * nmax + a], s);
    w.c1[o0 + a] = s;
  }
}

// -------------------------------------------------------------------------------------------------
// M-step objective + gradient kernels (device function in indiv_eval.cuh)
// -------------------------------------------------------------------------------------------------
size_t eval_smem_bytes(int nmax) { return sizeof(double) * eval_smem_doubles(nmax); }

// theta: [M][3] when theta_stride == 3, one shared [3] when theta_stride == 0.
// f[M], g[M][3] per individual; if sum4 != NULL also atomically accumulates (f, g) into sum4[0..3].
template <int E, int TT>
__global__ void __launch_bounds__(TT) indiv_eval_kernel(IndivData dd, IndivWork w, int nmax,
                                                        const double* __restrict__ theta, int theta_stride,
                                                        int want_grad, double* __restrict__ f,
                                                        double* __restrict__ g, double* __restrict__ sum4,
                                                        int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double out[4];
  const int i = blockIdx.x;
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  if (n > dd.small_max) return;   // global-memory path (indiv_big.cu)
  EvalSm s = eval_carve(sm, nmax);
  eval_load<TT>(s, n, dd.t + o0, dd.y + o0, w.c1 + o0, w.c2 + dd.woff[i]);
  const double* th = theta + (size_t)i * theta_stride;
  indiv_eval_v1<E, TT>(s, n, th[0], th[1], th[2], want_grad != 0, out, info);
  if (threadIdx.x == 0) {
    if (f) f[i] = out[0];
    if (g && want_grad) { g[3 * i] = out[1]; g[3 * i + 1] = out[2]; g[3 * i + 2] = out[3]; }
    if (sum4) {
      atomicAdd(sum4, out[0]);
      if (want_grad) { atomicAdd(sum4 + 1, out[1]); atomicAdd(sum4 + 2, out[2]); atomicAdd(sum4 + 3, out[3]); }
    }
  }
}

// Persistent M-step kernel for individual-specific hyper-parameters (CF.R:653-659). Each CTA pulls individuals
// from a global work queue and runs the WHOLE L-BFGS optimisation of theta_i in-kernel: corr2 stays in shared
// memory across all objective evaluations and there are no host round trips.
//
// SM reservation: the theta_k optimiser (shared theta_k, CF.R:666-669) does not depend on theta_i, so m_step_impl
// runs it concurrently on another stream. Its persistent slab inverse needs whole SMs (150 KB of shared memory per
// CTA), which this kernel would otherwise keep occupied for milliseconds; so the first `reserve` SMs that report in
// are left alone: CTAs that land on them exit immediately and their work flows to the others through the queue.
// ctl[0] = work queue, ctl[1] = number of SMs decided so far, ctl[2 + smid] = 0 undecided / 3 deciding / 1 work / 2 reserved.
template <int E, int TT>
__global__ void __launch_bounds__(TT) indiv_mstep_kernel(IndivData dd, IndivWork w, int nmax, int M,
                                                         const double* __restrict__ theta0,
                                                         double* __restrict__ theta_out, int maxit,
                                                         int* __restrict__ nfev, int* __restrict__ niter,
                                                         int* __restrict__ status, int* __restrict__ ctl,
                                                         int reserve, const int* __restrict__ order) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double out[4];
  __shared__ Lbfgs opt;
  __shared__ int st, cur;
  if (threadIdx.x == 0) {
    int role = 1;
    if (reserve > 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      int* flag = ctl + 2 + (smid & 1023);
      const int old = atomicCAS(flag, 0, 3);
      if (old == 0) {
        role = (atomicAdd(ctl + 1, 1) < reserve) ? 2 : 1;
        __threadfence();
        atomicExch(flag, role);
      } else {
        role = old;
        while (role == 3) role = *((volatile int*)flag);
      }
    }
    cur = (role == 2) ? -1 : 0;
  }
  tsync<TT>();
  if (cur < 0) return;
  EvalSm s = eval_carve(sm, nmax);
  int bad = 0;
  for (;;) {
    tsync<TT>();
    if (threadIdx.x == 0) cur = atomicAdd(ctl, 1);
    tsync<TT>();
    if (cur >= M) break;
    const int i = order ? order[cur] : cur;  // longest-expected-first order when the previous M-step's counts exist
    const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
    if (n > dd.small_max) continue;   // optimised by the host loop over dd.big_ids (api.cu)
    eval_load<TT>(s, n, dd.t + o0, dd.y + o0, w.c1 + o0, w.c2 + dd.woff[i]);
    if (threadIdx.x == 0) {
      lb_init(opt, 3, theta0 + 3 * (size_t)i, maxit);
      st = LB_EVAL;
    }
    tsync<TT>();
    while (true) {
      const double a = opt.x[0], b = opt.x[1], c = opt.x[2];
      indiv_eval_v1<E, TT>(s, n, a, b, c, true, out, &bad);
      if (threadIdx.x == 0) st = lb_advance(opt, out[0], out + 1);
      tsync<TT>();
      if (st != LB_EVAL) break;
    }
    if (threadIdx.x == 0) {
      theta_out[3 * (size_t)i] = opt.x[0];
      theta_out[3 * (size_t)i + 1] = opt.x[1];
      theta_out[3 * (size_t)i + 2] = opt.x[2];
      nfev[i] = opt.nfev;
      niter[i] = opt.iter;
      status[i] = st;
    }
  }
}

// The optimiser update is ~7 k SASS instructions of which ~900 run per call (thread 0 only): kept out of line so that
// the evaluation loop of the persistent kernels stays contiguous in the instruction cache.
__device__ __noinline__ int lb_advance_dev(Lbfgs& o, double f, const double* g) { return lb_advance(o, f, g); }

// ---- second design of the evaluator (indiv_eval2.cuh): same contracts as the two kernels above -------------------
// VER = 2: indiv_eval2.cuh (rank-1 sweep on 4 x 4 register blocks), VER = 3: indiv_eval3.cuh (blocked DMMA sweep)
template <int E, int TT, int NT, int VER>
__global__ void __launch_bounds__(TT, ((E > 1 || NT > 8) ? 256 : 512) / TT) indiv_eval2_kernel(IndivData dd, IndivWork w, int nmax,
                                                                   const double* __restrict__ theta, int theta_stride,
                                                                   int want_grad, double* __restrict__ f,
                                                                   double* __restrict__ g, double* __restrict__ sum4,
                                                                   int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double out[4];
  const int i = blockIdx.x;
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  if (n > dd.small_max) return;   // global-memory path (indiv_big.cu)
  EvalSm2<NT> s = eval2_carve<NT>(sm);
  eval2_load<TT, NT>(s, n, dd.t + o0, dd.y + o0, w.c1 + o0, w.c2 + dd.woff[i]);
  const double* th = theta + (size_t)i * theta_stride;
  if constexpr (VER == 3) {
    indiv_eval_v3<TT, NT>(s, n, th[0], th[1], th[2], want_grad != 0, out, info, i);
  } else {
    int bI[E], bJ[E];
    eval2_decode<E, TT>(n, bI, bJ);
    indiv_eval_v2<E, TT, NT>(s, n, bI, bJ, th[0], th[1], th[2], want_grad != 0, out, info, i);
  }
  if (threadIdx.x == 0) {
    if (f) f[i] = out[0];
    if (g && want_grad) { g[3 * i] = out[1]; g[3 * i + 1] = out[2]; g[3 * i + 2] = out[3]; }
    if (sum4) {
      atomicAdd(sum4, out[0]);
      if (want_grad) { atomicAdd(sum4 + 1, out[1]); atomicAdd(sum4 + 2, out[2]); atomicAdd(sum4 + 3, out[3]); }
    }
  }
}

// ctl words (kMstepCtlWords ints, cleared before the launch): [0] head of the work queue, [1] SMs that have reported,
// [2 + smid] role of the SM (0 undecided, 3 deciding, 1 shared, 2 reserved, 4 solo), [1026] tail counter of the queue,
// [1027 + smid] CTAs of a solo SM that have reported.
//
// Scheduling. The M-step of M individuals is bounded below by its longest optimisation (82 sequential evaluations at the
// C2 benchmark against a mean of 31), and an evaluation is 1.5x slower when four CTAs share an SM than when one has
// it alone. So the first `solo` SMs (after the `reserve` ones left to the theta_k chain) keep ONE working CTA, and the
// queue (longest-expected-first order) is popped from both ends: solo CTAs take the head, i.e. the `solo_items` longest
// optimisations, shared CTAs start behind them; each side moves on to the other's range when its own is exhausted.
template <int E, int TT, int NT, int VER>
__global__ void __launch_bounds__(TT, ((E > 1 || NT > 8) ? 256 : 512) / TT) indiv_mstep2_kernel(IndivData dd, IndivWork w, int nmax, int M,
                                                                    const double* __restrict__ theta0,
                                                                    double* __restrict__ theta_out, int maxit,
                                                                    int* __restrict__ nfev, int* __restrict__ niter,
                                                                    int* __restrict__ status, int* __restrict__ ctl,
                                                                    int reserve, const int* __restrict__ order,
                                                                    int solo, int solo_items, double* __restrict__ f_out,
                                                                    int solo_ctas) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double out[4];
  __shared__ Lbfgs opt;
  __shared__ int st, cur, my_role;
  if (threadIdx.x == 0) {
    int role = 1;
    if (reserve > 0 || solo > 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      smid &= 1023;
      int* flag = ctl + 2 + smid;
      const int old = atomicCAS(flag, 0, 3);
      if (old == 0) {
        const int ord = atomicAdd(ctl + 1, 1);
        role = (ord < reserve) ? 2 : (ord < reserve + solo ? 4 : 1);
        __threadfence();
        atomicExch(flag, role);
      } else {
        role = old;
        while (role == 3) role = *((volatile int*)flag);
      }
      if (role == 4 && atomicAdd(ctl + 1027 + smid, 1) >= solo_ctas) role = 2;   // a solo SM keeps its first solo_ctas CTAs only
    }
    my_role = role;
    cur = (role == 2) ? -1 : 0;
  }
  tsync<TT>();
  if (cur < 0) return;
  const bool is_solo = (my_role == 4);
  EvalSm2<NT> s = eval2_carve<NT>(sm);
  int bad = 0;
  for (;;) {
    tsync<TT>();
    if (threadIdx.x == 0) {
      int i;
      if (is_solo) {
        i = atomicAdd(ctl, 1);
        if (i >= solo_items) { i = solo_items + atomicAdd(ctl + 1026, 1); if (i >= M) i = -1; }
      } else {
        i = solo_items + atomicAdd(ctl + 1026, 1);
        if (i >= M) { i = atomicAdd(ctl, 1); if (i >= solo_items) i = -1; }
      }
      cur = i;
    }
    tsync<TT>();
    if (cur < 0) break;
    const int slot = cur;
    const int i = order ? order[slot] : slot;  // longest-expected-first order when the previous M-step's counts exist
    const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
    if (n > dd.small_max) continue;   // optimised by the host loop over dd.big_ids (api.cu)
    eval2_load<TT, NT>(s, n, dd.t + o0, dd.y + o0, w.c1 + o0, w.c2 + dd.woff[i]);
    int bI[E], bJ[E];
    if constexpr (VER != 3) eval2_decode<E, TT>(n, bI, bJ);
    if (threadIdx.x == 0) {
      lb_init(opt, 3, theta0 + 3 * (size_t)i, maxit);
      st = LB_EVAL;
    }
    tsync<TT>();
    while (true) {
      const double a = opt.x[0], b = opt.x[1], c = opt.x[2];
      if constexpr (VER == 3) indiv_eval_v3<TT, NT>(s, n, a, b, c, true, out, &bad, slot);
      else indiv_eval_v2<E, TT, NT>(s, n, bI, bJ, a, b, c, true, out, &bad, slot);
      if (threadIdx.x == 0) st = lb_advance_dev(opt, out[0], out + 1);
      tsync<TT>();
      if (st != LB_EVAL) break;
    }
    if (threadIdx.x == 0) {
      theta_out[3 * (size_t)i] = opt.x[0];
      theta_out[3 * (size_t)i + 1] = opt.x[1];
      theta_out[3 * (size_t)i + 2] = opt.x[2];
      nfev[i] = opt.nfev;
      niter[i] = opt.iter;
      status[i] = st;
      // objective at the returned point (the ELBO at the new hp needs exactly this); a start point without a finite
      // objective poisons the sum on purpose
      if (f_out) f_out[i] = (st == LB_NONFINITE) ? nan("") : opt.f;
    }
  }
}

// E-step, part 1 on the evaluator machinery: the same register-blocked sweep (or blocked DMMA sweep) that the M-step uses
// leaves W = Psi_i^-1, p = W y and y'Wy in shared memory (objective-only call with corr1 = 0 and corr2 = 0, for which
// F = (n log 2pi + log|Psi| + y'Wy) / 2 gives the log determinant back); then the tau-weighted scatter onto A_k, w_k as
// in indiv_factor_scatter_kernel, which remains the path for n_i > 80.
template <int E, int TT, int NT, int VER>
__global__ void __launch_bounds__(TT, ((E > 1 || NT > 8) ? 256 : 512) / TT) indiv_factor2_kernel(
    IndivData dd, const double* __restrict__ theta_i, const double* __restrict__ tau, int K, double* __restrict__ acc,
    int ldN, long long strideA, double* __restrict__ wk, IndivWork w, int* __restrict__ info) {
  using G = Eval2Geo<NT>;
  extern __shared__ __align__(16) double sm[];
  __shared__ double out[4];
  const int i = blockIdx.x;
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  if (n > dd.small_max) return;   // global-memory path (indiv_big.cu)
  EvalSm2<NT> s = eval2_carve<NT>(sm);
  {
    double2* z = reinterpret_cast<double2*>(s.W);
    for (int e = threadIdx.x; e < G::MAT; e += TT) z[e] = make_double2(0.0, 0.0);   // W and S (= 0 here)
    for (int e = threadIdx.x; e < 64; e += TT) s.red[64 + e] = 0.0;
    for (int a = threadIdx.x; a < n; a += TT) {
      s.t[a] = dd.t[o0 + a];
      s.y[a] = dd.y[o0 + a];
      s.c1[a] = 0.0;
    }
  }
  tsync<TT>();
  const double th1 = theta_i[3 * i], th2 = theta_i[3 * i + 1], sg = theta_i[3 * i + 2];
  if constexpr (VER == 3) {
    indiv_eval_v3<TT, NT>(s, n, th1, th2, sg, false, out, info, i);
  } else {
    int bI[E], bJ[E];
    eval2_decode<E, TT>(n, bI, bJ);
    indiv_eval_v2<E, TT, NT>(s, n, bI, bJ, th1, th2, sg, false, out, info, i);
  }
  // after the call: W in s.W (both triangles, ld = LDM), p in s.col, y'Wy in s.corner[0]
  int* idx = reinterpret_cast<int*>(s.S);   // the S buffer is free from here on
  for (int a = threadIdx.x; a < n; a += TT) idx[a] = dd.gidx[o0 + a];
  double* Wg = w.Winv + dd.woff[i];
  for (int e = threadIdx.x; e < n * n; e += TT) {
    const int a = e / n, b = e - a * n;
    Wg[e] = s.W[a * G::LDM + b];
  }
  for (int a = threadIdx.x; a < n; a += TT) w.pvec[o0 + a] = s.col[a];
  if (threadIdx.x == 0) {
    const double yWy = s.corner[0];
    w.yWy[i] = yWy;
    w.logdet[i] = 2.0 * out[0] - (double)n * kLog2Pi - yWy;
  }
  tsync<TT>();
  for (int k = 0; k < K; ++k) {
    const double tk = tau[(size_t)i * K + k];
    if (tk == 0.0) continue;  // exact zeros (one-hot initial tau) contribute nothing
    double* Ak = acc + k * strideA;
    for (int e = threadIdx.x; e < n * n; e += TT) {
      const int a = e / n, b = e - a * n;
      atomicAdd(Ak + (size_t)idx[a] * ldN + idx[b], tk * s.W[a * G::LDM + b]);
    }
    for (int a = threadIdx.x; a < n; a += TT) atomicAdd(wk + (size_t)k * ldN + idx[a], tk * s.col[a]);
  }
}

// -------------------------------------------------------------------------------------------------
// host launchers
// -------------------------------------------------------------------------------------------------
size_t factor_smem_bytes(int nmax) {
  const int ntot = nmax + 1, ld = ntot | 1;
  return sizeof(double) * ((size_t)ntot * ld + 2 * (size_t)nmax) + sizeof(int) * (size_t)nmax + 16;
}
size_t tau_smem_bytes(int nmax, int K) {
  return sizeof(double) * ((size_t)nmax * nmax + 2 * (size_t)nmax + 3 * (size_t)K) + sizeof(int) * (size_t)nmax + 16;
}

static cudaError_t set_smem(const void* fn, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_tau(cudaStream_t s, const IndivData& dd, int M, int nmax, int K, const double* mean,
                       const double* cov, int ldN, long long strideC, const double* pi, const IndivWork& w,
                       double* tau_new, double* logL, int fixed_tau) {
  if (dd.nbig > 0) {
    cudaError_t eb = launch_big_tau(s, dd, K, mean, cov, ldN, strideC, pi, w, tau_new, logL, fixed_tau);
    if (eb != cudaSuccess || dd.nbig == M) return eb;
  }
  static const bool v1 = getenv("MCB_TAU_V1") != nullptr;
  const size_t sm2 = tau2_smem_bytes(nmax, K);
  if (!v1 && K <= kTauKB && sm2 <= 200 * 1024) {
    cudaError_t e2 = set_smem((const void*)indiv_tau2_kernel, sm2);
    if (e2 != cudaSuccess) return e2;
    indiv_tau2_kernel<<<M, kT, sm2, s>>>(dd, K, mean, cov, ldN, strideC, pi, w, tau_new, logL, fixed_tau, nmax);
    return cudaGetLastError();
  }
  const size_t sm = tau_smem_bytes(nmax, K);
  cudaError_t e = set_smem((const void*)indiv_tau_kernel, sm);
  if (e != cudaSuccess) return e;
  indiv_tau_kernel<<<M, kT, sm, s>>>(dd, K, mean, cov, ldN, strideC, pi, w, tau_new, logL, fixed_tau);
  return cudaGetLastError();
}

template <class K>
static cudaError_t set_smem_t(K fn, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Threads per individual: 64 while the lower block triangle has at most 64 blocks (n <= 41), otherwise 128
// (measured: n = 30 -> 3.6 ms per 1e5 evaluations with 64 threads vs 4.0 with 128 and 5.5 with 32; n = 50 -> 5.4 ms per
// M-step with 128 vs 8.1 with 64). MCB_EVAL_THREADS=32|64|128 overrides (experiments).
static void eval_shape(int nmax, int* TT, int* E) {
  const int nblk = eval_num_blocks(nmax);
  int tt = (nblk <= 64) ? 64 : 128;
  static const char* ov = getenv("MCB_EVAL_THREADS");
  if (ov) tt = atoi(ov);
  if (tt != 32 && tt != 64 && tt != 128) tt = 128;
  while (tt < 128 && (nblk + tt - 1) / tt > 4) tt *= 2;
  *TT = tt;
  *E = (nblk + tt - 1) / tt;
}

#define MCB_DISPATCH_EVAL(MACRO)                       \
  do {                                                 \
    if (TT == 32) {                                    \
      if (E <= 1) MACRO(1, 32);                        \
      else if (E == 2) MACRO(2, 32);                   \
      else if (E <= 4) MACRO(4, 32);                   \
      else return cudaErrorInvalidValue;               \
    } else if (TT == 64) {                             \
      if (E <= 1) MACRO(1, 64);                        \
      else if (E == 2) MACRO(2, 64);                   \
      else if (E <= 4) MACRO(4, 64);                   \
      else return cudaErrorInvalidValue;               \
    } else {                                           \
      if (E <= 1) MACRO(1, 128);                       \
      else if (E == 2) MACRO(2, 128);                  \
      else if (E <= 4) MACRO(4, 128);                  \
      else return cudaErrorInvalidValue;               \
    }                                                  \
  } while (0)

// second design: TT = 64 while the lower block triangle fits 64 threads, else 128 (E = 2 above n = 56); NT = tile bucket
// E == 0 selects the third design (needs 8 NT >= nmax + 2): TT = 64 for NT = 4, else 128.
static bool eval2_shape(int nmax, int* TT, int* E, int* NT) {
  static const bool v1 = getenv("MCB_EVAL_V1") != nullptr;
  static const bool v2 = getenv("MCB_EVAL_V2") != nullptr;
  static const bool v3 = getenv("MCB_EVAL_V3") != nullptr;
  if (v1 || nmax > 80) return false;
  // measured (tools/dbg/eval_prof.cu, cycles per evaluation at full occupancy): n = 50: 87 k (second design) vs 116 k
  // (third), n = 30: 64 k vs 64 k, n = 72: 123 k vs 105 k -> the blocked DMMA sweep only pays for the largest bucket
  if (!v2 && nmax <= 78 && (v3 || nmax > 56)) {
    *NT = (nmax <= 30) ? 4 : (nmax <= 54 ? 7 : 10);
    *TT = (*NT == 4) ? 64 : 128;
    *E = 0;
    return true;
  }
  const int nbk = (nmax + 2 + 3) / 4, nblk = nbk * (nbk + 1) / 2;
  *TT = (nblk <= 64) ? 64 : 128;
  *E = (nblk + *TT - 1) / *TT;
  *NT = (nmax <= 32) ? 4 : (nmax <= 56 ? 7 : 10);
  return true;
}
static size_t eval2_smem_bytes(int NT) {
  return sizeof(double) * (size_t)(NT == 4 ? Eval2Geo<4>::DOUBLES : NT == 7 ? Eval2Geo<7>::DOUBLES : Eval2Geo<10>::DOUBLES);
}

#define MCB_DISPATCH_EVAL2(MACRO)                                   \
  do {                                                              \
    if (TT == 64 && E == 0 && NT == 4) MACRO(1, 64, 4, 3);          \
    else if (TT == 128 && E == 0 && NT == 7) MACRO(1, 128, 7, 3);   \
    else if (TT == 128 && E == 0 && NT == 10) MACRO(1, 128, 10, 3); \
    else if (TT == 64 && E == 1 && NT == 4) MACRO(1, 64, 4, 2);     \
    else if (TT == 64 && E == 1 && NT == 7) MACRO(1, 64, 7, 2);     \
    else if (TT == 128 && E == 1 && NT == 7) MACRO(1, 128, 7, 2);   \
    else if (TT == 128 && E == 2 && NT == 10) MACRO(2, 128, 10, 2); \
    else return cudaErrorInvalidValue;                              \
  } while (0)

cudaError_t launch_factor_scatter(cudaStream_t s, const IndivData& dd, int M, int nmax, const double* theta_i,
                                  const double* tau, int K, double* acc, int ldN, long long strideA, double* wk,
                                  const IndivWork& w, int* info) {
  if (dd.nbig > 0) {
    cudaError_t eb = launch_big_factor(s, dd, theta_i, tau, K, acc, ldN, strideA, wk, w, info);
    if (eb != cudaSuccess || dd.nbig == M) return eb;
  }
  {
    int TT, E, NT;
    static const bool old_path = getenv("MCB_FACTOR_V1") != nullptr;
    if (!old_path && eval2_shape(nmax, &TT, &E, &NT)) {
      const size_t sm2 = eval2_smem_bytes(NT);
      cudaError_t e2;
#define MCB_LAUNCH_FACTOR2(EE, T_, N_, V_)                                                                              \
  do {                                                                                                                 \
    if ((e2 = set_smem_t(indiv_factor2_kernel<EE, T_, N_, V_>, sm2)) != cudaSuccess) return e2;                        \
    indiv_factor2_kernel<EE, T_, N_, V_><<<M, T_, sm2, s>>>(dd, theta_i, tau, K, acc, ldN, strideA, wk, w, info);      \
  } while (0)
      MCB_DISPATCH_EVAL2(MCB_LAUNCH_FACTOR2);
#undef MCB_LAUNCH_FACTOR2
      return cudaGetLastError();
    }
  }
  const size_t sm = factor_smem_bytes(nmax);
  cudaError_t e = set_smem((const void*)indiv_factor_scatter_kernel, sm);
  if (e != cudaSuccess) return e;
  indiv_factor_scatter_kernel<<<M, kT, sm, s>>>(dd, theta_i, tau, K, acc, ldN, strideA, wk, w, info);
  return cudaGetLastError();
}

cudaError_t launch_indiv_eval(cudaStream_t s, const IndivData& dd, int M, int nmax, const IndivWork& w,
                              const double* theta, int theta_stride, int want_grad, double* f, double* g,
                              double* sum4, int* info) {
  if (dd.nbig > 0) {
    cudaError_t eb = launch_big_eval(s, dd, w, theta, theta_stride, 0, nullptr, want_grad, f, g, sum4, info);
    if (eb != cudaSuccess || dd.nbig == M) return eb;
  }
  {
    int TT, E, NT;
    if (eval2_shape(nmax, &TT, &E, &NT)) {
      const size_t sm2 = eval2_smem_bytes(NT);
      cudaError_t e;
#define MCB_LAUNCH_EVAL2(EE, T_, N_, V_)                                                                                   \
  do {                                                                                                                     \
    if ((e = set_smem_t(indiv_eval2_kernel<EE, T_, N_, V_>, sm2)) != cudaSuccess) return e;                                \
    indiv_eval2_kernel<EE, T_, N_, V_><<<M, T_, sm2, s>>>(dd, w, nmax, theta, theta_stride, want_grad, f, g, sum4, info);  \
  } while (0)
      MCB_DISPATCH_EVAL2(MCB_LAUNCH_EVAL2);
#undef MCB_LAUNCH_EVAL2
      return cudaGetLastError();
    }
  }
  const size_t sm = eval_smem_bytes(nmax);
  int TT, E;
  eval_shape(nmax, &TT, &E);
  cudaError_t e;
#define MCB_LAUNCH_EVAL(EE, T_)                                                                                  \
  do {                                                                                                           \
    if ((e = set_smem_t(indiv_eval_kernel<EE, T_>, sm)) != cudaSuccess) return e;                                \
    indiv_eval_kernel<EE, T_><<<M, T_, sm, s>>>(dd, w, nmax, theta, theta_stride, want_grad, f, g, sum4, info);  \
  } while (0)
  MCB_DISPATCH_EVAL(MCB_LAUNCH_EVAL);
#undef MCB_LAUNCH_EVAL
  return cudaGetLastError();
}

cudaError_t launch_indiv_mstep(cudaStream_t s, const IndivData& dd, int M, int nmax, const IndivWork& w,
                               const double* theta0, double* theta_out, int maxit, int* nfev, int* niter,
                               int* status, int* ctl, int reserve_sms, const int* order, double* f_out,
                               int* f_out_written) {
  if (f_out_written) *f_out_written = 0;
  const size_t sm = eval_smem_bytes(nmax);
  int TT, E;
  eval_shape(nmax, &TT, &E);
  cudaError_t e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if ((e = cudaMemsetAsync(ctl, 0, sizeof(int) * kMstepCtlWords, s)) != cudaSuccess) return e;
  {
    int NT;
    if (eval2_shape(nmax, &TT, &E, &NT)) {
      const size_t sm2 = eval2_smem_bytes(NT);
      static const int cap_per_sm = getenv("MCB_MSTEP_CTAS") ? atoi(getenv("MCB_MSTEP_CTAS")) : 0;
      static const int solo_sms = getenv("MCB_SOLO_SMS") ? atoi(getenv("MCB_SOLO_SMS")) : -1;
      static const int solo_mult = getenv("MCB_SOLO_MULT") ? atoi(getenv("MCB_SOLO_MULT")) : 1;
      static const int solo_ctas = getenv("MCB_SOLO_CTAS") ? std::max(1, atoi(getenv("MCB_SOLO_CTAS"))) : 1;
#define MCB_LAUNCH_MSTEP2(EE, T_, N_, V_)                                                                              \
  do {                                                                                                                 \
    if ((e = set_smem_t(indiv_mstep2_kernel<EE, T_, N_, V_>, sm2)) != cudaSuccess) return e;                           \
    int per_sm = 1;                                                                                                    \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, indiv_mstep2_kernel<EE, T_, N_, V_>, T_, sm2);              \
    if (per_sm < 1) per_sm = 1;                                                                                        \
    if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;                                                    \
    int grid = per_sm * sms;                                                                                           \
    int solo = 0, solo_items = 0;                                                                                      \
    if (order && per_sm > 1 && M >= 2 * sms) {                                                                         \
      /* solo SMs pay off while the optimisations outnumber the CTA slots only a few times (tail-bound regime) */      \
      solo = solo_sms >= 0 ? solo_sms : ((long long)M < 8LL * per_sm * sms ? (sms - reserve_sms + 7) / 14 : 0);        \
      if (solo > sms - reserve_sms - 1) solo = sms - reserve_sms - 1;                                                  \
      solo_items = solo_mult * solo * solo_ctas;                                                                       \
    }                                                                                                                  \
    if (grid > M && reserve_sms == 0 && solo == 0) grid = M;                                                           \
    indiv_mstep2_kernel<EE, T_, N_, V_><<<grid, T_, sm2, s>>>(dd, w, nmax, M, theta0, theta_out, maxit, nfev, niter,   \
                                                          status, ctl, reserve_sms, order, solo, solo_items, f_out,    \
                                                          solo_ctas);                                                  \
    if (f_out_written) *f_out_written = f_out ? 1 : 0;                                                                 \
  } while (0)
      MCB_DISPATCH_EVAL2(MCB_LAUNCH_MSTEP2);
#undef MCB_LAUNCH_MSTEP2
      return cudaGetLastError();
    }
    eval_shape(nmax, &TT, &E);
  }
#define MCB_LAUNCH_MSTEP(EE, T_)                                                                                  \
  do {                                                                                                            \
    if ((e = set_smem_t(indiv_mstep_kernel<EE, T_>, sm)) != cudaSuccess) return e;                                \
    int per_sm = 1;                                                                                               \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, indiv_mstep_kernel<EE, T_>, T_, sm);                   \
    if (per_sm < 1) per_sm = 1;                                                                                   \
    int grid = per_sm * sms;                                                                                      \
    if (grid > M && reserve_sms == 0) grid = M;                                                                   \
    indiv_mstep_kernel<EE, T_><<<grid, T_, sm, s>>>(dd, w, nmax, M, theta0, theta_out, maxit, nfev, niter, status, \
                                                    ctl, reserve_sms, order);                                     \
  } while (0)
  MCB_DISPATCH_EVAL(MCB_LAUNCH_MSTEP);
#undef MCB_LAUNCH_MSTEP
  return cudaGetLastError();
}

}  // namespace mcb
