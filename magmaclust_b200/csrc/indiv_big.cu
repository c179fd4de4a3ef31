// Per-individual kernels for individuals with MORE observations than the shared-memory kernels of indiv.cu hold
// (n_i > kSmallMax = 80). The reference has no size limit (kern_to_inv db branch CF.R:153-190, logL_clust_multi_GP
// CF.R:218-248, gr_clust_multi_GP CF.R:477-508 work on whatever filter(ID == i) returns), so these individuals take the
// same four operations with their matrices in a global-memory workspace (L2 resident at these sizes: (n + 2)^2 + n^2
// doubles, 0.64 MB at n = 200): one CTA of 256 threads each, plain FP64 FMAs, no tile buckets. Same algorithm as the
// first design of indiv.cu -- symmetric sweep of the bordered matrix [[Psi, y, c1], [., 0]] -- so results agree with the
// shared-memory path to rounding. A data set normally has none or few of them; the path is written for any n, not for speed.
#include "common.cuh"
#include "indiv.cuh"

namespace mcb {

constexpr int kTB = 256;

template <int NV>
__device__ __forceinline__ void big_cta_sum(double (&v)[NV], double* red /* [NV][kTB / 32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) red[q * (kTB / 32) + warp] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kTB / 32; ++w) s += red[q * (kTB / 32) + w];
    v[q] = s;
  }
}

// bordered SE covariance (kern_to_cov, CF.R:61-86: exp(th1 - exp(-th2)/2 d^2), diagonal exp(th1) + sigma^2) in global memory
__device__ void big_build(double* A, int ld, int n, int nbord, const double* t, const double* b0, const double* b1,
                          double th1, double th2, double sg) {
  const double hl = exp(-th2) / 2.0;
  const double dg = exp(th1) + sg * sg;
  const int ntot = n + nbord;
  for (int e = threadIdx.x; e < ntot * ntot; e += kTB) {
    const int a = e / ntot, b = e - a * ntot;
    double v = 0.0;
    if (a < n && b < n) {
      const double d = t[a] - t[b];
      v = (a == b) ? dg : exp(th1 - hl * (d * d));
    } else if (a < n) {
      v = (b == n) ? b0[a] : b1[a];
    } else if (b < n) {
      v = (a == n) ? b0[b] : b1[b];
    }
    A[(size_t)a * ld + b] = v;
  }
}

// Sweep of pivots 0..n-1 of the symmetric ntot x ntot matrix in GLOBAL memory (coherent within the CTA across __syncthreads;
// no read-only-cache loads on A). Returns the sum of the log pivots; *bad is set when a pivot is not positive.
__device__ double big_sweep(double* A, int n, int ntot, int ld, int* bad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double logdet = 0.0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double d = A[(size_t)k * ld + k];
    const double inv_d = 1.0 / d;
    logdet += log(d);
    if (!(d > 0.0) && threadIdx.x == 0) *bad = 1;
    const double* rk = A + (size_t)k * ld;
    for (int a = warp; a < ntot; a += kTB / 32) {
      if (a == k) continue;
      double* ra = A + (size_t)a * ld;
      const double ca = ra[k] * inv_d;
      for (int b = lane; b < ntot; b += 32)
        if (b != k) ra[b] -= ca * rk[b];
    }
    __syncthreads();
    for (int b = threadIdx.x; b < ntot; b += kTB) {
      if (b == k) continue;
      const double v = A[(size_t)k * ld + b] * inv_d;
      A[(size_t)k * ld + b] = v;
      A[(size_t)b * ld + k] = v;
    }
    if (threadIdx.x == 0) A[(size_t)k * ld + k] = -inv_d;
    __syncthreads();
  }
  return logdet;
}

// kern_to_inv (CF.R:153-190) + update_inv_VEM (CF.R:690-707) + update_mean_VEM (CF.R:709-731)
__global__ void __launch_bounds__(kTB) big_factor_kernel(IndivData dd, const double* __restrict__ theta_i,
                                                         const double* __restrict__ tau, int K, double* acc, int ldN,
                                                         long long strideA, double* wk, IndivWork w, int* info) {
  const int i = dd.big_ids[blockIdx.x];
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  const int ntot = n + 1, ld = n + 2;
  double* A = dd.big_ws + dd.big_wsoff[blockIdx.x];
  const double* t = dd.t + o0;
  const double* y = dd.y + o0;
  const int* idx = dd.gidx + o0;
  big_build(A, ld, n, 1, t, y, y, theta_i[3 * (size_t)i], theta_i[3 * (size_t)i + 1], theta_i[3 * (size_t)i + 2]);
  const double logdet = big_sweep(A, n, ntot, ld, info);
  double* Wg = w.Winv + dd.woff[i];
  for (int e = threadIdx.x; e < n * n; e += kTB) {
    const int a = e / n, b = e - a * n;
    Wg[e] = -A[(size_t)a * ld + b];
  }
  for (int a = threadIdx.x; a < n; a += kTB) w.pvec[o0 + a] = A[(size_t)a * ld + n];
  if (threadIdx.x == 0) {
    w.yWy[i] = -A[(size_t)n * ld + n];
    w.logdet[i] = logdet;
  }
  for (int k = 0; k < K; ++k) {
    const double tk = tau[(size_t)i * K + k];
    if (tk == 0.0) continue;
    double* Ak = acc + k * strideA;
    for (int e = threadIdx.x; e < n * n; e += kTB) {
      const int a = e / n, b = e - a * n;
      atomicAdd(Ak + (size_t)idx[a] * ldN + idx[b], -tk * A[(size_t)a * ld + b]);
    }
    for (int a = threadIdx.x; a < n; a += kTB) atomicAdd(wk + (size_t)k * ldN + idx[a], tk * A[(size_t)a * ld + n]);
  }
}

// update_tau_i_k_VEM (CF.R:733-762) via logL_GP_mod (CF.R:194-216); corr1 / corr2 (CF.R:236-245) and F_i with the new tau
__global__ void __launch_bounds__(kTB) big_tau_kernel(IndivData dd, int K, const double* __restrict__ mean,
                                                      const double* __restrict__ cov, int ldN, long long strideC,
                                                      const double* __restrict__ pi, IndivWork w, double* tau_new,
                                                      double* __restrict__ logL, int fixed_tau) {
  extern __shared__ double sm[];   // lk[K], ak[K], ek[K]
  __shared__ double red[3 * (kTB / 32)];
  const int i = dd.big_ids[blockIdx.x];
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  double* lk = sm;
  double* ak = lk + K;
  double* ek = ak + K;
  const double* Wg = w.Winv + dd.woff[i];
  const double* p = w.pvec + o0;
  const int* idx = dd.gidx + o0;
  const double base = (double)n * kLog2Pi + w.logdet[i] + w.yWy[i];
  for (int k = 0; k < K; ++k) {
    const double* mk = mean + (size_t)k * ldN;
    const double* Ck = cov + k * strideC;
    double v[3] = {0.0, 0.0, 0.0};
    for (int e = threadIdx.x; e < n * n; e += kTB) {
      const int a = e / n, b = e - a * n;
      const double wab = Wg[e];
      v[0] += wab * mk[idx[a]] * mk[idx[b]];
      v[1] += wab * Ck[(size_t)idx[a] * ldN + idx[b]];
    }
    for (int a = threadIdx.x; a < n; a += kTB) v[2] += p[a] * mk[idx[a]];
    big_cta_sum<3>(v, red);
    if (threadIdx.x == 0) {
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
      const double tk = fixed_tau ? tau_new[(size_t)i * K + k] : lk[k] / den;
      lk[k] = tk;
      if (!fixed_tau) tau_new[(size_t)i * K + k] = tk;
      Fi += tk * (0.5 * ak[k] - ek[k]);
    }
    w.Fi[i] = Fi;
  }
  __syncthreads();
  double* c2 = w.c2 + dd.woff[i];
  for (int e = threadIdx.x; e < n * n; e += kTB) {
    const int a = e / n, b = e - a * n;
    double s = 0.0;
    for (int k = 0; k < K; ++k) {
      const double* mk = mean + (size_t)k * ldN;
      s += lk[k] * (mk[idx[a]] * mk[idx[b]] + cov[k * strideC + (size_t)idx[a] * ldN + idx[b]]);
    }
    c2[e] = s;
  }
  for (int a = threadIdx.x; a < n; a += kTB) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += lk[k] * mean[(size_t)k * ldN + idx[a]];
    w.c1[o0 + a] = s;
  }
}

// logL_clust_multi_GP + gr_clust_multi_GP (CF.R:218-248, 477-508) at a trial theta. With W = Psi^-1, p = W y, q = W c1,
// u = p - 2 q, S = corr2, H = W S W:
//   f  = 0.5 (n log 2pi + log|Psi| + y'Wy) - y'W c1 + 0.5 sum(W o S)
//   g1 = -0.5 (sum(pu' o K) + sum(H o K) - sum(W o K)),  K = dPsi/dth1 (deriv_hp1, CF.R:436-439)
//   g2 = the same with dPsi/dth2 = exp(-th2)/2 d^2 K     (deriv_hp2, CF.R:441-446)
//   g3 = -sigma (p'u + tr H - tr W)                       (dPsi/dsigma = 2 sigma I, CF.R:500-503)
__global__ void __launch_bounds__(kTB) big_eval_kernel(IndivData dd, IndivWork w, const double* __restrict__ theta,
                                                       int theta_stride, int by_slot, const int* __restrict__ active,
                                                       int want_grad, double* f, double* g, double* sum4, int* info) {
  __shared__ double red[11 * (kTB / 32)];
  const int j = blockIdx.x;
  if (active && !active[j]) return;
  const int i = dd.big_ids[j];
  const int slot = by_slot ? j : i;
  const int o0 = dd.off[i], n = dd.off[i + 1] - o0;
  const int ntot = n + 2, ld = n + 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* A = dd.big_ws + dd.big_wsoff[j];
  double* X = A + (size_t)ntot * ld;
  const double* t = dd.t + o0;
  const double* S = w.c2 + dd.woff[i];
  const double* th = theta + (size_t)slot * theta_stride;
  const double th1 = th[0], th2 = th[1], sg = th[2];
  const double hl = exp(-th2) / 2.0, e1 = exp(th1);
  big_build(A, ld, n, 2, t, dd.y + o0, w.c1 + o0, th1, th2, sg);
  int bad = 0;
  const double logdet = big_sweep(A, n, ntot, ld, &bad);   // every thread sees the same pivots
  // after the sweep: A[0:n, 0:n] = -W, A[:, n] = p, A[:, n + 1] = q, A[n][n] = -y'Wy, A[n + 1][n] = -y'W c1
  double v[11];
#pragma unroll
  for (int q = 0; q < 11; ++q) v[q] = 0.0;
  for (int e = threadIdx.x; e < n * n; e += kTB) {
    const int a = e / n, b = e - a * n;
    const double wab = -A[(size_t)a * ld + b];
    v[0] += wab * S[e];
    if (want_grad) {
      const double d = t[a] - t[b];
      const double kv = (a == b) ? e1 : exp(th1 - hl * (d * d));
      const double k2 = hl * (d * d) * kv;
      const double pu = A[(size_t)a * ld + n] * (A[(size_t)b * ld + n] - 2.0 * A[(size_t)b * ld + n + 1]);
      v[1] += wab * kv;
      v[2] += wab * k2;
      v[4] += pu * kv;
      v[5] += pu * k2;
      if (a == b) {
        v[3] += wab;
        v[6] += pu;
      }
    }
  }
  if (want_grad) {
    // X = W S (S symmetric): one warp per row, lanes over the columns, the contraction index runs over rows of S
    for (int a = warp; a < n; a += kTB / 32) {
      const double* wa = A + (size_t)a * ld;
      for (int b0 = 0; b0 < n; b0 += 32) {
        const int b = b0 + lane;
        double s = 0.0;
        if (b < n)
          for (int k = 0; k < n; ++k) s = fma(-wa[k], S[(size_t)k * n + b], s);
        if (b < n) X[(size_t)a * n + b] = s;
      }
    }
    __syncthreads();
    // H = X W reduced on the fly against dPsi/dth1, dPsi/dth2 and I
    for (int a = warp; a < n; a += kTB / 32) {
      const double* xa = X + (size_t)a * n;
      for (int b0 = 0; b0 < n; b0 += 32) {
        const int b = b0 + lane;
        if (b < n) {
          double hv = 0.0;
          for (int k = 0; k < n; ++k) hv = fma(-xa[k], A[(size_t)k * ld + b], hv);
          const double d = t[a] - t[b];
          const double kv = (a == b) ? e1 : exp(th1 - hl * (d * d));
          v[9] += hv * kv;
          v[10] += hv * (hl * (d * d) * kv);
          if (a == b) v[8] += hv;
        }
      }
    }
  }
  big_cta_sum<11>(v, red);
  if (threadIdx.x == 0) {
    if (bad) *info = 1;
    const double yWy = -A[(size_t)n * ld + n], yWc = -A[(size_t)(n + 1) * ld + n];
    const double f0 = 0.5 * ((double)n * kLog2Pi + logdet + yWy) - yWc + 0.5 * v[0];
    const double g1 = -0.5 * (v[4] + v[9] - v[1]);
    const double g2 = -0.5 * (v[5] + v[10] - v[2]);
    const double g3 = -sg * (v[6] + v[8] - v[3]);
    if (f) f[slot] = bad ? NAN : f0;
    if (g && want_grad) { g[3 * (size_t)slot] = g1; g[3 * (size_t)slot + 1] = g2; g[3 * (size_t)slot + 2] = g3; }
    if (sum4) {
      atomicAdd(sum4, f0);
      if (want_grad) { atomicAdd(sum4 + 1, g1); atomicAdd(sum4 + 2, g2); atomicAdd(sum4 + 3, g3); }
    }
  }
}

cudaError_t launch_big_factor(cudaStream_t s, const IndivData& dd, const double* theta_i, const double* tau, int K,
                              double* acc, int ldN, long long strideA, double* wk, const IndivWork& w, int* info) {
  if (dd.nbig <= 0) return cudaSuccess;
  big_factor_kernel<<<dd.nbig, kTB, 0, s>>>(dd, theta_i, tau, K, acc, ldN, strideA, wk, w, info);
  return cudaGetLastError();
}

cudaError_t launch_big_tau(cudaStream_t s, const IndivData& dd, int K, const double* mean, const double* cov, int ldN,
                           long long strideC, const double* pi, const IndivWork& w, double* tau_new, double* logL,
                           int fixed_tau) {
  if (dd.nbig <= 0) return cudaSuccess;
  big_tau_kernel<<<dd.nbig, kTB, sizeof(double) * 3 * (size_t)K, s>>>(dd, K, mean, cov, ldN, strideC, pi, w, tau_new, logL,
                                                                      fixed_tau);
  return cudaGetLastError();
}

cudaError_t launch_big_eval(cudaStream_t s, const IndivData& dd, const IndivWork& w, const double* theta, int theta_stride,
                            int by_slot, const int* active, int want_grad, double* f, double* g, double* sum4, int* info) {
  if (dd.nbig <= 0) return cudaSuccess;
  big_eval_kernel<<<dd.nbig, kTB, 0, s>>>(dd, w, theta, theta_stride, by_slot, active, want_grad, f, g, sum4, info);
  return cudaGetLastError();
}

}  // namespace mcb
