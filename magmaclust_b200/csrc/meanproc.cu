// Small reductions around the dense N x N work of the mean processes: the scalar pieces of logL_GP_mod /
// gr_GP_mod (CF.R:194-216, 448-475) and their common-hp twins (CF.R:250-278, 510-537), pi_k (CF.R:682),
// pen_diag (CF.R:662). One warp per matrix row, warp-shuffle reductions, one atomicAdd per warp.
#include "common.cuh"

namespace mcb {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void group_reduce_kernel(int N, int ld, long long stride, const double* __restrict__ Kinv,
                                    const double* __restrict__ Csum, long long strideCs,
                                    const double* __restrict__ D1, const double* __restrict__ D2,
                                    double* __restrict__ redg) {
  const int g = blockIdx.z;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const double* kr = Kinv + g * stride + (size_t)row * ld;
  const double* cr = Csum ? Csum + g * strideCs + (size_t)row * ld : nullptr;
  const double* d1 = D1 ? D1 + g * stride + (size_t)row * ld : nullptr;
  const double* d2 = D2 ? D2 + g * stride + (size_t)row * ld : nullptr;
  double s = 0, t1 = 0, t2 = 0;
  for (int j = lane; j < N; j += 32) {
    const double k = kr[j];
    if (cr) s += k * cr[j];
    if (d1) { t1 += k * d1[j]; t2 += k * d2[j]; }
  }
  s = wsum(s); t1 = wsum(t1); t2 = wsum(t2);
  if (lane == 0) {
    atomicAdd(redg + 4 * g + 0, s);
    if (d1) { atomicAdd(redg + 4 * g + 1, t1); atomicAdd(redg + 4 * g + 2, t2); }
    atomicAdd(redg + 4 * g + 3, kr[row]);
  }
}

void group_reduce(cudaStream_t s, int N, int ld, long long stride, const double* Kinv, const double* Csum,
                  long long strideCs, const double* D1, const double* D2, double* redg, int G,
                  long long* launches) {
  group_reduce_kernel<<<dim3((N + 7) / 8, 1, G), 256, 0, s>>>(N, ld, stride, Kinv, Csum, strideCs, D1, D2, redg);
  if (launches) ++*launches;
}

__global__ void cluster_reduce_kernel(int N, int ld, long long stride, const double* __restrict__ r,
                                      const double* __restrict__ p, const double* __restrict__ D1,
                                      const double* __restrict__ D2, const int* __restrict__ gmap,
                                      double* __restrict__ redz) {
  const int z = blockIdx.z;
  const int g = gmap[z];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N || g < 0) return;   // g < 0: cluster not part of this evaluation
  const double* pz = p + (size_t)z * ld;
  double u1 = 0, u2 = 0;
  if (D1) {
    const double* d1 = D1 + g * stride + (size_t)row * ld;
    const double* d2 = D2 + g * stride + (size_t)row * ld;
    for (int j = lane; j < N; j += 32) {
      const double pj = pz[j];
      u1 += d1[j] * pj;
      u2 += d2[j] * pj;
    }
    u1 = wsum(u1); u2 = wsum(u2);
  }
  if (lane == 0) {
    const double pi = pz[row];
    atomicAdd(redz + 4 * z + 0, r[(size_t)z * ld + row] * pi);
    atomicAdd(redz + 4 * z + 1, pi * pi);
    if (D1) { atomicAdd(redz + 4 * z + 2, pi * u1); atomicAdd(redz + 4 * z + 3, pi * u2); }
  }
}

void cluster_reduce(cudaStream_t s, int N, int ld, long long stride, const double* r, const double* p,
                    const double* D1, const double* D2, const int* gmap, double* redz, int K,
                    long long* launches) {
  cluster_reduce_kernel<<<dim3((N + 7) / 8, 1, K), 256, 0, s>>>(N, ld, stride, r, p, D1, D2, gmap, redz);
  if (launches) ++*launches;
}

__global__ void symv_mapped_kernel(int N, int ld, long long stride, const double* __restrict__ Kinv,
                                   const int* __restrict__ gmap, const double* __restrict__ r,
                                   double* __restrict__ p) {
  const int z = blockIdx.z;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N || gmap[z] < 0) return;   // gmap < 0: cluster not part of this evaluation (its p is not used)
  const double* a = Kinv + gmap[z] * stride + (size_t)row * ld;
  const double* x = r + (size_t)z * ld;
  double acc = 0;
  for (int j = lane; j < N; j += 32) acc += a[j] * x[j];
  acc = wsum(acc);
  if (lane == 0) p[(size_t)z * ld + row] = acc;
}

void symv_mapped(cudaStream_t s, int N, int ld, long long stride, const double* Kinv, const int* gmap,
                 const double* r, double* p, int K, long long* launches) {
  symv_mapped_kernel<<<dim3((N + 7) / 8, 1, K), 256, 0, s>>>(N, ld, stride, Kinv, gmap, r, p);
  if (launches) ++*launches;
}

__global__ void sum_groups_kernel(long long nelem, long long stride, const double* __restrict__ src,
                                  const int* __restrict__ gmap, int K, double* __restrict__ dst) {
  const int g = blockIdx.y;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nelem;
       e += (long long)gridDim.x * blockDim.x) {
    double s = 0;
    for (int z = 0; z < K; ++z)
      if (gmap[z] == g) s += src[z * stride + e];
    dst[g * stride + e] = s;
  }
}

void sum_groups(cudaStream_t s, int N, int ld, long long stride, const double* src, const int* gmap, int K,
                double* dst, int G, long long* launches) {
  const long long nelem = (long long)N * ld;
  int blocks = (int)((nelem + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sum_groups_kernel<<<dim3(blocks, G), 256, 0, s>>>(nelem, stride, src, gmap, K, dst);
  if (launches) ++*launches;
}

__global__ void copy_mapped_kernel(long long nelem, long long stride, const double* __restrict__ src,
                                   const int* __restrict__ gmap, double* __restrict__ dst, double scale) {
  const int k = blockIdx.y;
  const double* sp = src + gmap[k] * stride;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nelem;
       e += (long long)gridDim.x * blockDim.x)
    dst[k * stride + e] = (scale == 0.0) ? 0.0 : scale * sp[e];
}

void copy_mapped(cudaStream_t s, int N, int ld, long long stride, const double* src, const int* gmap,
                 double* dst, int K, double scale, long long* launches) {
  const long long nelem = (long long)N * ld;
  int blocks = (int)((nelem + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  copy_mapped_kernel<<<dim3(blocks, K), 256, 0, s>>>(nelem, stride, src, gmap, dst, scale);
  if (launches) ++*launches;
}

__global__ void vec_sub_kernel(int n, const double* a, const double* b, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] - b[i];
}

void vec_sub(cudaStream_t s, int n, const double* a, const double* b, double* out, long long* launches) {
  vec_sub_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, a, b, out);
  if (launches) ++*launches;
}

// column sums of tau[M][K]: each CTA strides over individuals
__global__ void tau_colsum_kernel(int M, int K, const double* __restrict__ tau, double* __restrict__ out) {
  for (int k = 0; k < K; ++k) {
    double s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x)
      s += tau[(size_t)i * K + k];
    s = wsum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + k, s);
  }
}

void tau_colsum(cudaStream_t s, int M, int K, const double* tau, double* out, long long* launches) {
  int blocks = (M + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  cudaMemsetAsync(out, 0, sizeof(double) * K, s);
  tau_colsum_kernel<<<blocks, 256, 0, s>>>(M, K, tau, out);
  if (launches) ++*launches;
}

__global__ void strided_sum_kernel(int M, const double* __restrict__ v, int stride, int offset,
                                   double* __restrict__ out) {
  double s = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x)
    s += v[(size_t)i * stride + offset];
  s = wsum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

void theta3_sum(cudaStream_t s, int M, const double* theta, double* out, long long* launches) {
  int blocks = (M + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  cudaMemsetAsync(out, 0, sizeof(double), s);
  strided_sum_kernel<<<blocks, 256, 0, s>>>(M, theta, 3, 2, out);
  if (launches) ++*launches;
}

// out[0] += sum_i sum_k tau_ik (l_ik + log(pi_k / tau_ik)), the log term skipped when tau_ik == 0 or pi_k == 0
// (BIC's per-individual part, CF.R:384-385, with l_ik = mat_logL[i][k])
__global__ void bic_indiv_kernel(int M, int K, const double* __restrict__ tau, const double* __restrict__ logL,
                                 const double* __restrict__ pi, double* __restrict__ out) {
  double s = 0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M * K; e += gridDim.x * blockDim.x) {
    const int k = e % K;
    const double t = tau[e], p = pi[k];
    double v = t * logL[e];
    if (t != 0.0 && p != 0.0) v += t * log(p / t);
    s += v;
  }
  s = wsum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

void bic_indiv(cudaStream_t s, int M, int K, const double* tau, const double* logL, const double* pi, double* out,
               long long* launches) {
  int blocks = (M * K + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  bic_indiv_kernel<<<blocks, 256, 0, s>>>(M, K, tau, logL, pi, out);
  if (launches) ++*launches;
}

void vec_sum(cudaStream_t s, int M, const double* v, double* out, long long* launches) {
  int blocks = (M + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  strided_sum_kernel<<<blocks, 256, 0, s>>>(M, v, 1, 0, out);
  if (launches) ++*launches;
}

}  // namespace mcb
