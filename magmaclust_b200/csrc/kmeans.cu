// Seeded k-means for the initial responsibilities: the step BEFORE the hot path (ini_kmeans / ini_tau_i_k,
// MAGMAclust.R:562-607: stats::kmeans(centers = k, nstart = 50) on one feature row per individual). All `nstart` starts run
// side by side in the same launches (grid.y = start), Lloyd iterations to a fixed point, the start with the smallest total
// within-cluster sum of squares wins.
//
// stats::kmeans itself (Hartigan-Wong, R's RNG) is unpinned; what IS pinned is this file against
// oracle/kmeans_oracle.py, bit for bit, because every floating-point operation has a prescribed order:
//   distance   d2(i, k) = sum over d = 0 .. D-1, in that order, of fl(fl(x_id - c_kd)^2): separate multiply and add (no FMA)
//   label      the FIRST k with the smallest d2 (numpy argmin)
//   centre     per (k, d): points in chunks of kChunk consecutive rows, each chunk summed in row order, the chunk sums
//              added in chunk order, divided by the count; an empty cluster keeps its centre
//   cost       sum_i d2(i, label_i) with the same two-level order
// HBM-bound integer/FP64 streaming work: M x D doubles are read once per Lloyd pass and start (they stay in L2 across
// the starts: 2.4 MB at M = 1e5, D = 3).
#include <cstring>

#include "common.cuh"
#include "handle.h"

namespace mcb {

constexpr int kChunk = 1024;
constexpr int kTK = 256;

#define H_CHECK(cond, code, msg)  \
  do {                            \
    if (!(cond)) {                \
      h->err = (msg);             \
      return (code);              \
    }                             \
  } while (0)
#define H_CUDA(expr) MCB_CUDA(h, expr)

__device__ __forceinline__ double km_d2(const double* __restrict__ x, const double* c, int D) {
  double s = 0.0;
  for (int d = 0; d < D; ++d) {
    const double df = __dsub_rn(x[d], c[d]);
    s = __dadd_rn(s, __dmul_rn(df, df));
  }
  return s;
}

// labels[s][i] = argmin_k d2(i, k); changed[s] is raised when a label moved
__global__ void __launch_bounds__(kTK) km_assign_kernel(int M, int D, int K, const double* __restrict__ feat,
                                                        const double* __restrict__ cent, int* __restrict__ labels,
                                                        int* __restrict__ changed) {
  extern __shared__ double c[];   // [K][D] of this start
  const int s = blockIdx.y;
  for (int e = threadIdx.x; e < K * D; e += kTK) c[e] = cent[(size_t)s * K * D + e];
  __syncthreads();
  const int i = blockIdx.x * kTK + threadIdx.x;
  if (i >= M) return;
  const double* x = feat + (size_t)i * D;
  int best = 0;
  double bd = km_d2(x, c, D);
  for (int k = 1; k < K; ++k) {
    const double d = km_d2(x, c + (size_t)k * D, D);
    if (d < bd) { bd = d; best = k; }
  }
  int* l = labels + (size_t)s * M + i;
  if (*l != best) {
    *l = best;
    *changed = 1;   // any start: the host loop stops when no label of any start moved
  }
}

// partial[s][chunk][k][d] = row-ordered sum over the chunk's rows with label k; cnt[s][chunk][k]
__global__ void __launch_bounds__(128) km_partial_kernel(int M, int D, int K, const double* __restrict__ feat,
                                                         const int* __restrict__ labels, double* __restrict__ partial,
                                                         int* __restrict__ cnt) {
  const int ch = blockIdx.x, s = blockIdx.y, nch = gridDim.x;
  const int i0 = ch * kChunk, i1 = min(M, i0 + kChunk);
  const int* l = labels + (size_t)s * M;
  for (int e = threadIdx.x; e < K * D; e += blockDim.x) {
    const int k = e / D, d = e - k * D;
    double sum = 0.0;
    int n = 0;
    for (int i = i0; i < i1; ++i) {
      if (l[i] == k) {
        sum = __dadd_rn(sum, feat[(size_t)i * D + d]);
        ++n;
      }
    }
    partial[((size_t)s * nch + ch) * K * D + e] = sum;
    if (d == 0) cnt[((size_t)s * nch + ch) * K + k] = n;
  }
}

// centre[s][k][d] = (chunk sums in chunk order) / count
__global__ void __launch_bounds__(128) km_centre_kernel(int nch, int D, int K, const double* __restrict__ partial,
                                                        const int* __restrict__ cnt, double* __restrict__ cent) {
  const int s = blockIdx.x;
  for (int e = threadIdx.x; e < K * D; e += blockDim.x) {
    const int k = e / D;
    double sum = 0.0;
    long long n = 0;
    for (int ch = 0; ch < nch; ++ch) {
      sum = __dadd_rn(sum, partial[((size_t)s * nch + ch) * K * D + e]);
      n += cnt[((size_t)s * nch + ch) * K + k];
    }
    if (n > 0) cent[(size_t)s * K * D + e] = __ddiv_rn(sum, (double)n);
  }
}

// cost_part[s][chunk] = row-ordered sum of d2(i, label_i) over the chunk
__global__ void km_cost_kernel(int M, int D, int K, int nch, const double* __restrict__ feat, const double* __restrict__ cent,
                               const int* __restrict__ labels, double* __restrict__ cost_part) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (ch >= nch) return;
  const int i0 = ch * kChunk, i1 = min(M, i0 + kChunk);
  double sum = 0.0;
  for (int i = i0; i < i1; ++i)
    sum = __dadd_rn(sum, km_d2(feat + (size_t)i * D, cent + ((size_t)s * K + labels[(size_t)s * M + i]) * D, D));
  cost_part[(size_t)s * nch + ch] = sum;
}

__global__ void km_cost_sum_kernel(int nch, const double* __restrict__ cost_part, double* __restrict__ cost) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= gridDim.x * blockDim.x) return;
  double sum = 0.0;
  for (int ch = 0; ch < nch; ++ch) sum = __dadd_rn(sum, cost_part[(size_t)s * nch + ch]);
  cost[s] = sum;
}

}  // namespace mcb

using namespace mcb;

extern "C" int mcb_kmeans(mcb_handle* h, int M, int D, const double* feat, int K, int nstart, const int* start_rows,
                          int max_iter, int* out_labels, double* out_centers, double* out_cost, int* out_iters) {
  if (!h || !feat || !start_rows || !out_labels) return MCB_EINVAL;
  H_CHECK(M >= 1 && D >= 1 && K >= 1 && K <= M && nstart >= 1 && max_iter >= 1, MCB_EINVAL, "mcb_kmeans: bad shape");
  H_CHECK((size_t)K * D * sizeof(double) <= 200 * 1024, MCB_EINVAL, "mcb_kmeans: K * D centres exceed the shared-memory budget");
  for (int e = 0; e < nstart * K; ++e) H_CHECK(start_rows[e] >= 0 && start_rows[e] < M, MCB_EINVAL, "mcb_kmeans: start row out of range");
  H_CUDA(cudaSetDevice(h->device));
  const int nch = (M + kChunk - 1) / kChunk;
  DevBuf<double> d_feat, d_cent, d_part, d_costp, d_cost;
  DevBuf<int> d_lab, d_cnt, d_changed;
  auto cleanup = [&]() {
    d_feat.release(); d_cent.release(); d_part.release(); d_costp.release(); d_cost.release();
    d_lab.release(); d_cnt.release(); d_changed.release();
  };
  cudaError_t e = cudaSuccess;
  const size_t KD = (size_t)K * D;
  if ((e = d_feat.ensure((size_t)M * D)) || (e = d_cent.ensure(nstart * KD)) || (e = d_part.ensure((size_t)nstart * nch * KD)) ||
      (e = d_costp.ensure((size_t)nstart * nch)) || (e = d_cost.ensure(nstart)) || (e = d_lab.ensure((size_t)nstart * M)) ||
      (e = d_cnt.ensure((size_t)nstart * nch * K)) || (e = d_changed.ensure(1))) {
    cleanup();
    h->err = std::string("mcb_kmeans: ") + cudaGetErrorString(e);
    return MCB_ENOMEM;
  }
  // initial centres: the rows named by start_rows (kmeans(centers = k) draws k distinct rows, MC.R:585)
  std::vector<double> c0(nstart * KD);
  for (int s = 0; s < nstart; ++s)
    for (int k = 0; k < K; ++k)
      memcpy(&c0[((size_t)s * K + k) * D], feat + (size_t)start_rows[s * K + k] * D, sizeof(double) * D);
  cudaStream_t st = h->st;
  cudaMemcpyAsync(d_feat.p, feat, sizeof(double) * (size_t)M * D, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_cent.p, c0.data(), sizeof(double) * nstart * KD, cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(d_lab.p, 0xff, sizeof(int) * (size_t)nstart * M, st);   // -1: every label "moves" in the first pass
  const size_t smem = sizeof(double) * KD;
  if (smem > 48 * 1024) cudaFuncSetAttribute(km_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  h->launches = 0;
  int iters = 0;
  int* changed = reinterpret_cast<int*>(h->hbuf);
  for (int it = 0; it < max_iter; ++it) {
    cudaMemsetAsync(d_changed.p, 0, sizeof(int), st);
    km_assign_kernel<<<dim3((M + kTK - 1) / kTK, nstart), kTK, smem, st>>>(M, D, K, d_feat.p, d_cent.p, d_lab.p, d_changed.p);
    cudaMemcpyAsync(changed, d_changed.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    ++h->launches;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    ++iters;
    if (!*changed) break;   // fixed point of every start (a converged start repeats itself exactly in the meantime)
    km_partial_kernel<<<dim3(nch, nstart), 128, 0, st>>>(M, D, K, d_feat.p, d_lab.p, d_part.p, d_cnt.p);
    km_centre_kernel<<<nstart, 128, 0, st>>>(nch, D, K, d_part.p, d_cnt.p, d_cent.p);
    h->launches += 2;
  }
  if (e == cudaSuccess) {
    km_cost_kernel<<<dim3((nch + 63) / 64, nstart), 64, 0, st>>>(M, D, K, nch, d_feat.p, d_cent.p, d_lab.p, d_costp.p);
    km_cost_sum_kernel<<<nstart, 1, 0, st>>>(nch, d_costp.p, d_cost.p);
    h->launches += 2;
    std::vector<double> cost(nstart);
    cudaMemcpyAsync(cost.data(), d_cost.p, sizeof(double) * nstart, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) == cudaSuccess) {
      int best = 0;
      for (int s = 1; s < nstart; ++s)
        if (cost[s] < cost[best]) best = s;   // the first of equal costs
      cudaMemcpyAsync(out_labels, d_lab.p + (size_t)best * M, sizeof(int) * M, cudaMemcpyDeviceToHost, st);
      if (out_centers) cudaMemcpyAsync(out_centers, d_cent.p + best * KD, sizeof(double) * KD, cudaMemcpyDeviceToHost, st);
      if (out_cost) *out_cost = cost[best];
      if (out_iters) *out_iters = iters;
      e = cudaStreamSynchronize(st);
    }
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) {
    h->err = std::string("mcb_kmeans: ") + cudaGetErrorString(e);
    return MCB_ECUDA;
  }
  return MCB_OK;
}
