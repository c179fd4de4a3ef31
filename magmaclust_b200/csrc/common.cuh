// Internal declarations shared by the translation units of libmagmaclust_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/mcb.h"

namespace mcb {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;

// ---- error plumbing: nothing throws across the C ABI -----------------------------------------------
struct Status {
  int code = MCB_OK;
  std::string msg;
};

#define MCB_CUDA(h, expr)                                                                  \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      char b_[512];                                                                        \
      snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      (h)->err = b_;                                                                       \
      return MCB_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)

#define MCB_TRY(expr)              \
  do {                             \
    int rc_ = (expr);              \
    if (rc_ != MCB_OK) return rc_; \
  } while (0)

// ---- growable device buffer ----------------------------------------------------------------------
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e != cudaSuccess) return e;
    cap = n;
    // zero fill: dense buffers rely on zero padding of every row up to ld (ld % 16 == 0). cudaMemset runs on the legacy
    // default stream, asynchronously to the host, and the library's streams are NON-BLOCKING: without the synchronisation
    // below the fill can land after the first kernel has already written the new buffer (seen at N = 4096: a 268 MB
    // fill overtaken by se_build). Allocation is rare (growth only), so the full synchronisation costs nothing.
    if ((e = cudaMemset(p, 0, n * sizeof(T))) != cudaSuccess) return e;
    return cudaDeviceSynchronize();
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// ---- dense N x N toolkit (dense.cu) ----------------------------------------------------------------
// All matrices row-major with leading dimension ld; batches are addressed as base + z * stride.
constexpr int kNB = 32;  // pivot block of the blocked symmetric sweep

struct DenseWs {
  DevBuf<double> Pbuf, Pbuf2, Cold;  // panels of the blocked sweep: [Z][rows][256] P (ping-pong pair) and V
  DevBuf<double> red;         // scratch for reductions
  DevBuf<double> slabV, slabP, slabL;  // panel exchange buffers of the persistent slab inverse (V, P triple-buffered)
  DevBuf<unsigned> bar;                // its per-matrix arrival counter and published-block flag
  int max_ctas = 0;                    // SMs the slab inverse may count on (0: all); see spd_inverse_slab
};

// K[z] = SE covariance on `grid` with theta[z] = (a, b, sigma): exp(a - exp(-b)/2 d^2) off-diagonal,
// exp(a) + sigma^2 on the diagonal (kern_to_cov, CF.R:61-86). Optionally also D2[z] = dK/db (deriv_hp2,
// CF.R:441-446; zero diagonal). theta is a DEVICE pointer, theta_stride doubles between problems.
// D1[z] = dK/da (deriv_hp1, CF.R:436-439: the kernel itself, diagonal exp(a) without the noise term).
// zj (optional): buffers that block 0 clears on the way (accumulators of the kernels that follow in the same graph)
struct ZeroJob {
  double* d; int n_d;
  double* d2; int n_d2;
  double* d3; int n_d3;
  unsigned* u; int n_u;
};
void se_build(cudaStream_t s, double* Kmat, double* D1, double* D2, int N, int ld, long long stride,
              const double* grid, const double* theta, int theta_stride, int Z, long long* launches,
              const ZeroJob* zj = nullptr);

// Side jobs of the fused objective / gradient of a theta_k shared by all K clusters (logL_GP_mod_common_hp_k,
// gr_GP_mod_common_hp_k, CF.R:250-278, 510-537), carried by the two GEMM stages so that one evaluation is four kernels.
// red[0..3] = sum Kinv o Csum, tr(Kinv dK/da), tr(Kinv dK/db), tr Kinv; red[4..7] = sum_z (r'p, p'p, p'dK/da p,
// p'dK/db p); red[8..10] = tr(G dK/da), tr(G dK/db), tr G with G = Kinv Csum Kinv. red must be zero on entry.
struct MeanFuse {
  int K;            // clusters
  int ldp;          // stride of r and p
  const double* r;  // [K][ldp] m_hat_k - m_k
  double* p;        // [K][ldp] Kinv r_k (accumulated by stage 1: zero on entry; read by stage 2); K <= 8
  double* red;      // [12]
};
void mean_stage1(cudaStream_t s, int N, const double* Kinv, const double* Csum, double* X, int ld, const double* D1,
                 const double* D2, const MeanFuse& mf, long long* launches);
void mean_stage2(cudaStream_t s, int N, const double* X, const double* Kinv, int ld, const double* D1, const double* D2,
                 const MeanFuse& mf, long long* launches);

// In-place inverse of Z symmetric positive definite matrices by a blocked symmetric sweep (block
// Gauss-Jordan). logdet[z] (device) receives log det of the ORIGINAL matrix; *info (device int) is set
// non-zero if a non-positive pivot is met. Both triangles are read and written.
cudaError_t spd_inverse_reserve(DenseWs& ws, int N, int Z);  // pre-size the workspace (no allocation later)
// dense_slab.cu: persistent shared-memory-resident variant for N <= 768 (chosen automatically by spd_inverse)
bool spd_inverse_slab_ok(int N, int ld);
cudaError_t spd_inverse_slab_reserve(DenseWs& ws, int N, int Z);
// pre_zeroed: logdet[0..Z) and the 2 Z sync words (slab_sync_words) were cleared by the caller (no memset nodes)
cudaError_t spd_inverse_slab(cudaStream_t s, double* A, int N, int ld, long long stride, int Z, double* logdet,
                             int* info, DenseWs& ws, long long* launches, bool pre_zeroed = false);
unsigned* slab_sync_words(DenseWs& ws);
cudaError_t spd_inverse(cudaStream_t s, double* A, int N, int ld, long long stride, int Z, double* logdet,
                        int* info, DenseWs& ws, long long* launches);

// C[z] = alpha * A[z] * B[z]^T + beta * C[z]  (M x N x K, row-major, DMMA m8n8k4)
void gemm_nt(cudaStream_t s, int M, int N, int K, const double* A, int lda, long long sA, const double* B,
             int ldb, long long sB, double* C, int ldc, long long sC, double alpha, double beta, int Z,
             long long* launches);

// out[z][0..2] += (sum_ij G_ij D1_ij, sum_ij G_ij D2_ij, trace G) with G = A[z] * B[z]^T never stored.
void gemm_nt_trace(cudaStream_t s, int N, const double* A, int lda, long long sA, const double* B, int ldb,
                   long long sB, const double* D1, const double* D2, int ldd, long long sD, double* out,
                   int out_stride, int Z, long long* launches);

// y[z] = A[z] x[z]; optional dot[z] += x[z] . y[z]
void symv(cudaStream_t s, int N, const double* A, int lda, long long sA, const double* x, long long sx,
          double* y, long long sy, int Z, long long* launches);

// ---- mean-process reductions (meanproc.cu) ---------------------------------------------------------
// redg[g][0..3] += (sum Kinv_g o Csum_g, sum Kinv_g o D1_g, sum Kinv_g o D2_g, trace Kinv_g); D1/D2 may be NULL
void group_reduce(cudaStream_t s, int N, int ld, long long stride, const double* Kinv, const double* Csum,
                  long long strideCs, const double* D1, const double* D2, double* redg, int G, long long* launches);
// redz[z][0..3] += (r_z.p_z, p_z.p_z, p_z'D1_g p_z, p_z'D2_g p_z) with g = gmap[z]; D1/D2 may be NULL
void cluster_reduce(cudaStream_t s, int N, int ld, long long stride, const double* r, const double* p,
                    const double* D1, const double* D2, const int* gmap, double* redz, int K, long long* launches);
// p[z] = Kinv[gmap[z]] r[z]
void symv_mapped(cudaStream_t s, int N, int ld, long long stride, const double* Kinv, const int* gmap,
                 const double* r, double* p, int K, long long* launches);
// dst[g] = sum over z with gmap[z] == g of src[z]   (N x ld matrices)
void sum_groups(cudaStream_t s, int N, int ld, long long stride, const double* src, const int* gmap, int K,
                double* dst, int G, long long* launches);
// dst[k] = src[gmap[k]] scaled by `scale` (0 -> zero fill): initialises A_k with K_k^-1
void copy_mapped(cudaStream_t s, int N, int ld, long long stride, const double* src, const int* gmap,
                 double* dst, int K, double scale, long long* launches);
// r[k][j] = mean[k][j] - mk[k][j]
void vec_sub(cudaStream_t s, int n, const double* a, const double* b, double* out, long long* launches);
// out[k] = (sum_i tau[i][k])   (caller divides by M_total after any all-reduce)
void tau_colsum(cudaStream_t s, int M, int K, const double* tau, double* out, long long* launches);
// out[0] = sum_i theta[i][2]
void theta3_sum(cudaStream_t s, int M, const double* theta, double* out, long long* launches);
// out[0] += sum_i sum_k tau_ik (logL_ik + log(pi_k / tau_ik))   (BIC, CF.R:384-385)
void bic_indiv(cudaStream_t s, int M, int K, const double* tau, const double* logL, const double* pi, double* out,
               long long* launches);
// out[0] += sum_i v[i]
void vec_sum(cudaStream_t s, int M, const double* v, double* out, long long* launches);
// symmetrise/mirror helpers are not needed: every producer writes both triangles.

}  // namespace mcb
