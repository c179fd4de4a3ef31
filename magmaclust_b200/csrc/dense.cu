// Dense N x N FP64 toolkit for the mean-process side of the VEM path (sm_100a).
//
//   se_build        kern_to_cov on the union grid                      (CF.R:61-86, 436-446)
//   spd_inverse     solve(K_k), solve(new_inv)                         (CF.R:142, 612; MC.R:218)
//   gemm_nt(_trace) inv %*% (new_cov %*% inv) and the trace-by-product (CF.R:456-466, 526-529)
//   symv            cov_k %*% weighted_mean, inv %*% (y - mean)        (CF.R:620, 455, 525)
//
// FP64 has no tcgen05 path on Blackwell: the tensor-core route for doubles is the legacy
// mma.sync.m8n8k4.f64 (SASS DMMA), measured at 37.1 TFLOP/s on this pool's B200 (same as cuBLAS DGEMM,
// profiles/r01_fp64_peaks.json), so every contraction below is tiled onto DMMA.
//
// The inverse is a blocked symmetric sweep (block Gauss-Jordan on an SPD matrix) in Cholesky form: for each
// 32-wide pivot block b:  D = A[b,b] = L L';  V = A[:,b] L^-T;  P = V L^-1;  A[i,j] -= V[i] V[j]' (i,j not in
// b);  A[:,b] = P;  A[b,b] = -D^-1.  After the last block the matrix holds -A^-1 (the sign is flipped in the
// last epilogue). Total N^3 FMAs in uniform rank-32 DMMA updates, 2 launches per block, log det = sum of the
// log pivots. The un-swept part evolves exactly as in a right-looking blocked Cholesky.
#include "common.cuh"

namespace mcb {

// -------------------------------------------------------------------------------------------------
// SE covariance on the grid
// -------------------------------------------------------------------------------------------------
__global__ void se_build_kernel(double* __restrict__ Kmat, double* __restrict__ D1, double* __restrict__ D2, int N, int ld,
                                long long stride, const double* __restrict__ grid,
                                const double* __restrict__ theta, int theta_stride) {
  const int z = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= N || j >= N) return;
  const double a = theta[(size_t)z * theta_stride + 0];
  const double b = theta[(size_t)z * theta_stride + 1];
  const double sg = theta[(size_t)z * theta_stride + 2];
  const double hl = exp(-b) / 2.0;
  const double d = grid[i] - grid[j];
  const double d2 = d * d;
  double k, dk1, dk2;
  if (i == j) {
    dk1 = exp(a);
    k = dk1 + sg * sg;
    dk2 = 0.0;
  } else {
    k = exp(a - hl * d2);
    dk1 = k;
    const double g = 0.5 * exp(-b) * d2;
    dk2 = exp(a) * g * exp(-g);
  }
  Kmat[z * stride + (size_t)i * ld + j] = k;
  if (D1) D1[z * stride + (size_t)i * ld + j] = dk1;
  if (D2) D2[z * stride + (size_t)i * ld + j] = dk2;
}

void se_build(cudaStream_t s, double* Kmat, double* D1, double* D2, int N, int ld, long long stride,
              const double* grid, const double* theta, int theta_stride, int Z, long long* launches) {
  dim3 blk(32, 8), grd((N + 31) / 32, (N + 7) / 8, Z);
  se_build_kernel<<<grd, blk, 0, s>>>(Kmat, D1, D2, N, ld, stride, grid, theta, theta_stride);
  if (launches) ++*launches;
}

// -------------------------------------------------------------------------------------------------
// DMMA NT GEMM with pluggable epilogue
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

enum { EPI_STORE = 0, EPI_SWEEP = 1, EPI_TRACE = 2 };

struct GemmParams {
  int M, N, K;
  const double* A; int lda; long long sA;
  const double* B; int ldb; long long sB;
  double* C; int ldc; long long sC;
  double alpha, beta;
  // sweep epilogue
  const double* Pbuf; long long sP; int c0, nb; double sign;
  // trace epilogue
  const double* D1; const double* D2; int ldd; long long sD; double* out; int out_stride;
};

constexpr int kBK = 16;
constexpr int kPad = 4;  // row stride kBK+kPad doubles: (4r + c) mod 16 distinct over a DMMA fragment

// CTA tile (2*WT) x (2*WT), 4 warps as 2 x 2, each warp WT x WT (WT = 32 or 16) in m8n8k4 tiles.
template <int WT, int MODE>
__global__ void __launch_bounds__(128) gemm_nt_kernel(GemmParams p) {
  constexpr int BT = 2 * WT;
  constexpr int F = WT / 8;  // fragments per warp per dimension
  __shared__ double As[BT][kBK + kPad];
  __shared__ double Bs[BT][kBK + kPad];
  const int z = blockIdx.z;
  const int m0 = blockIdx.y * BT, n0 = blockIdx.x * BT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * WT, wn = (warp & 1) * WT;
  const double* A = p.A + z * p.sA;
  const double* B = p.B + z * p.sB;

  double acc[F][F][2];
#pragma unroll
  for (int i = 0; i < F; ++i)
#pragma unroll
    for (int j = 0; j < F; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int k0 = 0; k0 < p.K; k0 += kBK) {
#pragma unroll
    for (int e = tid; e < BT * kBK; e += 128) {
      const int r = e / kBK, c = e % kBK;
      const int gk = k0 + c;
      const int ga = m0 + r, gb = n0 + r;
      As[r][c] = (ga < p.M && gk < p.K) ? A[(size_t)ga * p.lda + gk] : 0.0;
      Bs[r][c] = (gb < p.N && gk < p.K) ? B[(size_t)gb * p.ldb + gk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; kk += 4) {
      double af[F], bf[F];
#pragma unroll
      for (int i = 0; i < F; ++i) af[i] = As[wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < F; ++j) bf[j] = Bs[wn + 8 * j + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int i = 0; i < F; ++i)
#pragma unroll
        for (int j = 0; j < F; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }

  double t1 = 0, t2 = 0, t0 = 0;
#pragma unroll
  for (int i = 0; i < F; ++i) {
#pragma unroll
    for (int j = 0; j < F; ++j) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int gi = m0 + wm + 8 * i + (lane >> 2);
        const int gj = n0 + wn + 8 * j + (lane & 3) * 2 + q;
        if (gi >= p.M || gj >= p.N) continue;
        const double v = acc[i][j][q];
        if (MODE == EPI_STORE) {
          double* c = p.C + z * p.sC + (size_t)gi * p.ldc + gj;
          *c = (p.beta == 0.0) ? p.alpha * v : p.alpha * v + p.beta * *c;
        } else if (MODE == EPI_SWEEP) {
          double* c = p.C + z * p.sC + (size_t)gi * p.ldc + gj;
          const bool ip = gi >= p.c0 && gi < p.c0 + p.nb;
          const bool jp = gj >= p.c0 && gj < p.c0 + p.nb;
          double r;
          if (!ip && !jp) r = *c - v;
          else if (jp) r = p.Pbuf[z * p.sP + (size_t)gi * kNB + (gj - p.c0)];
          else r = p.Pbuf[z * p.sP + (size_t)gj * kNB + (gi - p.c0)];
          *c = p.sign * r;
        } else {
          const size_t o = z * p.sD + (size_t)gi * p.ldd + gj;
          double d1 = p.D1[o];
          if (gi == gj) { t0 += v; }
          t1 += v * d1;
          t2 += v * p.D2[o];
        }
      }
    }
  }
  if (MODE == EPI_TRACE) {
    __shared__ double red[3][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, o);
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (lane == 0) { red[0][warp] = t1; red[1][warp] = t2; red[2][warp] = t0; }
    __syncthreads();
    if (tid < 3) {
      double s = red[tid][0] + red[tid][1] + red[tid][2] + red[tid][3];
      atomicAdd(p.out + (size_t)z * p.out_stride + tid, s);
    }
  }
}

template <int MODE>
static void launch_gemm(cudaStream_t s, const GemmParams& p, int Z, long long* launches) {
  // small problems: 32x32 CTA tiles so that more SMs take part
  long long ctas64 = (long long)((p.M + 63) / 64) * ((p.N + 63) / 64) * Z;
  if (ctas64 >= 2 * 148) {
    dim3 grd((p.N + 63) / 64, (p.M + 63) / 64, Z);
    gemm_nt_kernel<32, MODE><<<grd, 128, 0, s>>>(p);
  } else {
    dim3 grd((p.N + 31) / 32, (p.M + 31) / 32, Z);
    gemm_nt_kernel<16, MODE><<<grd, 128, 0, s>>>(p);
  }
  if (launches) ++*launches;
}

void gemm_nt(cudaStream_t s, int M, int N, int K, const double* A, int lda, long long sA, const double* B,
             int ldb, long long sB, double* C, int ldc, long long sC, double alpha, double beta, int Z,
             long long* launches) {
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.sA = sA;
  p.B = B; p.ldb = ldb; p.sB = sB;
  p.C = C; p.ldc = ldc; p.sC = sC;
  p.alpha = alpha; p.beta = beta;
  launch_gemm<EPI_STORE>(s, p, Z, launches);
}

void gemm_nt_trace(cudaStream_t s, int N, const double* A, int lda, long long sA, const double* B, int ldb,
                   long long sB, const double* D1, const double* D2, int ldd, long long sD, double* out,
                   int out_stride, int Z, long long* launches) {
  GemmParams p{};
  p.M = N; p.N = N; p.K = N;
  p.A = A; p.lda = lda; p.sA = sA;
  p.B = B; p.ldb = ldb; p.sB = sB;
  p.D1 = D1; p.D2 = D2; p.ldd = ldd; p.sD = sD; p.out = out; p.out_stride = out_stride;
  launch_gemm<EPI_TRACE>(s, p, Z, launches);
}

// -------------------------------------------------------------------------------------------------
// blocked symmetric sweep
// -------------------------------------------------------------------------------------------------
// Panel step for pivot block b (nb <= 32 columns starting at c0). One CTA per 128 rows, thread = row.
//   D = L L'            Cholesky of the pivot block (warp 0; pivots feed log det and the PD check)
//   V[i] = A[i,b] L^-T  forward substitution per row   -> trailing update uses V V' (= A[:,b] D^-1 A[b,:])
//   P[i] = V[i] L^-1    back substitution per row      -> new block column A[:,b] = A[:,b] D^-1
// Rows of the pivot block itself carry the identity instead of A[b,b], so their P is D^-1 (stored negated).
// The triangular solves keep the update as stable as a blocked Cholesky: multiplying by an explicitly
// inverted D loses cond(D) digits, which is fatal for smooth SE kernels with small jitter.
constexpr int kPanelRows = 128;

__global__ void __launch_bounds__(kPanelRows) sweep_panel_kernel(const double* __restrict__ Aall, int N, int ld,
                                                                 long long stride, int c0, int nb,
                                                                 double* __restrict__ Vbuf,
                                                                 double* __restrict__ Pbuf, long long sP,
                                                                 double* __restrict__ logdet,
                                                                 int* __restrict__ info) {
  __shared__ double L[kNB][kNB + 1];
  __shared__ double Rw[kPanelRows][kNB + 1];
  const int z = blockIdx.z;
  const int r0 = blockIdx.x * kPanelRows;
  const int tid = threadIdx.x, lane = tid & 31;
  const double* A = Aall + z * stride;
  for (int e = tid; e < kNB * kNB; e += kPanelRows) {
    const int a = e >> 5, b = e & 31;
    L[a][b] = (a < nb && b < nb) ? A[(size_t)(c0 + a) * ld + c0 + b] : 0.0;
  }
  // own rows of the block column, coalesced over the 32 columns
  for (int e = tid; e < kPanelRows * kNB; e += kPanelRows) {
    const int a = e >> 5, b = e & 31;
    const int gr = r0 + a;
    double v = 0.0;
    if (gr < N && b < nb) {
      if (gr >= c0 && gr < c0 + nb) v = (gr - c0 == b) ? 1.0 : 0.0;
      else v = A[(size_t)gr * ld + c0 + b];
    }
    Rw[a][b] = v;
  }
  __syncthreads();
  if (tid < 32) {
    // right-looking Cholesky, lane = row; lower triangle of L
    double ldsum = 0.0;
    bool bad = false;
    for (int k = 0; k < nb; ++k) {
      const double dkk = L[k][k];
      if (!(dkk > 0.0)) bad = true;
      const double sk = sqrt(dkk), inv = 1.0 / sk;
      ldsum += log(dkk);
      __syncwarp();
      double l = 0.0;
      if (lane > k && lane < nb) { l = L[lane][k] * inv; L[lane][k] = l; }
      if (lane == k) L[k][k] = sk;
      __syncwarp();
      if (lane > k && lane < nb)
        for (int j = k + 1; j <= lane; ++j) L[lane][j] -= l * L[j][k];
      __syncwarp();
    }
    if (lane == 0) {
      if (bad) *info = 1;
      if (r0 <= c0 && c0 < r0 + kPanelRows) logdet[z] += ldsum;  // exactly one CTA owns the pivot rows
    }
  }
  __syncthreads();
  const int gr = r0 + tid;
  double* row = Rw[tid];
  // forward: v L' = r  ->  v_k = (r_k - sum_{j<k} v_j L[k][j]) / L[k][k]
  for (int k = 0; k < nb; ++k) {
    double acc = row[k];
    for (int j = 0; j < k; ++j) acc -= row[j] * L[k][j];
    row[k] = acc / L[k][k];
  }
  if (gr < N) {
    double* vo = Vbuf + z * sP + (size_t)gr * kNB;
    for (int k = 0; k < kNB; ++k) vo[k] = (k < nb) ? row[k] : 0.0;
  }
  // backward: p L = v  ->  p_k = (v_k - sum_{j>k} p_j L[j][k]) / L[k][k]
  for (int k = nb - 1; k >= 0; --k) {
    double acc = row[k];
    for (int j = k + 1; j < nb; ++j) acc -= row[j] * L[j][k];
    row[k] = acc / L[k][k];
  }
  if (gr < N) {
    const double sgn = (gr >= c0 && gr < c0 + nb) ? -1.0 : 1.0;
    double* po = Pbuf + z * sP + (size_t)gr * kNB;
    for (int k = 0; k < kNB; ++k) po[k] = (k < nb) ? sgn * row[k] : 0.0;
  }
}

cudaError_t spd_inverse(cudaStream_t s, double* A, int N, int ld, long long stride, int Z, double* logdet,
                        int* info, DenseWs& ws, long long* launches) {
  const int nblk = (N + kNB - 1) / kNB;
  const long long sP = (long long)nblk * kNB * kNB;
  cudaError_t e;
  if ((e = ws.Pbuf.ensure((size_t)sP * Z)) != cudaSuccess) return e;
  if ((e = ws.Cold.ensure((size_t)sP * Z)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(logdet, 0, sizeof(double) * Z, s)) != cudaSuccess) return e;
  for (int b = 0; b < nblk; ++b) {
    const int c0 = b * kNB;
    const int nb = (N - c0 < kNB) ? N - c0 : kNB;
    sweep_panel_kernel<<<dim3((N + kPanelRows - 1) / kPanelRows, 1, Z), kPanelRows, 0, s>>>(
        A, N, ld, stride, c0, nb, ws.Cold.p, ws.Pbuf.p, sP, logdet, info);
    GemmParams p{};
    p.M = N; p.N = N; p.K = nb;
    p.A = ws.Cold.p; p.lda = kNB; p.sA = sP;   // V
    p.B = ws.Cold.p; p.ldb = kNB; p.sB = sP;   // V
    p.C = A; p.ldc = ld; p.sC = stride;
    p.Pbuf = ws.Pbuf.p; p.sP = sP; p.c0 = c0; p.nb = nb;
    p.sign = (b == nblk - 1) ? -1.0 : 1.0;
    launch_gemm<EPI_SWEEP>(s, p, Z, nullptr);
    if (launches) *launches += 2;
  }
  return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// y = A x, one warp per row
// -------------------------------------------------------------------------------------------------
__global__ void symv_kernel(int N, const double* __restrict__ A, int lda, long long sA,
                            const double* __restrict__ x, long long sx, double* __restrict__ y, long long sy) {
  const int z = blockIdx.z;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const double* a = A + z * sA + (size_t)row * lda;
  const double* xv = x + z * sx;
  double acc = 0.0;
  for (int j = lane; j < N; j += 32) acc += a[j] * xv[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[z * sy + row] = acc;
}

void symv(cudaStream_t s, int N, const double* A, int lda, long long sA, const double* x, long long sx,
          double* y, long long sy, int Z, long long* launches) {
  symv_kernel<<<dim3((N + 7) / 8, 1, Z), 256, 0, s>>>(N, A, lda, sA, x, sx, y, sy);
  if (launches) ++*launches;
}

}  // namespace mcb
