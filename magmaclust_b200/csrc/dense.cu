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
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "factor32.cuh"

namespace mcb {

// -------------------------------------------------------------------------------------------------
// SE covariance on the grid
// -------------------------------------------------------------------------------------------------
__global__ void se_build_kernel(double* __restrict__ Kmat, double* __restrict__ D1, double* __restrict__ D2, int N, int ld,
                                long long stride, const double* __restrict__ grid,
                                const double* __restrict__ theta, int theta_stride, ZeroJob zj) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    // side job of the fused theta_k evaluation: clear the accumulators that the following kernels add into
    const int t = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    for (int e = t; e < zj.n_d; e += nt) zj.d[e] = 0.0;
    for (int e = t; e < zj.n_d2; e += nt) zj.d2[e] = 0.0;
    for (int e = t; e < zj.n_d3; e += nt) zj.d3[e] = 0.0;
    for (int e = t; e < zj.n_u; e += nt) zj.u[e] = 0u;
  }
  const int z = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= N || j >= N) return;
  const double a = theta[(size_t)z * theta_stride + 0];
  const double b = theta[(size_t)z * theta_stride + 1];
  const double sg = theta[(size_t)z * theta_stride + 2];
  const double hl = exp(-b) / 2.0;
  const double d = grid[i] - grid[j];
  const double g = hl * (d * d);
  // one exp per element: dK/db = e^a g e^-g = g K off the diagonal (the per-individual kernels use the same form)
  double k, dk1, dk2;
  if (i == j) {
    dk1 = exp(a);
    k = dk1 + sg * sg;
    dk2 = 0.0;
  } else {
    k = exp(a - g);
    dk1 = k;
    dk2 = g * k;
  }
  Kmat[z * stride + (size_t)i * ld + j] = k;
  if (D1) D1[z * stride + (size_t)i * ld + j] = dk1;
  if (D2) D2[z * stride + (size_t)i * ld + j] = dk2;
}

void se_build(cudaStream_t s, double* Kmat, double* D1, double* D2, int N, int ld, long long stride,
              const double* grid, const double* theta, int theta_stride, int Z, long long* launches, const ZeroJob* zj) {
  dim3 blk(32, 8), grd((N + 31) / 32, (N + 7) / 8, Z);
  ZeroJob none{};
  se_build_kernel<<<grd, blk, 0, s>>>(Kmat, D1, D2, N, ld, stride, grid, theta, theta_stride, zj ? *zj : none);
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

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
// 16-byte copy that writes zeros instead when !valid (the source address is not dereferenced)
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// EPI_STORE_F / EPI_TRACE_F: the two stages of the fused shared-theta_k objective (MeanFuse in common.cuh)
enum { EPI_STORE = 0, EPI_SWEEP = 1, EPI_TRACE = 2, EPI_STORE_F = 3, EPI_TRACE_F = 4 };

struct GemmParams {
  int M, N, K;
  const double* A; int lda; long long sA;
  const double* B; int ldb; long long sB;
  double* C; int ldc; long long sC;
  double alpha, beta;
  // sweep epilogue: pivot rows [r0p, r0p + nb), pivot columns [c0, c0 + nb) of C; column j of C is row j + cro of the
  // full matrix (0 unless C is a column panel); Pbuf row stride ldp
  const double* Pbuf; long long sP; int c0, nb; double sign; int r0p, cro, ldp;
  int nofast;  // debug: disable the batched epilogue
  int sym;   // C symmetric and A == B: only the tiles on and below the diagonal are computed, results are mirrored
  // trace epilogue
  const double* D1; const double* D2; int ldd; long long sD; double* out; int out_stride;
  // fused mean-process stages
  MeanFuse mf;
};

constexpr int kBK = 32;
constexpr int kPad = 4;  // row stride kBK+kPad doubles: (4r + c) mod 16 distinct over a DMMA fragment

// C tile BM x BN per CTA, 4 warps as 2 x 2, each warp (BM/2) x (BN/2) in m8n8k4 tiles; the K loop streams
// 32-wide slabs of A and B rows through an ST-stage cp.async pipeline. The matrices of this path are small
// (N = 401 at the headline configuration), so a CTA's K loop is a chain of L2 round trips unless ST - 1 tiles
// (~100 k-columns) are in flight.
// Contract: rows of A and B are zero-padded up to a multiple of 16 columns (every dense buffer of this
// library has ld % 16 == 0 and zero pads), so the K loop needs no column guards below roundup16(K) (the last
// half tile is zero-filled); row indices past M / N are clamped (their results are discarded by the epilogue).
template <int BM, int BN, int ST>
constexpr size_t gemm_smem_bytes() { return sizeof(double) * ST * (BM + BN) * (kBK + kPad); }

template <int BM, int BN, int MODE, int ST>
__global__ void __launch_bounds__(128) gemm_nt_kernel(GemmParams p) {
  constexpr int FM = BM / 16, FN = BN / 16;  // fragments per warp per dimension
  constexpr int LDT = kBK + kPad;
  extern __shared__ __align__(16) double gsm[];
  double* As = gsm;                       // [ST][BM][LDT]
  double* Bs = gsm + ST * BM * LDT;       // [ST][BN][LDT]
  const int z = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * (BM / 2), wn = (warp & 1) * (BN / 2);
  const double* A = p.A + z * p.sA;
  const double* B = p.B + z * p.sB;
  const int K16 = (p.K + 15) & ~15;
  if (MODE == EPI_TRACE_F && n0 > m0 + BM - 1) return;   // G = X Kinv is symmetric: strictly upper tiles are not needed
  if (MODE == EPI_SWEEP && p.sym && n0 > m0 + BM - 1) return;

  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_tile = [&](int st, int k0) {
#pragma unroll
    for (int c = tid; c < BM * (kBK / 2); c += 128) {
      const int r = c / (kBK / 2), c2 = (c % (kBK / 2)) * 2;
      const int gr = min(m0 + r, p.M - 1);
      const bool ok = k0 + c2 < K16;
      cp_async16_zfill(As + (st * BM + r) * LDT + c2, A + (size_t)gr * p.lda + (ok ? k0 + c2 : 0), ok);
    }
#pragma unroll
    for (int c = tid; c < BN * (kBK / 2); c += 128) {
      const int r = c / (kBK / 2), c2 = (c % (kBK / 2)) * 2;
      const int gr = min(n0 + r, p.N - 1);
      const bool ok = k0 + c2 < K16;
      cp_async16_zfill(Bs + (st * BN + r) * LDT + c2, B + (size_t)gr * p.ldb + (ok ? k0 + c2 : 0), ok);
    }
  };

  const int nk = (p.K + kBK - 1) / kBK;
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) {
    if (s < nk) load_tile(s, s * kBK);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<ST - 2>();   // tile kt has landed
    __syncthreads();           // ... for everyone, and everyone is done with tile kt - 1 (its stage is refilled now)
    {
      const int nx = kt + ST - 1;
      if (nx < nk) load_tile(nx % ST, nx * kBK);
      cp_async_commit();
    }
    const double* at = As + (kt % ST) * BM * LDT;
    const double* bt = Bs + (kt % ST) * BN * LDT;
#pragma unroll
    for (int kk = 0; kk < kBK; kk += 4) {
      double af[FM], bf[FN];
#pragma unroll
      for (int i = 0; i < FM; ++i) af[i] = at[(wm + 8 * i + (lane >> 2)) * LDT + kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < FN; ++j) bf[j] = bt[(wn + 8 * j + (lane >> 2)) * LDT + kk + (lane & 3)];
#pragma unroll
      for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }

  if constexpr (MODE == EPI_SWEEP) {
    // Fast path for the tiles of the rank-k update that touch neither the pivot rows / columns nor the matrix edge (nor,
    // in the symmetric case, the diagonal): all C values of the thread are requested in one batch of independent
    // 16-byte loads (the generic loop below interleaves loads, tests and stores and was the longer half of the kernel)
    const bool clean = (m0 + BM <= p.r0p || m0 >= p.r0p + p.nb) && (n0 + BN <= p.c0 || n0 >= p.c0 + p.nb) &&
                       m0 + BM <= p.M && n0 + BN <= p.N && (!p.sym || n0 + BN - 1 < m0) && (p.ldc % 2 == 0) && !p.nofast;
    if (clean) {
      double* Cz = p.C + z * p.sC;
      double2 cv[FM][FN];
#pragma unroll
      for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j)
          cv[i][j] = *reinterpret_cast<const double2*>(Cz + (size_t)(m0 + wm + 8 * i + (lane >> 2)) * p.ldc + n0 + wn +
                                                       8 * j + (lane & 3) * 2);
#pragma unroll
      for (int i = 0; i < FM; ++i) {
        const int gi = m0 + wm + 8 * i + (lane >> 2);
#pragma unroll
        for (int j = 0; j < FN; ++j) {
          const int gj = n0 + wn + 8 * j + (lane & 3) * 2;
          const double r0 = p.sign * (cv[i][j].x - acc[i][j][0]);
          const double r1 = p.sign * (cv[i][j].y - acc[i][j][1]);
          *reinterpret_cast<double2*>(Cz + (size_t)gi * p.ldc + gj) = make_double2(r0, r1);
          if (p.sym) {
            Cz[(size_t)gj * p.ldc + gi] = r0;
            Cz[(size_t)(gj + 1) * p.ldc + gi] = r1;
          }
        }
      }
      return;
    }
  }

  double t1 = 0, t2 = 0, t0 = 0, u1 = 0, u2 = 0;
#pragma unroll
  for (int i = 0; i < FM; ++i) {
    const int gi = m0 + wm + 8 * i + (lane >> 2);
#pragma unroll
    for (int j = 0; j < FN; ++j) {
      const int gj = n0 + wn + 8 * j + (lane & 3) * 2;
      if (gi >= p.M || gj >= p.N) continue;
      const bool two = (gj + 1 < p.N);
      if (MODE == EPI_STORE || MODE == EPI_STORE_F) {
        double* c = p.C + z * p.sC + (size_t)gi * p.ldc + gj;
        if (p.beta == 0.0) {
          c[0] = p.alpha * acc[i][j][0];
          if (two) c[1] = p.alpha * acc[i][j][1];
        } else {
          c[0] = p.alpha * acc[i][j][0] + p.beta * c[0];
          if (two) c[1] = p.alpha * acc[i][j][1] + p.beta * c[1];
        }
      } else if (MODE == EPI_SWEEP) {
        double* c = p.C + z * p.sC + (size_t)gi * p.ldc + gj;
        const bool ip = gi >= p.r0p && gi < p.r0p + p.nb;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !two) break;
          const int gq = gj + q;
          if (p.sym && gq > gi) continue;   // owned by the tile that holds (gq, gi): it mirrors its result up here
          const bool jp = gq >= p.c0 && gq < p.c0 + p.nb;
          double r;
          if (!ip && !jp) r = c[q] - acc[i][j][q];
          else if (jp) r = p.Pbuf[z * p.sP + (size_t)gi * p.ldp + (gq - p.c0)];
          else r = p.Pbuf[z * p.sP + (size_t)(gq + p.cro) * p.ldp + (gi - p.r0p)];
          c[q] = p.sign * r;
          if (p.sym && gq < gi) p.C[z * p.sC + (size_t)gq * p.ldc + gi] = p.sign * r;
        }
      } else if (MODE == EPI_TRACE) {
        const size_t o = z * p.sD + (size_t)gi * p.ldd + gj;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !two) break;
          const double v = acc[i][j][q];
          if (gi == gj + q) t0 += v;
          t1 += v * p.D1[o + q];
          t2 += v * p.D2[o + q];
        }
      } else {
        // lower triangle with weight 2 (G and dK are symmetric); the quadratic forms sum_z p_z' dK p_z ride along
        const size_t o = (size_t)gi * p.ldd + gj;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !two) break;
          const int gq = gj + q;
          if (gq > gi) continue;
          const double wgt = (gq == gi) ? 1.0 : 2.0;
          const double v = acc[i][j][q];
          const double d1 = p.D1[o + q], d2 = p.D2[o + q];
          double pp = 0.0;
          for (int zz = 0; zz < p.mf.K; ++zz) pp += p.mf.p[(size_t)zz * p.mf.ldp + gi] * p.mf.p[(size_t)zz * p.mf.ldp + gq];
          if (gq == gi) t0 += v;
          t1 += wgt * v * d1;
          t2 += wgt * v * d2;
          u1 += wgt * pp * d1;
          u2 += wgt * pp * d2;
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
  if (MODE == EPI_TRACE_F) {
    // red[8..10] += tr(G dK/da), tr(G dK/db), tr G ; red[6..7] += sum_z p_z' dK/da p_z, p_z' dK/db p_z
    double v5[5] = {t1, t2, t0, u1, u2};
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v5[q] += __shfl_xor_sync(0xffffffffu, v5[q], o);
    if (lane == 0) {
      atomicAdd(p.mf.red + 8, v5[0]);
      atomicAdd(p.mf.red + 9, v5[1]);
      atomicAdd(p.mf.red + 10, v5[2]);
      atomicAdd(p.mf.red + 6, v5[3]);
      atomicAdd(p.mf.red + 7, v5[4]);
    }
    if (blockIdx.x == 0) {
      // red[4] += sum_z r_z' p_z, red[5] += sum_z p_z' p_z over this CTA's rows
      double q1 = 0, q2 = 0;
      for (int e = tid; e < BM * p.mf.K; e += 128) {
        const int zz = e / BM, gi = m0 + (e - zz * BM);
        if (gi < p.M) {
          const double pv = p.mf.p[(size_t)zz * p.mf.ldp + gi];
          q1 += p.mf.r[(size_t)zz * p.mf.ldp + gi] * pv;
          q2 += pv * pv;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        q2 += __shfl_xor_sync(0xffffffffu, q2, o);
      }
      if (lane == 0) { atomicAdd(p.mf.red + 4, q1); atomicAdd(p.mf.red + 5, q2); }
    }
  }
  if constexpr (MODE == EPI_STORE_F) {
    // side job, spread over all tiles: over THIS tile's BM x BN patch of Kinv (= A), the element-wise sums against
    // Csum (= B), dK/da, dK/db and the partial products p_z[row] += Kinv[row, patch] r_z[patch] (p is cleared by the
    // build kernel). All loads of a thread are independent: one L2 round trip.
    static_assert(BM * BN == 128 * 4, "one thread per 4 consecutive columns");
    const int rr = tid / (BN / 4), cc = (tid % (BN / 4)) * 4;
    const int gi = m0 + rr, gj = n0 + cc;
    double g4[4] = {0, 0, 0, 0};
    double pz[8];
#pragma unroll
    for (int zz = 0; zz < 8; ++zz) pz[zz] = 0.0;
    if (gi < p.M && gj < p.N) {   // pads beyond N inside the row are zero in every operand
      const double2 k01 = *reinterpret_cast<const double2*>(A + (size_t)gi * p.lda + gj);
      const double2 k23 = *reinterpret_cast<const double2*>(A + (size_t)gi * p.lda + gj + 2);
      const double2 c01 = *reinterpret_cast<const double2*>(B + (size_t)gi * p.ldb + gj);
      const double2 c23 = *reinterpret_cast<const double2*>(B + (size_t)gi * p.ldb + gj + 2);
      const double2 a01 = *reinterpret_cast<const double2*>(p.D1 + (size_t)gi * p.ldd + gj);
      const double2 a23 = *reinterpret_cast<const double2*>(p.D1 + (size_t)gi * p.ldd + gj + 2);
      const double2 b01 = *reinterpret_cast<const double2*>(p.D2 + (size_t)gi * p.ldd + gj);
      const double2 b23 = *reinterpret_cast<const double2*>(p.D2 + (size_t)gi * p.ldd + gj + 2);
      g4[0] = k01.x * c01.x + k01.y * c01.y + k23.x * c23.x + k23.y * c23.y;
      g4[1] = k01.x * a01.x + k01.y * a01.y + k23.x * a23.x + k23.y * a23.y;
      g4[2] = k01.x * b01.x + k01.y * b01.y + k23.x * b23.x + k23.y * b23.y;
      if (gi >= gj && gi < gj + 4) g4[3] = (gi == gj) ? k01.x : (gi == gj + 1) ? k01.y : (gi == gj + 2) ? k23.x : k23.y;
#pragma unroll
      for (int zz = 0; zz < 8; ++zz) {
        if (zz < p.mf.K) {
          const double2 r01 = *reinterpret_cast<const double2*>(p.mf.r + (size_t)zz * p.mf.ldp + gj);
          const double2 r23 = *reinterpret_cast<const double2*>(p.mf.r + (size_t)zz * p.mf.ldp + gj + 2);
          pz[zz] = k01.x * r01.x + k01.y * r01.y + k23.x * r23.x + k23.y * r23.y;
        }
      }
    }
    // the BN / 4 = 8 threads of a row are consecutive lanes
#pragma unroll
    for (int zz = 0; zz < 8; ++zz) {
      if (zz < p.mf.K) {
        double a = pz[zz];
#pragma unroll
        for (int o = BN / 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((tid % (BN / 4)) == 0 && gi < p.M) atomicAdd(p.mf.p + (size_t)zz * p.mf.ldp + gi, a);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) g4[q] += __shfl_xor_sync(0xffffffffu, g4[q], o);
      if (lane == 0 && g4[q] != 0.0) atomicAdd(p.mf.red + q, g4[q]);
    }
  }
}

template <int BM, int BN, int MODE, int ST>
static void launch_gemm_cfg(cudaStream_t s, const GemmParams& p, dim3 grd) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_nt_kernel<BM, BN, MODE, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)gemm_smem_bytes<BM, BN, ST>());
    attr_set = true;
  }
  gemm_nt_kernel<BM, BN, MODE, ST><<<grd, 128, gemm_smem_bytes<BM, BN, ST>(), s>>>(p);
}

template <int MODE>
static void launch_gemm(cudaStream_t s, const GemmParams& p, int Z, long long* launches) {
  // The DMMA pipe of one SM sustains ~250 GFLOP/s, so a CTA tile is a fixed amount of serial time (BM * BN * K * 2 flop
  // at that rate): small problems get small tiles so that the work spreads over all 148 SMs in near-equal shares.
  const long long ctas64 = (long long)((p.M + 63) / 64) * ((p.N + 63) / 64) * Z;
  const long long ctas32 = (long long)((p.M + 31) / 32) * ((p.N + 63) / 64) * Z;
  if (ctas64 >= 2 * 148) {
    launch_gemm_cfg<64, 64, MODE, 3>(s, p, dim3((p.N + 63) / 64, (p.M + 63) / 64, Z));   // 108 KB: 2 CTAs / SM
  } else if (ctas32 >= 2 * 148) {
    launch_gemm_cfg<32, 64, MODE, 4>(s, p, dim3((p.N + 63) / 64, (p.M + 31) / 32, Z));   // 108 KB
  } else {
    launch_gemm_cfg<16, 32, MODE, 4>(s, p, dim3((p.N + 31) / 32, (p.M + 15) / 16, Z));   // 54 KB: 4 CTAs / SM
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

// Fused stages of the shared-theta_k objective (see MeanFuse). Stage 1: X = Kinv Csum, plus the row-wise group sums
// and p_z = Kinv r_z. Stage 2: traces of G = X Kinv against dK/da, dK/db, I on the lower triangle, plus the quadratic
// forms of p_z. Both use the small-tile configuration: they run next to the persistent theta_i kernel on few SMs.
void mean_stage1(cudaStream_t s, int N, const double* Kinv, const double* Csum, double* X, int ld, const double* D1,
                 const double* D2, const MeanFuse& mf, long long* launches) {
  GemmParams p{};
  p.M = N; p.N = N; p.K = N;
  p.A = Kinv; p.lda = ld; p.B = Csum; p.ldb = ld; p.C = X; p.ldc = ld;
  p.alpha = 1.0; p.beta = 0.0;
  p.D1 = D1; p.D2 = D2; p.ldd = ld;
  p.mf = mf;
  launch_gemm_cfg<16, 32, EPI_STORE_F, 4>(s, p, dim3((N + 31) / 32, (N + 15) / 16, 1));
  if (launches) ++*launches;
}

void mean_stage2(cudaStream_t s, int N, const double* X, const double* Kinv, int ld, const double* D1, const double* D2,
                 const MeanFuse& mf, long long* launches) {
  GemmParams p{};
  p.M = N; p.N = N; p.K = N;
  p.A = X; p.lda = ld; p.B = Kinv; p.ldb = ld;
  p.D1 = D1; p.D2 = D2; p.ldd = ld;
  p.mf = mf;
  launch_gemm_cfg<16, 32, EPI_TRACE_F, 4>(s, p, dim3((N + 31) / 32, (N + 15) / 16, 1));
  if (launches) ++*launches;
}

// -------------------------------------------------------------------------------------------------
// blocked symmetric sweep
// -------------------------------------------------------------------------------------------------
// Panel step for pivot block b (nb <= 32 columns starting at c0); one CTA per 32 rows, 256 threads.
//   D = L L'            every CTA eliminates the 32 x 32 pivot block redundantly (32 steps, one barrier each):
//                       D = Lu Delta Lu' with unit-lower Lu, carrying W = Lu^-1 along, so L^-1 = Delta^-1/2 W
//   V[i] = A[i,b] L^-T  -> trailing update uses V V' (= A[:,b] D^-1 A[b,:])
//   P[i] = V[i] L^-1    -> new block column A[:,b] = A[:,b] D^-1
// Rows of the pivot block itself carry the identity instead of A[b,b], so their P is D^-1 (stored negated).
// Working with L^-1 (condition sqrt(cond D)) and the V V' form keeps the un-swept part evolving exactly like a
// right-looking blocked Cholesky; multiplying by an explicitly inverted D instead loses cond(D) digits, which is
// fatal for smooth SE kernels with a small jitter.
constexpr int kPanelRows = 32;

// The pivot block sits at rows [c0, c0 + nb) and COLUMNS [cc0, cc0 + nb) of the matrix it is given: the two differ when
// the matrix is a column panel of the full one (wide path below). V goes to Vbuf with row stride ldv.
__global__ void __launch_bounds__(256) sweep_panel_kernel(const double* __restrict__ Aall, int N, int ld,
                                                          long long stride, int c0, int cc0, int nb,
                                                          double* __restrict__ Vbuf, int ldv, long long sV,
                                                          double* __restrict__ Pbuf,
                                                          long long sP, double* __restrict__ logdet,
                                                          int* __restrict__ info) {
  __shared__ double D[kNB][kNB + 1];
  __shared__ double W[kNB][kNB + 1];
  __shared__ double R[kPanelRows][kNB + 1];
  __shared__ double V[kPanelRows][kNB + 1];
  __shared__ double piv[kNB];
  const int z = blockIdx.z;
  const int r0 = blockIdx.x * kPanelRows;
  const int tid = threadIdx.x;
  const int ti = tid >> 3, tc = (tid & 7) * 4;  // this thread's row and first of its 4 columns
  const double* A = Aall + z * stride;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = tc + q;
    // identity padding keeps a partial last block a no-op for the extra pivots
    D[ti][c] = (ti < nb && c < nb) ? A[(size_t)(c0 + ti) * ld + cc0 + c] : (ti == c ? 1.0 : 0.0);
    W[ti][c] = (ti == c) ? 1.0 : 0.0;
    const int gr = r0 + ti;
    double v = 0.0;
    if (gr < N && c < nb) {
      if (gr >= c0 && gr < c0 + nb) v = (gr - c0 == c) ? 1.0 : 0.0;
      else v = A[(size_t)gr * ld + cc0 + c];
    }
    R[ti][c] = v;
  }
  __syncthreads();
  for (int k = 0; k < kNB; ++k) {
    const double d = D[k][k];
    if (tid == 0) piv[k] = d;
    if (ti > k) {
      const double m = D[ti][k] / d;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = tc + q;
        if (c <= k) W[ti][c] -= m * W[k][c];
        else if (c <= ti) D[ti][c] -= m * D[c][k];
      }
    }
    __syncthreads();
  }
  if (tid < 32) {
    const double d = piv[tid];
    double lg = log(d);
    const bool bad = !(d > 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
    if (__any_sync(0xffffffffu, bad) && tid == 0) *info = 1;
    if (tid == 0 && r0 == c0) logdet[z] += lg;  // exactly one CTA owns the pivot rows
  }
  {
    // L^-1 = Delta^-1/2 W (lower triangular), kept in D's storage
    const double s = rsqrt(piv[ti]);
    const double sfix = s * (1.5 - 0.5 * piv[ti] * s * s);  // one Newton step: full double accuracy
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = tc + q;
      D[ti][c] = (c <= ti) ? W[ti][c] * sfix : 0.0;
    }
  }
  __syncthreads();
  const int gr = r0 + ti;
  // V = R L^-T : V[r][c] = sum_{b<=c} R[r][b] Linv[c][b]
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = tc + q;
    double acc = 0.0;
    for (int b = 0; b <= c; ++b) acc += R[ti][b] * D[c][b];
    V[ti][c] = acc;
    if (gr < N) Vbuf[z * sV + (size_t)gr * ldv + c] = acc;
  }
  __syncthreads();
  // P = V L^-1 : P[r][c] = sum_{b>=c} V[r][b] Linv[b][c]
  const double sgn = (gr >= c0 && gr < c0 + nb) ? -1.0 : 1.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = tc + q;
    double acc = 0.0;
    for (int b = c; b < kNB; ++b) acc += V[ti][b] * D[b][c];
    if (gr < N) Pbuf[z * sP + (size_t)gr * kNB + c] = sgn * acc;
  }
}

constexpr int kWide = 256;   // outer pivot block of the large-grid path (8 inner blocks of kNB)

// -------------------------------------------------------------------------------------------------
// Panel phase of the wide path, round 2: ONE kernel per inner pivot block (instead of sweep_panel_kernel + a K = 32 GEMM),
// no dependency between CTAs inside a launch. CTA = 32 rows of the column panel, 256 threads, for inner block j:
//   A  (j > 0) its 32 x 256 slice of the panel takes the rank-32 update of block j - 1 on DMMA (own V rows from shared
//      memory, the V rows of the 256 pivot-panel rows streamed from L2), pivot rows / columns of block j - 1 replaced from
//      P_{j-1}; the slice is read from the OLD panel buffer and written to the NEW one (ping-pong), so every CTA of the launch
//      sees the same old values.   (j == 0: the slice is gathered from the matrix itself.)
//   B  (j < nin) the 32 x 32 pivot block j is recomputed by EVERY CTA from old values (D - Vp Vp', one DMMA tile product),
//      factorised with factor32 (two-level, register-resident 8 x 8 tiles: 6.4 k cycles against ~15 k for the 32-barrier
//      elimination of sweep_panel_kernel), V_j = R L^-T and P_j = V_j L^-1 on DMMA for the CTA's own rows.
// Same arithmetic as the two-kernel version (sweep_panel_kernel + gemm_nt_kernel<EPI_SWEEP>); 9 launches per outer block
// instead of 17 and no redundant 32-step elimination. Measured (tools/dbg/dense_time, one matrix): N = 8192 33.6 -> 26.8 ms per
// inversion, N = 4096 7.89 -> 5.65 ms, N = 2000 2.45 -> 1.74 ms; C4 E-step 253 -> 216 ms. A launch takes 20 - 27 us
// (profiles/r02s2_panel_launches.txt): 256 CTAs are 1.7 per SM, each a chain of L2 round trips, so the panel phase is still
// ~7 of the 26.8 ms; MCB_OLD_PANEL=1 selects the two-kernel version.
constexpr int kPLd = kF32Ld;   // 36: row stride of the 32 x 32 shared-memory tiles

__device__ __forceinline__ void pdmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct PanelStep {
  int N, rows, ld;             // matrix size, padded rows of the panel buffers, leading dimension of the matrix
  long long strideA;           // between the matrices of the batch
  int c0, nbw, j, nin;         // outer block start / width, inner block index, number of inner blocks
  const double* A;             // the matrix (j == 0: source of the panel)
  const double* Pold;          // [Z][rows][kWide] panel before the update of block j - 1 (j > 0)
  double* Pnew;                // [Z][rows][kWide] panel after it
  double* Vc;                  // [Z][rows][kWide] V_0 .. V_7 side by side
  const double* Pin_prev;      // [Z][rows][kNB] P_{j-1}
  double* Pin;                 // [Z][rows][kNB] P_j
  double* logdet; int* info;
};

__global__ void __launch_bounds__(256) panel_step_kernel(PanelStep p) {
  extern __shared__ __align__(16) double psm[];
  double* Vown = psm;                      // V_{j-1} of the CTA's rows
  double* Vp = Vown + kNB * kPLd;          // V_{j-1} of the rows of pivot block j
  double* Dm = Vp + kNB * kPLd;            // pivot block j (destroyed by factor32)
  double* Li = Dm + kNB * kPLd;            // L^-1
  double* Rs = Li + kNB * kPLd;            // the CTA's rows of the pivot columns of block j
  double* Vs = Rs + kNB * kPLd;            // V_j of the CTA's rows
  double* scratch = Vs + kNB * kPLd;
  const int z = blockIdx.z, r0 = blockIdx.x * kNB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gr = lane >> 2, gc = lane & 3;
  const long long sW = (long long)p.rows * kWide, sI = (long long)p.rows * kNB;
  const double* A = p.A + z * p.strideA;
  const double* Pold = p.Pold + z * sW;
  double* Pnew = p.Pnew + z * sW;
  double* Vc = p.Vc + z * sW;
  const double* Pq = p.Pin_prev + z * sI;
  const int j = p.j;
  const bool have_prev = j > 0, do_sweep = j < p.nin;
  const int cp = (j - 1) * kNB, rp = p.c0 + cp;            // pivot columns (panel) / rows (matrix) of block j - 1
  const int cpj = j * kNB, rpj = p.c0 + cpj;               // ... of block j
  const int nbp = have_prev ? min(kNB, p.nbw - cp) : 0;
  const int nbj = do_sweep ? min(kNB, p.nbw - cpj) : 0;

  if (have_prev) {
    for (int e = tid; e < kNB * kNB; e += 256) {
      const int r = e >> 5, c = e & 31;
      Vown[r * kPLd + c] = (r0 + r < p.N) ? Vc[(size_t)(r0 + r) * kWide + cp + c] : 0.0;
      if (do_sweep) Vp[r * kPLd + c] = (r < nbj) ? Vc[(size_t)(rpj + r) * kWide + cp + c] : 0.0;
    }
  }
  if (do_sweep) {
    // old value of pivot block j (identity padded); the update of block j - 1 is subtracted below
    for (int e = tid; e < kNB * kNB; e += 256) {
      const int r = e >> 5, c = e & 31;
      double v = (r == c) ? 1.0 : 0.0;
      if (r < nbj && c < nbj)
        v = have_prev ? Pold[(size_t)(rpj + r) * kWide + cpj + c] : A[(size_t)(rpj + r) * p.ld + p.c0 + cpj + c];
      Dm[r * kPLd + c] = v;
    }
  }
  __syncthreads();

  // ---- A: the CTA's slice of the panel --------------------------------------------------------------------------------
  // warp w owns the column tiles ct = w, w + 8, w + 16, w + 24 (8 columns each) for all 4 row tiles. Two column tiles per
  // round: their V rows (L2) and the old values of the slice (L2) are requested together, ahead of the DMMAs
#pragma unroll 1
  for (int tt = 0; tt < 2; ++tt) {
    double bf[2][8];
    double2 oldv[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ct = warp + 8 * (2 * tt + u);
      if (have_prev) {
        const int vrow = p.c0 + 8 * ct + gr;               // panel column c is row c0 + c of the matrix
        const bool live = 8 * ct + gr < p.nbw && vrow < p.N;
        const double* vj = Vc + (size_t)(live ? vrow : 0) * kWide + cp + gc;
#pragma unroll
        for (int k = 0; k < 8; ++k) bf[u][k] = live ? __ldcg(vj + 4 * k) : 0.0;
      }
#pragma unroll
      for (int rt = 0; rt < 4; ++rt) {
        const int grow = r0 + 8 * rt + gr, c = 8 * ct + 2 * gc;
        double2 v = make_double2(0.0, 0.0);
        if (grow < p.N) {
          if (have_prev) {
            v = __ldcg(reinterpret_cast<const double2*>(Pold + (size_t)grow * kWide + c));
          } else {
            // gather from the matrix (ld and c0 + c are even only when ld is: two scalar loads)
            if (c < p.nbw) v.x = A[(size_t)grow * p.ld + p.c0 + c];
            if (c + 1 < p.nbw) v.y = A[(size_t)grow * p.ld + p.c0 + c + 1];
          }
        }
        oldv[u][rt] = v;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ct = warp + 8 * (2 * tt + u);
#pragma unroll
      for (int rt = 0; rt < 4; ++rt) {
        double a0 = 0.0, a1 = 0.0;
        if (have_prev) {
#pragma unroll
          for (int k = 0; k < 8; ++k) pdmma(a0, a1, Vown[(8 * rt + gr) * kPLd + 4 * k + gc], bf[u][k]);
        }
        const int r = 8 * rt + gr, grow = r0 + r;
        double out2[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = 8 * ct + 2 * gc + q;
          const double ov = q ? oldv[u][rt].y : oldv[u][rt].x;
          double v = 0.0;
          if (grow < p.N && c < p.nbw) {
            if (!have_prev) {
              v = ov;
            } else {
              const bool ip = grow >= rp && grow < rp + nbp, jp = c >= cp && c < cp + nbp;
              if (!ip && !jp) v = ov - (q ? a1 : a0);
              else if (jp) v = Pq[(size_t)grow * kNB + (c - cp)];
              else v = Pq[(size_t)(p.c0 + c) * kNB + (grow - rp)];
            }
          }
          out2[q] = v;
          if (do_sweep && c >= cpj && c < cpj + kNB) {
            // R: the CTA's rows of the pivot columns; identity rows for the pivot rows themselves (their P is D^-1)
            const int cc = c - cpj;
            double rv = (cc < nbj) ? v : 0.0;
            if (grow >= rpj && grow < rpj + nbj) rv = (grow - rpj == cc) ? 1.0 : 0.0;
            Rs[r * kPLd + cc] = rv;
          }
        }
        *reinterpret_cast<double2*>(Pnew + (size_t)grow * kWide + 8 * ct + 2 * gc) = make_double2(out2[0], out2[1]);
      }
    }
  }
  if (!do_sweep) return;

  // ---- B: pivot block j ------------------------------------------------------------------------------------------------
  if (have_prev) {
    // D -= Vp Vp' (the rank-32 update of block j - 1 restricted to the pivot block; every CTA recomputes it)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int tile = warp * 2 + t, ti = tile >> 2, tj = tile & 3;
      double n0 = 0.0, n1 = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) pdmma(n0, n1, Vp[(8 * ti + gr) * kPLd + 4 * k + gc], Vp[(8 * tj + gr) * kPLd + 4 * k + gc]);
      const int r = 8 * ti + gr, c = 8 * tj + 2 * gc;
      if (r < nbj && c < nbj) Dm[r * kPLd + c] -= n0;
      if (r < nbj && c + 1 < nbj) Dm[r * kPLd + c + 1] -= n1;
    }
  }
  __syncthreads();
  double logdet_part = 0.0;
  int bad = 0;
  factor32(Dm, Li, scratch, &logdet_part, &bad);   // ends with a barrier; Li = L^-1 (lower triangular)
  if (r0 == (rpj / kNB) * kNB) {                    // exactly one CTA owns the pivot rows: log det and the PD check
    if (warp == 7) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) logdet_part += __shfl_xor_sync(0xffffffffu, logdet_part, o);
      if (lane == 0) atomicAdd(p.logdet + z, logdet_part);
    }
    if (bad) *p.info = 1;
  }
  // V = R L^-T: V[r][c] = sum_b R[r][b] Linv[c][b]
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int tile = warp * 2 + t, ti = tile >> 2, tj = tile & 3;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) pdmma(c0, c1, Rs[(8 * ti + gr) * kPLd + 4 * k + gc], Li[(8 * tj + gr) * kPLd + 4 * k + gc]);
    Vs[(8 * ti + gr) * kPLd + 8 * tj + 2 * gc] = c0;
    Vs[(8 * ti + gr) * kPLd + 8 * tj + 2 * gc + 1] = c1;
    const int grow = r0 + 8 * ti + gr;
    if (grow < p.N) *reinterpret_cast<double2*>(Vc + (size_t)grow * kWide + cpj + 8 * tj + 2 * gc) = make_double2(c0, c1);
  }
  __syncthreads();
  // P = V L^-1: P[r][c] = sum_b V[r][b] Linv[b][c]; negated on the pivot rows (-D^-1)
  double* Pout = p.Pin + z * sI;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int tile = warp * 2 + t, ti = tile >> 2, tj = tile & 3;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) pdmma(c0, c1, Vs[(8 * ti + gr) * kPLd + 4 * k + gc], Li[(4 * k + gc) * kPLd + 8 * tj + gr]);
    const int grow = r0 + 8 * ti + gr;
    const double sgn = (grow >= rpj && grow < rpj + nbj) ? -1.0 : 1.0;
    if (grow < p.N) *reinterpret_cast<double2*>(Pout + (size_t)grow * kNB + 8 * tj + 2 * gc) = make_double2(sgn * c0, sgn * c1);
  }
}

cudaError_t spd_inverse_reserve(DenseWs& ws, int N, int Z) {
  const int ld16 = (N + 15) & ~15;
  if (spd_inverse_slab_ok(N, ld16)) return spd_inverse_slab_reserve(ws, N, Z);
  const int nblk = (N + kNB - 1) / kNB;
  const size_t rows = (size_t)nblk * kNB;
  cudaError_t e;
  if ((e = ws.Pbuf.ensure(rows * kWide * Z)) != cudaSuccess) return e;
  if ((e = ws.Pbuf2.ensure(rows * kWide * Z)) != cudaSuccess) return e;
  if ((e = ws.Cold.ensure(rows * kWide * Z)) != cudaSuccess) return e;
  return ws.red.ensure(2 * rows * kNB * Z);
}

// Pn[r][c] = A[r][c0 + c] for c < nbw (zero beyond, and for the pad rows): the column panel of an outer block
__global__ void panel_gather_kernel(const double* __restrict__ A, int N, int ld, long long stride, int c0, int nbw,
                                    double* __restrict__ Pn, long long sPn, int rows) {
  const int z = blockIdx.z;
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)rows * kWide) return;
  const int r = (int)(e / kWide), c = (int)(e - (long long)r * kWide);
  Pn[z * sPn + e] = (r < N && c < nbw) ? A[z * stride + (size_t)r * ld + c0 + c] : 0.0;
}

// dense_tma.cu: TMA-fed persistent version of the symmetric full-matrix update (returns false when it does not apply)
bool sweep_update_tma(cudaStream_t s, int N, int K, const double* V, int ldv, int rows, int Z, double* C, int ldc,
                      long long sC, const double* P, int ldp, long long sP, int c0, int nb, double sign, cudaError_t* err);

static void launch_sweep_update(cudaStream_t s, int M, int Ncols, int K, const double* Va, int lda, long long sA,
                                const double* Vb, int ldb, long long sB, double* C, int ldc, long long sC,
                                const double* P, int ldp, long long sP, int r0p, int c0, int nb, int cro, double sign,
                                int Z, int sym = 0) {
  static const bool verify = getenv("MCB_TMA_VERIFY") != nullptr;   // debug: run both kernels and compare
  double* Cold = nullptr;
  if (sym && Va == Vb && lda == ldb && sA == sB && M == Ncols && r0p == c0 && cro == 0 && sA % lda == 0) {
    cudaError_t e;
    if (verify) {
      cudaMalloc(&Cold, sizeof(double) * (size_t)sC * Z);
      cudaMemcpyAsync(Cold, C, sizeof(double) * (size_t)sC * Z, cudaMemcpyDeviceToDevice, s);
      if (atoi(getenv("MCB_TMA_VERIFY")) >= 2) {   // evict C from L2 again: the TMA kernel then sees what it sees in the real flow
        static void* scratch = nullptr;
        if (!scratch) cudaMalloc(&scratch, 512u << 20);
        cudaMemsetAsync(scratch, 1, 512u << 20, s);
      }
    }
    if (sweep_update_tma(s, M, K, Va, lda, (int)(sA / lda), Z, C, ldc, sC, P, ldp, sP, c0, nb, sign, &e)) {
      if (!verify) return;
      std::swap(C, Cold);   // the cp.async kernel below now works on the copy of the input; compared afterwards
    } else if (Cold) {
      cudaFree(Cold);
      Cold = nullptr;
    }
  }
  GemmParams p{};
  p.M = M; p.N = Ncols; p.K = K;
  p.A = Va; p.lda = lda; p.sA = sA;
  p.B = Vb; p.ldb = ldb; p.sB = sB;
  p.C = C; p.ldc = ldc; p.sC = sC;
  p.Pbuf = P; p.sP = sP; p.ldp = ldp; p.r0p = r0p; p.c0 = c0; p.nb = nb; p.cro = cro;
  p.sign = sign;
  static const bool nosym = getenv("MCB_SWEEP_NOSYM") != nullptr;
  static const bool nofast = getenv("MCB_SWEEP_NOFAST") != nullptr;
  p.sym = nosym ? 0 : sym;
  p.nofast = nofast ? 1 : 0;
  launch_gemm<EPI_SWEEP>(s, p, Z, nullptr);
  if (Cold) {
    // C (here: the copy) holds the cp.async result, Cold (the real matrix) the TMA result
    cudaStreamSynchronize(s);
    std::vector<double> a((size_t)sC * Z), b((size_t)sC * Z);
    cudaMemcpy(a.data(), C, sizeof(double) * a.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), Cold, sizeof(double) * b.size(), cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0; long long bad = 0; int wi = -1, wj = -1;
    for (int i = 0; i < M; ++i) for (int j = 0; j < Ncols; ++j) {
      const double x = a[(size_t)i * ldc + j], y = b[(size_t)i * ldc + j];
      scale = std::max(scale, std::fabs(x));
      const double d = std::fabs(x - y);
      if (!(d <= 1e-9 * (1.0 + std::fabs(x)))) { ++bad; if (!(d <= worst)) { worst = d; wi = i; wj = j; } }
    }
    printf("[tma verify] N %d K %d c0 %d nb %d sign %+.0f: mismatches %lld worst %.3e at (%d, %d), |C| max %.3e\n", M, K, c0, nb, sign,
           bad, worst, wi, wj, scale);
    cudaFree(C);   // the copy; the matrix keeps the TMA result
  }
}

cudaError_t spd_inverse(cudaStream_t s, double* A, int N, int ld, long long stride, int Z, double* logdet,
                        int* info, DenseWs& ws, long long* launches) {
  static const bool no_slab = getenv("MCB_NO_SLAB") != nullptr;
  static const bool narrow = getenv("MCB_NARROW_SWEEP") != nullptr;
  if (!no_slab && ld % 16 == 0 && spd_inverse_slab_ok(N, ld))
    return spd_inverse_slab(s, A, N, ld, stride, Z, logdet, info, ws, launches);
  const int nblk = (N + kNB - 1) / kNB;
  const int rows = nblk * kNB;
  cudaError_t e;
  if ((e = ws.Pbuf.ensure((size_t)rows * kWide * Z)) != cudaSuccess) return e;
  if ((e = ws.Pbuf2.ensure((size_t)rows * kWide * Z)) != cudaSuccess) return e;
  if ((e = ws.Cold.ensure((size_t)rows * kWide * Z)) != cudaSuccess) return e;
  if ((e = ws.red.ensure(2 * (size_t)rows * kNB * Z)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(logdet, 0, sizeof(double) * Z, s)) != cudaSuccess) return e;
  if (narrow) {
    // first version: one rank-32 update of the whole matrix per 32 pivots (each re-reads and re-writes the matrix)
    const long long sP = (long long)rows * kNB;
    for (int b = 0; b < nblk; ++b) {
      const int c0 = b * kNB;
      const int nb = (N - c0 < kNB) ? N - c0 : kNB;
      sweep_panel_kernel<<<dim3(nblk, 1, Z), 256, 0, s>>>(A, N, ld, stride, c0, c0, nb, ws.Cold.p, kNB, sP, ws.Pbuf.p,
                                                          sP, logdet, info);
      launch_sweep_update(s, N, N, kNB, ws.Cold.p, kNB, sP, ws.Cold.p, kNB, sP, A, ld, stride, ws.Pbuf.p, kNB, sP, c0,
                          c0, nb, 0, (b == nblk - 1) ? -1.0 : 1.0, Z);
      if (launches) *launches += 2;
    }
    return cudaGetLastError();
  }
  // Wide path: 128 pivots per pass over the matrix. The column panel A[:, B] of the outer block B is copied out and
  // swept on its own (4 inner blocks of 32 pivots: factorisation, V_j, P_j and a rank-32 update of the 128 panel
  // columns only, 8 MB at N = 8192); the V_j side by side are the operand of ONE rank-128 update of the whole matrix,
  // whose epilogue takes the pivot rows / columns from the finished panel. Same arithmetic as 4 narrow steps (the
  // updates of the columns outside B are additive), a quarter of the matrix traffic, and a GEMM with K = 128.
  double* Pn = ws.Pbuf.p;     // [Z][rows][kWide] the panel, becomes P (= A[:, B] D^-1, -D^-1 on the pivot rows)
  double* Vc = ws.Cold.p;     // [Z][rows][kWide] V_1 .. V_4
  double* Pin = ws.red.p;     // [Z][rows][kNB]   P_j of the current inner block
  const long long sW = (long long)rows * kWide, sI = (long long)rows * kNB;
  const int nout = (N + kWide - 1) / kWide;
  static const bool old_panel = getenv("MCB_OLD_PANEL") != nullptr;
  const size_t psmem = sizeof(double) * (6 * (size_t)kNB * kPLd + kF32Scratch + 8);
  if (!old_panel) {
    static bool attr_set = false;
    if (!attr_set) {
      if ((e = cudaFuncSetAttribute(panel_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem)) != cudaSuccess) return e;
      attr_set = true;
    }
  }
  for (int ob = 0; ob < nout; ++ob) {
    const int c0 = ob * kWide;
    const int nbw = (N - c0 < kWide) ? N - c0 : kWide;
    const int nin = (nbw + kNB - 1) / kNB;
    if (!old_panel) {
      // one fused launch per inner block (+ one that applies the last update): the panel ping-pongs between Pn and Pn2,
      // P_j between the two halves of Pin
      double* Pn2 = ws.Pbuf2.p;
      double* cur = Pn;          // written by the launch in flight
      double* prev = Pn2;
      for (int j = 0; j <= nin; ++j) {
        PanelStep ps{};
        ps.N = N; ps.rows = rows; ps.ld = ld; ps.strideA = stride;
        ps.c0 = c0; ps.nbw = nbw; ps.j = j; ps.nin = nin;
        ps.A = A; ps.Pold = prev; ps.Pnew = cur; ps.Vc = Vc;
        ps.Pin_prev = Pin + (size_t)((j + 1) & 1) * sI * Z;
        ps.Pin = Pin + (size_t)(j & 1) * sI * Z;
        ps.logdet = logdet; ps.info = info;
        panel_step_kernel<<<dim3(nblk, 1, Z), 256, psmem, s>>>(ps);
        if (launches) ++*launches;
        std::swap(cur, prev);
      }
      // `prev` now is the buffer the last launch wrote: the finished panel P
      launch_sweep_update(s, N, N, nin * kNB, Vc, kWide, sW, Vc, kWide, sW, A, ld, stride, prev, kWide, sW, c0, c0, nbw, 0,
                          (ob == nout - 1) ? -1.0 : 1.0, Z, 1);
      if (launches) ++*launches;
      continue;
    }
    const long long nel = (long long)rows * kWide;
    panel_gather_kernel<<<dim3((unsigned)((nel + 255) / 256), 1, Z), 256, 0, s>>>(A, N, ld, stride, c0, nbw, Pn, sW, rows);
    for (int j = 0; j < nin; ++j) {
      const int cj = j * kNB;
      const int nbj = (nbw - cj < kNB) ? nbw - cj : kNB;
      sweep_panel_kernel<<<dim3(nblk, 1, Z), 256, 0, s>>>(Pn, N, kWide, sW, c0 + cj, cj, nbj, Vc + cj, kWide, sW, Pin, sI,
                                                          logdet, info);
      // panel columns: Pn[i][c] -= V_j[i] . V_j[c0 + c] (column c of the panel is row c0 + c of the matrix)
      launch_sweep_update(s, N, nbw, kNB, Vc + cj, kWide, sW, Vc + (size_t)c0 * kWide + cj, kWide, sW, Pn, kWide, sW, Pin,
                          kNB, sI, c0 + cj, cj, nbj, c0, 1.0, Z);
      if (launches) *launches += 2;
    }
    launch_sweep_update(s, N, N, nin * kNB, Vc, kWide, sW, Vc, kWide, sW, A, ld, stride, Pn, kWide, sW, c0, c0, nbw, 0,
                        (ob == nout - 1) ? -1.0 : 1.0, Z, 1);
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
