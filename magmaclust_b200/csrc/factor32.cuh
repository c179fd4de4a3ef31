// L^-1 of a 32 x 32 SPD pivot block (D = L L'), whole CTA of 256 threads, two-level blocked:
//   * the 4 diagonal 8 x 8 tiles are eliminated by ONE WARP each time (lane = row/column-pair, the pivot vector travels
//     through shared memory with __syncwarp only): 8 sequential pivots at ~150 cycles instead of ~470 with a CTA barrier
//   * everything between them is 8 x 8 x 8 DMMA tile products: panel L_ip = D_ip Linv_pp', trailing D_ij -= L_ip L_jp',
//     and the off-diagonal tiles of L^-1 by block forward substitution
// The N sequential pivots of the N x N factorisation are the critical path of every theta_k objective evaluation
// (dense_slab.cu); this routine is that path's inner loop.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

namespace mcb {

constexpr int kF32Ld = 36;          // row stride of the 32 x 32 shared-memory tiles (conflict-free DMMA fragments)
constexpr int kF32Scratch = 2 * 12 + 8;  // travelling pivot vectors (2 x (8 + 1), padded) + pivots

__device__ __forceinline__ double f32_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  return r;
}

// 1 / d: hardware seed (~2^-20) and one cubic (Halley) step, r (1 + e + e^2) with e = 1 - d r
__device__ __forceinline__ double f32_rcp3(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

// sqrt(r) for r = 1 / d > 0 (= 1 / sqrt(d))
__device__ __forceinline__ double f32_rsqrt_of_rcp(double r) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r));   // ~ 1 / sqrt(r)
  // Newton on y -> 1 / sqrt(r), then sqrt(r) = r y
  const double hr = 0.5 * r;
  y = fma(y, fma(-hr * y, y, 0.5), y);
  y = fma(y, fma(-hr * y, y, 0.5), y);
  const double g = r * y;
  return fma(fma(-g, g, r), 0.5 * y, g);
}

__device__ __forceinline__ double f32_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double hd = 0.5 * d;
  y = fma(y, fma(-hd * y, y, 0.5), y);
  y = fma(y, fma(-hd * y, y, 0.5), y);
  return y;
}

__device__ __forceinline__ void f32_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// C (8x8, accumulator layout in c0, c1) += sign * A (8x8) * op(B) (8x8), tiles in shared memory with stride kF32Ld.
// BT = true : C += A * B'   (B fragment element (k, n) = B[n][k])
// BT = false: C += A * B    (B fragment element (k, n) = B[k][n])
template <bool BT>
__device__ __forceinline__ void f32_mm(double& c0, double& c1, const double* A, const double* B, double sign, int lane) {
  const int gr = lane >> 2, gc = lane & 3;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double a = sign * A[gr * kF32Ld + 4 * k + gc];
    const double b = BT ? B[gr * kF32Ld + 4 * k + gc] : B[(4 * k + gc) * kF32Ld + gr];
    f32_dmma(c0, c1, a, b);
  }
}

__device__ __forceinline__ void f32_store(double* C, double c0, double c1, int lane) {
  const int gr = lane >> 2, gc = lane & 3;
  *reinterpret_cast<double2*>(C + gr * kF32Ld + 2 * gc) = make_double2(c0, c1);
}

// Dm: 32 x 32 SPD block (full symmetric, identity-padded), destroyed. Li: receives L^-1 (lower triangular, zeros above).
// logdet_acc (lanes of warp 0): += log pivots. bad: set if a pivot is not positive. All 256 threads call.
#ifdef F32_PROFILE
#define F32_T(i) if (threadIdx.x == 0) f32_prof[i] += clock64() - f32_t0, f32_t0 = clock64();
__device__ long long f32_prof[8];
#else
#define F32_T(i)
#endif
static __device__ void factor32(double* Dm, double* Li, double* scratch, double* logdet_acc, int* bad) {
#ifdef F32_PROFILE
  long long f32_t0 = clock64();
#endif
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  (void)scratch;
  for (int e = tid; e < 32 * kF32Ld; e += 256) Li[e] = 0.0;
  __syncthreads();
  // lower tiles (1,1) (2,1) (2,2) (3,1) (3,2) (3,3) <- warps 0..5
  const int ti = (warp >= 3) ? 3 : (warp >= 1 ? 2 : 1);
  const int tj = (warp == 0) ? 1 : (warp < 3 ? warp : warp - 2);
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    double* Dpp = Dm + (8 * p) * kF32Ld + 8 * p;
    double* Lpp = Li + (8 * p) * kF32Ld + 8 * p;
    if (warp == 0) {
      // Every lane factorises the whole 8 x 8 tile in registers (redundantly): no communication, no divergence, and
      // one copy of straight-line code (the p loop is kept rolled so that it stays in the instruction cache).
      double a[8][8], w[8][8], sk[8];
#ifdef F32_PROFILE
      long long q0 = clock64();
#endif
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) a[i][j] = Dpp[i * kF32Ld + j];
      bool ok = true;
#ifdef F32_PROFILE
      if (a[7][7] == 123.456) q0 = 0;
      long long q1 = clock64();
#endif
      // LDL' with unscaled columns. The next pivot column takes the short route a_ij -= (a_ik a_jk) / d_k (products
      // formed while 1 / d_k is in flight), the rest the cheap one m_i = a_ik / d_k, a_ij -= m_i a_jk. W = Lu^-1 rides
      // along (W_ic -= m_i W_kc, as if the identity were appended), so L^-1 = Delta^-1/2 W is complete with the last pivot.
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double d = a[k][k];
        ok = ok && (d > 0.0);
        const double r = f32_rcp3(d);
        sk[k] = f32_rsqrt_of_rcp(r);
        if (k + 1 < 8) {
#pragma unroll
          for (int i = k + 1; i < 8; ++i) a[i][k + 1] = fma(-(a[i][k] * a[k + 1][k]), r, a[i][k + 1]);
        }
#pragma unroll
        for (int i = k + 1; i < 8; ++i) {
          const double m = a[i][k] * r;
#pragma unroll
          for (int j = k + 2; j <= i; ++j) a[i][j] = fma(-m, a[j][k], a[i][j]);
          w[i][k] = -m;
#pragma unroll
          for (int c = 0; c < k; ++c) w[i][c] = fma(-m, w[k][c], w[i][c]);
        }
      }
#ifdef F32_PROFILE
      if (w[7][0] * sk[7] == 123.456) q0 = 0;
      long long q2 = clock64();
#endif
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int c = 0; c < i; ++c) Lpp[i * kF32Ld + c] = w[i][c] * sk[i];
        Lpp[i * kF32Ld + i] = sk[i];
      }
#ifdef F32_PROFILE
      if (tid == 0) { long long q3 = clock64(); f32_prof[5] += q1 - q0; f32_prof[6] += q2 - q1; f32_prof[7] += q3 - q2; }
#endif
      if (!ok && warp == 0) *bad = 1;
    }
    __syncthreads();
    F32_T(0)
    // panel: L_ip = D_ip Linv_pp'  (i > p), written over D_ip
    if (warp >= 1 && warp <= 3 - p) {
      const int i = p + warp;
      double c0 = 0.0, c1 = 0.0;
      double* Dip = Dm + (8 * i) * kF32Ld + 8 * p;
      f32_mm<true>(c0, c1, Dip, Lpp, 1.0, lane);
      __syncwarp();
      f32_store(Dip, c0, c1, lane);
    }
    __syncthreads();
    F32_T(1)
    // trailing update: D_ij -= L_ip L_jp'  for p < j <= i; warp w < 6 owns lower tile (ti, tj) for the whole call
    if (warp < 6 && tj > p) {
      double* Dij = Dm + (8 * ti) * kF32Ld + 8 * tj;
      const int gr = lane >> 2, gc = lane & 3;
      double2 c = *reinterpret_cast<double2*>(Dij + gr * kF32Ld + 2 * gc);
      f32_mm<true>(c.x, c.y, Dm + (8 * ti) * kF32Ld + 8 * p, Dm + (8 * tj) * kF32Ld + 8 * p, -1.0, lane);
      *reinterpret_cast<double2*>(Dij + gr * kF32Ld + 2 * gc) = c;
    }
    __syncthreads();
    F32_T(2)
  }
  // off-diagonal tiles of L^-1, two levels (the strictly upper tiles of Dm are dead and serve as staging):
  //   level 1: Linv_10 = -Linv_11 (L_10 Linv_00), Linv_32 = -Linv_33 (L_32 Linv_22)          (2 warps)
  //   level 2: the 16 x 16 block [Linv_20 Linv_21; Linv_30 Linv_31] = -Linv_22' (L_21' Linv_11') on 2 x 2 tiles (4 warps)
  // warp 7 meanwhile takes the log of the pivots (1 / diag(Linv)^2).
  if (warp == 7) {
    const double li = Li[lane * kF32Ld + lane];
    if (!(li > 0.0) || !(li < 1e300)) *bad = 1;
    *logdet_acc -= 2.0 * log(li);
  }
  if (warp < 2) {
    const int j = 2 * warp, i = j + 1;
    double s0 = 0.0, s1 = 0.0;
    f32_mm<false>(s0, s1, Dm + (8 * i) * kF32Ld + 8 * j, Li + (8 * j) * kF32Ld + 8 * j, 1.0, lane);
    double* T = Dm + (8 * j) * kF32Ld + 8 * i;
    f32_store(T, s0, s1, lane);
    __syncwarp();
    double c0 = 0.0, c1 = 0.0;
    f32_mm<false>(c0, c1, Li + (8 * i) * kF32Ld + 8 * i, T, -1.0, lane);
    f32_store(Li + (8 * i) * kF32Ld + 8 * j, c0, c1, lane);
  }
  __syncthreads();
  F32_T(3)
  if (warp < 4) {
    const int i = 2 + (warp >> 1), j = warp & 1;
    // T_ij = sum_{k=j..1} L_ik Linv_kj
    double s0 = 0.0, s1 = 0.0;
    for (int k = j; k < 2; ++k) f32_mm<false>(s0, s1, Dm + (8 * i) * kF32Ld + 8 * k, Li + (8 * k) * kF32Ld + 8 * j, 1.0, lane);
    f32_store(Dm + (8 * j) * kF32Ld + 8 * i, s0, s1, lane);
  }
  __syncthreads();
  if (warp < 4) {
    const int i = 2 + (warp >> 1), j = warp & 1;
    // Linv_ij = -sum_{k=2..i} Linv_ik T_kj
    double c0 = 0.0, c1 = 0.0;
    for (int k = 2; k <= i; ++k) f32_mm<false>(c0, c1, Li + (8 * i) * kF32Ld + 8 * k, Dm + (8 * j) * kF32Ld + 8 * k, -1.0, lane);
    f32_store(Li + (8 * i) * kF32Ld + 8 * j, c0, c1, lane);
  }
  __syncthreads();
  F32_T(4)
}

}  // namespace mcb
