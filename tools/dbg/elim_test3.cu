// Debug: unified-vector elimination of a 32 x 32 SPD block with TPR threads per row (256 / 128 / 64 / 32 threads):
// tests the "BAR drains every pending STS at k(nwarps) cycles each" model.
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kNB = 32;
__device__ __forceinline__ double fast_rcp(double d) {
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r); r = fma(fma(-d, r, 1.0), r, r); return r;
}
// TPR threads per row, CW = 32 / TPR columns per thread
template <int TPR>
__global__ void elim(const double* Din, double* Linv, long long* cyc) {
  constexpr int CW = kNB / TPR;
  constexpr int NT = kNB * TPR;
  __shared__ __align__(16) double U[2][kNB + 4];
  __shared__ double piv[kNB];
  const int tid = threadIdx.x, ti = tid / TPR, tc = (tid % TPR) * CW;
  double reg[CW];
#pragma unroll
  for (int q = 0; q < CW; ++q) reg[q] = Din[ti * kNB + tc + q];
  if (tc == 0) {
    if (ti > 0) U[0][ti] = reg[0];
    else { U[0][0] = 1.0; U[0][kNB] = fast_rcp(reg[0]); piv[0] = reg[0]; }
  }
  if (NT > 32) __syncthreads(); else __syncwarp();
  long long t0 = clock64();
  for (int k = 0; k < kNB; ++k) {
    const double* u = U[k & 1];
    if (ti > k) {
      const double m = u[ti] * u[kNB];
#pragma unroll
      for (int q = 0; q < CW; q += 2) {
        const double2 uu = *reinterpret_cast<const double2*>(u + tc + q);
        reg[q] = fma(-m, uu.x, (tc + q == k) ? 0.0 : reg[q]);
        reg[q + 1] = fma(-m, uu.y, (tc + q + 1 == k) ? 0.0 : reg[q + 1]);
      }
    }
    const int k1 = k + 1;
    if (k1 < kNB) {
      double* un = U[k1 & 1];
      const int q1 = k1 - tc;
      double v = 0.0;
#pragma unroll
      for (int q = 0; q < CW; ++q) if (q == q1) v = reg[q];
      if (ti == k1) {
        // row k1: Lu^-1[k1][c < k1]; the entries c >= k1 of this chunk are written by the column threads
#pragma unroll
        for (int q = 0; q < CW; q += 2) {
          const int c = tc + q;
          if (c + 1 < k1) *reinterpret_cast<double2*>(un + c) = make_double2(reg[q], reg[q + 1]);
          else if (c < k1) un[c] = reg[q];
        }
        if (q1 >= 0 && q1 < CW) { un[k1] = 1.0; un[kNB] = fast_rcp(v); piv[k1] = v; }
      } else if (ti > k1 && q1 >= 0 && q1 < CW) {
        un[ti] = v;
      }
    }
    if (NT > 32) __syncthreads(); else __syncwarp();
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  const double sc = 1.0 / sqrt(piv[ti]);
#pragma unroll
  for (int q = 0; q < CW; ++q) {
    const int c = tc + q;
    Linv[ti * kNB + c] = (c == ti) ? sc : (c < ti ? reg[q] * sc : 0.0);
  }
}
int main() {
  double h[kNB * kNB], L[kNB * kNB] = {0}, Li[kNB * kNB] = {0}, out[kNB * kNB];
  for (int i = 0; i < kNB; ++i) for (int j = 0; j < kNB; ++j) h[i * kNB + j] = (i == j) ? 3.0 : exp(-0.05 * (i - j) * (i - j));
  for (int j = 0; j < kNB; ++j) { double s = h[j * kNB + j]; for (int k = 0; k < j; ++k) s -= L[j * kNB + k] * L[j * kNB + k]; L[j * kNB + j] = sqrt(s);
    for (int i = j + 1; i < kNB; ++i) { double t = h[i * kNB + j]; for (int k = 0; k < j; ++k) t -= L[i * kNB + k] * L[j * kNB + k]; L[i * kNB + j] = t / L[j * kNB + j]; } }
  for (int c = 0; c < kNB; ++c) for (int i = c; i < kNB; ++i) { double s = (i == c) ? 1.0 : 0.0; for (int k = c; k < i; ++k) s -= L[i * kNB + k] * Li[k * kNB + c]; Li[i * kNB + c] = s / L[i * kNB + i]; }
  double *d, *o; long long* c; cudaMalloc(&d, sizeof h); cudaMalloc(&o, sizeof h); cudaMalloc(&c, 8);
  cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
  long long cy;
#define RUN(TPR) elim<TPR><<<1, 32 * TPR>>>(d, o, c); elim<TPR><<<1, 32 * TPR>>>(d, o, c); cudaDeviceSynchronize(); \
  cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(out, o, sizeof out, cudaMemcpyDeviceToHost); \
  { double err = 0; for (int i = 0; i < kNB * kNB; ++i) err = fmax(err, fabs(out[i] - Li[i])); \
  printf("threads %3d: %6lld cycles (%.0f per pivot), max |Linv - host| = %.3e\n", 32 * TPR, cy, cy / 32.0, err); }
  RUN(16) RUN(8) RUN(4)
  return 0;
}
