// Debug: branch-free unified-vector elimination of a 32 x 32 SPD block -> L^-1; timing + check against host.
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kNB = 32;
__device__ __forceinline__ double fast_rcp(double d) {
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r); r = fma(fma(-d, r, 1.0), r, r); return r;
}
template <int V>
__global__ void __launch_bounds__(256, 1) elim(const double* Din, double* Linv, long long* cyc) {
  __shared__ __align__(16) double Uf[2 * (kNB + 4)];
  __shared__ double piv[kNB];
  const int tid = threadIdx.x, ti = tid >> 3, tc = (tid & 7) * 4;
  double reg[4];
  int zero = 0;
  asm volatile("" : "+r"(zero));  // opaque 0: keeps the broadcast loads on per-thread addresses (no S2UR)
  for (int q = 0; q < 4; ++q) reg[q] = Din[ti * kNB + tc + q];
  if (tc == 0) {
    if (ti > 0) Uf[ti] = reg[0];
    else { Uf[0] = 1.0; Uf[kNB] = fast_rcp(reg[0]); piv[0] = reg[0]; }
  }
  __syncthreads();
  long long t0 = clock64();
  for (int k = 0; k < kNB; ++k) {
    const int ub = (k & 1) * (kNB + 4);
    if (ti > k) {
      const double m = Uf[ub + ti] * Uf[ub + kNB + zero];
      const double2 ua = *reinterpret_cast<const double2*>(&Uf[ub + tc]);
      const double2 ubv = *reinterpret_cast<const double2*>(&Uf[ub + tc + 2]);
      const int kq = k - tc;
      if (V == 2) { reg[0] += m + ua.x + ubv.y; } else {
      reg[0] = fma(-m, ua.x, kq == 0 ? 0.0 : reg[0]);
      reg[1] = fma(-m, ua.y, kq == 1 ? 0.0 : reg[1]);
      reg[2] = fma(-m, ubv.x, kq == 2 ? 0.0 : reg[2]);
      reg[3] = fma(-m, ubv.y, kq == 3 ? 0.0 : reg[3]); }
    }
    const int k1 = k + 1;
    if (k1 < kNB) {
      const int un = (k1 & 1) * (kNB + 4);
      const int q1 = k1 - tc;
      if (V != 3 && ti == k1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (tc + q < k1) Uf[un + tc + q] = reg[q];
      }
      if (q1 >= 0 && q1 < 4) {
        const double v = q1 == 0 ? reg[0] : q1 == 1 ? reg[1] : q1 == 2 ? reg[2] : reg[3];
        if (V != 3 && ti > k1) Uf[un + ti] = v;
        else if (ti == k1) { Uf[un + k1] = 1.0; Uf[un + kNB] = (V == 1) ? 0.25 : fast_rcp(v); piv[k1] = v; }
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  const double sc = 1.0 / sqrt(piv[ti]);
  for (int q = 0; q < 4; ++q) {
    const int c = tc + q;
    Linv[ti * kNB + c] = (c == ti) ? sc : (c < ti ? reg[q] * sc : 0.0);
  }
}
int main() {
  double h[kNB * kNB], L[kNB * kNB] = {0}, Li[kNB * kNB] = {0}, out[kNB * kNB];
  for (int i = 0; i < kNB; ++i) for (int j = 0; j < kNB; ++j) h[i * kNB + j] = (i == j) ? 3.0 : exp(-0.05 * (i - j) * (i - j));
  // host Cholesky + inverse
  for (int j = 0; j < kNB; ++j) { double s = h[j * kNB + j]; for (int k = 0; k < j; ++k) s -= L[j * kNB + k] * L[j * kNB + k]; L[j * kNB + j] = sqrt(s);
    for (int i = j + 1; i < kNB; ++i) { double t = h[i * kNB + j]; for (int k = 0; k < j; ++k) t -= L[i * kNB + k] * L[j * kNB + k]; L[i * kNB + j] = t / L[j * kNB + j]; } }
  for (int c = 0; c < kNB; ++c) for (int i = c; i < kNB; ++i) { double s = (i == c) ? 1.0 : 0.0; for (int k = c; k < i; ++k) s -= L[i * kNB + k] * Li[k * kNB + c]; Li[i * kNB + c] = s / L[i * kNB + i]; }
  double *d, *o; long long* c; cudaMalloc(&d, sizeof h); cudaMalloc(&o, sizeof h); cudaMalloc(&c, 8);
  cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
  long long cy;
#define RUN(V) elim<V><<<1, 256>>>(d, o, c); elim<V><<<1, 256>>>(d, o, c); cudaDeviceSynchronize(); \
  cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(out, o, sizeof out, cudaMemcpyDeviceToHost); \
  { double err = 0; for (int i = 0; i < kNB * kNB; ++i) err = fmax(err, fabs(out[i] - Li[i])); \
  printf("variant %d: %lld cycles (%.0f per pivot), max |Linv - host| = %.3e\n", V, cy, cy / 32.0, err); }
  RUN(0) RUN(1) RUN(2) RUN(3)
  for (int th : {64, 128}) { elim<0><<<1, th>>>(d, o, c); cudaDeviceSynchronize(); }
  return 0;
}
