// Debug: variants of the 32 x 32 pivot-block elimination loop, cycles per variant (one CTA of 256 threads).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kNB = 32;
__device__ __forceinline__ double fast_rcp(double d) {
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r); r = fma(fma(-d, r, 1.0), r, r); return r;
}
template <int V>
__global__ void __launch_bounds__(256, 1) elim(const double* Din, double* out, long long* cyc) {
  __shared__ double colD[2][kNB + 1], rowW[2][kNB], piv[kNB];
  const int tid = threadIdx.x, ti = tid >> 3, tc = (tid & 7) * 4;
  double dreg[4], wreg[4];
  for (int q = 0; q < 4; ++q) { dreg[q] = Din[ti * kNB + tc + q]; wreg[q] = (ti == tc + q) ? 1.0 : 0.0; }
  if (tc == 0) { colD[0][ti] = dreg[0]; if (ti == 0) { colD[0][kNB] = 1.0 / dreg[0]; piv[0] = dreg[0]; } }
  if (ti == 0) for (int q = 0; q < 4; ++q) rowW[0][tc + q] = wreg[q];
  __syncthreads();
  long long t0 = clock64();
  for (int k = 0; k < kNB; ++k) {
    const double* cd = colD[k & 1];
    const double* rw = rowW[k & 1];
    if (ti > k) {
      const double m = cd[ti] * cd[kNB];
      if (V != 3) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = tc + q;
          if (V == 4) {  // branch-free: select operand and target
            const bool isw = c <= k;
            const double o = isw ? rw[c] : cd[c];
            const double upd = (isw || c <= ti) ? m * o : 0.0;
            if (isw) wreg[q] -= upd; else dreg[q] -= upd;
          } else {
            if (c <= k) { if (V != 2) wreg[q] = fma(-m, rw[c], wreg[q]); }
            else if (c <= ti) dreg[q] = fma(-m, cd[c], dreg[q]);
          }
        }
      }
    }
    const int k1 = k + 1;
    if (k1 < kNB) {
      double* cdn = colD[k1 & 1];
      double* rwn = rowW[k1 & 1];
      if (tc <= k1 && k1 < tc + 4 && ti >= k1) {
        const int q = k1 - tc;
        const double v = q == 0 ? dreg[0] : q == 1 ? dreg[1] : q == 2 ? dreg[2] : dreg[3];
        cdn[ti] = v;
        if (ti == k1) { cdn[kNB] = (V == 1) ? 0.5 : fast_rcp(v); piv[k1] = v; }
      }
      if (ti == k1) for (int q = 0; q < 4; ++q) rwn[tc + q] = wreg[q];
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  for (int q = 0; q < 4; ++q) out[ti * kNB + tc + q] = dreg[q] + wreg[q] + piv[ti];
}
int main() {
  double h[kNB * kNB];
  for (int i = 0; i < kNB; ++i) for (int j = 0; j < kNB; ++j) h[i * kNB + j] = (i == j) ? 40.0 : 1.0 / (1 + abs(i - j));
  double *d, *o; long long* c; cudaMalloc(&d, sizeof h); cudaMalloc(&o, sizeof h); cudaMalloc(&c, 8);
  cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
  long long cy;
#define RUN(V) elim<V><<<1, 256>>>(d, o, c); elim<V><<<1, 256>>>(d, o, c); cudaDeviceSynchronize(); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("variant %d: %lld cycles (%.0f per pivot)\n", V, cy, cy / 32.0);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
  return 0;
}
