// Debug harness for the 2-level factorisation of the 32 x 32 pivot block (tools only; the device function lives in
// magmaclust_b200/csrc/factor32.cuh and is shared with dense_slab.cu).
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>
#define F32_PROFILE
#include "../../magmaclust_b200/csrc/factor32.cuh"
using namespace mcb;
__global__ void __launch_bounds__(256, 1) k(const double* Din, int nb, double* Linv, double* logdet, int* bad, long long* cyc) {
  __shared__ __align__(16) double Dm[32 * kF32Ld];
  __shared__ __align__(16) double Li[32 * kF32Ld];
  __shared__ __align__(16) double scratch[kF32Scratch];
  const int tid = threadIdx.x;
  for (int e = tid; e < 1024; e += 256) {
    const int r = e >> 5, c = e & 31;
    Dm[r * kF32Ld + c] = (r < nb && c < nb) ? Din[r * 32 + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  long long t0 = clock64();
  double ld = 0.0;
  __shared__ int b; if (tid == 0) b = 0; __syncthreads();
  factor32(Dm, Li, scratch, &ld, &b);
  if (tid == 0) *bad = b;
  long long t1 = clock64();
  if (tid == 0) { cyc[0] = t1 - t0; }
  if (tid >= 224) {
    for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
    if (tid == 224) { *logdet = ld; }
  }
  for (int e = tid; e < 1024; e += 256) Linv[e] = Li[(e >> 5) * kF32Ld + (e & 31)];
}
int main() {
  const int NB = 32;
  for (int nb : {32, 17, 3}) {
    double h[NB * NB], L[NB * NB] = {0}, Li[NB * NB] = {0}, out[NB * NB];
    for (int i = 0; i < NB; ++i) for (int j = 0; j < NB; ++j) h[i * NB + j] = (i == j) ? 1.04 + 2.718 : 2.718 * exp(-0.002 * (i - j) * (i - j));
    for (int i = 0; i < NB; ++i) for (int j = 0; j < NB; ++j) if (i >= nb || j >= nb) h[i * NB + j] = (i == j) ? 1.0 : 0.0;
    double ldet = 0;
    for (int j = 0; j < NB; ++j) { double s = h[j * NB + j]; for (int k = 0; k < j; ++k) s -= L[j * NB + k] * L[j * NB + k]; L[j * NB + j] = sqrt(s); ldet += 2 * log(L[j * NB + j]);
      for (int i = j + 1; i < NB; ++i) { double t = h[i * NB + j]; for (int k = 0; k < j; ++k) t -= L[i * NB + k] * L[j * NB + k]; L[i * NB + j] = t / L[j * NB + j]; } }
    for (int c = 0; c < NB; ++c) for (int i = c; i < NB; ++i) { double s = (i == c) ? 1.0 : 0.0; for (int k = c; k < i; ++k) s -= L[i * NB + k] * Li[k * NB + c]; Li[i * NB + c] = s / L[i * NB + i]; }
    double *d, *o, *ld; long long* c; int* bad;
    cudaMalloc(&d, sizeof h); cudaMalloc(&o, sizeof h); cudaMalloc(&c, 8); cudaMalloc(&ld, 8); cudaMalloc(&bad, 4);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    k<<<1, 256>>>(d, nb, o, ld, bad, c); k<<<1, 256>>>(d, nb, o, ld, bad, c);
    cudaError_t e = cudaDeviceSynchronize();
    long long cy; double ldg; int b;
    cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(out, o, sizeof out, cudaMemcpyDeviceToHost);
    cudaMemcpy(&ldg, ld, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&b, bad, 4, cudaMemcpyDeviceToHost);
    double err = 0, mx = 0;
    for (int i = 0; i < NB * NB; ++i) { err = fmax(err, fabs(out[i] - Li[i])); mx = fmax(mx, fabs(Li[i])); }
    long long pr[8]; cudaMemcpyFromSymbol(pr, f32_prof, sizeof pr); printf("  phases (2 runs): diag %lld panel %lld trail %lld linv1 %lld linv2 %lld | tile: load %lld elim %lld store %lld\n", pr[0], pr[1], pr[2], pr[3], pr[4], pr[5], pr[6], pr[7]);
    printf("nb=%2d: %s  %lld cycles, max |Linv - host| = %.3e (max |Linv| %.2f), logdet err %.2e, bad %d\n", nb, cudaGetErrorString(e), cy, err, mx, fabs(ldg - ldet), b);
  }
  return 0;
}
