// Debug: segment profile of the persistent slab inverse (dense_slab.cu compiled with SLAB_PROF) at N = 401.
//   segments: 0 wait L^-1 flag | 1 load L^-1 + stage | 2 V | 3 P | 4 store V,P + arrive | 5 barrier wait |
//             6 look-ahead: diagonal update | 7 look-ahead: factor + publish | 8 trailing: operand loads | 9 rest of phase 3
#include <cstdio>
#include <vector>
#define SLAB_PROF 1
#include "../../magmaclust_b200/csrc/dense_slab.cu"
using namespace mcb;
int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 401;
  const int reps = 20;
  cudaStream_t st; cudaStreamCreate(&st);
  const int ld = (N + 15) & ~15; const long long sN = (long long)N * ld;
  std::vector<double> grid(N); for (int i = 0; i < N; ++i) grid[i] = 10.0 * i / (N - 1);
  const double th[3] = {1.0, 1.0, 0.2};
  double *dg, *dth, *K, *ldv; int* info;
  cudaMalloc(&dg, N * 8); cudaMalloc(&dth, 24); cudaMalloc(&ldv, 8); cudaMalloc(&info, 16); cudaMemset(info, 0, 16);
  cudaMalloc(&K, sN * 8); cudaMemset(K, 0, sN * 8);
  cudaMemcpy(dg, grid.data(), N * 8, cudaMemcpyHostToDevice); cudaMemcpy(dth, th, 24, cudaMemcpyHostToDevice);
  DenseWs ws;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float tot = 0;
  long long zero[48] = {0};
  for (int r = 0; r < reps + 2; ++r) {
    se_build(st, K, nullptr, nullptr, N, ld, sN, dg, dth, 3, 1, nullptr);
    if (r == 2) cudaMemcpyToSymbolAsync(g_slab_prof, zero, sizeof zero, 0, cudaMemcpyHostToDevice, st);
    cudaEventRecord(e0, st);
    cudaError_t e = spd_inverse_slab(st, K, N, ld, sN, 1, ldv, info, ws, nullptr);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 2) tot += ms;
  }
  long long p[48]; cudaMemcpyFromSymbol(p, g_slab_prof, sizeof p);
  const int nblk = (N + 31) / 32;
  printf("N = %d: %.1f us per inverse (with profiling stamps), %d pivot blocks\n", N, tot * 1e3 / reps, nblk);
  const char* cls[3] = {"look-ahead CTA (critical path), per step", "other pivot-column CTAs, per CTA-step      ", "other CTAs, per CTA-step                  "};
  for (int c = 0; c < 3; ++c) {
    printf("%s:", cls[c]);
    long long s = 0;
    for (int k = 0; k < 10; ++k) { printf(" [%d] %7.0f", k, (double)p[c * 16 + k] / reps / nblk); s += p[c * 16 + k]; }
    printf("  sum/step %7.0f\n", (double)s / reps / nblk);
  }
  return 0;
}
