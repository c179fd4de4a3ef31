// Debug: per-phase clock64 stamps of the slab inverse (build with -DMCB_SLAB_PROFILE, includes the kernel source)
#define MCB_SLAB_PROFILE
#include "../../magmaclust_b200/csrc/dense_slab.cu"
#include <cstdio>
#include <vector>
using namespace mcb;
int main() {
  int N = 401, Z = 1, ld = 416; long long sN = (long long)N * ld;
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  std::vector<double> A((size_t)sN, 0.0);
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { double d = 10.0 * (i - j) / 400.0; A[(size_t)i * ld + j] = (i == j) ? 2.718 + 0.04 : exp(1.0 - 0.5 * exp(-1.0) * d * d); }
  double *dA, *ldv; int* info; cudaMalloc(&dA, sN * 8); cudaMalloc(&ldv, 8); cudaMalloc(&info, 4); cudaMemset(info, 0, 4);
  DenseWs ws;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(dA, A.data(), sN * 8, cudaMemcpyHostToDevice);
    cudaError_t e = spd_inverse_slab(st, dA, N, ld, sN, Z, ldv, info, ws, nullptr);
    cudaStreamSynchronize(st);
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long prof[64][12];
  cudaMemcpyFromSymbol(prof, g_slab_prof, sizeof prof);
  printf("lookahead CTA (I = b + 1) timeline, cycles: ph2 | barrier B | diag update | factor32 | rest of ph3 | barrier A (to next ph2)\n");
  for (int b = 0; b < 12; ++b)
    printf("b=%2d  ph2 %6lld  barB %6lld  upd %6lld  factor %6lld  rest %6lld  barA %6lld | step %6lld\n", b,
           prof[b][2] - prof[b][1], prof[b][3] - prof[b][2], prof[b][4] - prof[b][3], prof[b][5] - prof[b][4],
           prof[b][7] - prof[b][5], prof[b + 1][1] - prof[b][7], prof[b + 1][1] - prof[b][1]);
  long long prof2[64][12];
  cudaMemcpyFromSymbol(prof2, g_slab_prof2, sizeof prof2);
  printf("trailing-update CTA (I = b + 2): ph2 | barrier B | ph3\n");
  for (int b = 0; b < 11; ++b)
    printf("b=%2d  ph2 %6lld  barB %6lld  ph3 %6lld\n", b, prof2[b][2] - prof2[b][1], prof2[b][3] - prof2[b][2], prof2[b][7] - prof2[b][3]);
  return 0;
}
