// Debug harness: device timings of the dense pieces of one theta_k evaluation at N = 401.
#include <cstdio>
#include <vector>
#include <random>
#include "../../magmaclust_b200/csrc/common.cuh"
using namespace mcb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
template <class F> float time_us(cudaStream_t st, F f, int reps = 20) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(e1, st); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / reps;
}
int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 401;
  int Z = argc > 2 ? atoi(argv[2]) : 1;
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int ld = (N + 15) & ~15; long long sN = (long long)N * ld;
  std::vector<double> grid(N); for (int i = 0; i < N; ++i) grid[i] = 10.0 * i / (N - 1);
  std::vector<double> th(3 * Z); for (int z = 0; z < Z; ++z) { th[3*z] = 1.0; th[3*z+1] = 1.0; th[3*z+2] = 0.2; }
  double *dg, *dth, *K, *D1, *D2, *X, *C, *ldv, *red; int* info;
  CK(cudaMalloc(&dg, N * 8)); CK(cudaMalloc(&dth, 3 * Z * 8)); CK(cudaMalloc(&ldv, Z * 8)); CK(cudaMalloc(&red, 64 * 8)); CK(cudaMalloc(&info, 16));
  CK(cudaMemset(info, 0, 16));
  for (double** p : {&K, &D1, &D2, &X, &C}) { CK(cudaMalloc(p, Z * sN * 8)); CK(cudaMemset(*p, 0, Z * sN * 8)); }
  CK(cudaMemcpy(dg, grid.data(), N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dth, th.data(), 3 * Z * 8, cudaMemcpyHostToDevice));
  DenseWs ws;
  CK(spd_inverse_reserve(ws, N, Z));
  printf("N=%d Z=%d\n", N, Z);
  printf("se_build            %8.1f us\n", time_us(st, [&] { se_build(st, K, D1, D2, N, ld, sN, dg, dth, 3, Z, nullptr); }));
  printf("build + inverse     %8.1f us\n", time_us(st, [&] { se_build(st, K, nullptr, nullptr, N, ld, sN, dg, dth, 3, Z, nullptr); spd_inverse(st, K, N, ld, sN, Z, ldv, info, ws, nullptr); }));
  se_build(st, C, nullptr, nullptr, N, ld, sN, dg, dth, 3, Z, nullptr);
  printf("gemm_nt             %8.1f us\n", time_us(st, [&] { gemm_nt(st, N, N, N, K, ld, sN, C, ld, sN, X, ld, sN, 1.0, 0.0, Z, nullptr); }));
  printf("gemm_nt_trace       %8.1f us\n", time_us(st, [&] { gemm_nt_trace(st, N, X, ld, sN, K, ld, sN, D1, D2, ld, sN, red, 4, Z, nullptr); }));
  printf("group_reduce        %8.1f us\n", time_us(st, [&] { group_reduce(st, N, ld, sN, K, C, sN, D1, D2, red, Z, nullptr); }));
  int flag; CK(cudaMemcpy(&flag, info, 4, cudaMemcpyDeviceToHost)); printf("info %d\n", flag);
  return 0;
}
