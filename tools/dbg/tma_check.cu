// Debug harness: the TMA-fed rank-256 update (dense_tma.cu) against a host evaluation of the same formula, and a
// one-hot probe of the contraction index pairing.   usage: tma_check [N]
#include <cstdio>
#include <cmath>
#include <vector>
#include <random>
#include "../../magmaclust_b200/csrc/common.cuh"
namespace mcb {
bool sweep_update_tma(cudaStream_t s, int N, int K, const double* V, int ldv, int rows, int Z, double* C, int ldc,
                      long long sC, const double* P, int ldp, long long sP, int c0, int nb, double sign, cudaError_t* err);
}
using namespace mcb;
__global__ void cmp_kernel(const double* a, const double* b, long long n, unsigned long long* cnt, long long* first) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  unsigned long long bad = 0;
  for (; i < n; i += (long long)gridDim.x * blockDim.x)
    if (a[i] != b[i]) { ++bad; atomicMin((unsigned long long*)first, (unsigned long long)i); }
  if (bad) atomicAdd(cnt, bad);
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
int main(int argc, char** argv) {
  setenv("MCB_TMA_MIN_N", "0", 1);   // the harness exercises the kernel at every size
  const int N = argc > 1 ? atoi(argv[1]) : 1024, K = 256, Z = 1;
  const int rows = ((N + 31) / 32) * 32, ld = (N + 15) & ~15;
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  std::mt19937_64 rng(1); std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> V((size_t)rows * K, 0.0), C((size_t)N * ld, 0.0), P((size_t)rows * K);
  for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) V[(size_t)r * K + k] = nd(rng);
  for (auto& v : P) v = nd(rng);
  for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) C[(size_t)i * ld + j] = C[(size_t)j * ld + i] = nd(rng);
  double *dV, *dC, *dP;
  CK(cudaMalloc(&dV, V.size() * 8)); CK(cudaMalloc(&dC, C.size() * 8)); CK(cudaMalloc(&dP, P.size() * 8));
  for (int probe = -1; probe < (argc > 2 ? atoi(argv[2]) : 16); ++probe) {
    std::vector<double> Vp = V;
    if (probe >= 0) { for (auto& v : Vp) v = 0.0; for (int r = 0; r < N; ++r) Vp[(size_t)r * K + probe] = 1.0 + r; }
    CK(cudaMemcpy(dV, Vp.data(), Vp.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dP, P.data(), P.size() * 8, cudaMemcpyHostToDevice));
    const int c0 = 256, nb = 256; const double sign = 1.0;
    cudaError_t e;
    bool used = sweep_update_tma(st, N, K, dV, K, rows, Z, dC, ld, (long long)N * ld, dP, K, (long long)rows * K, c0, nb, sign, &e);
    CK(e); CK(cudaStreamSynchronize(st));
    if (!used) { printf("TMA path not taken\n"); return 1; }
    std::vector<double> out(C.size());
    CK(cudaMemcpy(out.data(), dC, out.size() * 8, cudaMemcpyDeviceToHost));
    double worst = 0; int wi = -1, wj = -1; long long nbad = 0;
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) {
      const int gi = i > j ? i : j, gq = i > j ? j : i;   // lower representative
      const bool ip = gi >= c0 && gi < c0 + nb, jp = gq >= c0 && gq < c0 + nb;
      double r;
      if (!ip && !jp) { double s = 0; for (int k = 0; k < K; ++k) s += Vp[(size_t)gi * K + k] * Vp[(size_t)gq * K + k]; r = C[(size_t)gi * ld + gq] - s; }
      else if (jp) r = P[(size_t)gi * K + (gq - c0)];
      else r = P[(size_t)gq * K + (gi - c0)];
      r *= sign;
      const double d = fabs(out[(size_t)i * ld + j] - r);
      if (d > 1e-9 * (1.0 + fabs(r))) { ++nbad; if (d > worst) { worst = d; wi = i; wj = j; } }
    }
    printf("probe %2d: bad %lld worst %.3e at (%d, %d)\n", probe, nbad, worst, wi, wj);
    if (probe >= 0 && nbad) {
      // which contraction column did row i actually pair with? out = C - v_i v_j if paired correctly
      int i = 700 < N ? 700 : N - 1, j = 3;
      printf("   (700,3): got %.6f want %.6f C %.6f\n", out[(size_t)i * ld + j], C[(size_t)i * ld + j] - (1.0 + i) * (1.0 + j), C[(size_t)i * ld + j]);
    }
  }
  // ---- stress: many launches back to back on the same input; every result must equal the first one bit for bit ----
  {
    const int reps = argc > 3 ? atoi(argv[3]) : 0;
    const int flush = argc > 4 ? atoi(argv[4]) : 0;
    if (reps > 0) {
      double *dC0, *dRef; CK(cudaMalloc(&dC0, C.size() * 8)); CK(cudaMalloc(&dRef, C.size() * 8));
      void* scratch = nullptr; if (flush) CK(cudaMalloc(&scratch, 512u << 20));
      unsigned long long* dCnt; long long* dFirst;
      CK(cudaMalloc(&dCnt, 8 * reps)); CK(cudaMalloc(&dFirst, 8 * reps));
      CK(cudaMemset(dCnt, 0, 8 * reps)); CK(cudaMemset(dFirst, 0x7f, 8 * reps));
      CK(cudaMemcpy(dV, V.data(), V.size() * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dC0, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
      for (int r = 0; r < reps; ++r) {
        cudaError_t e;
        CK(cudaMemcpyAsync(dC, dC0, C.size() * 8, cudaMemcpyDeviceToDevice, st));
        if (flush) CK(cudaMemsetAsync(scratch, r & 255, 512u << 20, st));
        sweep_update_tma(st, N, K, dV, K, rows, Z, dC, ld, (long long)N * ld, dP, K, (long long)rows * K, 256, 256, 1.0, &e);
        CK(e);
        if (r == 0) CK(cudaMemcpyAsync(dRef, dC, C.size() * 8, cudaMemcpyDeviceToDevice, st));
        else cmp_kernel<<<592, 256, 0, st>>>(dC, dRef, (long long)C.size(), dCnt + r, dFirst + r);
      }
      CK(cudaStreamSynchronize(st));
      std::vector<unsigned long long> cnt(reps); std::vector<long long> first(reps);
      CK(cudaMemcpy(cnt.data(), dCnt, 8 * reps, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(first.data(), dFirst, 8 * reps, cudaMemcpyDeviceToHost));
      int bad_runs = 0;
      for (int r = 1; r < reps; ++r) if (cnt[r]) { if (++bad_runs <= 12) printf("run %d: %llu elements differ, first at (%lld, %lld)\n", r, cnt[r], first[r] / ld, first[r] % ld); }
      printf("stress (flush %d): %d of %d runs differ from the first\n", flush, bad_runs, reps - 1);
    }
  }
  return 0;
}
