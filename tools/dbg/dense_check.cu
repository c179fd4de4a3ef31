// Debug harness (not part of the product): checks the dense toolkit and the per-individual sweep against
// straightforward host code. Build: see tools/dbg/Makefile.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
#include "../../magmaclust_b200/csrc/common.cuh"
#include "../../magmaclust_b200/csrc/indiv.cuh"
using namespace mcb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

static void host_inv(std::vector<double> A, int n, std::vector<double>& out, double& logdet) {
  // Gauss-Jordan with partial pivoting
  out.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) out[(size_t)i * n + i] = 1.0;
  logdet = 0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i) if (fabs(A[(size_t)i * n + k]) > fabs(A[(size_t)p * n + k])) p = i;
    if (p != k) for (int j = 0; j < n; ++j) { std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]); std::swap(out[(size_t)k * n + j], out[(size_t)p * n + j]); }
    double d = A[(size_t)k * n + k];
    logdet += log(fabs(d));
    for (int j = 0; j < n; ++j) { A[(size_t)k * n + j] /= d; out[(size_t)k * n + j] /= d; }
    for (int i = 0; i < n; ++i) if (i != k) {
      double f = A[(size_t)i * n + k];
      if (f == 0) continue;
      for (int j = 0; j < n; ++j) { A[(size_t)i * n + j] -= f * A[(size_t)k * n + j]; out[(size_t)i * n + j] -= f * out[(size_t)k * n + j]; }
    }
  }
}

int main() {
  cudaStream_t st; CK(cudaStreamCreate(&st));
  std::mt19937_64 rng(1);
  std::normal_distribution<double> nd;
  DenseWs ws;
  int* info; CK(cudaMalloc(&info, 16)); CK(cudaMemset(info, 0, 16));
  // ---- gemm_nt
  for (int N : {8, 33, 70, 401}) {
    int ld = (N + 15) & ~15;
    std::vector<double> A((size_t)N * ld, 0.0), B((size_t)N * ld, 0.0), C((size_t)N * ld, 0.0);
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { A[(size_t)i * ld + j] = nd(rng); B[(size_t)i * ld + j] = nd(rng); }
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, A.size() * 8)); CK(cudaMalloc(&dC, A.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), A.size() * 8, cudaMemcpyHostToDevice));
    gemm_nt(st, N, N, N, dA, ld, 0, dB, ld, 0, dC, ld, 0, 1.0, 0.0, 1, nullptr);
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(C.data(), dC, A.size() * 8, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) {
      double s = 0; for (int k = 0; k < N; ++k) s += A[(size_t)i * ld + k] * B[(size_t)j * ld + k];
      err = fmax(err, fabs(s - C[(size_t)i * ld + j]));
    }
    printf("gemm_nt N=%d max err %.3e\n", N, err);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
  }
  // ---- spd_inverse
  for (int N : {5, 20, 32, 33, 64, 100, 195, 401, 513, 700, 768, 800, 1000, 1111}) {
    for (int Z : {1, 3}) {
      if (N > 768 && Z == 3 && N != 800) continue;   // the large-grid (multi-launch) path: one batched case is enough
      int ld = (N + 15) & ~15;
      long long sN = (long long)N * ld;
      std::vector<double> A((size_t)Z * sN, 0.0);
      std::vector<std::vector<double>> dense(Z);
      for (int z = 0; z < Z; ++z) {
        std::vector<double> R((size_t)N * N);
        for (auto& v : R) v = nd(rng);
        dense[z].assign((size_t)N * N, 0.0);
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) {
          double s = (i == j) ? 0.5 * N : 0.0;
          for (int k = 0; k < N; ++k) s += R[(size_t)i * N + k] * R[(size_t)j * N + k];
          dense[z][(size_t)i * N + j] = s;
          A[z * sN + (size_t)i * ld + j] = s;
        }
      }
      double *dA, *dl; CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dl, Z * 8));
      CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
      CK(spd_inverse(st, dA, N, ld, sN, Z, dl, info, ws, nullptr));
      CK(cudaStreamSynchronize(st));
      std::vector<double> out(A.size()), ldv(Z);
      CK(cudaMemcpy(out.data(), dA, A.size() * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(ldv.data(), dl, Z * 8, cudaMemcpyDeviceToHost));
      int flag; CK(cudaMemcpy(&flag, info, 4, cudaMemcpyDeviceToHost));
      double err = 0, lerr = 0;
      for (int z = 0; z < Z; ++z) {
        std::vector<double> ref; double ldr;
        host_inv(dense[z], N, ref, ldr);
        double mx = 0;
        for (auto v : ref) mx = fmax(mx, fabs(v));
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j)
          err = fmax(err, fabs(ref[(size_t)i * N + j] - out[z * sN + (size_t)i * ld + j]) / mx);
        lerr = fmax(lerr, fabs(ldr - ldv[z]));
      }
      printf("spd_inverse N=%d Z=%d rel err %.3e logdet err %.3e info %d\n", N, Z, err, lerr, flag);
      cudaFree(dA); cudaFree(dl);
    }
  }
  // ---- per-individual factor kernel (no scatter): Psi^-1 against host inverse
  {
    const int M = 4; int ns[M] = {1, 7, 30, 50};
    std::vector<int> off(M + 1, 0), gidx; std::vector<long long> woff(M + 1, 0);
    std::vector<double> t, y, th(3 * M);
    for (int i = 0; i < M; ++i) {
      off[i + 1] = off[i] + ns[i]; woff[i + 1] = woff[i] + (long long)ns[i] * ns[i];
      for (int a = 0; a < ns[i]; ++a) { t.push_back(0.2 * a + 0.01 * i); y.push_back(nd(rng)); gidx.push_back(a); }
      th[3 * i] = 1.0 + 0.3 * i; th[3 * i + 1] = 0.5; th[3 * i + 2] = 0.3;
    }
    int P = off[M];
    int *doff, *dg; long long* dw; double *dt, *dy, *dth, *dW, *dp, *dyWy, *dld, *dtau;
    CK(cudaMalloc(&doff, (M + 1) * 4)); CK(cudaMalloc(&dg, P * 4)); CK(cudaMalloc(&dw, (M + 1) * 8));
    CK(cudaMalloc(&dt, P * 8)); CK(cudaMalloc(&dy, P * 8)); CK(cudaMalloc(&dth, 3 * M * 8));
    CK(cudaMalloc(&dW, woff[M] * 8)); CK(cudaMalloc(&dp, P * 8)); CK(cudaMalloc(&dyWy, M * 8)); CK(cudaMalloc(&dld, M * 8));
    CK(cudaMalloc(&dtau, M * 8));
    CK(cudaMemcpy(doff, off.data(), (M + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dg, gidx.data(), P * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, woff.data(), (M + 1) * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dt, t.data(), P * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy, y.data(), P * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dth, th.data(), 3 * M * 8, cudaMemcpyHostToDevice));
    IndivData dd{doff, dw, dg, dt, dy, nullptr, nullptr, nullptr, 0, kSmallMax};
    IndivWork w{dW, dp, dyWy, dld, nullptr, nullptr, nullptr};
    CK(launch_factor_scatter(st, dd, M, 50, dth, dtau, 0, nullptr, 0, 0, nullptr, w, info));
    CK(cudaStreamSynchronize(st));
    std::vector<double> W(woff[M]), ldv(M), yWy(M);
    CK(cudaMemcpy(W.data(), dW, woff[M] * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ldv.data(), dld, M * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(yWy.data(), dyWy, M * 8, cudaMemcpyDeviceToHost));
    int flag; CK(cudaMemcpy(&flag, info, 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < M; ++i) {
      int n = ns[i];
      std::vector<double> Psi((size_t)n * n), ref; double ldr;
      for (int a = 0; a < n; ++a) for (int b = 0; b < n; ++b) {
        double d = t[off[i] + a] - t[off[i] + b];
        Psi[(size_t)a * n + b] = (a == b) ? exp(th[3 * i]) + th[3 * i + 2] * th[3 * i + 2] : exp(th[3 * i] - exp(-th[3 * i + 1]) / 2 * d * d);
      }
      host_inv(Psi, n, ref, ldr);
      double err = 0, mx = 0, q = 0;
      for (auto v : ref) mx = fmax(mx, fabs(v));
      for (int e = 0; e < n * n; ++e) err = fmax(err, fabs(ref[e] - W[woff[i] + e]) / mx);
      for (int a = 0; a < n; ++a) for (int b = 0; b < n; ++b) q += y[off[i] + a] * ref[(size_t)a * n + b] * y[off[i] + b];
      printf("indiv n=%d Winv rel err %.3e logdet err %.3e yWy err %.3e info %d\n", n, err, fabs(ldr - ldv[i]), fabs(q - yWy[i]) / fabs(q), flag);
    }
  }
  return 0;
}
