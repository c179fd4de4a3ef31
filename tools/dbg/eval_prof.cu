// Debug: per-phase cycle profile of indiv_eval_v1 (build / sweep / post / X = W S / H = X W / reduce) and of the
// serial L-BFGS update, for one n x n individual replicated over the whole GPU at a chosen CTAs-per-SM.
//   nvcc -DMCB_PROF ... ; ./eval_prof [n] [evals]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#define MCB_PROF 1
#include "../../magmaclust_b200/csrc/indiv_eval.cuh"
#include "../../magmaclust_b200/csrc/indiv_eval2.cuh"
#include "../../magmaclust_b200/csrc/indiv_eval3.cuh"
#include "../../magmaclust_b200/csrc/lbfgs.h"

using namespace mcb;

template <int E, int TT>
__global__ void __launch_bounds__(TT) prof_kernel(int n, int nmax, const double* t, const double* y, const double* c1,
                                                  const double* c2, int evals, double* sink, long long* total) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double out[4];
  __shared__ Lbfgs opt;
  __shared__ int st;
  EvalSm s = eval_carve(sm, nmax);
  int bad = 0;
  eval_load<TT>(s, n, t, y, c1, c2);
  const double x0[3] = {1.0, 1.0, 0.2};
  if (threadIdx.x == 0) { lb_init(opt, 3, x0, 100); st = LB_EVAL; }
  __syncthreads();
  long long t0 = clock64(), tl = 0;
  for (int it = 0; it < evals; ++it) {
    const double a = opt.x[0], b = opt.x[1], c = opt.x[2];
    indiv_eval_v1<E, TT>(s, n, a, b, c, true, out, &bad);
    if (it == 0 && threadIdx.x == 0 && blockIdx.x == 0) { sink[4000] = out[0]; sink[4001] = out[1]; sink[4002] = out[2]; sink[4003] = out[3]; }
    long long q0 = clock64();
    if (threadIdx.x == 0) {
      st = lb_advance(opt, out[0], out + 1);
      if (st != LB_EVAL) { lb_init(opt, 3, x0, 100); st = LB_EVAL; }
    }
    __syncthreads();
    tl += clock64() - q0;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { total[0] = t1 - t0; total[1] = tl; }
  if (threadIdx.x == 0) sink[blockIdx.x] = out[0] + out[1] + bad;
}

template <int E, int TT, int NT, int MINB, int VER = 2>
__global__ void __launch_bounds__(TT, MINB) prof2_kernel(int n, int nmax, const double* t, const double* y, const double* c1,
                                                         const double* c2, int evals, double* sink, long long* total) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double out[4];
  __shared__ Lbfgs opt;
  __shared__ int st;
  EvalSm2<NT> s = eval2_carve<NT>(sm);
  int bad = 0;
  eval2_load<TT, NT>(s, n, t, y, c1, c2);
  int bI[E], bJ[E];
  eval2_decode<E, TT>(n, bI, bJ);
  const double x0[3] = {1.0, 1.0, 0.2};
  if (threadIdx.x == 0) { lb_init(opt, 3, x0, 100); st = LB_EVAL; }
  __syncthreads();
  long long t0 = clock64(), tl = 0;
  for (int it = 0; it < evals; ++it) {
    const double a = opt.x[0], b = opt.x[1], c = opt.x[2];
    if constexpr (VER == 3) indiv_eval_v3<TT, NT>(s, n, a, b, c, true, out, &bad, blockIdx.x);
    else indiv_eval_v2<E, TT, NT>(s, n, bI, bJ, a, b, c, true, out, &bad, blockIdx.x);
    if (it == 0 && threadIdx.x == 0 && blockIdx.x == 0) { sink[4000] = out[0]; sink[4001] = out[1]; sink[4002] = out[2]; sink[4003] = out[3]; }
    long long q0 = clock64();
    if (threadIdx.x == 0) {
      st = lb_advance(opt, out[0], out + 1);
      if (st != LB_EVAL) { lb_init(opt, 3, x0, 100); st = LB_EVAL; }
    }
    __syncthreads();
    tl += clock64() - q0;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { total[0] = t1 - t0; total[1] = tl; }
  if (threadIdx.x == 0) sink[blockIdx.x] = out[0] + out[1] + bad;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 50;
  const int evals = argc > 2 ? atoi(argv[2]) : 40;
  std::vector<double> t(n), y(n), c1(n), c2((size_t)n * n);
  srand(1);
  for (int i = 0; i < n; ++i) { t[i] = 10.0 * i / n + 0.05 * (rand() % 3); y[i] = 20 + sin(t[i]) * 3; c1[i] = 20 + sin(t[i]) * 2.9; }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) c2[(size_t)i * n + j] = c1[i] * c1[j] + exp(1 - 0.18 * (t[i] - t[j]) * (t[i] - t[j])) + (i == j ? 0.04 : 0);
  double *dt, *dy, *dc1, *dc2, *sink; long long* tot;
  cudaMalloc(&dt, n * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&dc1, n * 8); cudaMalloc(&dc2, (size_t)n * n * 8);
  cudaMalloc(&sink, 8 * 4096); cudaMemset(sink, 0, 8 * 4096); cudaMalloc(&tot, 16);
  cudaMemcpy(dt, t.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dc1, c1.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dc2, c2.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice);
  size_t smem = eval_smem_doubles(n) * 8;
  const int nblk = eval_num_blocks(n);
  const int E = (nblk + 127) / 128;
  printf("n = %d, blocks = %d, E = %d, smem = %zu B\n", n, nblk, E, smem);
  auto run = [&](auto kern, int ctas, const char* label, int threads = 128) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long zero[16] = {0};
    kern<<<ctas, threads, smem>>>(n, n, dt, dy, dc1, dc2, 2, sink, tot);
    cudaDeviceSynchronize();
    cudaMemcpyToSymbol(g_eval_prof, zero, sizeof zero);
    kern<<<ctas, threads, smem>>>(n, n, dt, dy, dc1, dc2, evals, sink, tot);
    cudaError_t e = cudaDeviceSynchronize();
    long long p[16], tt[2];
    cudaMemcpyFromSymbol(p, g_eval_prof, sizeof p);
    cudaMemcpy(tt, tot, 16, cudaMemcpyDeviceToHost);
    double fg[4];
    cudaMemcpy(fg, sink + 4000, 32, cudaMemcpyDeviceToHost);
    printf("   f,g = %.12e %.12e %.12e %.12e\n", fg[0], fg[1], fg[2], fg[3]);
    printf("%-18s %s  cycles/eval %8.0f | build %6.0f sweep %6.0f post %6.0f X=WS %6.0f H=XW %6.0f reduce %6.0f tail %5.0f | lbfgs %6.0f\n",
           label, cudaGetErrorString(e), (double)tt[0] / evals, (double)p[0] / evals, (double)p[1] / evals, (double)p[2] / evals,
           (double)p[3] / evals, (double)p[4] / evals, (double)p[5] / evals, 0.0, (double)tt[1] / evals);
  };
  if (E == 1) {
    run(prof_kernel<1, 128>, 1, "1 CTA");
    run(prof_kernel<1, 128>, 148, "1 CTA/SM");
    run(prof_kernel<1, 128>, 296, "2 CTA/SM");
  } else if (E == 2) {
    run(prof_kernel<2, 128>, 1, "1 CTA");
    run(prof_kernel<2, 128>, 148, "1 CTA/SM");
    run(prof_kernel<2, 128>, 296, "2 CTA/SM");
  }
  smem = 8 * (size_t)(n <= 32 ? Eval2Geo<4>::DOUBLES : n <= 56 ? Eval2Geo<7>::DOUBLES : Eval2Geo<10>::DOUBLES);
  printf("v2: smem = %zu B\n", smem);
  if (n <= 32) {
    run(prof2_kernel<1, 64, 4, 8>, 1, "v2 1 CTA", 64);
    run(prof2_kernel<1, 64, 4, 8>, 148 * 8, "v2 8 CTA/SM", 64);
  } else if (n <= 56) {
    run(prof2_kernel<1, 128, 7, 4>, 1, "v2 1 CTA");
    run(prof2_kernel<1, 128, 7, 4>, 148, "v2 1 CTA/SM");
    run(prof2_kernel<1, 128, 7, 4>, 296, "v2 2 CTA/SM");
    run(prof2_kernel<1, 128, 7, 4>, 592, "v2 4 CTA/SM");
  } else {
    run(prof2_kernel<2, 128, 10, 2>, 1, "v2 1 CTA");
    run(prof2_kernel<2, 128, 10, 2>, 296, "v2 2 CTA/SM");
  }
  printf("v3:\n");
  if (n <= 30) {
    run(prof2_kernel<1, 64, 4, 8, 3>, 1, "v3 1 CTA", 64);
    run(prof2_kernel<1, 64, 4, 8, 3>, 148 * 8, "v3 8 CTA/SM", 64);
  } else if (n <= 54) {
    run(prof2_kernel<1, 128, 7, 4, 3>, 1, "v3 1 CTA");
    run(prof2_kernel<1, 128, 7, 4, 3>, 148, "v3 1 CTA/SM");
    run(prof2_kernel<1, 128, 7, 4, 3>, 296, "v3 2 CTA/SM");
    run(prof2_kernel<1, 128, 7, 4, 3>, 592, "v3 4 CTA/SM");
  } else if (n <= 78) {
    run(prof2_kernel<1, 128, 10, 2, 3>, 1, "v3 1 CTA");
    run(prof2_kernel<1, 128, 10, 2, 3>, 296, "v3 2 CTA/SM");
  }
  return 0;
}
