// latency microbenchmarks (single warp / single CTA): dependent DFMA, DMMA, LDS, BAR, rcp, div
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, int iters) {
  __shared__ double sm[1024];
  double x = threadIdx.x * 1e-9 + 1.0, a = 1.0000001, b = 1e-9;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i + 1) % 1024;
  __syncthreads();
  long long t0, t1;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, a, b);
  t1 = clock64(); if (threadIdx.x == 0) t[0] = t1 - t0;
  double c0 = x, c1 = b;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  t1 = clock64(); if (threadIdx.x == 0) t[1] = t1 - t0;
  x += c0 + c1;
  int idx = threadIdx.x & 31;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) idx = (int)sm[idx];
  t1 = clock64(); if (threadIdx.x == 0) t[2] = t1 - t0;
  x += idx;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) t[3] = t1 - t0;
  double y = x + 3.0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) y = 1.0 / y + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) t[4] = t1 - t0;
  double w = x + 3.0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w)); r = fma(fma(-w, r, 1.0), r, r); r = fma(fma(-w, r, 1.0), r, r); w = r + 1.5; }
  t1 = clock64(); if (threadIdx.x == 0) t[5] = t1 - t0;
  double s = x + 3.0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) s = sqrt(s) + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) t[6] = t1 - t0;
  double e = 0.5;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) e = exp(-e);
  t1 = clock64(); if (threadIdx.x == 0) t[7] = t1 - t0;
  double l = 2.5;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) l = log(l) + 2.0;
  t1 = clock64(); if (threadIdx.x == 0) t[8] = t1 - t0;
  double v = x;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { sm[threadIdx.x] = v; __syncthreads(); v = sm[(threadIdx.x + 1) % blockDim.x] + 1.0; __syncthreads(); }
  t1 = clock64(); if (threadIdx.x == 0) t[9] = t1 - t0;
  out[threadIdx.x] = x + y + w + s + e + l + v;
}
int main() {
  double* out; long long* t; cudaMalloc(&out, 8 * 1024); cudaMalloc(&t, 8 * 16);
  const char* names[] = {"DFMA dep", "DMMA dep", "LDS chase", "BAR", "1.0/y", "fast rcp", "sqrt", "exp", "log", "STS+BAR+LDS+BAR"};
  for (int threads : {32, 128, 256}) {
    int iters = 1000;
    k<<<1, threads>>>(out, t, iters); cudaDeviceSynchronize();
    k<<<1, threads>>>(out, t, iters); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, t, 8 * 16, cudaMemcpyDeviceToHost);
    printf("threads=%d:", threads);
    for (int i = 0; i < 10; ++i) printf("  %s %.1f", names[i], (double)h[i] / iters);
    printf("\n");
  }
  return 0;
}
