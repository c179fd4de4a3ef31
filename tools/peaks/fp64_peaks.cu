// FP64 roofline denominators for B200 (sm_100a): DFMA loop, DMMA (mma.sync f64) loops,
// cuBLAS DGEMM 8192^3, red.global.add.f64 throughput. Prints one JSON object.
// MEASURED_PEAKS.json (driver-written) has no FP64 entry; this fills that gap.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs per lane. 8*8*4 = 256 FMA per warp instr.
__global__ void dmma884_kernel(double* out, int iters) {
  double c[8][2];
  double a = threadIdx.x * 1e-3, b = 1e-3;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8 (sm_90+): A 4 regs, B 2 regs, C 4 regs. 16*8*8 = 1024 FMA per warp instr.
__global__ void dmma1688_kernel(double* out, int iters) {
  double c[8][4];
  double a0 = threadIdx.x * 1e-3, a1 = 1e-4, a2 = 2e-4, a3 = 3e-4, b0 = 1e-3, b1 = 2e-3;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16 (sm_90+): A 8 regs, B 4 regs, C 4 regs. 2048 FMA per warp instr.
__global__ void dmma16816_kernel(double* out, int iters) {
  double c[6][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i * 1e-4;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (i + 1);
#pragma unroll
  for (int i = 0; i < 6; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// red.global.add.f64 into `span` distinct addresses (spread), per-thread pseudo-random address
__global__ void red_kernel(double* acc, unsigned span, int iters) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    atomicAdd(&acc[(s >> 8) % span], 1.0);
  }
}

__global__ void exp_kernel(double* out, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-6, s = 0;
  for (int it = 0; it < iters; ++it) { s += exp(-x); x += 1e-3; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  {
    int iters = 20000, blocks = sms * 8, thr = 256;
    float ms = time_ms([&] { dfma_kernel<<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 16 * iters * (double)blocks * thr;
    printf(", \"dfma_tflops\": %.2f", fl / ms / 1e9);
  }
  {
    int iters = 20000, blocks = sms * 4, thr = 256;
    float ms = time_ms([&] { dmma884_kernel<<<blocks, thr>>>(out, iters); });
    double fl = 2.0 * 256 * 8 * iters * (double)blocks * (thr / 32);
    printf(", \"dmma_m8n8k4_tflops\": %.2f", fl / ms / 1e9);
  }
  {
    int iters = 10000, blocks = sms * 4, thr = 256;
    float ms = time_ms([&] { dmma1688_kernel<<<blocks, thr>>>(out, iters); });
    double fl = 2.0 * 1024 * 8 * iters * (double)blocks * (thr / 32);
    printf(", \"dmma_m16n8k8_tflops\": %.2f", fl / ms / 1e9);
  }
  {
    int iters = 5000, blocks = sms * 4, thr = 256;
    float ms = time_ms([&] { dmma16816_kernel<<<blocks, thr>>>(out, iters); });
    double fl = 2.0 * 2048 * 6 * iters * (double)blocks * (thr / 32);
    printf(", \"dmma_m16n8k16_tflops\": %.2f", fl / ms / 1e9);
  }
  {
    int iters = 2000, blocks = sms * 8, thr = 256;
    float ms = time_ms([&] { exp_kernel<<<blocks, thr>>>(out, iters, 0.5); });
    printf(", \"exp_f64_gops\": %.2f", (double)iters * blocks * thr / ms / 1e6);
  }
  for (unsigned span : {1u << 10, 1u << 16, 1u << 20, 1u << 24}) {
    double* acc; CK(cudaMalloc(&acc, sizeof(double) * span)); CK(cudaMemset(acc, 0, sizeof(double) * span));
    int iters = 200, blocks = sms * 16, thr = 256;
    float ms = time_ms([&] { red_kernel<<<blocks, thr>>>(acc, span, iters); });
    printf(", \"red_f64_gatomics_span%u\": %.2f", span, (double)iters * blocks * thr / ms / 1e6);
    CK(cudaFree(acc));
  }
  {
    int n = 8192; double *A, *B, *C;
    CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
    CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
    cublasHandle_t h; cublasCreate(&h); double al = 1, be = 0;
    float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 3);
    printf(", \"cublas_dgemm_8192_tflops\": %.2f", 2.0 * n * n * (double)n / ms / 1e9);
    n = 512;
    ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 5);
    printf(", \"cublas_dgemm_512_tflops\": %.3f, \"cublas_dgemm_512_us\": %.1f", 2.0 * n * n * (double)n / ms / 1e9, ms * 1e3);
    cublasDestroy(h);
  }
  printf("}\n");
  return 0;
}
