// FP64 roofline denominators measured on the device the handle is bound to. MEASURED_PEAKS.json (written by
// the driver) carries HBM GB/s and bf16 TFLOP/s only; the kernels of this library are bounded by the FP64
// pipe, so bench.py measures a DFMA loop and a DMMA (mma.sync.m8n8k4.f64) loop and quotes "of measured".
#include "handle.h"

namespace mcb {

__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
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

__global__ void dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
  const double a = threadIdx.x * 1e-3, b = 1e-3;
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

}  // namespace mcb

extern "C" int mcb_measure_fp64_peaks(mcb_handle* h, double out[2]) {
  if (!h || !out) return MCB_EINVAL;
  MCB_CUDA(h, cudaSetDevice(h->device));
  cudaDeviceProp prop;
  MCB_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
  const int sms = prop.multiProcessorCount;
  mcb::DevBuf<double> buf;
  MCB_CUDA(h, buf.ensure((size_t)sms * 8 * 256));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 2; ++which) {
    const int iters = 8000;
    const int blocks = which == 0 ? sms * 8 : sms * 4;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0, h->st);
      if (which == 0) mcb::dfma_peak_kernel<<<blocks, 256, 0, h->st>>>(buf.p, iters, 1.0000001, 1e-9);
      else mcb::dmma_peak_kernel<<<blocks, 256, 0, h->st>>>(buf.p, iters);
      cudaEventRecord(e1, h->st);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double flops = which == 0 ? 2.0 * 16 * iters * (double)blocks * 256
                                    : 2.0 * 256 * 8 * iters * (double)blocks * 8;
    out[which] = flops / best / 1e9;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  buf.release();
  MCB_CUDA(h, cudaGetLastError());
  return MCB_OK;
}
