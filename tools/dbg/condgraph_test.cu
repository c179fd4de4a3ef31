// Feasibility probe: WHILE conditional graph node whose body holds a cooperative-launch kernel (grid barrier) and a
// controller kernel that decides whether to iterate again. Debug tool only.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void coop_kernel(double* x, int n) {
  cg::grid_group g = cg::this_grid();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] += 1.0;
  g.sync();
  if (i == 0) x[n] = x[0] * n;
}
__global__ void control_kernel(double* x, int n, int* iters, cudaGraphConditionalHandle h) {
  ++*iters;
  cudaGraphSetConditional(h, x[n] < 10.0 * n ? 1u : 0u);
}
#define CK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(_e)); return 1; } } while (0)
int main() {
  const int n = 13 * 256;
  double* x; int* it;
  CK(cudaMalloc(&x, (n + 1) * 8)); CK(cudaMemset(x, 0, (n + 1) * 8)); CK(cudaMalloc(&it, 4)); CK(cudaMemset(it, 0, 4));
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {}; np.type = cudaGraphNodeTypeConditional; np.conditional.handle = h;
  np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
  cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, nullptr, 0, &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  { int nn = n; void* args[] = {&x, &nn}; CK(cudaLaunchCooperativeKernel((void*)coop_kernel, dim3(13), dim3(256), args, 0, st)); }
  control_kernel<<<1, 1, 0, st>>>(x, n, it, h);
  CK(cudaStreamEndCapture(st, nullptr));
  cudaGraphExec_t ge; CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemsetAsync(x, 0, (n + 1) * 8, st)); CK(cudaMemsetAsync(it, 0, 4, st));
    cudaEventRecord(e0, st);
    CK(cudaGraphLaunch(ge, st));
    cudaEventRecord(e1, st);
    CK(cudaStreamSynchronize(st));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int hi; double s; CK(cudaMemcpy(&hi, it, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&s, x + n, 8, cudaMemcpyDeviceToHost));
    printf("rep %d: iterations %d (expect 10), sum %.0f (expect %d), %.1f us per iteration\n", rep, hi, s, 10 * n, ms * 1e3 / hi);
  }
  return 0;
}
