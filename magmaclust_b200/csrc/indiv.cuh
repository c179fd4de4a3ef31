// Launchers of the per-individual kernels (indiv.cu).
#pragma once
#include <cuda_runtime.h>

namespace mcb {

// the training set in CSR form (device pointers)
struct IndivData {
  const int* off;          // [M+1]
  const long long* woff;   // [M+1] offsets into n_i^2-sized per-individual blocks
  const int* gidx;         // [P] grid index of each row
  const double* t;         // [P] timestamps
  const double* y;         // [P] outputs
  // individuals with more than small_max observations do not fit the shared-memory kernels: those kernels skip them and
  // the global-memory kernels of indiv_big.cu (one CTA each, workspace in HBM/L2) take them. nbig == 0: none.
  const int* big_ids;          // [nbig] their indices, ascending
  const long long* big_wsoff;  // [nbig + 1] offsets into big_ws: (n + 2)^2 + n^2 doubles each
  double* big_ws;
  int nbig, small_max;
};
constexpr int kSmallMax = 80;   // tile buckets of the evaluators (indiv.cu: eval2_shape)

// per-individual results kept on the device between kernels
struct IndivWork {
  double* Winv;    // [sum n_i^2] Psi_i^-1 at the E-step's theta_i
  double* pvec;    // [P] Psi_i^-1 y_i
  double* yWy;     // [M]
  double* logdet;  // [M] log|Psi_i|
  double* c1;      // [P]  corr1 (CF.R:236-245)
  double* c2;      // [sum n_i^2] corr2
  double* Fi;      // [M] F_i at the E-step's theta_i with the new tau
};

size_t factor_smem_bytes(int nmax);
size_t tau_smem_bytes(int nmax, int K);
size_t eval_smem_bytes(int nmax);

cudaError_t launch_factor_scatter(cudaStream_t s, const IndivData& dd, int M, int nmax, const double* theta_i,
                                  const double* tau, int K, double* acc, int ldN, long long strideA, double* wk,
                                  const IndivWork& w, int* info);
cudaError_t launch_tau(cudaStream_t s, const IndivData& dd, int M, int nmax, int K, const double* mean,
                       const double* cov, int ldN, long long strideC, const double* pi, const IndivWork& w,
                       double* tau_new, double* logL, int fixed_tau);
cudaError_t launch_indiv_eval(cudaStream_t s, const IndivData& dd, int M, int nmax, const IndivWork& w,
                              const double* theta, int theta_stride, int want_grad, double* f, double* g,
                              double* sum4, int* info);
constexpr int kMstepCtlWords = 2 + 1024 + 1 + 1024;
// f_out (optional, [M]): objective at the returned theta; *f_out_written tells whether the kernel variant filled it
cudaError_t launch_indiv_mstep(cudaStream_t s, const IndivData& dd, int M, int nmax, const IndivWork& w,
                               const double* theta0, double* theta_out, int maxit, int* nfev, int* niter,
                               int* status, int* ctl /* [kMstepCtlWords] ints */, int reserve_sms,
                               const int* order /* [M] processing order or NULL */, double* f_out = nullptr,
                               int* f_out_written = nullptr);

// indiv_big.cu -- the same four operations for the individuals listed in dd.big_ids (any n_i). theta / f / g are indexed by
// the individual, or by its position in big_ids when by_slot; active (optional, [nbig]) masks finished optimisations.
cudaError_t launch_big_factor(cudaStream_t s, const IndivData& dd, const double* theta_i, const double* tau, int K,
                              double* acc, int ldN, long long strideA, double* wk, const IndivWork& w, int* info);
cudaError_t launch_big_tau(cudaStream_t s, const IndivData& dd, int K, const double* mean, const double* cov, int ldN,
                           long long strideC, const double* pi, const IndivWork& w, double* tau_new, double* logL,
                           int fixed_tau);
cudaError_t launch_big_eval(cudaStream_t s, const IndivData& dd, const IndivWork& w, const double* theta, int theta_stride,
                            int by_slot, const int* active, int want_grad, double* f, double* g, double* sum4, int* info);

}  // namespace mcb
