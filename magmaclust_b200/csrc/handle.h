// The opaque handle behind the C ABI: one GPU, its stream, the resident training set, model state and
// workspaces.
#pragma once
#include <map>

#include "common.cuh"
#include "indiv.cuh"

struct mcb_handle {
  int device = 0;
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;              // second stream: theta_i optimisations overlapped with theta_k
  cudaStream_t st_copy = nullptr;          // result copies of mcb_vem_iteration_host behind the M-step
  cudaEvent_t ev_copy0 = nullptr, ev_copy1 = nullptr;
  cudaStream_t st_hi = nullptr;            // high-priority stream: the theta_k chain next to host-driven theta_i evaluations
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  mcb::DevBuf<int> mstep_ctl;               // work queue + SM reservation flags of indiv_mstep_kernel
  std::string err;

  // ---- communicator (NCCL, loaded lazily; nranks == 1 means single GPU) ----
  int nranks = 1, rank = 0;
  void* comm = nullptr;
  // peer-memory mailboxes for the scalar all-reduces of the optimiser loops (comm.cu: tiny_allreduce_kernel)
  double* mbox = nullptr;                  // this GPU's mailbox
  double** mbox_peers = nullptr;           // device array [nranks]: every rank's mailbox as seen from this GPU
  std::vector<void*> mbox_opened;          // pointers obtained with cudaIpcOpenMemHandle (closed in comm_destroy)
  unsigned* mbox_epoch_dev = nullptr;      // call counter (device resident: the all-reduce also runs inside graph loops)
  int* mbox_err = nullptr;                 // device flag: a peer did not show up in time

  // ---- training set (local shard) ----
  bool has_data = false;
  int M = 0, N = 0, ld = 0, nmax = 0;       // nmax: most observations among the individuals of the shared-memory kernels
  int nmax_all = 0, nbig = 0;               // individuals above mcb::kSmallMax observations take the global-memory kernels
  std::vector<int> h_big_ids;
  mcb::DevBuf<int> big_ids, big_active;
  mcb::DevBuf<long long> big_wsoff;
  mcb::DevBuf<double> big_ws, big_theta, big_f, big_g;
  long long P = 0, M_total = 0, W2 = 0;
  mcb::DevBuf<int> off, gidx;
  mcb::DevBuf<long long> woff;
  mcb::DevBuf<double> t, y, grid;

  // ---- model state ----
  bool has_state = false, has_post = false, has_cand = false;
  int K = 0;
  mcb::DevBuf<double> theta_i, theta_k, pi, mk, tau;               // current hp / prior means / "old" tau
  mcb::DevBuf<double> theta_i_new, theta_k_new, pi_new;            // M-step candidate
  mcb::DevBuf<double> tau_new, logL;                               // E-step outputs
  std::vector<double> h_theta_k, h_theta_k_new, h_pi, h_pi_new;    // host mirrors (tiny)

  // ---- per-individual work ----
  mcb::DevBuf<double> Winv, pvec, yWy, logdet_i, c1, c2, Fi, fi_tmp, gi_tmp;
  mcb::DevBuf<int> nfev_i, niter_i, status_i, order_i;
  mcb::DevBuf<double> kscr;   // scratch of the stand-alone kern_to_cov / kern_to_inv / dmvnorm entry points
  mcb::DevBuf<double> gscr;   // scratch of mcb_gp_mod (explicit-argument logL_GP_mod / gr_GP_mod)
  mcb::DevBuf<int> gscr_i;
  mcb::DevBuf<double> Fnew;   // F_i at theta_i_new as left by the per-individual M-step kernel
  bool fnew_valid = false;
  double fsum_new = 0.0;     // sum over ALL individuals (all ranks) of F_i at theta_i_new as left by the shared-theta_i optimiser
  bool fsum_valid = false;
  bool nfev_valid = false;   // nfev_i holds the counts of a previous per-individual M-step on this data set

  // ---- dense side ----
  int G = 0;                       // distinct theta_k at the last E-step
  std::vector<int> gmap;           // cluster -> group
  mcb::DevBuf<int> gmap_d;
  mcb::DevBuf<double> theta_g;     // [G][3] device
  mcb::DevBuf<double> Kinv, cov, wk, mean, rvec, pz, logdetK, logdetA;
  std::vector<double> h_logdetK, h_logdetA;
  // evaluator workspace (trial theta_k)
  mcb::DevBuf<double> eK, eD1, eD2, eCsum, eX, etheta, ered, elogdet;
  mcb::DevBuf<int> egmap;
  bool csum_valid = false;
  std::vector<int> csum_gmap;
  struct EvalGraph { cudaGraphExec_t exec; long long kernels; };
  std::map<int, EvalGraph> eval_graphs;   // captured evaluation sequences, keyed by (G, want_grad, resident)
  // device-driven theta_k optimisation (common theta_k): a WHILE graph whose body is one objective evaluation
  cudaGraphExec_t mopt_exec = nullptr;
  int mopt_budget = 0;                    // ws.max_ctas it was captured with
  long long mopt_pre = 0, mopt_body = 0;  // kernels before the loop / per evaluation
  mcb::DevBuf<double> mopt_state;         // the L-BFGS state (Lbfgs struct) + result slots
  // device-driven shared-theta_i optimisation (common_hp_i): the same kind of WHILE graph around the batched evaluation
  cudaGraphExec_t iopt_exec = nullptr;
  bool iopt_failed = false;               // the graph could not be built once: stay on the host loop
  long long iopt_pre = 0, iopt_body = 0;
  mcb::DevBuf<double> iopt_state;
  cudaStream_t st_cap = nullptr;          // capture-only stream for conditional-node bodies
  mcb::DenseWs ws;       // training path only: its buffers are baked into the captured graphs (sized by alloc_state, never grown later)
  mcb::DenseWs ws_aux;   // posterior_mu_k on another grid and the stand-alone kern_to_inv / dmvnorm: may grow at any time
  mcb::DevBuf<int> info;
  mcb::DevBuf<double> scal;        // small device scalars

  // ---- prediction posterior ----
  bool has_pred = false;
  int T = 0, ldT = 0, Kp = 0;      // Kp: clusters of the prediction posterior (== K after mcb_posterior_mu_k)
  mcb::DevBuf<double> pgrid, pKinv, pcov, pwk, pmean, plogdet;
  std::vector<double> h_pi_pred;   // pi_k = mean_i tau_ik of the frozen tau (MC.R:141)
  mcb::DevBuf<int> pmapidx;        // training row -> index in the prediction grid

  // ---- pinned staging + timing ----
  double* hbuf = nullptr;          // pinned, kHbuf doubles
  static constexpr size_t kHbuf = 1 << 16;
  std::vector<cudaEvent_t> evpool;
  struct Span { int phase; int e0, e1; };
  std::vector<Span> spans;
  int ev_next = 0;
  long long launches = 0;
  long long ms_stats[6] = {0, 0, 0, 0, 0, 0};
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;   // mcb_timer_start / mcb_timer_stop
  mcb::DevBuf<unsigned char> flush;            // mcb_flush_l2
  int flush_byte = 0;

  mcb::IndivData idata() const {
    return mcb::IndivData{off.p, woff.p, gidx.p, t.p, y.p, big_ids.p, big_wsoff.p, big_ws.p, nbig, mcb::kSmallMax};
  }
  mcb::IndivWork iwork() const {
    return mcb::IndivWork{Winv.p, pvec.p, yWy.p, logdet_i.p, c1.p, c2.p, Fi.p};
  }
};

namespace mcb {
enum Phase { PH_INDIV = 0, PH_DENSE = 1, PH_TAU = 2, PH_MSTEP_I = 3, PH_MSTEP_K = 4, PH_ELBO = 5, PH_PRED = 6, PH_COMM = 7 };

// NCCL shim (comm.cu): no-ops when the handle is single-rank
int comm_allreduce_sum(mcb_handle* h, double* buf, size_t count);
bool comm_allreduce_small_on(mcb_handle* h, cudaStream_t s, double* buf, int count);   // peer-memory version, any stream
// cluster ownership for large grids: block z lives on rank z % nranks
int comm_reduce_to_owners(mcb_handle* h, double* buf, size_t count, long long stride, int Z);
int comm_bcast_from_owners(mcb_handle* h, double* buf, size_t count, long long stride, int Z);
void comm_destroy(mcb_handle* h);
}  // namespace mcb
