// NCCL plumbing for the sharded path: individuals are partitioned over ranks (one process per GPU), the
// per-cluster grid accumulators A_k, w_k (update_inv_VEM / update_mean_VEM, CF.R:690-731) and the scalar
// sums (common-hp objectives, ELBO, pi_k, pen_diag) are all-reduced over NVLink.
//
// NCCL is bound with dlopen at mcb_comm_init time so that the single-GPU path has no NCCL dependency and
// the library loads on hosts without it. When the process has already loaded a libnccl.so.2 (e.g. through
// torch), the dynamic loader hands back that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>
#include <vector>

#include "handle.h"

namespace mcb {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static bool nccl_load(std::string* why) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    if (why) *why = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
    return false;
  }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
  g_nccl.Reduce = (decltype(g_nccl.Reduce))dlsym(lib, "ncclReduce");
  g_nccl.Broadcast = (decltype(g_nccl.Broadcast))dlsym(lib, "ncclBroadcast");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(lib, "ncclAllGather");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(lib, "ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(lib, "ncclGroupEnd");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.Reduce ||
      !g_nccl.Broadcast || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
    if (why) *why = "libnccl is missing required symbols";
    return false;
  }
  g_nccl.lib = lib;
  return true;
}

// ---- scalar all-reduce over NVLink peer memory ---------------------------------------------------------------------------
// The optimiser loops of the sharded path all-reduce 1 - 5 doubles per objective evaluation (common-hp objective + gradient,
// ELBO, pen_diag, pi_k). Through NCCL that is a kernel of its own with ~20 us of latency, 25 - 45 times per M-step: at
// M = 100,000 individuals on 8 GPUs a quarter of the step. Here every rank owns a small mailbox in device memory, mapped into
// all peers with CUDA IPC: one warp writes this rank's values straight into the peers' mailboxes (P2P stores over NVLink /
// NVSwitch), raises its flag there, waits for the peers' flags in its own mailbox and adds the nranks contributions in RANK
// ORDER, so every rank gets the bit-identical sum. Two slots by epoch parity: a rank can run at most one epoch ahead of the
// slowest (it needs that rank's contribution to finish the next one), so a slot is never overwritten before it was read.
constexpr int kMboxVals = 8;                       // doubles per contribution
constexpr int kMboxMaxRanks = 32;
// mailbox layout: vals[2][kMboxMaxRanks][kMboxVals] doubles, then flags[2][kMboxMaxRanks] unsigned
constexpr size_t kMboxDoubles = 2 * kMboxMaxRanks * kMboxVals + 2 * kMboxMaxRanks;

__global__ void tiny_allreduce_kernel(double* buf, int count, double* const* peers, int rank, int nranks,
                                      unsigned* epoch_ctr, int* err) {
  __shared__ double got[kMboxMaxRanks][kMboxVals];
  __shared__ unsigned epoch_s;
  const int r = threadIdx.x;
  // the call counter lives on the device (every rank makes the same sequence of calls, also from inside graph loops)
  if (r == 0) epoch_s = ++(*epoch_ctr);
  __syncthreads();
  const unsigned epoch = epoch_s;
  const int par = epoch & 1;
  if (r < nranks) {
    double* mb = peers[r];
    volatile double* vals = mb + ((size_t)par * kMboxMaxRanks + rank) * kMboxVals;
    for (int c = 0; c < count; ++c) vals[c] = buf[c];
    __threadfence_system();
    volatile unsigned* flags = reinterpret_cast<volatile unsigned*>(mb + 2 * kMboxMaxRanks * kMboxVals);
    flags[par * kMboxMaxRanks + rank] = epoch;
    // receive the contribution of rank r in the own mailbox
    double* own = peers[rank];
    volatile unsigned* myflags = reinterpret_cast<volatile unsigned*>(own + 2 * kMboxMaxRanks * kMboxVals);
    const long long t0 = clock64();
    bool ok = true;
    while (myflags[par * kMboxMaxRanks + r] != epoch) {
      if (clock64() - t0 > 20000000000LL) { ok = false; break; }   // ~10 s: a peer never arrived; do not hang the GPU
    }
    __threadfence_system();
    volatile double* in = own + ((size_t)par * kMboxMaxRanks + r) * kMboxVals;
    for (int c = 0; c < count; ++c) got[r][c] = ok ? in[c] : nan("");
    if (!ok) *err = 1;
  }
  __syncthreads();
  if (r < count) {
    double s = 0.0;
    for (int q = 0; q < nranks; ++q) s += got[q][r];
    buf[r] = s;
  }
}

static int mbox_setup(mcb_handle* h) {
  static const bool off = getenv("MCB_NO_PEER_ALLREDUCE") != nullptr;
  if (off || !g_nccl.AllGather || h->nranks > kMboxMaxRanks) return MCB_OK;
  const int R = h->nranks;
  cudaIpcMemHandle_t mine;
  cudaIpcMemHandle_t* d_all = nullptr;
  std::vector<cudaIpcMemHandle_t> all(R);
  auto fail = [&](const char*) {   // any failure leaves the NCCL path in charge
    if (d_all) cudaFree(d_all);
    if (h->mbox) { cudaFree(h->mbox); h->mbox = nullptr; }
    cudaGetLastError();
    return MCB_OK;
  };
  if (cudaMalloc(&h->mbox, sizeof(double) * kMboxDoubles) != cudaSuccess) return fail("malloc");
  if (cudaMemset(h->mbox, 0, sizeof(double) * kMboxDoubles) != cudaSuccess) return fail("memset");
  if (cudaIpcGetMemHandle(&mine, h->mbox) != cudaSuccess) return fail("ipc handle");
  if (cudaMalloc(&d_all, sizeof(cudaIpcMemHandle_t) * R) != cudaSuccess) return fail("malloc");
  if (cudaMemcpy(d_all + h->rank, &mine, sizeof mine, cudaMemcpyHostToDevice) != cudaSuccess) return fail("copy");
  if (g_nccl.AllGather(d_all + h->rank, d_all, sizeof mine, ncclChar, (ncclComm_t)h->comm, h->st) != ncclSuccess) return fail("allgather");
  if (cudaStreamSynchronize(h->st) != cudaSuccess) return fail("sync");
  if (cudaMemcpy(all.data(), d_all, sizeof(cudaIpcMemHandle_t) * R, cudaMemcpyDeviceToHost) != cudaSuccess) return fail("copy");
  cudaFree(d_all);
  d_all = nullptr;
  std::vector<double*> ptrs(R, nullptr);
  bool ok = true;
  for (int r = 0; r < R && ok; ++r) {
    if (r == h->rank) { ptrs[r] = h->mbox; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
    h->mbox_opened.push_back(p);
    ptrs[r] = static_cast<double*>(p);
  }
  // every rank must take the same decision: one more (NCCL) all-reduce of the outcome
  double* d_ok = nullptr;
  double v = ok ? 0.0 : 1.0;
  if (cudaMalloc(&d_ok, sizeof(double)) == cudaSuccess) {
    cudaMemcpy(d_ok, &v, sizeof v, cudaMemcpyHostToDevice);
    g_nccl.AllReduce(d_ok, d_ok, 1, ncclDouble, ncclSum, (ncclComm_t)h->comm, h->st);
    cudaStreamSynchronize(h->st);
    cudaMemcpy(&v, d_ok, sizeof v, cudaMemcpyDeviceToHost);
    cudaFree(d_ok);
  } else {
    v = 1.0;
  }
  if (v != 0.0 || cudaMalloc(&h->mbox_peers, sizeof(double*) * R) != cudaSuccess ||
      cudaMalloc(&h->mbox_err, sizeof(int)) != cudaSuccess || cudaMalloc(&h->mbox_epoch_dev, sizeof(unsigned)) != cudaSuccess) {
    for (void* p : h->mbox_opened) cudaIpcCloseMemHandle(p);
    h->mbox_opened.clear();
    if (h->mbox_peers) { cudaFree(h->mbox_peers); h->mbox_peers = nullptr; }
    return fail("peer access");
  }
  cudaMemcpy(h->mbox_peers, ptrs.data(), sizeof(double*) * R, cudaMemcpyHostToDevice);
  cudaMemset(h->mbox_err, 0, sizeof(int));
  cudaMemset(h->mbox_epoch_dev, 0, sizeof(unsigned));
  return MCB_OK;
}

// the peer-memory all-reduce on a given stream (a capture stream of a graph body); false when it is not available
bool comm_allreduce_small_on(mcb_handle* h, cudaStream_t s, double* buf, int count) {
  if (h->nranks <= 1) return true;
  if (!h->mbox_peers || count > kMboxVals) return false;
  tiny_allreduce_kernel<<<1, kMboxMaxRanks, 0, s>>>(buf, count, h->mbox_peers, h->rank, h->nranks, h->mbox_epoch_dev, h->mbox_err);
  return cudaGetLastError() == cudaSuccess;
}

int comm_allreduce_sum(mcb_handle* h, double* buf, size_t count) {
  if (h->nranks <= 1 || count == 0) return MCB_OK;
  if (h->mbox_peers && count <= (size_t)kMboxVals) {
    tiny_allreduce_kernel<<<1, kMboxMaxRanks, 0, h->st>>>(buf, (int)count, h->mbox_peers, h->rank, h->nranks,
                                                          h->mbox_epoch_dev, h->mbox_err);
    if (cudaGetLastError() != cudaSuccess) { h->err = "tiny_allreduce_kernel: launch failed"; return MCB_ECUDA; }
    return MCB_OK;
  }
  ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)h->comm, h->st);
  if (r != ncclSuccess) {
    h->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return MCB_ENCCL;
  }
  return MCB_OK;
}

// Z blocks of `count` doubles at buf + z * stride: block z is summed onto rank owner(z) = z % nranks (in place)
int comm_reduce_to_owners(mcb_handle* h, double* buf, size_t count, long long stride, int Z) {
  if (h->nranks <= 1) return MCB_OK;
  ncclResult_t r = g_nccl.GroupStart();
  for (int z = 0; z < Z && r == ncclSuccess; ++z)
    r = g_nccl.Reduce(buf + z * stride, buf + z * stride, count, ncclDouble, ncclSum, z % h->nranks, (ncclComm_t)h->comm, h->st);
  if (r == ncclSuccess) r = g_nccl.GroupEnd();
  if (r != ncclSuccess) {
    h->err = std::string("ncclReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return MCB_ENCCL;
  }
  return MCB_OK;
}

// block z is broadcast from rank owner(z) = z % nranks to every rank (in place)
int comm_bcast_from_owners(mcb_handle* h, double* buf, size_t count, long long stride, int Z) {
  if (h->nranks <= 1) return MCB_OK;
  ncclResult_t r = g_nccl.GroupStart();
  for (int z = 0; z < Z && r == ncclSuccess; ++z)
    r = g_nccl.Broadcast(buf + z * stride, buf + z * stride, count, ncclDouble, z % h->nranks, (ncclComm_t)h->comm, h->st);
  if (r == ncclSuccess) r = g_nccl.GroupEnd();
  if (r != ncclSuccess) {
    h->err = std::string("ncclBroadcast: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return MCB_ENCCL;
  }
  return MCB_OK;
}

void comm_destroy(mcb_handle* h) {
  for (void* p : h->mbox_opened) cudaIpcCloseMemHandle(p);
  h->mbox_opened.clear();
  if (h->mbox_peers) cudaFree(h->mbox_peers);
  if (h->mbox_err) cudaFree(h->mbox_err);
  if (h->mbox_epoch_dev) cudaFree(h->mbox_epoch_dev);
  h->mbox_epoch_dev = nullptr;
  if (h->mbox) cudaFree(h->mbox);
  h->mbox_peers = nullptr; h->mbox_err = nullptr; h->mbox = nullptr;
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->comm);
  h->comm = nullptr;
  h->nranks = 1;
  h->rank = 0;
}

}  // namespace mcb

extern "C" int mcb_comm_unique_id(unsigned char out_id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::string why;
  if (!mcb::nccl_load(&why)) return MCB_ENCCL;
  ncclUniqueId id;
  if (mcb::g_nccl.GetUniqueId(&id) != ncclSuccess) return MCB_ENCCL;
  memcpy(out_id, &id, 128);
  return MCB_OK;
}

extern "C" int mcb_comm_init(mcb_handle* h, int nranks, int rank, const unsigned char id[128]) {
  if (!h || nranks < 1 || rank < 0 || rank >= nranks || !id) return MCB_EINVAL;
  if (nranks == 1) return MCB_OK;
  std::string why;
  if (!mcb::nccl_load(&why)) {
    h->err = why;
    return MCB_ENCCL;
  }
  MCB_CUDA(h, cudaSetDevice(h->device));
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclComm_t c;
  ncclResult_t r = mcb::g_nccl.CommInitRank(&c, nranks, uid, rank);
  if (r != ncclSuccess) {
    h->err = std::string("ncclCommInitRank: ") + (mcb::g_nccl.GetErrorString ? mcb::g_nccl.GetErrorString(r) : "error");
    return MCB_ENCCL;
  }
  h->comm = c;
  h->nranks = nranks;
  h->rank = rank;
  // a shared-theta_i loop graph captured while the handle was single-rank has no all-reduce in its body
  if (h->iopt_exec) { cudaGraphExecDestroy(h->iopt_exec); h->iopt_exec = nullptr; }
  h->iopt_failed = false;
  return mcb::mbox_setup(h);
}
