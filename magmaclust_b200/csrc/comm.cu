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

#include "handle.h"

namespace mcb {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
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

int comm_allreduce_sum(mcb_handle* h, double* buf, size_t count) {
  if (h->nranks <= 1 || count == 0) return MCB_OK;
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
  return MCB_OK;
}
