// Multi-GPU plumbing: one process per GPU, slab decomposition along kernel axis 0
// (SURVEY.md §8e).  NCCL is loaded at run time (dlopen of the libnccl.so.2 that torch already
// has in the process) so the library itself has no link-time NCCL dependency.
//
// Per CG iteration:  all-reduce {d.Ad}  ->  phase B  ->  send/recv of r's first/last owned
// plane into the neighbours' ghost planes (grouped)  +  all-reduce {r.r, |dx|^2 interior,
// |dx|^2 shell}.  d's ghost planes are never exchanged: phase A recomputes d_new = r + beta*d
// on them from the exchanged r ghosts, bit-identically to the owner.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace pa {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool ok = false;
};

static inline NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.error = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?");
    return api;
  }
#define PA_SYM(field, name)                                                       \
  *(void**)(&api.field) = dlsym(api.handle, name);                                \
  if (!api.field) {                                                               \
    api.error = std::string("missing NCCL symbol ") + name;                       \
    return api;                                                                   \
  }
  PA_SYM(GetUniqueId, "ncclGetUniqueId")
  PA_SYM(CommInitRank, "ncclCommInitRank")
  PA_SYM(CommDestroy, "ncclCommDestroy")
  PA_SYM(AllReduce, "ncclAllReduce")
  PA_SYM(Send, "ncclSend")
  PA_SYM(Recv, "ncclRecv")
  PA_SYM(GroupStart, "ncclGroupStart")
  PA_SYM(GroupEnd, "ncclGroupEnd")
  PA_SYM(GetErrorString, "ncclGetErrorString")
#undef PA_SYM
  api.ok = true;
  return api;
}

struct Dist {
  ncclComm_t comm;
  int rank, nranks;
  int ring = 0;  // axis 0 is periodic: rank 0 and rank P-1 are neighbours (wrap-around ghost planes)
};

// First NCCL failure of the current solve (the collectives are enqueued from deep inside the
// iteration builders; the drivers check this once per solve and report PA_ERR_NCCL).
static inline ncclResult_t& nccl_first_error() {
  static thread_local ncclResult_t rc = ncclSuccess;
  return rc;
}
static inline ncclResult_t nccl_note(ncclResult_t rc) {
  if (rc != ncclSuccess && nccl_first_error() == ncclSuccess) nccl_first_error() = rc;
  return rc;
}

// sum-all-reduce `count` doubles in place (device memory)
static inline ncclResult_t dist_allreduce(const Dist& d, double* buf, int count, cudaStream_t s) {
  return nccl_note(nccl_api().AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, d.comm, s));
}

// exchange the boundary planes of a slab-decomposed vector: first/last OWNED plane -> the
// neighbour's ghost plane.  plane_elems = n1*n2.  Planes are contiguous runs, no packing.
// Order inside the group: [first plane -> lower neighbour, upper ghost <- upper neighbour], then
// [last plane -> upper neighbour, lower ghost <- lower neighbour]; NCCL matches the operations
// between two ranks in issue order, which keeps the pairing right when both neighbours are the
// same rank (ring of two).
template <typename T>
static inline ncclResult_t dist_halo_exchange(const Dist& d, T* v, long long plane_elems, int olo0,
                                              int ohi0, cudaStream_t s) {
  NcclApi& a = nccl_api();
  const ncclDataType_t dt = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;
  const int lower = d.rank > 0 ? d.rank - 1 : (d.ring ? d.nranks - 1 : -1);
  const int upper = d.rank < d.nranks - 1 ? d.rank + 1 : (d.ring ? 0 : -1);
  ncclResult_t rc = nccl_note(a.GroupStart());
  if (rc != ncclSuccess) return rc;
  if (lower >= 0) nccl_note(a.Send(v + (long long)olo0 * plane_elems, (size_t)plane_elems, dt, lower, d.comm, s));
  if (upper >= 0) nccl_note(a.Recv(v + (long long)ohi0 * plane_elems, (size_t)plane_elems, dt, upper, d.comm, s));
  if (upper >= 0) nccl_note(a.Send(v + (long long)(ohi0 - 1) * plane_elems, (size_t)plane_elems, dt, upper, d.comm, s));
  if (lower >= 0) nccl_note(a.Recv(v + (long long)(olo0 - 1) * plane_elems, (size_t)plane_elems, dt, lower, d.comm, s));
  return nccl_note(a.GroupEnd());
}

// point-to-point helpers of the slab-periodic boundary condition (api.cu launch_bcs)
template <typename T>
static inline ncclResult_t dist_send(const Dist& d, const T* p, long long n, int peer, cudaStream_t s) {
  return nccl_note(nccl_api().Send(p, (size_t)n, sizeof(T) == 8 ? ncclFloat64 : ncclFloat32, peer, d.comm, s));
}
template <typename T>
static inline ncclResult_t dist_recv(const Dist& d, T* p, long long n, int peer, cudaStream_t s) {
  return nccl_note(nccl_api().Recv(p, (size_t)n, sizeof(T) == 8 ? ncclFloat64 : ncclFloat32, peer, d.comm, s));
}

}  // namespace pa
