// NCCL communicator of an event-sharded fit (include/brie_b200.h, "multi-GPU").
//
// The only exchange step of the path: with gene features (Kg > 0) or a per-cell intercept the
// per-cell gradients d loss / d Wg[c, :], d b[c], d log sigma[c] are sums over ALL events
// (brie/models/model_TFProb.py:85, 124-125: Wg is shared by the events; model_wrap.py:241-260:
// events are the shard axis), so every optimisation step all-reduces the (M, Nc, Kg + 2) buffer
// G between the fused step kernel and the per-cell Adam update.  The all-reduce is enqueued on
// the fit's own stream from inside brie_fit_run_steps -- no host round trip per step.
//
// NCCL is bound at run time with dlopen (first the copy already loaded into the process, i.e.
// torch's libnccl.so.2, so there is exactly one NCCL in the address space), which keeps
// libbrie_b200.so loadable on machines without NCCL; the entry points below then return
// BRIE_ERR_UNSUPPORTED.
#include <dlfcn.h>
#include <string.h>

#include <new>

#include "brie_host.h"

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId is 128 opaque bytes)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;     // ncclSuccess = 0
enum { kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {            // the copy torch has loaded, if any
    a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
    if (a.handle) break;
  }
  for (const char* n : names) {
    if (a.handle) break;
    a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);   // NB a process that imports torch must do so BEFORE this point:
                                                   // the loader resolves torch's libnccl.so.2 by soname to whatever is loaded
  }
  if (!a.handle) return a;
  *(void**)&a.GetUniqueId = dlsym(a.handle, "ncclGetUniqueId");
  *(void**)&a.CommInitRank = dlsym(a.handle, "ncclCommInitRank");
  *(void**)&a.CommDestroy = dlsym(a.handle, "ncclCommDestroy");
  *(void**)&a.AllReduce = dlsym(a.handle, "ncclAllReduce");
  *(void**)&a.GetErrorString = dlsym(a.handle, "ncclGetErrorString");
  *(void**)&a.GetVersion = dlsym(a.handle, "ncclGetVersion");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
  return a;
}

}  // namespace

struct brie_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  int64_t n_allreduce = 0;
};

#define BRIE_NCCL(call)                                                                             \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != 0)                                                                                    \
      return brie::fail(BRIE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, api().GetErrorString(r_),    \
                        __FILE__, __LINE__);                                                        \
  } while (0)

namespace brie {

// used by brie_fit_step_phase / brie_fit_run_steps (brie_abi.cu)
int comm_allreduce_sum_f32(brie_comm* c, float* buf, int64_t n, cudaStream_t s) {
  if (!c || !c->comm) return fail(BRIE_ERR_ARG, "no communicator");
  if (n <= 0) return BRIE_OK;
  BRIE_NCCL(api().AllReduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, c->comm, s));
  c->n_allreduce += 1;
  return BRIE_OK;
}

}  // namespace brie

extern "C" {

int brie_comm_nccl_version(void) {
  if (!api().ok || !api().GetVersion) return 0;
  int v = 0;
  return api().GetVersion(&v) == 0 ? v : 0;
}

int brie_comm_unique_id(void* id_host) {
  if (!id_host) return brie::fail(BRIE_ERR_ARG, "null argument");
  if (!api().ok) return brie::fail(BRIE_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) is not available: %s", dlerror());
  ncclUniqueId id;
  BRIE_NCCL(api().GetUniqueId(&id));
  memcpy(id_host, &id, sizeof id);
  return BRIE_OK;
}

int brie_comm_create(const void* id_host, int32_t rank, int32_t world, brie_comm** out) {
  if (!id_host || !out) return brie::fail(BRIE_ERR_ARG, "null argument");
  if (world < 1 || rank < 0 || rank >= world) return brie::fail(BRIE_ERR_ARG, "rank %d / world %d out of range", rank, world);
  if (!api().ok) return brie::fail(BRIE_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) is not available");
  brie_comm* c = new (std::nothrow) brie_comm();
  if (!c) return brie::fail(BRIE_ERR_ARG, "out of host memory");
  ncclUniqueId id;
  memcpy(&id, id_host, sizeof id);
  ncclResult_t r = api().CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    delete c;
    return brie::fail(BRIE_ERR_CUDA, "ncclCommInitRank failed: %s", api().GetErrorString(r));
  }
  c->rank = rank;
  c->world = world;
  *out = c;
  return BRIE_OK;
}

int brie_comm_destroy(brie_comm* c) {
  if (c && c->comm && api().ok) api().CommDestroy(c->comm);
  delete c;
  return BRIE_OK;
}

int brie_comm_allreduce_f32(brie_comm* c, float* buf, int64_t n, void* stream) {
  return brie::comm_allreduce_sum_f32(c, buf, n, (cudaStream_t)stream);
}

int brie_comm_allreduce_f64(brie_comm* c, double* buf, int64_t n, void* stream) {
  if (!c || !c->comm) return brie::fail(BRIE_ERR_ARG, "no communicator");
  if (n <= 0) return BRIE_OK;
  BRIE_NCCL(api().AllReduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, c->comm, (cudaStream_t)stream));
  c->n_allreduce += 1;
  return BRIE_OK;
}

int64_t brie_comm_allreduce_count(const brie_comm* c) { return c ? c->n_allreduce : -1; }

}  // extern "C"
