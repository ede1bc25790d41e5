#include "tp_comm.h"

#include <dlfcn.h>

#include <mutex>
#include <string>

#include "host_util.h"

namespace oasr {
namespace {

struct NcclUniqueId {
  char internal[TP_UNIQUE_ID_BYTES];
};
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(void**, int, NcclUniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int /*ncclDataType_t*/, int /*ncclRedOp_t*/, void*, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);

struct NcclApi {
  void* so = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  GetErrorStringFn error_string = nullptr;
  std::string load_error;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    // a host that already uses NCCL (PyTorch) has libnccl.so.2 mapped: the same SONAME resolves to that copy
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.so) break;
    }
    if (!a.so) {
      a.load_error = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?");
      return;
    }
    a.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(a.so, "ncclGetUniqueId"));
    a.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(a.so, "ncclCommInitRank"));
    a.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(a.so, "ncclCommDestroy"));
    a.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(a.so, "ncclAllReduce"));
    a.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(a.so, "ncclGetErrorString"));
    if (!a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_reduce)
      a.load_error = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllReduce";
  });
  return a;
}

int nccl_fail(const char* what, int rc) {
  NcclApi& a = api();
  return fail(OASR_ERR_CUDA, std::string(what) + ": " + (a.error_string ? a.error_string(rc) : "NCCL error") + " (" +
                                 std::to_string(rc) + ")");
}

}  // namespace

int tp_unique_id(void* out128) {
  NcclApi& a = api();
  if (!a.load_error.empty()) return fail(OASR_ERR_STATE, a.load_error);
  NcclUniqueId id;
  const int rc = a.get_unique_id(&id);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(out128, &id, sizeof(id));
  return OASR_OK;
}

int tp_comm_create(void** comm, int rank, int world, const void* id128) {
  NcclApi& a = api();
  if (!a.load_error.empty()) return fail(OASR_ERR_STATE, a.load_error);
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  const int rc = a.comm_init_rank(comm, world, id, rank);
  if (rc != 0) return nccl_fail("ncclCommInitRank", rc);
  return OASR_OK;
}

void tp_comm_destroy(void* comm) {
  if (comm && api().comm_destroy) api().comm_destroy(comm);
}

int tp_allreduce_f32(void* comm, float* buf, size_t count, cudaStream_t stream) {
  const int rc = api().all_reduce(buf, buf, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, stream);
  if (rc != 0) return nccl_fail("ncclAllReduce", rc);
  return OASR_OK;
}

}  // namespace oasr
