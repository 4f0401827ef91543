// trx_comm.h -- the one collective of the path: an all-reduce (sum) of {loss, gradient} over the ranks that share a
// rollout batch (rollouts shard across GPUs, the policy-parameter gradient is summed), issued
// on the objective's stream from inside the library so that the device-resident solver loops run sharded.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the one torch ships is already in a Python host's process; a
// C++ host links or preloads its own), so the engine has no link-time dependency on it and single-GPU users never
// touch it.  Only the five entry points used here are declared; types follow nccl.h (2.x ABI).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

namespace trxcomm {

struct UniqueId {
  char internal[128];  // ncclUniqueId
};
typedef void *Comm;  // ncclComm_t
enum { kNcclSuccess = 0, kNcclFloat64 = 8, kNcclSum = 0 };

struct Api {
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string error;  // non-empty: NCCL is not available
};

inline const Api &api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      a.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
      return;
    }
    auto sym = [&](const char *n) -> void * {
      void *p = dlsym(h, n);
      if (!p) a.error = std::string("NCCL symbol missing: ") + n;
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return a;
}

}  // namespace trxcomm

struct trx_comm {
  trxcomm::Comm comm = nullptr;
  int n_ranks = 1, rank = 0, device = 0;
};
