// gbp_comm.cuh — transport of the shard group: point-to-point transfers between shards.
//
// The ONLY inter-GPU communication of the engine is neighbour send/recv:
//   * nccl  — one process per GPU; ncclSend/ncclRecv grouped per exchange on the
//             shard's own stream (NVLink 5 / NVSwitch on a B200 box).  libnccl is
//             dlopen'ed on first use so that a single-GPU world needs no NCCL.
//   * local — every shard of the group lives in this process on ONE device and shares
//             one stream; a transfer is a device-to-device copy.  Same sharding code
//             path, used by the 1-GPU parity tests of the multi-GPU logic.
// An exchange is described by per-shard lists of sends and receives; the k-th send
// from shard a to shard b pairs with the k-th receive of b from a (NCCL semantics).
#pragma once
#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace gbp {

struct Xfer {
  int peer;
  void *ptr;
  size_t bytes;
};
struct XferPlan {
  std::vector<Xfer> sends, recvs;
  void clear() {
    sends.clear();
    recvs.clear();
  }
};

// Minimal NCCL surface (types as in nccl.h 2.x; the ABI of these entry points is stable).
struct NcclApi {
  typedef struct ncclComm *comm_t;
  struct unique_id {
    char internal[128];
  };
  int (*GetUniqueId)(unique_id *) = nullptr;
  int (*CommInitRank)(comm_t *, int, unique_id, int) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
  void *handle = nullptr;
  std::string error;

  bool load() {
    if (handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
      return false;
    }
    auto sym = [&](const char *n) {
      void *p = dlsym(handle, n);
      if (!p) error = std::string("NCCL symbol missing: ") + n;
      return p;
    };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    Send = reinterpret_cast<decltype(Send)>(sym("ncclSend"));
    Recv = reinterpret_cast<decltype(Recv)>(sym("ncclRecv"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    GetVersion = reinterpret_cast<decltype(GetVersion)>(sym("ncclGetVersion"));
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !Send || !Recv || !GroupStart || !GroupEnd ||
        !GetErrorString) {
      dlclose(handle);
      handle = nullptr;
      return false;
    }
    return true;
  }
};

inline NcclApi &nccl_api() {
  static NcclApi api;
  return api;
}

constexpr int kNcclUint8 = 1;  // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

}  // namespace gbp
