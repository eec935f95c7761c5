// nccl_dyn.h -- NCCL entry points resolved with dlopen at first multi-GPU use, so that the
// single-GPU library has no NCCL dependency and, inside a Python process that already
// imported torch, the very same libnccl.so.2 instance torch uses is picked up.
// Only the types/enums of <nccl.h> are taken from the header; no symbol is linked.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace fvb {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;

  // returns empty string on success, else the reason
  std::string load() {
    if (lib) return "";
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
#define FVB_SYM(field, name)                                         \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name));       \
  if (!field) return std::string("NCCL symbol missing: ") + name;
    FVB_SYM(GetUniqueId, "ncclGetUniqueId")
    FVB_SYM(CommInitRank, "ncclCommInitRank")
    FVB_SYM(CommDestroy, "ncclCommDestroy")
    FVB_SYM(AllReduce, "ncclAllReduce")
    FVB_SYM(Send, "ncclSend")
    FVB_SYM(Recv, "ncclRecv")
    FVB_SYM(GroupStart, "ncclGroupStart")
    FVB_SYM(GroupEnd, "ncclGroupEnd")
    FVB_SYM(GetErrorString, "ncclGetErrorString")
#undef FVB_SYM
    return "";
  }
};

inline NcclApi &nccl() {
  static NcclApi api;
  return api;
}

struct Comm {
  ncclComm_t comm = nullptr;
};

}  // namespace fvb
