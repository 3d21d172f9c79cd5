// dcs_nccl.h — NCCL reached through dlopen, so that (a) a process that already carries a NCCL
// (torch bundles its own libnccl.so.2) shares that copy instead of loading a second one, and
// (b) single-GPU use has no NCCL dependency at all.  Only the handful of stable entry points the
// PCG / LM exchange steps need.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types / enums only; no symbol of libnccl is linked

namespace dcs {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
#define DCS_NCCL_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(lib, sym)); if (!field) return false
    DCS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    DCS_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    DCS_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    DCS_NCCL_SYM(AllReduce, "ncclAllReduce");
    DCS_NCCL_SYM(AllGather, "ncclAllGather");
    DCS_NCCL_SYM(Send, "ncclSend");
    DCS_NCCL_SYM(Recv, "ncclRecv");
    DCS_NCCL_SYM(GroupStart, "ncclGroupStart");
    DCS_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    DCS_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef DCS_NCCL_SYM
    return true;
  }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

}  // namespace dcs
