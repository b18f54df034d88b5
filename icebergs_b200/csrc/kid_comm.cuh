// kid_comm.cuh -- NCCL, loaded at run time, and the device-side pack/unpack of migrating
// bergs (send_bergs_to_other_pes F:2997, pack_berg_into_buffer2 F:3250,
// unpack_berg_from_buffer2 F:3468).
#pragma once
#include <dlfcn.h>

#include "kid_device.cuh"

namespace kid {

// the slice of the NCCL API this library uses (types per nccl.h 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclInt32_ = 2, ncclInt64_ = 4, ncclFloat64_ = 8 };

struct NcclApi {
  void* dl = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok() const { return dl && GetUniqueId && CommInitRank && Send && Recv && GroupStart && GroupEnd && AllGather; }
};

inline NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // prefer the NCCL already in the process (torch's), then the system one
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.dl = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.dl) break; }
    if (!api.dl) for (const char* n : names) { api.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.dl) break; }
    if (api.dl) {
      api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(api.dl, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.dl, "ncclCommInitRank");
      api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.dl, "ncclCommDestroy");
      api.GroupStart = (int (*)())dlsym(api.dl, "ncclGroupStart");
      api.GroupEnd = (int (*)())dlsym(api.dl, "ncclGroupEnd");
      api.Send = (int (*)(const void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.dl, "ncclSend");
      api.Recv = (int (*)(void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.dl, "ncclRecv");
      api.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(api.dl, "ncclAllGather");
      api.GetErrorString = (const char* (*)(int))dlsym(api.dl, "ncclGetErrorString");
    }
  }
  return api;
}

// ---- exchange record: one migrating berg = PACK_W fp64 words.  Integers travel as
// their own bit patterns (the reference round-trips them through float()/nint(),
// F:3401/F:3425, exact below 2^53; the id is split in two halves F:3299-3301 -- here
// the 64-bit id is carried whole).
enum PackSlot : int {
  PK_F64_0 = 0,                       // C_LON .. C_FL_K in BergCol order (C_NBASE words)
  PK_ID = C_NBASE, PK_INE_JNE, PK_YEAR_FLAGS,
  PACK_W
};

}  // namespace kid
