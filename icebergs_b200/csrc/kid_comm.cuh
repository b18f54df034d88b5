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

namespace kid {

// layout of the ranks over the global grid (mpp_define_layout / mpp_compute_extent restated in
// kid_define_domain): rank = px + lx*py owns columns xs[px]..xs[px+1]-1 and rows ys[py]..ys[py+1]-1
#define KID_MAX_DIV 64
struct DevLayout {
  int32_t lx, ly, gni, gnj, cyclic_x, cyclic_y, rank, nranks;
  int32_t xs[KID_MAX_DIV + 1], ys[KID_MAX_DIV + 1];
};

// owner of global cell (i,j); -1 = outside the model (NULL_PE).  i may be one period off.
__device__ __forceinline__ int owner_rank(const DevLayout& L, int i, int j) {
  if (i < 1 || i > L.gni) { if (!L.cyclic_x) return -1; i = ((i - 1) % L.gni + L.gni) % L.gni + 1; }
  if (j < 1 || j > L.gnj) { if (!L.cyclic_y) return -1; j = ((j - 1) % L.gnj + L.gnj) % L.gnj + 1; }
  int px = 0, py = 0;
  while (px + 1 < L.lx && i >= L.xs[px + 1]) px++;
  while (py + 1 < L.ly && j >= L.ys[py + 1]) py++;
  return px + L.lx * py;
}

// pass 1: how many bergs leave for each rank (send_bergs_to_other_pes F:3024-3050, all directions at once)
__global__ void k_count_leavers(const __grid_constant__ DevLayout L, const uint8_t* __restrict__ flags,
                                const int32_t* __restrict__ ine, const int32_t* __restrict__ jne, long long n_slots,
                                int32_t* __restrict__ counts /* [nranks] */) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  if (!(flags[s] & BF_LEAVER)) return;
  int d = owner_rank(L, ine[s], jne[s]);
  if (d >= 0 && d != L.rank) atomicAdd(&counts[d], 1);
}

// pass 2: pack_berg_into_buffer2 (F:3250) into the per-destination regions; the slot is freed
__global__ void k_pack_leavers(const __grid_constant__ DevLayout L, const __grid_constant__ DevBergs b,
                               long long n_slots, const int32_t* __restrict__ offsets /* [nranks] */,
                               int32_t* __restrict__ cursor /* [nranks], zeroed */, double* __restrict__ sendbuf) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t f = b.flags[s];
  if (!(f & BF_LEAVER)) return;
  int d = owner_rank(L, b.ine[s], b.jne[s]);
  b.flags[s] = 0;
  if (d < 0 || d == L.rank) return;          // left the model through an open boundary
  int pos = offsets[d] + atomicAdd(&cursor[d], 1);
  double* rec = sendbuf + (size_t)pos * PACK_W;
#pragma unroll
  for (int c = 0; c < C_NBASE; c++) rec[PK_F64_0 + c] = b.f64[c][s];
  rec[PK_ID] = __longlong_as_double(b.id[s]);
  rec[PK_INE_JNE] = __longlong_as_double(((long long)(unsigned)b.ine[s] << 32) | (unsigned)b.jne[s]);
  rec[PK_YEAR_FLAGS] = __longlong_as_double(((long long)(unsigned)b.start_year[s] << 32) | (unsigned)(f & ~BF_LEAVER));
}

}  // namespace kid
