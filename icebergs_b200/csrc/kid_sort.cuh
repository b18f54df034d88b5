// kid_sort.cuh -- the cell-binned sort of the berg store.
//
// Replaces the reference's per-cell linked lists and move_berg_between_cells (F:1758-1797, sorted
// insert F:4270-4359): slots are kept ordered by the linear index of the berg's cell, equal cells in
// ascending previous slot (stable), dead slots and leavers dropped.
//
//   k_sort_keys      key = cell index (n2 = dropped) per slot + the per-cell histogram; runs of one cell
//                    inside a warp are aggregated so a cell receives one atomic per run.
//   k_radix_hist / k_radix_scatter   hand-written stable LSD radix sort of (key, slot) pairs, <= 8-bit
//                    digits, 2048 pairs per CTA; ranks inside a CTA come from warp match + per-warp
//                    digit counters, so equal keys keep their order (no per-cell repair pass).
//   k_gather_f64 / k_gather_misc / k_gather_bonds   the permutation applied to up to 8 fp64 columns
//                    (and to id / ine / jne / start_year / flags / halo_code, and to all bond planes) per
//                    launch: each thread reads perm[k] once and moves every column of its berg.
#pragma once
#include "kid_kernels.cuh"

namespace kid {

#define KID_RADIX_THREADS 256
#define KID_RADIX_ITEMS 8
#define KID_RADIX_TILE (KID_RADIX_THREADS * KID_RADIX_ITEMS)     // pairs per CTA
#define KID_RADIX_WARPS (KID_RADIX_THREADS / 32)
#define KID_RADIX_MAXBINS 256

// key of every slot, identity payload, per-cell counts.  A warp's 32 consecutive slots mostly share a
// handful of cells (the store is nearly sorted): one atomic per run of equal keys.
__global__ void __launch_bounds__(256)
k_sort_keys(const __grid_constant__ DevGrid g, const uint8_t* __restrict__ flags, const int32_t* __restrict__ ine,
            const int32_t* __restrict__ jne, long long n_slots, int32_t dead_key, int32_t* __restrict__ keys,
            int32_t* __restrict__ vals, int32_t* __restrict__ cell_count) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t key = dead_key;
  if (s < n_slots) {
    uint8_t f = flags[s];
    if ((f & BF_ALIVE) && !(f & BF_LEAVER)) key = gidx(g, ine[s], jne[s]);
    keys[s] = key;
    vals[s] = (int32_t)s;
  }
  const int lane = threadIdx.x & 31;
  int32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
  bool head = (lane == 0) || (key != prev);
  unsigned heads = __ballot_sync(0xffffffffu, head);
  if (head && key != dead_key) {
    unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
    int end = above ? (lane + 1 + (__ffs(above) - 1)) : 32;
    atomicAdd(&cell_count[key], end - lane);
  }
}

// digit histogram of one tile: ghist[d * ntiles + tile]
__global__ void __launch_bounds__(KID_RADIX_THREADS)
k_radix_hist(const int32_t* __restrict__ keys, long long n, int shift, int nbins, int ntiles,
             int32_t* __restrict__ ghist) {
  __shared__ int32_t hist[KID_RADIX_MAXBINS];
  for (int d = threadIdx.x; d < nbins; d += blockDim.x) hist[d] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long base = (long long)blockIdx.x * KID_RADIX_TILE + warp * (32 * KID_RADIX_ITEMS);
  const int mask = nbins - 1;
#pragma unroll
  for (int r = 0; r < KID_RADIX_ITEMS; r++) {
    long long k = base + r * 32 + lane;
    const bool valid = k < n;                       // the padding of the last tile is not part of the sort
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      int d = (keys[k] >> shift) & mask;
      unsigned peers = __match_any_sync(act, d);
      if (lane == __ffs(peers) - 1) atomicAdd(&hist[d], __popc(peers));
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < nbins; d += blockDim.x) ghist[(long long)d * ntiles + blockIdx.x] = hist[d];
}

// stable scatter of one tile by the current digit; gbase = exclusive scan of ghist
__global__ void __launch_bounds__(KID_RADIX_THREADS)
k_radix_scatter(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, int32_t* __restrict__ keys_out,
                int32_t* __restrict__ vals_out, long long n, int shift, int nbins, int ntiles,
                const int32_t* __restrict__ gbase) {
  __shared__ int32_t cnt[KID_RADIX_WARPS][KID_RADIX_MAXBINS];     // per-warp digit counts, then start positions
  for (int d = threadIdx.x; d < KID_RADIX_WARPS * KID_RADIX_MAXBINS; d += blockDim.x) (&cnt[0][0])[d] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long base = (long long)blockIdx.x * KID_RADIX_TILE + warp * (32 * KID_RADIX_ITEMS);
  const int mask = nbins - 1;
  const unsigned lt = (1u << lane) - 1u;
  int32_t key[KID_RADIX_ITEMS], val[KID_RADIX_ITEMS], rank[KID_RADIX_ITEMS];
#pragma unroll
  for (int r = 0; r < KID_RADIX_ITEMS; r++) {
    long long k = base + r * 32 + lane;
    key[r] = (k < n) ? keys_in[k] : 0;
    val[r] = (k < n) ? vals_in[k] : -1;
  }
#pragma unroll
  for (int r = 0; r < KID_RADIX_ITEMS; r++) {
    const bool valid = base + r * 32 + lane < n;    // the padding of the last tile is not part of the sort
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    rank[r] = 0;
    if (valid) {
      int d = (key[r] >> shift) & mask;
      unsigned peers = __match_any_sync(act, d);
      int leader = __ffs(peers) - 1;
      int32_t old = 0;
      if (lane == leader) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(peers); }
      old = __shfl_sync(peers, old, leader);
      rank[r] = old + __popc(peers & lt);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int d = threadIdx.x; d < nbins; d += blockDim.x) {
    int32_t run = gbase[(long long)d * ntiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < KID_RADIX_WARPS; w++) { int32_t c = cnt[w][d]; cnt[w][d] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < KID_RADIX_ITEMS; r++) {
    long long k = base + r * 32 + lane;
    if (k < n) {
      int d = (key[r] >> shift) & mask;
      int32_t pos = cnt[warp][d] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

// ------------------------------------------------------------------ gathers
#define KID_GATHER_NC 8
struct GatherF64 { const double* src[KID_GATHER_NC]; double* dst[KID_GATHER_NC]; int nc; };

template <int NC>
__global__ void __launch_bounds__(256)
k_gather_f64(const __grid_constant__ GatherF64 a, const int32_t* __restrict__ perm, long long n) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t s = perm[k];
  double v[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) v[c] = a.src[c][s];
#pragma unroll
  for (int c = 0; c < NC; c++) a.dst[c][k] = v[c];
}

struct GatherMisc {
  const int64_t* id_src; int64_t* id_dst;
  const int32_t* i32_src[3]; int32_t* i32_dst[3];       // ine, jne, start_year
  const uint8_t* u8_src[2]; uint8_t* u8_dst[2];         // flags, halo_code
  const int32_t* aux_src[2]; int32_t* aux_dst[2];       // conglom_id, n_bonds (interactive runs; else null)
};
// covers [0, n_old): destination slots at or beyond n_new are marked dead
__global__ void __launch_bounds__(256)
k_gather_misc(const __grid_constant__ GatherMisc a, const int32_t* __restrict__ perm, long long n_new, long long n_old) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_old) return;
  if (k >= n_new) { a.u8_dst[0][k] = 0; a.u8_dst[1][k] = 0; return; }
  const int32_t s = perm[k];
  int64_t id = a.id_src[s];
  int32_t i0 = a.i32_src[0][s], i1 = a.i32_src[1][s], i2 = a.i32_src[2][s];
  uint8_t f = a.u8_src[0][s], hc = a.u8_src[1][s];
  a.id_dst[k] = id;
  a.i32_dst[0][k] = i0; a.i32_dst[1][k] = i1; a.i32_dst[2][k] = i2;
  a.u8_dst[0][k] = f; a.u8_dst[1][k] = hc;
#pragma unroll
  for (int q = 0; q < 2; q++) if (a.aux_src[q]) a.aux_dst[q][k] = a.aux_src[q][s];
}

// every plane of the half-bond arrays (entry k of slot s at [k*capacity + s]); other_slot is re-resolved by
// connect_all_bonds after the sort and is not moved
struct GatherBonds {
  const int64_t* oid_src; int64_t* oid_dst;
  const int32_t* i32_src[3]; int32_t* i32_dst[3];       // other_ine, other_jne, broken (dem, else null)
  const double* f64_src[1 + BD_N]; double* f64_dst[1 + BD_N];   // length, dem history and saved pair forces (dem, else null)
  long long capacity; int max_bonds;
};
__global__ void __launch_bounds__(256)
k_gather_bonds(const __grid_constant__ GatherBonds a, const int32_t* __restrict__ perm, long long n) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t s = perm[k];
  for (int b = 0; b < a.max_bonds; b++) {
    const long long so = (long long)b * a.capacity + s, dk = (long long)b * a.capacity + k;
    a.oid_dst[dk] = a.oid_src[so];
#pragma unroll
    for (int q = 0; q < 3; q++) if (a.i32_src[q]) a.i32_dst[q][dk] = a.i32_src[q][so];
#pragma unroll
    for (int q = 0; q < 1 + BD_N; q++) if (a.f64_src[q]) a.f64_dst[q][dk] = a.f64_src[q][so];
  }
}

}  // namespace kid
