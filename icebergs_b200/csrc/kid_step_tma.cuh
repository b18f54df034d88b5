// kid_step_tma.cuh -- k_step_fast as a persistent kernel fed by bulk copies (TMA, cp.async.bulk) of the berg columns.
//
// k_step_fast (kid_kernels.cuh) is latency-bound: a warp loads its 17 columns, waits a DRAM round trip, gathers the
// grid records of its cells, waits again, and then runs ~1400 dependent fp64 instructions with nothing in flight; at
// 96 registers only 20 warps per SM are there to cover for each other (profiles/r2_notes.md: 45 % of the stall samples
// are long-scoreboard waits at those two places).  Here the columns of a 128-berg tile travel by cp.async.bulk into
// shared memory, one tile AHEAD of the tile being computed, and complete on an mbarrier:
//   * the DRAM round trip of tile t+1 overlaps the arithmetic of tile t -- without holding 34 registers of loaded
//     values per thread the way a register-level software pipeline would;
//   * the columns are read from shared memory at their point of use (lon and the thermodynamics-only columns late),
//     so they are not live across the momentum solve;
//   * half-way through tile t every thread prefetches the grid records of ITS berg of tile t+1 (cell indices are
//     already in shared memory) into L1, so the gather of the next tile is an L1 hit.
// CTA b takes tiles b, b + G, b + 2G, ... (G = CTAs in the grid): at any time the resident CTAs work on G consecutive
// tiles of the cell-sorted store, i.e. on neighbouring cells (L2 locality of the grid records as before).
// The arithmetic is k_step_fast's, statement for statement: which kernel steps a berg does not change its result.
#pragma once
#include <cstdint>
#include "kid_kernels.cuh"

namespace kid {

// the columns a tile brings in, in the order of their first use
enum StagedCol : int { SC_LAT, SC_UVEL, SC_VVEL, SC_AXN, SC_AYN, SC_BXN, SC_BYN, SC_XI, SC_YJ, SC_MASS, SC_THICK, SC_WIDTH,
                       SC_LENGTH, SC_LON, SC_MSCAL, SC_MBITS, SC_HEAT, SC_NCOL };
#define KID_STAGED_COL_IDS {C_LAT, C_UVEL, C_VVEL, C_AXN, C_AYN, C_BXN, C_BYN, C_XI, C_YJ, C_MASS, C_THICKNESS, C_WIDTH, C_LENGTH, \
                            C_LON, C_MASS_SCALING, C_MASS_OF_BITS, C_HEAT_DENSITY}
#ifndef KID_TMA_PREFETCH
#define KID_TMA_PREFETCH 0      // 0: no record prefetch for the next tile, 1: the cell's own records, 2: + the SSH-slope neighbours
#endif                          // (measured: the prefetches cost more issue slots / L1 tag traffic than the gather latency they hide)
#ifndef KID_TMA_COLS
#define KID_TMA_COLS 13         // 17: every column a berg reads; 13: lon and the thermodynamics-only columns stay in global memory
#endif                          // (prefetched into L1 at the start of the tile), which leaves shared memory for more resident CTAs

struct __align__(128) TileStage {
  double col[KID_TMA_COLS][KID_BLOCK];
  int32_t ine[KID_BLOCK], jne[KID_BLOCK];
  uint8_t flags[KID_BLOCK];
};
constexpr uint32_t kTileBytes = KID_TMA_COLS * KID_BLOCK * 8 + 2 * KID_BLOCK * 4 + KID_BLOCK;
constexpr int kTmaStages = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) {}
}
// non-blocking (try_wait may suspend the thread until the phase completes or a time limit passes)
__device__ __forceinline__ bool mbar_poll(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}

// the grid records interp_flds<LEAN> / pos_within_cell / interp_thermo gather for cell (i,j): into L1
__device__ __forceinline__ void prefetch_cell_records(const DevGrid& g, int i, int j, double xi, double yj) {
  if (KID_TMA_PREFETCH == 0 || !cell_on_pe(g, i, j)) return;
  const int ne = gidx(g, i, j), nid = g.nid;
  prefetch_l1(&g.corner[ne]); prefetch_l1(&g.corner[ne - 1]);
  prefetch_l1(&g.corner[ne - nid]); prefetch_l1(&g.corner[ne - nid - 1]);
  prefetch_l1(&g.cell[ne]);
  prefetch_l1(&g.rect[ne]);
  if (KID_TMA_PREFETCH >= 2) {          // the neighbours interp_flds takes the SSH slopes from (I:4830-4860)
    const int rj = (yj >= 0.5) ? ne + nid : ne - nid, ri = (xi >= 0.5) ? ne + 1 : ne - 1;
    prefetch_l1(&g.cell[ne - 1]); prefetch_l1(&g.cell[rj]); prefetch_l1(&g.cell[rj - 1]);
    prefetch_l1(&g.cell[ne - nid]); prefetch_l1(&g.cell[ri]); prefetch_l1(&g.cell[ri - nid]);
  }
}

template <bool DENSE>
__global__ void __launch_bounds__(KID_BLOCK, KID_FAST_MINBLOCKS)
k_step_tma(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
           DevCounters* __restrict__ cnt, long long n_slots, long long s_base, const __grid_constant__ SlowList slow) {
  constexpr bool LEAN = true;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  TileStage* stage = reinterpret_cast<TileStage*>(tma_smem);
  __shared__ __align__(8) uint64_t full[kTmaStages];
  const int tid = threadIdx.x;
  const long long ntiles = (n_slots - s_base + KID_BLOCK - 1) / KID_BLOCK;
  const long long G = gridDim.x;
  if (tid == 0) {
    for (int q = 0; q < kTmaStages; q++) mbar_init(&full[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // one thread queues the copies of a tile: 17 x 1 KB + 2 x 512 B + 128 B, all completing on the stage's barrier
  auto issue = [&](long long tile, int q) {
    const long long s0 = s_base + tile * KID_BLOCK;
    TileStage& st = stage[q];
    mbar_expect_tx(&full[q], kTileBytes);
    constexpr int ids[SC_NCOL] = KID_STAGED_COL_IDS;
#pragma unroll
    for (int c = 0; c < KID_TMA_COLS; c++) bulk_g2s(st.col[c], b.f64[ids[c]] + s0, KID_BLOCK * 8, &full[q]);
    bulk_g2s(st.ine, b.ine + s0, KID_BLOCK * 4, &full[q]);
    bulk_g2s(st.jne, b.jne + s0, KID_BLOCK * 4, &full[q]);
    bulk_g2s(st.flags, b.flags + s0, KID_BLOCK, &full[q]);
  };
  if (tid == 0) {
    for (int q = 0; q < kTmaStages; q++)
      if ((long long)blockIdx.x + q * G < ntiles) issue((long long)blockIdx.x + q * G, q);
  }
  uint32_t parity = 0;          // bit q: the phase stage q completes next
  int q = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += G, q ^= 1) {
    const TileStage& st = stage[q];
    mbar_wait(&full[q], (parity >> q) & 1u);
    parity ^= (1u << q);
    const long long s = s_base + tile * KID_BLOCK + tid;
    const bool in_range = s < n_slots;
#if KID_TMA_COLS < 17
    prefetch_l1(&b.f64[C_LON][s]); prefetch_l1(&b.f64[C_MASS_SCALING][s]);
    prefetch_l1(&b.f64[C_MASS_OF_BITS][s]); prefetch_l1(&b.f64[C_HEAT_DENSITY][s]);
#define KID_LATE(sc, c) b.f64[c][s]
#else
#define KID_LATE(sc, c) st.col[sc][tid]
#endif
    uint8_t flags = in_range ? st.flags[tid] : (uint8_t)0;
    int i = st.ine[tid], j = st.jne[tid];
    const double lat = st.col[SC_LAT][tid];
    const bool owned = (flags & BF_ALIVE) && !(flags & (BF_HALO | BF_LEAVER));
    // 0 = done here, 1 = whole berg to the slow kernel, 2 = from the position update on
    int defer = 0;
    if (owned && ((flags & BF_STATIC) || !(fabs(lat) <= 89.) || !cell_on_pe(g, i, j))) defer = 1;
    Scatter sc;
    sc.key = -1;
    sc.fx.floating_melt = sc.fx.calving_hflx = sc.fx.berg_melt = sc.fx.bergy_src = sc.fx.bergy_melt = 0.;
    sc.fx.net_heat = 0.;
    bool melted = false;
    const bool have_next = tile + G < ntiles;
    bool next_prefetched = false;
    if (owned && !defer) {
      const int ne = gidx(g, i, j);
      const double dt = p.dt, dt_2 = 0.5 * dt;
      double uvel = st.col[SC_UVEL][tid], vvel = st.col[SC_VVEL][tid];
      double axn = st.col[SC_AXN][tid], ayn = st.col[SC_AYN][tid], bxn = st.col[SC_BXN][tid], byn = st.col[SC_BYN][tid];
      double xi = st.col[SC_XI][tid], yj = st.col[SC_YJ][tid];
      const double M = st.col[SC_MASS][tid], T = st.col[SC_THICK][tid], W = st.col[SC_WIDTH][tid], L = st.col[SC_LENGTH][tid];
      // ---- verlet_stepping I:7203-7328 (same statements as step_berg)
      double sin_lat, cos_lat;
      sincos_halfpi_nofallback(p.pi_180 * lat, &sin_lat, &cos_lat);
      b.f64[C_UVEL_PREV][s] = uvel - dt_2 * bxn;
      b.f64[C_VVEL_PREV][s] = vvel - dt_2 * byn;
      const double uvel3 = uvel + (dt_2 * axn);
      const double vvel3 = vvel + (dt_2 * ayn);
      Env e;
      if (!interp_flds<LEAN>(g, p, i, j, xi, yj, e)) atomicOr(&cnt->error_flags, 64u);
      const double f_cori = p.omega2 * sin_lat;
      double ax1, ay1, un_l, vn_l;
      IAcc ia0 = {0., 0., 0., 0., 0., 0., 0., 0.};
      accel_core<false, LEAN>(p, M, T, W, L, f_cori, uvel, vvel, dt, e, 1.0, ia0,
                              [](double, double, IAcc&) {}, ax1, ay1, axn, ayn, bxn, byn, un_l, vn_l);
      uvel = uvel3 + (dt * ax1); vvel = vvel3 + (dt * ay1);      // evolve_icebergs I:7157-7162
      b.f64[C_AXN][s] = axn; b.f64[C_AYN][s] = ayn; b.f64[C_BXN][s] = bxn; b.f64[C_BYN][s] = byn;
      b.f64[C_UVEL][s] = uvel; b.f64[C_VVEL][s] = vvel;
      // the next tile has landed by now (its copies were queued a whole tile ago): its grid records into L1
      if (have_next && mbar_test(&full[q ^ 1], (parity >> (q ^ 1)) & 1u)) {
        const TileStage& nx = stage[q ^ 1];
        prefetch_cell_records(g, nx.ine[tid], nx.jne[tid], nx.col[SC_XI][tid], nx.col[SC_YJ][tid]);
        next_prefetched = true;
      }
      // ---- update_verlet_position I:7684-7764
      const double uvel2 = uvel + (dt_2 * axn) + (dt_2 * bxn);
      const double vvel2 = vvel + (dt_2 * ayn) + (dt_2 * byn);
      const double dxdl1 = p.r180_pi * rcp_nr(p.Rearth * cos_lat);
      const double u2 = uvel2 * dxdl1, v2 = vvel2 * p.dlat_dy;
      const double lonn = KID_LATE(SC_LON, C_LON) + (dt * u2), latn = lat + (dt * v2);
      // ---- adjust_index_and_ground I:7819: in the cell, or one hop to a wet neighbour; anything else is deferred
      const double lo = KID_EDGE_BAND, hi = 1. - KID_EDGE_BAND;
      const double Lx = p.Lx, Lx_2 = Lx * 0.5;
      const RectCell rc = g.rect[ne];
      double a = 0., bb = 0.;
      bool okpos = rc.ralpha == rc.ralpha;
      {
        double xm = lonn;
        if (Lx > 0.) {                                   // apply_modulo_around_point F:6558, in-range case
          const double yy = KSUB(rc.x1, Lx_2), t = KSUB(lonn, yy);
          okpos = okpos && (t >= 0. && t < Lx);
          xm = KADD(t, yy);
        }
        a = KADD(KMUL(KSUB(xm, rc.x1), rc.ralpha), p.rect_add);
        bb = KADD(KMUL(KSUB(latn, rc.y1), rc.reps), p.rect_add);
      }
      const bool inside = a > lo && a < hi && bb > lo && bb < hi;
      if (okpos && !inside) {
        // strictly outside the cell (not within the edge band, where the reference's sign test decides)
        okpos = (a < -lo || a > 1. + lo || bb < -lo || bb > 1. + lo);
        const double* __restrict__ msk = g.msk;
        if (okpos) {
          if (a < 0.) { okpos = (i > g.isd + 1) && (msk[gidx(g, i - 1, j)] > 0.); i = i - 1; }
          else if (a >= 1.) { okpos = (i < g.ied) && (msk[gidx(g, i + 1, j)] > 0.); i = i + 1; }
        }
        if (okpos) {
          if (bb < 0.) { okpos = (j > g.jsd + 1) && (msk[gidx(g, i, j - 1)] > 0.); j = j - 1; }
          else if (bb >= 1.) { okpos = (j < g.jed) && (msk[gidx(g, i, j + 1)] > 0.); j = j + 1; }
        }
        okpos = okpos && cell_on_pe(g, i, j) && !(i > g.iec || i < g.isc || j > g.jec || j < g.jsc);
        if (okpos) {
          const RectCell r2 = g.rect[gidx(g, i, j)];
          okpos = r2.ralpha == r2.ralpha;
          double xm = lonn;
          if (Lx > 0.) {
            const double yy = KSUB(r2.x1, Lx_2), t = KSUB(lonn, yy);
            okpos = okpos && (t >= 0. && t < Lx);
            xm = KADD(t, yy);
          }
          a = KADD(KMUL(KSUB(xm, r2.x1), r2.ralpha), p.rect_add);
          bb = KADD(KMUL(KSUB(latn, r2.y1), r2.reps), p.rect_add);
          okpos = okpos && (a > lo && a < hi && bb > lo && bb < hi);
        }
      }
      if (!okpos) {
        defer = 2;
      } else {
        xi = a; yj = bb;
        b.f64[C_LON][s] = lonn; b.f64[C_LAT][s] = latn;
        b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
        b.ine[s] = i; b.jne[s] = j;
        // ---- thermodynamics I:2844-3300 at the new position
#ifdef KID_DBG_NOTHERMO
        int outcome = TH_KEEP; sc.fx.floating_melt = M * xi;
        if (false)
#else
        int outcome;
#endif
        outcome = thermo_slot<false, LEAN>(g, b, p, s, flags, i, j, xi, yj, uvel, vvel, M, T, W, L,
                                               KID_LATE(SC_MSCAL, C_MASS_SCALING), KID_LATE(SC_MBITS, C_MASS_OF_BITS), KID_LATE(SC_HEAT, C_HEAT_DENSITY), sc, cnt);
        if (outcome == TH_DELETE) { melted = true; b.flags[s] = 0; }
      }
    }
    if (have_next && !next_prefetched) {        // (deferred / dead slots: they still fetch for their berg of the next tile)
      mbar_wait(&full[q ^ 1], (parity >> (q ^ 1)) & 1u);
      const TileStage& nx = stage[q ^ 1];
      prefetch_cell_records(g, nx.ine[tid], nx.jne[tid], nx.col[SC_XI][tid], nx.col[SC_YJ][tid]);
    }
    if (defer) {
      unsigned long long k = atomicAdd(slow.count, 1ull);
      if ((long long)k < slow.cap) slow.slots[k] = (uint32_t)s | (defer == 2 ? KID_SLOW_SPLIT : 0u);
      else atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_CAPACITY);
    }
#ifndef KID_DBG_NOSCATTER
    scatter_fluxes<false, false, DENSE>(g, sc);
#else
    if (sc.fx.floating_melt == 1.2345e-300) atomicOr(&cnt->error_flags, 1u << 30);
#endif
    if (__any_sync(0xffffffffu, melted)) warp_count_add(&cnt->nbergs_melted, melted);
    double nh = sc.fx.net_heat;
    if (__any_sync(0xffffffffu, nh != 0.)) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) nh += shfl_down_d(nh, d);
      if ((threadIdx.x & 31) == 0 && nh != 0.) atomicAdd(&cnt->net_heat_to_ocean, nh);
    }
    // every thread is done with this stage: it takes the tile after next
    __syncthreads();
    if (tid == 0 && tile + kTmaStages * G < ntiles) issue(tile + kTmaStages * G, q);
  }
}

}  // namespace kid
