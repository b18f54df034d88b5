// kid_kernels.cuh -- the kernels of the B200 KID hot path (single translation unit:
// included by kid_b200.cu).
//
//   k_step            fused evolve_icebergs (I:7081) + send_bergs_to_other_pes wrap
//                     (F:2997, single-rank cyclic-x) + thermodynamics (I:2844): one
//                     thread per berg slot, each column read once and written once,
//                     melt fluxes scattered with warp-aggregated fp64 atomics.
//   k_thermo_range    thermodynamics alone for a slot range (bergs that arrived by
//                     migration after the fused kernel ran).
//   k_scan_* / k_cell_order   scans and the keyed in-cell order used by the cell-binned sort
//                     (kid_sort.cuh) that replaces move_berg_between_cells (F:1758) and the per-cell lists (F:416).
//   k_ingest_* / k_pack_*   forcing ingest (I:5203-5383) and the packed grid records.
//   k_accumulate_calving / k_calve   I:6153 / I:6225.
#pragma once
#include "kid_physics.cuh"

namespace kid {

#ifndef KID_BLOCK
#ifndef KID_BLOCK
#define KID_BLOCK 128
#endif
#endif
#ifndef KID_MINBLOCKS
#define KID_MINBLOCKS 5
#endif

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// Warp-aggregated scatter-add: lanes that hold the same key in a run of consecutive
// lanes are reduced first (segmented suffix sum), the head lane of each run issues
// one atomic.  All 32 lanes must call; key < 0 = nothing to add.
struct SegInfo { int seg_end; int rounds; bool head; };
__device__ __forceinline__ SegInfo seg_info(long long key) {
  int lane = threadIdx.x & 31;
  long long prev = __shfl_up_sync(0xffffffffu, key, 1);
  bool head = (lane == 0) || (key != prev);
  unsigned heads = __ballot_sync(0xffffffffu, head);
  unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
  SegInfo s;
  s.seg_end = above ? (lane + 1 + (__ffs(above) - 1)) : 32;
  s.head = head;
  // longest run in the warp decides how many doubling rounds the suffix sums need
  int len = head ? (s.seg_end - lane) : 0;
  s.rounds = __reduce_max_sync(0xffffffffu, len);
  return s;
}
__device__ __forceinline__ double seg_sum(double v, const SegInfo& s) {
  int lane = threadIdx.x & 31;
  for (int d = 1; d < s.rounds; d <<= 1) {
    double o = shfl_down_d(v, d);
    if (lane + d < s.seg_end) v += o;
  }
  return v;
}
__device__ __forceinline__ void seg_scatter(double* __restrict__ fld, long long key, double v, const SegInfo& s) {
  if (!__any_sync(0xffffffffu, v != 0.)) return;     // e.g. calving_hflx when no berg carries heat
  double t = seg_sum(v, s);
  if (s.head && key >= 0 && t != 0.) atomicAdd(&fld[key], t);
}

__device__ __forceinline__ void warp_count_add(unsigned long long* ctr, bool pred) {
  unsigned m = __ballot_sync(0xffffffffu, pred);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(ctr, (unsigned long long)__popc(m));
}

// send_bergs_to_other_pes (F:2997) seen from one berg.  Returns 0 = stays owned,
// 1 = leaves to another rank, 2 = leaves the model (NULL_PE).  The cyclic-x wrap to
// this same rank re-homes the berg as the receiving side's unpack does
// (check_and_find_cell F:5973 then pos_within_cell, *_old reset F:3574-3577).
__device__ __forceinline__ int route_berg(const DevGrid& g, const DevParams& p, double lon, double lat, int& i,
                                          int& j, double& xi, double& yj, unsigned int* err,
                                          unsigned long long* n_wrapped) {
  if (j > g.jec && g.fold_north) {
    // FOLD_NORTH_EDGE (F:3138-3147): the berg belongs to the cell on the other side of the fold, (gni+1-i, 2 gnj+1-j).
    // On this rank: re-homed here as the receiver's unpack would (F:3629-3641); otherwise it leaves for the owner of
    // that cell (owner_rank and k_pack_leavers fold the index)
    int oi = g.gni + 1 - i, oj = 2 * g.gnj + 1 - j;
    oi = ((oi - 1) % g.gni + g.gni) % g.gni + 1;
    if (oi < g.isc || oi > g.iec || oj < g.jsc) return 1;
    bool found = is_point_in_cell(g, p, lon, lat, oi, oj, err);
    if (!found) found = find_cell_wide(g, p, lon, lat, &oi, &oj, err);
    if (!found) { atomicOr(err, (unsigned)KID_DEVERR_LOST_BERG); return 2; }
    i = oi; j = oj;
    pos_within_cell(g, p, lon, lat, i, j, &xi, &yj, err);
    atomicAdd(n_wrapped, 1ull);
    return 0;
  }
  if (i > g.iec || i < g.isc) {
    bool east = i > g.iec;
    bool self = east ? g.pe_E_self : g.pe_W_self;
    bool has = east ? g.has_E : g.has_W;
    if (self) {
      int oi = i + (east ? -g.gni : g.gni), oj = j;
      bool found = false;
      if (cell_on_pe(g, oi, oj)) found = is_point_in_cell(g, p, lon, lat, oi, oj, err);
      if (!found) found = find_cell_wide(g, p, lon, lat, &oi, &oj, err);
      if (!found) { atomicOr(err, (unsigned)KID_DEVERR_LOST_BERG); return 2; }
      i = oi; j = oj;
      pos_within_cell(g, p, lon, lat, i, j, &xi, &yj, err);
      atomicAdd(n_wrapped, 1ull);
    } else if (has) {
      return 1;
    } else {
      return 2;
    }
  }
  if (j > g.jec || j < g.jsc) {
    bool has = (j > g.jec) ? g.has_N : g.has_S;
    return has ? 1 : 2;
  }
  return 0;
}

// per-berg scatter payload collected by the fused kernel
struct Scatter { long long key; ThermoFlux fx; };

template <bool FOOTLOOSE, bool DIAG, bool dense = true>
__device__ __forceinline__ void scatter_fluxes(const DevGrid& g, const Scatter& sc) {
  // Two aggregation variants; the tuned (lean) kernel instance exists in both and sort_bergs picks one from the
  // population density (kid_t::scatter_dense), every other instance uses the density-robust first one:
  //  - dense (> ~20 bergs per occupied cell: the weak-scaling tiles): runs of one cell are reduced with segmented
  //    suffix sums and the head lane of each run issues ONE atomic per field.  Per-berg reductions to one address
  //    from this warp and its neighbours pile up in one L2 slice (measured 1.21 ms vs 1.04 ms at 125 bergs/cell).
  //  - sparse (~13 bergs per cell, the 10M-berg workload): a warp whose 32 bergs all sit in one cell reduces with a
  //    butterfly; a warp that straddles cells issues one reduction (RED.ADD.F64) per berg and field, cheaper than
  //    the shuffle tree there (0.894 ms vs 0.929 ms).
  SegInfo si; si.seg_end = 32; si.rounds = 0; si.head = false;
  {
    double v[5] = {sc.fx.floating_melt, sc.fx.calving_hflx, sc.fx.berg_melt, sc.fx.bergy_src, sc.fx.bergy_melt};
    double* f[5] = {g.floating_melt, g.calving_hflx, g.berg_melt, g.bergy_src, g.bergy_melt};
    if (dense) {
      si = seg_info(sc.key);
#pragma unroll
      for (int q = 0; q < 5; q++) seg_scatter(f[q], sc.key, v[q], si);
    } else {
      long long k0 = __shfl_sync(0xffffffffu, sc.key, 0);
      bool uniform = __all_sync(0xffffffffu, sc.key == k0);
      if (uniform) {
        if (k0 >= 0) {
#pragma unroll
          for (int q = 0; q < 5; q++) {
            double t = v[q];
            if (!__any_sync(0xffffffffu, t != 0.)) continue;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) t += shfl_down_d(t, d);
            if ((threadIdx.x & 31) == 0 && t != 0.) atomicAdd(&f[q][k0], t);
          }
        }
      } else if (sc.key >= 0) {
#pragma unroll
        for (int q = 0; q < 5; q++)
          if (v[q] != 0.) atomicAdd(&f[q][sc.key], v[q]);
      }
    }
  }
#define KID_SEG(fld, v) seg_scatter(fld, sc.key, v, si)
  if ((FOOTLOOSE || DIAG) && !dense) si = seg_info(sc.key);
  if (FOOTLOOSE) {
    KID_SEG(g.fl_bits_melt, sc.fx.fl_bits_melt);
    KID_SEG(g.fl_bits_src, sc.fx.fl_bits_src);
  }
  if (DIAG) {
    KID_SEG(g.fl_parent_melt, sc.fx.fl_parent_melt);
    KID_SEG(g.fl_child_melt, sc.fx.fl_child_melt);
    KID_SEG(g.melt_buoy, sc.fx.melt_buoy);
    KID_SEG(g.melt_eros, sc.fx.melt_eros);
    KID_SEG(g.melt_conv, sc.fx.melt_conv);
    KID_SEG(g.melt_buoy_fl, sc.fx.melt_buoy_fl);
    KID_SEG(g.melt_eros_fl, sc.fx.melt_eros_fl);
    KID_SEG(g.melt_conv_fl, sc.fx.melt_conv_fl);
  }
#undef KID_SEG
}

// thermodynamics of the berg in slot s located at (i,j,xi,yj) moving with (uvel,vvel);
// loads and stores only the columns thermodynamics touches.  Fills sc; returns the
// TH_* outcome.
// Berg columns are streamed: read once and written once per step (3 GB per launch at 10 M bergs), while the grid records
// the bergs gather (~0.2 GB) are what should stay in the 126 MB L2 from one step to the next.  With KID_STREAM_COLS the
// column traffic of the fast step kernel is marked evict-first (ld/st.global.cs).
#ifdef KID_STREAM_COLS
#define KLD(ptr) __ldcs(ptr)
#define KST(ptr, v) __stcs((ptr), (v))
#else
#define KLD(ptr) (*(ptr))
#define KST(ptr, v) (*(ptr) = (v))
#endif
template <bool FOOTLOOSE, bool LEAN = false>
__device__ __forceinline__ int thermo_slot(const DevGrid& g, const DevBergs& b, const DevParams& p, long long s,
                                           uint8_t flags, int i, int j, double xi, double yj, double uvel,
                                           double vvel, double M, double T, double W, double L, double mass_scaling,
                                           double mass_of_bits, double heat_density, Scatter& sc,
                                           DevCounters* cnt, const EnvThermo* pre = nullptr) {
  const int cidx = gidx(g, i, j);
  EnvThermo e;
  if (pre) e = *pre;                       // the caller interpolated already (k_step_fast's cached cell)
  else interp_thermo<LEAN>(g, p, cidx, xi, yj, e);
  if ((e.uo != e.uo) || (e.vo != e.vo) || (e.ua != e.ua) || (e.va != e.va) || (e.sst != e.sst) || (e.cn != e.cn))
    atomicOr(&cnt->error_flags, 64u);
  if (e.rarea == 0.) { atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_GROUNDED); return TH_KEEP; }
  ThermoState st;
  st.mass = M; st.thickness = T; st.width = W; st.length = L;
  st.mass_scaling = mass_scaling;
  st.mass_of_bits = mass_of_bits;
  st.heat_density = heat_density;
  if (FOOTLOOSE) {
    st.mass_of_fl_bits = b.f64[C_MASS_OF_FL_BITS][s];
    st.mass_of_fl_bergy_bits = b.f64[C_MASS_OF_FL_BERGY_BITS][s];
    st.fl_k = b.f64[C_FL_K][s];
  } else {
    st.mass_of_fl_bits = 0.; st.mass_of_fl_bergy_bits = 0.; st.fl_k = 0.;
  }
  st.start_day = 0.; st.start_year = 0;
  double N_bonds = 0.;
  ShelfIn sh = {0., 0., 0.};
  if (!LEAN && (p.melt_icebergs_as_ice_shelf || p.use_mixed_melting)) {
    sh.lat = b.f64[C_LAT][s]; sh.sss = g.cell[cidx].sss; sh.ocean_depth = g.ocean_depth[cidx];
  }
  if (PF(allow_bergs_to_roll, 1) || (!LEAN && p.use_mixed_melting)) {
    // N_bonds: I:2928-2944 (this%n_bonds = the length of the bond list, assign_n_bonds F:4617)
    for (int k = 0; k < b.max_bonds; k++) if (b.bond_other_id[(long long)k * b.capacity + s] != 0) N_bonds += 1.0;
    if (flags & BF_STATIC) N_bonds = p.hexagonal_icebergs ? 6.0 : 4.0;
  }
  int outcome = thermo_berg<LEAN>(p, e, uvel, vvel, N_bonds, st, sc.fx, sh);
  sc.key = (long long)cidx;
  KST(&b.f64[C_MASS][s], st.mass);
  KST(&b.f64[C_THICKNESS][s], st.thickness);
  KST(&b.f64[C_WIDTH][s], st.width);
  KST(&b.f64[C_LENGTH][s], st.length);
  KST(&b.f64[C_MASS_OF_BITS][s], st.mass_of_bits);
  if (FOOTLOOSE) {
    b.f64[C_MASS_OF_FL_BITS][s] = st.mass_of_fl_bits;
    b.f64[C_MASS_OF_FL_BERGY_BITS][s] = st.mass_of_fl_bergy_bits;
    b.f64[C_FL_K][s] = st.fl_k;
    if (outcome == TH_BECAME_FL) {
      b.f64[C_MASS_SCALING][s] = st.mass_scaling;
      b.f64[C_START_DAY][s] = st.start_day;
      b.start_year[s] = st.start_year;
    }
  }
  return outcome;
}

// ------------------------------------------------------------ fused step
// One thread per berg slot; each warp is independent (no block-level synchronisation).  Memory
// latency is the first-order cost of this kernel (profiles/r1 notes): every column load is issued
// before the first use, the gathers of a berg's grid records depend only on (ine,jne) and are
// issued or prefetched as soon as those arrive, so a warp pays two DRAM/L2 round trips
// (columns, then grid records) instead of one per routine.
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// what one berg brings into the step
// (lon and the thermodynamics-only columns are prefetched into L1 by k_step and loaded at their
// first use: they would otherwise sit in registers through the momentum solve, where the register
// file is the limit on resident warps)
struct BergIn {
  double lat, uvel, vvel, axn, ayn, bxn, byn, xi, yj, M, T, W, L;
  int i, j;
  uint8_t flags;
};

// evolve_icebergs (I:7081) + send_bergs_to_other_pes (F:2997) + thermodynamics (I:2844) for one berg
// SPLIT: the velocity solve already ran (k_ia_velocity, interactions on): this is the second sweep of
// evolve_icebergs I:7182-7197 (position, *_old refresh) followed by send_bergs and thermodynamics
template <bool FOOTLOOSE, bool SPLIT, bool LEAN = false>
__device__ __forceinline__ void step_berg(const DevGrid& g, const DevBergs& b, const DevParams& p,
                                          DevCounters* __restrict__ cnt, long long s, const BergIn& in, Scatter& sc,
                                          bool& melted, bool& became_fl, bool& bounced, bool& speeding, bool& left) {
  const double dt = p.dt, dt_2 = 0.5 * dt;
  uint8_t flags = in.flags;
  int i = in.i, j = in.j;
  double lon = 0., lat = in.lat, uvel = in.uvel, vvel = in.vvel, xi = in.xi, yj = in.yj;
  double M = in.M, T = in.T, W = in.W, L = in.L;
  if (!(flags & BF_STATIC)) {
    double axn = in.axn, ayn = in.ayn, bxn = in.bxn, byn = in.byn;
    double sin_lat = 0., cos_lat = 1.;
    if (PF(grid_is_latlon, 1)) sincos_halfpi(p.pi_180 * lat, &sin_lat, &cos_lat);
    const bool tang = (lat > 89.) && PF(grid_is_latlon, 1);
    if (!SPLIT) {
      // ---- verlet_stepping I:7203-7328
      b.f64[C_UVEL_PREV][s] = uvel - dt_2 * bxn;
      b.f64[C_VVEL_PREV][s] = vvel - dt_2 * byn;
      double uvel3 = uvel + (dt_2 * axn);
      double vvel3 = vvel + (dt_2 * ayn);
      Env e;
      if (!interp_flds<LEAN>(g, p, i, j, xi, yj, e)) atomicOr(&cnt->error_flags, 64u);
      double f_cori = (PF(grid_is_latlon, 1) && !PF(use_f_plane, 0)) ? p.omega2 * sin_lat : p.f_cori_plane;
      double ax1, ay1, un_l, vn_l;
      IAcc ia0 = {0., 0., 0., 0., 0., 0., 0., 0.};
      accel_core<false, LEAN>(p, M, T, W, L, f_cori, uvel, vvel, dt, e, 1.0, ia0,
                        [](double, double, IAcc&) {}, ax1, ay1, axn, ayn, bxn, byn, un_l, vn_l);
      if ((PF(speed_limit, 0.) > 0.) || (PF(speed_limit, 0.) == -1.)) {   // I:2304-2323: only the ticket survives
        double speed = sqrt(un_l * un_l + vn_l * vn_l);
        if (speed > 0.) {
          size_t c = gidx(g, i, j);
          double loc_dx = fmin(0.5 * (g.dx[c] + g.dx[c - g.nid]), 0.5 * (g.dy[c] + g.dy[c - 1]));
          double new_speed = loc_dx / dt * PF(speed_limit, 0.);
          if (new_speed < speed && PF(speed_limit, 0.) > 0.) speeding = true;
        }
      }
      double uveln, vveln;
      lon = b.f64[C_LON][s];
      if (tang) tang_velocity(p, lon, uvel3, vvel3, ax1, ay1, dt, uveln, vveln);
      else { uveln = uvel3 + (dt * ax1); vveln = vvel3 + (dt * ay1); }
      if (PF(override_iceberg_velocities, 0)) { uveln = p.u_override; vveln = p.v_override; }
      uvel = uveln; vvel = vveln;      // evolve_icebergs I:7157-7162
    } else {
      lon = b.f64[C_LON][s];
    }
    if (SPLIT && !LEAN && p.runge_not_verlet) {
      // Runge_Kutta_stepping already placed the berg (k_step_rk<.., STEP_ONLY> / k_step_rk_ia): the second loop of
      // evolve_icebergs only refreshes *_old (I:7178-7197)
      lon = b.f64[C_LON][s];
      if (b.f64[C_UVEL_OLD]) {
        b.f64[C_UVEL_OLD][s] = uvel; b.f64[C_VVEL_OLD][s] = vvel;
        b.f64[C_LON_OLD][s] = lon; b.f64[C_LAT_OLD][s] = lat;
      }
    } else {
    // ---- update_verlet_position I:7684-7764 (uses the NEW velocity and accelerations)
    double uvel2 = uvel + (dt_2 * axn) + (dt_2 * bxn);
    double vvel2 = vvel + (dt_2 * ayn) + (dt_2 * byn);
    if (!SPLIT) {
      b.f64[C_AXN][s] = axn; b.f64[C_AYN][s] = ayn; b.f64[C_BXN][s] = bxn; b.f64[C_BYN][s] = byn;
      b.f64[C_UVEL][s] = uvel; b.f64[C_VVEL][s] = vvel;
    }
    double lonn, latn;
    if (tang) {
      tang_position(p, lon, lat, uvel2, vvel2, dt, lonn, latn);
    } else {
      double dxdl1 = 1., dydl = 1.;
      if (PF(grid_is_latlon, 1)) { dxdl1 = p.r180_pi * rcp_nr(p.Rearth * cos_lat); dydl = p.dlat_dy; }
      double u2 = uvel2 * dxdl1, v2 = vvel2 * dydl;
      lonn = lon + (dt * u2); latn = lat + (dt * v2);
    }
    bounced = adjust_index_and_ground(g, p, lonn, latn, i, j, xi, yj, &cnt->error_flags, &cnt->warn_adjust);
    lon = lonn; lat = latn;
    b.f64[C_LON][s] = lon; b.f64[C_LAT][s] = lat;
    if (SPLIT && b.f64[C_UVEL_OLD]) {      // I:7189-7194 (columns of interactive runs)
      b.f64[C_UVEL_OLD][s] = uvel; b.f64[C_VVEL_OLD][s] = vvel;
      b.f64[C_LON_OLD][s] = lon; b.f64[C_LAT_OLD][s] = lat;
    }
    }
  } else if (i > g.iec || i < g.isc || j > g.jec || j < g.jsc) {
    lon = b.f64[C_LON][s];
  }
  // ---- send_bergs_to_other_pes, F:2997
  int route = 0;
  if (i > g.iec || i < g.isc || j > g.jec || j < g.jsc)
    route = route_berg(g, p, lon, lat, i, j, xi, yj, &cnt->error_flags, &cnt->n_wrapped);
  if (!(flags & BF_STATIC) || route != 0) {
    b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
    b.ine[s] = i; b.jne[s] = j;
  }
  if (route == 1) {
    left = true;
    b.flags[s] = flags | BF_LEAVER;
    unsigned long long k = atomicAdd(b.leaver_count, 1ull);      // rare: a berg in a few thousand per step
    if ((long long)k < b.leaver_cap) b.leaver_list[k] = (int32_t)s;
    else atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_CAPACITY);
  }
  else if (route == 2) { b.flags[s] = 0; }
  else if (!FOOTLOOSE) {      // (with footloose on, thermodynamics follows footloose_calving: I:5455, I:5497)
    // ---- thermodynamics I:2844-3300 at the new position
    int outcome = thermo_slot<FOOTLOOSE, LEAN>(g, b, p, s, flags, i, j, xi, yj, uvel, vvel, M, T, W, L,
                                         b.f64[C_MASS_SCALING][s], b.f64[C_MASS_OF_BITS][s], b.f64[C_HEAT_DENSITY][s],
                                         sc, cnt);
    if (outcome == TH_DELETE) { melted = true; b.flags[s] = 0; }
    else if (outcome == TH_BECAME_FL) { melted = true; became_fl = true; }
  }
}


template <bool FOOTLOOSE, bool DIAG, bool SPLIT = false, bool LEAN = false, bool DENSE = true>
__global__ void __launch_bounds__(KID_BLOCK, KID_MINBLOCKS)
k_step(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
       const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt, long long n_slots, long long s_base) {
  // s_base: first slot of the launch (0, or the tile-aligned start of the bergs that arrived while the main
  // launch of this step was already running, see step_core)
  long long s = s_base + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = s < n_slots;
  const long long sl = in_range ? s : 0;          // out-of-range lanes read slot 0 and are masked by flags
  // all column loads are issued unconditionally, ahead of the flags test (dead slots are rare and
  // only live until the next sort)
  BergIn in;
  in.flags = b.flags[sl];
  in.i = b.ine[sl]; in.j = b.jne[sl];
  in.lat = b.f64[C_LAT][sl];
  in.uvel = b.f64[C_UVEL][sl]; in.vvel = b.f64[C_VVEL][sl];
  in.axn = b.f64[C_AXN][sl]; in.ayn = b.f64[C_AYN][sl]; in.bxn = b.f64[C_BXN][sl]; in.byn = b.f64[C_BYN][sl];
  in.xi = b.f64[C_XI][sl]; in.yj = b.f64[C_YJ][sl];
  in.M = b.f64[C_MASS][sl]; in.T = b.f64[C_THICKNESS][sl]; in.W = b.f64[C_WIDTH][sl]; in.L = b.f64[C_LENGTH][sl];
  prefetch_l1(&b.f64[C_LON][sl]); prefetch_l1(&b.f64[C_MASS_SCALING][sl]);
  prefetch_l1(&b.f64[C_MASS_OF_BITS][sl]); prefetch_l1(&b.f64[C_HEAT_DENSITY][sl]);
  if (!in_range) in.flags = 0;
  bool owned = (in.flags & BF_ALIVE) && !(in.flags & (BF_HALO | BF_LEAVER));   // leavers wait for the (overlapped) exchange
  if (owned && cell_on_pe(g, in.i, in.j)) {
    // corner positions of the berg's cell (pos_within_cell, a few hundred instructions from now)
    int ne = gidx(g, in.i, in.j);
    prefetch_l1(&g.rect[ne]);
  }
  Scatter sc;
  sc.key = -1;
  sc.fx.floating_melt = sc.fx.calving_hflx = sc.fx.berg_melt = sc.fx.bergy_src = sc.fx.bergy_melt = 0.;
  sc.fx.fl_bits_melt = sc.fx.fl_bits_src = sc.fx.net_heat = 0.;
  sc.fx.fl_parent_melt = sc.fx.fl_child_melt = sc.fx.melt_buoy = sc.fx.melt_eros = sc.fx.melt_conv = 0.;
  sc.fx.melt_buoy_fl = sc.fx.melt_eros_fl = sc.fx.melt_conv_fl = 0.;
  bool melted = false, became_fl = false, bounced = false, speeding = false, left = false;
  if (owned) step_berg<FOOTLOOSE, SPLIT, LEAN>(g, b, p, cnt, s, in, sc, melted, became_fl, bounced, speeding, left);
  scatter_fluxes<FOOTLOOSE, DIAG, DENSE>(g, sc);
  // event counters: one vote decides whether the warp has anything to report at all
  if (__any_sync(0xffffffffu, melted | bounced | speeding | left)) {
    warp_count_add(&cnt->nbergs_melted, melted);
    if (FOOTLOOSE) warp_count_add(&cnt->nbergs_calved_fl, became_fl);
    warp_count_add(&cnt->n_bounced, bounced);
    warp_count_add(&cnt->nspeeding, speeding);
    warp_count_add(&cnt->n_leavers, left);
  }
  // net_heat_to_ocean I:3130
  double nh = sc.fx.net_heat;
  if (__any_sync(0xffffffffu, nh != 0.)) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nh += shfl_down_d(nh, d);
    if ((threadIdx.x & 31) == 0 && nh != 0.) atomicAdd(&cnt->net_heat_to_ocean, nh);
  }
}

// ------------------------------------------------- fast path + slow list
// The benchmark configuration (lean_config(): free-drifting Verlet bergs on a regular lat-lon grid with the
// default namelist switches) spends its time on bergs that do the ordinary thing: stay in their cell or hop to
// a wet neighbour cell, stay on the tile, survive the melt.  k_step_fast carries exactly that case in straight-
// line code -- no out-of-line calls, hence no call-clobbered registers and none of the general cell walk, coast
// bounce, tangent-plane, exchange or footloose code in its instruction stream -- and hands every other berg to
// k_step_slow through a device-side list of slots.  A list entry says how far the fast kernel got:
//   SLOW_FULL   nothing done (static berg, |lat| > 89, cell off the tile's data domain): the whole step_berg
//   SLOW_SPLIT  verlet_stepping done and stored (velocity, accelerations); position update, cell walk,
//               send_bergs and thermodynamics to do (the SPLIT instance of step_berg, which reads the stored
//               post-solve values back as update_verlet_position I:7727 uses them)
// Both kernels evaluate the same inlined functions, so which one handles a berg does not change its result.
#define KID_SLOW_SPLIT 0x80000000u
struct SlowList { uint32_t* slots; unsigned long long* count; long long cap; };

#ifndef KID_FAST_MINBLOCKS
#define KID_FAST_MINBLOCKS 5
#endif

template <bool DENSE>
__global__ void __launch_bounds__(KID_BLOCK, KID_FAST_MINBLOCKS)
k_step_fast(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
            DevCounters* __restrict__ cnt, long long n_slots, long long s_base, const __grid_constant__ SlowList slow) {
  constexpr bool LEAN = true;
  long long s = s_base + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = s < n_slots;
  const long long sl = in_range ? s : 0;
  uint8_t flags = KLD(&b.flags[sl]);
  int i = KLD(&b.ine[sl]), j = KLD(&b.jne[sl]);
  double lat = KLD(&b.f64[C_LAT][sl]);
  double uvel = KLD(&b.f64[C_UVEL][sl]), vvel = KLD(&b.f64[C_VVEL][sl]);
  double axn = KLD(&b.f64[C_AXN][sl]), ayn = KLD(&b.f64[C_AYN][sl]), bxn = KLD(&b.f64[C_BXN][sl]), byn = KLD(&b.f64[C_BYN][sl]);
  double xi = KLD(&b.f64[C_XI][sl]), yj = KLD(&b.f64[C_YJ][sl]);
  double M = KLD(&b.f64[C_MASS][sl]), T = KLD(&b.f64[C_THICKNESS][sl]), W = KLD(&b.f64[C_WIDTH][sl]), L = KLD(&b.f64[C_LENGTH][sl]);
  prefetch_l1(&b.f64[C_LON][sl]); prefetch_l1(&b.f64[C_MASS_SCALING][sl]);
  prefetch_l1(&b.f64[C_MASS_OF_BITS][sl]); prefetch_l1(&b.f64[C_HEAT_DENSITY][sl]);
  if (!in_range) flags = 0;
  const bool owned = (flags & BF_ALIVE) && !(flags & (BF_HALO | BF_LEAVER));
  // 0 = done here, 1 = whole berg to the slow kernel, 2 = from the position update on
  int defer = 0;
  if (owned && ((flags & BF_STATIC) || !(fabs(lat) <= 89.) || !cell_on_pe(g, i, j))) defer = 1;
  Scatter sc;
  sc.key = -1;
  sc.fx.floating_melt = sc.fx.calving_hflx = sc.fx.berg_melt = sc.fx.bergy_src = sc.fx.bergy_melt = 0.;
  sc.fx.net_heat = 0.;
  bool melted = false;
  if (owned && !defer) {
    const int ne = gidx(g, i, j);
    prefetch_l1(&g.rect[ne]);
    const double dt = p.dt, dt_2 = 0.5 * dt;
    // ---- verlet_stepping I:7203-7328 (same statements as step_berg)
    double sin_lat, cos_lat;
    sincos_halfpi_nofallback(p.pi_180 * lat, &sin_lat, &cos_lat);
    KST(&b.f64[C_UVEL_PREV][s], uvel - dt_2 * bxn);
    KST(&b.f64[C_VVEL_PREV][s], vvel - dt_2 * byn);
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
    KST(&b.f64[C_AXN][s], axn); KST(&b.f64[C_AYN][s], ayn); KST(&b.f64[C_BXN][s], bxn); KST(&b.f64[C_BYN][s], byn);
    KST(&b.f64[C_UVEL][s], uvel); KST(&b.f64[C_VVEL][s], vvel);
    // ---- update_verlet_position I:7684-7764
    const double uvel2 = uvel + (dt_2 * axn) + (dt_2 * bxn);
    const double vvel2 = vvel + (dt_2 * ayn) + (dt_2 * byn);
    const double dxdl1 = p.r180_pi * rcp_nr(p.Rearth * cos_lat);
    const double u2 = uvel2 * dxdl1, v2 = vvel2 * p.dlat_dy;
    const double lonn = KLD(&b.f64[C_LON][s]) + (dt * u2), latn = lat + (dt * v2);
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
      KST(&b.f64[C_LON][s], lonn); KST(&b.f64[C_LAT][s], latn);
      KST(&b.f64[C_XI][s], xi); KST(&b.f64[C_YJ][s], yj);
      KST(&b.ine[s], i); KST(&b.jne[s], j);
      // ---- thermodynamics I:2844-3300 at the new position
#ifdef KID_FAST_BARRIER
      // (compiler barrier: the corner records interp_flds gathered are gathered again here -- an L1 hit -- instead
      // of being carried in registers, i.e. spilled, across the momentum solve)
      asm volatile("" ::: "memory");
#endif
      int outcome = thermo_slot<false, LEAN>(g, b, p, s, flags, i, j, xi, yj, uvel, vvel, M, T, W, L,
                                             KLD(&b.f64[C_MASS_SCALING][s]), KLD(&b.f64[C_MASS_OF_BITS][s]), KLD(&b.f64[C_HEAT_DENSITY][s]),
                                             sc, cnt);
      if (outcome == TH_DELETE) { melted = true; b.flags[s] = 0; }
    }
  }
  if (defer) {
    unsigned long long k = atomicAdd(slow.count, 1ull);
    if ((long long)k < slow.cap) slow.slots[k] = (uint32_t)s | (defer == 2 ? KID_SLOW_SPLIT : 0u);
    else atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_CAPACITY);
  }
  scatter_fluxes<false, false, DENSE>(g, sc);
  if (__any_sync(0xffffffffu, melted)) warp_count_add(&cnt->nbergs_melted, melted);
  double nh = sc.fx.net_heat;
  if (__any_sync(0xffffffffu, nh != 0.)) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nh += shfl_down_d(nh, d);
    if ((threadIdx.x & 31) == 0 && nh != 0.) atomicAdd(&cnt->net_heat_to_ocean, nh);
  }
}

// the bergs k_step_fast deferred: warp w takes list entries [32 w, 32 w + 32), grid-stride
__global__ void __launch_bounds__(KID_BLOCK)
k_step_slow(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
            DevCounters* __restrict__ cnt, const __grid_constant__ SlowList slow) {
  constexpr bool LEAN = true;
  const unsigned long long n = min(*slow.count, (unsigned long long)slow.cap);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < (long long)n; base += stride) {
    const long long k = base + (threadIdx.x & 31);
    Scatter sc;
    sc.key = -1;
    sc.fx.floating_melt = sc.fx.calving_hflx = sc.fx.berg_melt = sc.fx.bergy_src = sc.fx.bergy_melt = 0.;
    sc.fx.fl_bits_melt = sc.fx.fl_bits_src = sc.fx.net_heat = 0.;
    sc.fx.fl_parent_melt = sc.fx.fl_child_melt = sc.fx.melt_buoy = sc.fx.melt_eros = sc.fx.melt_conv = 0.;
    sc.fx.melt_buoy_fl = sc.fx.melt_eros_fl = sc.fx.melt_conv_fl = 0.;
    bool melted = false, became_fl = false, bounced = false, speeding = false, left = false;
    if (k < (long long)n) {
      const uint32_t ent = slow.slots[k];
      const long long s = (long long)(ent & ~KID_SLOW_SPLIT);
      BergIn in;
      in.flags = b.flags[s];
      in.i = b.ine[s]; in.j = b.jne[s];
      in.lat = b.f64[C_LAT][s];
      in.uvel = b.f64[C_UVEL][s]; in.vvel = b.f64[C_VVEL][s];
      in.axn = b.f64[C_AXN][s]; in.ayn = b.f64[C_AYN][s]; in.bxn = b.f64[C_BXN][s]; in.byn = b.f64[C_BYN][s];
      in.xi = b.f64[C_XI][s]; in.yj = b.f64[C_YJ][s];
      in.M = b.f64[C_MASS][s]; in.T = b.f64[C_THICKNESS][s]; in.W = b.f64[C_WIDTH][s]; in.L = b.f64[C_LENGTH][s];
      if (ent & KID_SLOW_SPLIT) step_berg<false, true, LEAN>(g, b, p, cnt, s, in, sc, melted, became_fl, bounced, speeding, left);
      else step_berg<false, false, LEAN>(g, b, p, cnt, s, in, sc, melted, became_fl, bounced, speeding, left);
    }
    scatter_fluxes<false, false, true>(g, sc);
    if (__any_sync(0xffffffffu, melted | bounced | speeding | left)) {
      warp_count_add(&cnt->nbergs_melted, melted);
      warp_count_add(&cnt->n_bounced, bounced);
      warp_count_add(&cnt->nspeeding, speeding);
      warp_count_add(&cnt->n_leavers, left);
    }
    double nh = sc.fx.net_heat;
    if (__any_sync(0xffffffffu, nh != 0.)) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) nh += shfl_down_d(nh, d);
      if ((threadIdx.x & 31) == 0 && nh != 0.) atomicAdd(&cnt->net_heat_to_ocean, nh);
    }
  }
}

// ------------------------------------------------------- Runge-Kutta step
// Runge_Kutta_stepping I:7331-7679 (the namelist default, F:733) + send_bergs + thermodynamics for one
// berg per thread: four accel evaluations at the stage positions (interp_flds at each stage's cell),
// adjust_index_and_ground after every stage, tangent plane above 89N.  Not the benchmarked path: kept
// straightforward (plain IEEE arithmetic in accel_rk) rather than tuned.
struct RkStage { double lon, lat, uvel, vvel, u, v, ax, ay, axn, ayn, xdot, ydot, xddot, yddot, xddotn, yddotn; };

// The four stages and the combination for the (non-static) berg in slot s; stores the accelerations, velocity and
// position and hands back the final cell.  IAFN(u0, v0, u1, v1, IAcc&) evaluates interactive_force for this berg
// (kid_interact.cuh) when INTERACTIVE: the berg's own position enters through lon_old / lat_old and its stored cell,
// so it is the same at every stage (I:611-655); the other bergs are read through their *_old columns only, which
// nobody writes in this sweep (evolve_icebergs refreshes them in its second loop, I:7178-7197).
template <bool INTERACTIVE, class IAFN>
__device__ __forceinline__ void rk_stepping(const DevGrid& g, const DevBergs& b, const DevParams& p, DevCounters* __restrict__ cnt,
                                            long long s, int& i, int& j, double& xi, double& yj, double& lon, double& lat,
                                            double& uvel, double& vvel, double M, double T, double W, double L, double dragfrac,
                                            IAFN&& iafn, bool& any_bounce, bool& speeding) {
  const double dt = p.dt, dt_2 = 0.5 * dt, dt_6 = dt / 6.;
  const int i1 = i, j1 = j;
  const double xi0 = xi, yj0 = yj;
  const bool tang = (lat > 89.) && p.grid_is_latlon;
  const double axn0 = b.f64[C_AXN][s], ayn0 = b.f64[C_AYN][s];
  double bxn = 0., byn = 0., x1 = 0., y1 = 0., dydl;
  RkStage st[4];
  double lonn = lon, latn = lat, uveln = uvel, vveln = vvel, axn = 0., ayn = 0.;
  for (int k = 0; k < 4; k++) {
    RkStage& q = st[k];
    const double h = (k == 3) ? dt : dt_2;          // stage step to reach this stage: dt_2, dt_2, dt
    if (k == 0) {
      q.lon = lon; q.lat = lat; q.uvel = uvel; q.vvel = vvel;
      if (tang) { rotpos_to_tang(p, q.lon, q.lat, x1, y1); rotvec_to_tang(p, q.lon, q.uvel, q.vvel, q.xdot, q.ydot); }
    } else {
      const RkStage& r = st[k - 1];
      if (tang) {
        double x = x1 + h * r.xdot, y = y1 + h * r.ydot;
        q.xdot = st[0].xdot + h * r.xddot; q.ydot = st[0].ydot + h * r.yddot;
        rotpos_from_tang(p, x, y, q.lon, q.lat);
        rotvec_from_tang(p, q.lon, q.xdot, q.ydot, q.uvel, q.vvel);
      } else {
        q.lon = lon + h * r.u; q.lat = lat + h * r.v;
        q.uvel = uvel + h * r.ax; q.vvel = vvel + h * r.ay;
      }
      i = i1; j = j1; xi = xi0; yj = yj0;
      any_bounce |= adjust_index_and_ground(g, p, q.lon, q.lat, i, j, xi, yj, &cnt->error_flags, &cnt->warn_adjust);
    }
    double dxdl;
    convert_from_meters_to_grid(p, q.lat, dxdl, dydl);
    q.u = q.uvel * dxdl; q.v = q.vvel * dydl;
    Env e;
    if (!interp_flds(g, p, i, j, xi, yj, e)) atomicOr(&cnt->error_flags, 64u);
    double loc_dx = 0.;
    if ((p.speed_limit > 0.) || (p.speed_limit == -1.)) {
      int c = gidx(g, i, j);
      loc_dx = fmin(0.5 * (g.dx[c] + g.dx[c - g.nid]), 0.5 * (g.dy[c] + g.dy[c - 1]));
    }
    q.axn = axn0; q.ayn = ayn0;
    IAcc ia = {0., 0., 0., 0., 0., 0., 0., 0.};
    const double u0 = uvel, v0 = vvel;
    if (INTERACTIVE) iafn(u0, v0, u0, v0, ia);                                                    // I:2153
    accel_rk<INTERACTIVE>(p, M, T, W, L, q.lat, q.uvel, q.vvel, uvel, vvel, (k < 2) ? dt_2 : dt, e, loc_dx, dragfrac, ia,
                          [&](double us, double vs, IAcc& a) { iafn(u0, v0, us, vs, a); },            // I:2217
                          q.ax, q.ay, q.axn, q.ayn, bxn, byn, speeding);
    if (tang) { rotvec_to_tang(p, q.lon, q.ax, q.ay, q.xddot, q.yddot); rotvec_to_tang(p, q.lon, q.axn, q.ayn, q.xddotn, q.yddotn); }
  }
  if (tang) {
    double xn = x1 + dt_6 * ((st[0].xdot + st[3].xdot) + 2. * (st[1].xdot + st[2].xdot));
    double yn = y1 + dt_6 * ((st[0].ydot + st[3].ydot) + 2. * (st[1].ydot + st[2].ydot));
    double xdotn = st[0].xdot + dt_6 * ((st[0].xddot + st[3].xddot) + 2. * (st[1].xddot + st[2].xddot));
    double ydotn = st[0].ydot + dt_6 * ((st[0].yddot + st[3].yddot) + 2. * (st[1].yddot + st[2].yddot));
    double xddotn = ((st[0].xddotn + st[3].xddotn) + 2. * (st[1].xddotn + st[2].xddotn)) / 6.;
    double yddotn = ((st[0].yddotn + st[3].yddotn) + 2. * (st[1].yddotn + st[2].yddotn)) / 6.;
    rotpos_from_tang(p, xn, yn, lonn, latn);
    rotvec_from_tang(p, lonn, xdotn, ydotn, uveln, vveln);
    rotvec_from_tang(p, lonn, xddotn, yddotn, axn, ayn);
  } else {
    lonn = lon + dt_6 * ((st[0].u + st[3].u) + 2. * (st[1].u + st[2].u));
    latn = lat + dt_6 * ((st[0].v + st[3].v) + 2. * (st[1].v + st[2].v));
    uveln = uvel + dt_6 * ((st[0].ax + st[3].ax) + 2. * (st[1].ax + st[2].ax));
    vveln = vvel + dt_6 * ((st[0].ay + st[3].ay) + 2. * (st[1].ay + st[2].ay));
    axn = ((st[0].axn + st[3].axn) + 2. * (st[1].axn + st[2].axn)) / 6.;
    ayn = ((st[0].ayn + st[3].ayn) + 2. * (st[1].ayn + st[2].ayn)) / 6.;
    bxn = (((st[0].ax + st[3].ax) + 2. * (st[1].ax + st[2].ax)) / 6) - (axn / 2);
    byn = (((st[0].ay + st[3].ay) + 2. * (st[1].ay + st[2].ay)) / 6) - (ayn / 2);
  }
  i = i1; j = j1; xi = xi0; yj = yj0;
  any_bounce |= adjust_index_and_ground(g, p, lonn, latn, i, j, xi, yj, &cnt->error_flags, &cnt->warn_adjust);
  if (p.override_iceberg_velocities) { uveln = p.u_override; vveln = p.v_override; }
  lon = lonn; lat = latn; uvel = uveln; vvel = vveln;
  b.f64[C_AXN][s] = axn; b.f64[C_AYN][s] = ayn; b.f64[C_BXN][s] = bxn; b.f64[C_BYN][s] = byn;
  b.f64[C_UVEL][s] = uvel; b.f64[C_VVEL][s] = vvel;
  b.f64[C_LON][s] = lon; b.f64[C_LAT][s] = lat;
}

// STEP_ONLY: stop after the stepping (position, velocity and cell stored): send_bergs and thermodynamics follow in
// the SPLIT instance of k_step (footloose runs, whose thermodynamics comes after footloose_calving)
template <bool DIAG, bool STEP_ONLY = false>
__global__ void __launch_bounds__(KID_BLOCK)
k_step_rk(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
          DevCounters* __restrict__ cnt, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t flags = (s < n_slots) ? b.flags[s] : (uint8_t)0;
  bool owned = (flags & BF_ALIVE) && !(flags & BF_HALO);
  Scatter sc;
  sc.key = -1;
  sc.fx.floating_melt = sc.fx.calving_hflx = sc.fx.berg_melt = sc.fx.bergy_src = sc.fx.bergy_melt = 0.;
  sc.fx.fl_bits_melt = sc.fx.fl_bits_src = sc.fx.net_heat = 0.;
  sc.fx.fl_parent_melt = sc.fx.fl_child_melt = sc.fx.melt_buoy = sc.fx.melt_eros = sc.fx.melt_conv = 0.;
  sc.fx.melt_buoy_fl = sc.fx.melt_eros_fl = sc.fx.melt_conv_fl = 0.;
  bool melted = false, any_bounce = false, speeding = false, left = false;
  if (owned) {
    int i = b.ine[s], j = b.jne[s];
    double xi = b.f64[C_XI][s], yj = b.f64[C_YJ][s];
    double lon = b.f64[C_LON][s], lat = b.f64[C_LAT][s], uvel = b.f64[C_UVEL][s], vvel = b.f64[C_VVEL][s];
    const double M = b.f64[C_MASS][s], T = b.f64[C_THICKNESS][s], W = b.f64[C_WIDTH][s], L = b.f64[C_LENGTH][s];
    if (!(flags & BF_STATIC))
      rk_stepping<false>(g, b, p, cnt, s, i, j, xi, yj, lon, lat, uvel, vvel, M, T, W, L, 1.0,
                         [](double, double, double, double, IAcc&) {}, any_bounce, speeding);
    if (STEP_ONLY) {
      if (!(flags & BF_STATIC)) { b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj; b.ine[s] = i; b.jne[s] = j; }
    } else {
      int route = 0;
      if (i > g.iec || i < g.isc || j > g.jec || j < g.jsc)
        route = route_berg(g, p, lon, lat, i, j, xi, yj, &cnt->error_flags, &cnt->n_wrapped);
      if (!(flags & BF_STATIC) || route != 0) {
        b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
        b.ine[s] = i; b.jne[s] = j;
      }
      if (route == 1) {
        left = true;
        b.flags[s] = flags | BF_LEAVER;
        unsigned long long k = atomicAdd(b.leaver_count, 1ull);
        if ((long long)k < b.leaver_cap) b.leaver_list[k] = (int32_t)s;
        else atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_CAPACITY);
      } else if (route == 2) {
        b.flags[s] = 0;
      } else {
        int outcome = thermo_slot<false>(g, b, p, s, flags, i, j, xi, yj, uvel, vvel, M, T, W, L, b.f64[C_MASS_SCALING][s],
                                         b.f64[C_MASS_OF_BITS][s], b.f64[C_HEAT_DENSITY][s], sc, cnt);
        if (outcome == TH_DELETE) { melted = true; b.flags[s] = 0; }
      }
    }
  }
  if (!STEP_ONLY) scatter_fluxes<false, DIAG>(g, sc);
  warp_count_add(&cnt->nbergs_melted, melted);
  warp_count_add(&cnt->n_bounced, any_bounce);
  warp_count_add(&cnt->nspeeding, speeding);
  warp_count_add(&cnt->n_leavers, left);
  double nh = sc.fx.net_heat;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) nh += shfl_down_d(nh, d);
  if ((threadIdx.x & 31) == 0 && nh != 0.) atomicAdd(&cnt->net_heat_to_ocean, nh);
}

// thermodynamics alone over [s0, s1): bergs flagged BF_ARRIVAL (migration) or, with
// all_owned, every owned berg (the split path used when interactions are on).
template <bool FOOTLOOSE, bool DIAG>
__global__ void __launch_bounds__(KID_BLOCK)
k_thermo_range(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
               const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt, long long s0, long long s1,
               int all_owned) {
  long long s = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t flags = (s < s1) ? b.flags[s] : (uint8_t)0;
  bool todo = (flags & BF_ALIVE) && !(flags & BF_LEAVER) && (all_owned || (flags & BF_ARRIVAL));
  if (!p.mts && !p.dem) { /* halo copies melt too, I:2890 */ } else if (flags & BF_HALO) todo = false;
  Scatter sc;
  sc.key = -1;
  sc.fx.floating_melt = sc.fx.calving_hflx = sc.fx.berg_melt = sc.fx.bergy_src = sc.fx.bergy_melt = 0.;
  sc.fx.fl_bits_melt = sc.fx.fl_bits_src = sc.fx.net_heat = 0.;
  sc.fx.fl_parent_melt = sc.fx.fl_child_melt = sc.fx.melt_buoy = sc.fx.melt_eros = sc.fx.melt_conv = 0.;
  sc.fx.melt_buoy_fl = sc.fx.melt_eros_fl = sc.fx.melt_conv_fl = 0.;
  bool melted = false, became_fl = false;
  if (todo) {
    int i = b.ine[s], j = b.jne[s];
    // thermodynamics covers jsc-1:jec+1 x isc-1:iec+1 only (I:2885)
    if (i >= g.isc - 1 && i <= g.iec + 1 && j >= g.jsc - 1 && j <= g.jec + 1) {
      double xi = b.f64[C_XI][s], yj = b.f64[C_YJ][s];
      double uvel = b.f64[C_UVEL][s], vvel = b.f64[C_VVEL][s];
      double M = b.f64[C_MASS][s], T = b.f64[C_THICKNESS][s], W = b.f64[C_WIDTH][s], L = b.f64[C_LENGTH][s];
      int outcome = thermo_slot<FOOTLOOSE>(g, b, p, s, flags, i, j, xi, yj, uvel, vvel, M, T, W, L,
                                           b.f64[C_MASS_SCALING][s], b.f64[C_MASS_OF_BITS][s],
                                           b.f64[C_HEAT_DENSITY][s], sc, cnt);
      if (outcome == TH_DELETE) { melted = true; flags = 0; }
      else if (outcome == TH_BECAME_FL) { melted = true; became_fl = true; }
    }
    if (flags) flags &= ~BF_ARRIVAL;
    b.flags[s] = flags;
  }
  scatter_fluxes<FOOTLOOSE, DIAG>(g, sc);
  warp_count_add(&cnt->nbergs_melted, melted);
  if (FOOTLOOSE) warp_count_add(&cnt->nbergs_calved_fl, became_fl);
  double nh = sc.fx.net_heat;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) nh += shfl_down_d(nh, d);
  if ((threadIdx.x & 31) == 0 && nh != 0.) atomicAdd(&cnt->net_heat_to_ocean, nh);
}

// ------------------------------------------------- scans and in-cell order (the sort itself: kid_sort.cuh)
// exclusive scan of int32 counts, three phases; 1024 items per block
#define KID_SCAN_ITEMS 1024
__global__ void __launch_bounds__(256) k_scan_block(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                    int32_t* __restrict__ block_sums, long long n) {
  __shared__ int32_t sh[256];
  long long base = (long long)blockIdx.x * KID_SCAN_ITEMS + threadIdx.x * 4;
  int32_t v[4];
  int32_t t = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { v[k] = (base + k < n) ? in[base + k] : 0; t += v[k]; }
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    int32_t o = (threadIdx.x >= d) ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += o;
    __syncthreads();
  }
  int32_t excl = sh[threadIdx.x] - t;
  if (threadIdx.x == 255) block_sums[blockIdx.x] = sh[255];
#pragma unroll
  for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
}
// cells holding at least one berg (decides the scatter variant, see scatter_fluxes)
__global__ void k_count_occupied(const int32_t* __restrict__ cell_count, long long n2, int32_t* __restrict__ out) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned m = __ballot_sync(0xffffffffu, c < n2 && cell_count[c] > 0);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, __popc(m));
}

__global__ void __launch_bounds__(1024) k_scan_sums(int32_t* __restrict__ block_sums, int nb, int32_t* total) {
  // single block; serial over chunks of 1024
  __shared__ int32_t sh[1024];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int idx = base + threadIdx.x;
    int32_t t = (idx < nb) ? block_sums[idx] : 0;
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      int32_t o = (threadIdx.x >= d) ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += o;
      __syncthreads();
    }
    if (idx < nb) block_sums[idx] = carry + sh[threadIdx.x] - t;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void k_scan_add(int32_t* __restrict__ out, const int32_t* __restrict__ block_sums, long long n) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] += block_sums[k / KID_SCAN_ITEMS];
}

// the reference's list key, inorder() F:4318-4358: (start_year, start_day, start_mass, start_lon,
// start_lat) ascending; equal keys keep their previous order
__device__ __forceinline__ bool key_before(const DevBergs& b, int32_t x, int32_t y) {
  int32_t ya = b.start_year[x], yb = b.start_year[y];
  if (ya != yb) return ya < yb;
  const int cols[4] = {C_START_DAY, C_START_MASS, C_START_LON, C_START_LAT};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    double va = b.f64[cols[q]][x], vb = b.f64[cols[q]][y];
    if (va != vb) return va < vb;
  }
  return x < y;
}

// Where the reference's per-cell list order decides an integer (ids of footloose children, the pair order of
// the DEM sweeps) the slots of a cell are put in that order: insertion sort by the list key of a run that the
// stable radix sort left in ascending previous slot.  Populations of these configurations are small.
__global__ void k_cell_order(const __grid_constant__ DevBergs b, const int32_t* __restrict__ cell_start,
                             const int32_t* __restrict__ cell_count, long long ncell, int32_t* __restrict__ perm) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int32_t n = cell_count[c];
  if (n < 2) return;
  int32_t* a = perm + cell_start[c];
  for (int32_t k = 1; k < n; k++) {
    int32_t v = a[k];
    int32_t m = k - 1;
    while (m >= 0 && key_before(b, v, a[m])) { a[m + 1] = a[m]; m--; }
    a[m + 1] = v;
  }
}

// ------------------------------------------------------------ grid kernels
struct FieldList { double* f[24]; int n; };
__global__ void k_zero_fields(const __grid_constant__ FieldList fl, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  for (int q = 0; q < fl.n; q++) fl.f[q][k] = 0.;
}

// mpp_update_domains for one rank that is its own E/W neighbour: cyclic wrap of the
// compute rows (see oracle/kid_oracle.c halo_update for the same single-rank semantics)
__global__ void k_halo_wrap_x(const __grid_constant__ DevGrid g, const __grid_constant__ FieldList fl) {
  int halo = g.isc - g.isd;
  int nic = g.iec - g.isc + 1, njc = g.jec - g.jsc + 1;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long per = (long long)2 * halo * njc;
  if (k >= per) return;
  int jj = (int)(k / (2 * halo)), hh = (int)(k % (2 * halo));
  int j = g.jsc + jj;
  int i, src;
  if (hh < halo) { i = g.isd + hh; src = i + nic; } else { i = g.iec + 1 + (hh - halo); src = i - nic; }
  for (int q = 0; q < fl.n; q++) fl.f[q][gidx(g, i, j)] = fl.f[q][gidx(g, src, j)];
}

// copy a caller array into the data-domain field: ring=0 -> (isc:iec,jsc:jec),
// ring=1 -> (isc-1:iec+1,jsc-1:jec+1); optional multiply by msk
__global__ void k_copy_in(const __grid_constant__ DevGrid g, const double* __restrict__ src, double* __restrict__ dst,
                          int ring, int mul_msk, double add) {
  int ni = g.iec - g.isc + 1 + 2 * ring, nj = g.jec - g.jsc + 1 + 2 * ring;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)ni * nj) return;
  int ii = (int)(k % ni), jj = (int)(k / ni);
  size_t c = gidx(g, g.isc - ring + ii, g.jsc - ring + jj);
  double v = src ? src[k] + add : add;
  if (mul_msk) v = v * g.msk[c];
  dst[c] = v;
}

// add_iceberg_thickness_to_SSH I:5330-5337: over the whole data domain the sea surface height is REPLACED by the
// freeboard-equivalent of the berg mass spread in the previous step (cells of zero area keep the input value)
__global__ void k_ssh_from_spread_mass(const __grid_constant__ DevGrid g, const double* __restrict__ spread_mass,
                                       double rho_ratio, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  const double a = g.area[k];
  if (a > 0.) g.ssh[k] = ((spread_mass[k] / a) * rho_ratio);
}

// I:5221, I:5227: whole data domain
// get_running_mean_calving I:5999-6038 on the whole data-domain arrays: the means start from the first field they see
// (I:6010-6017), then rmean = beta*field + alpha*rmean and the model goes on with the mean (I:6035-6036)
__global__ void k_rmean_calving(const __grid_constant__ DevGrid g, double* __restrict__ rm, double* __restrict__ rmh,
                                int init_c, int init_h, double alpha, double beta, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  double c = g.calving[k], hf = g.calving_hflx[k];
  double a = init_c ? rm[k] : c, b = init_h ? rmh[k] : hf;
  if (alpha != 0.) {
    a = beta * c + alpha * a; b = beta * hf + alpha * b;
    g.calving[k] = a; g.calving_hflx[k] = b;
  }
  rm[k] = a; rmh[k] = b;
}
__global__ void k_calving_units(const __grid_constant__ DevGrid g, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  g.calving[k] = g.calving[k] * g.msk[k] * g.area[k];
  g.calving_hflx[k] = g.calving_hflx[k] * g.msk[k];
}

// C-grid velocities to B-grid corners, I:5244-5259 ((nic+2)-sized inputs: offsets 0)
__global__ void k_cgrid_vel(const __grid_constant__ DevGrid g, const double* __restrict__ uo, const double* __restrict__ vo,
                            const double* __restrict__ ui, const double* __restrict__ vi) {
  int ni = g.iec - g.isc + 2, nj = g.jec - g.jsc + 2;   // isc-1..iec, jsc-1..jec
  int nin = g.iec - g.isc + 3;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)ni * nj) return;
  int ii = (int)(k % ni), jj = (int)(k / ni);
  int i = g.isc - 1 + ii, j = g.jsc - 1 + jj;
  size_t c = gidx(g, i, j);
  double mask = fmin(fmin(g.msk[c], g.msk[c + 1]), fmin(g.msk[c + g.nid], g.msk[c + g.nid + 1]));
  size_t a = (size_t)ii + (size_t)jj * nin;
  g.uo[c] = mask * 0.5 * (uo[a] + uo[a + nin]);
  g.ui[c] = mask * 0.5 * (ui[a] + ui[a + nin]);
  g.vo[c] = mask * 0.5 * (vo[a] + vo[a + 1]);
  g.vi[c] = mask * 0.5 * (vi[a] + vi[a + 1]);
}

// wind stress from C-grid / A-grid temporaries (data-domain arrays ut, vt), I:5268-5313
__global__ void k_stress_to_corners(const __grid_constant__ DevGrid g, const double* __restrict__ ut,
                                    const double* __restrict__ vt, int agrid) {
  int ni = g.iec - g.isc + 2, nj = g.jec - g.jsc + 2;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)ni * nj) return;
  int i = g.isc - 1 + (int)(k % ni), j = g.jsc - 1 + (int)(k / ni);
  size_t c = gidx(g, i, j), nid = g.nid;
  double mask = fmin(fmin(g.msk[c], g.msk[c + 1]), fmin(g.msk[c + nid], g.msk[c + nid + 1]));
  if (!agrid) {
    g.ua[c] = mask * 0.5 * (ut[c] + ut[c + nid]);
    g.va[c] = mask * 0.5 * (vt[c] + vt[c + 1]);
  } else {
    g.ua[c] = mask * 0.25 * ((ut[c] + ut[c + nid + 1]) + (ut[c + 1] + ut[c + nid]));
    g.va[c] = mask * 0.25 * ((vt[c] + vt[c + nid + 1]) + (vt[c + 1] + vt[c + nid]));
  }
}

// invert_tau_for_du I:8272-8296
__global__ void k_invert_tau(const __grid_constant__ DevGrid g, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  const double cd = 0.0015;
  double u = g.ua[k], v = g.va[k];
  double tau2 = u * u + v * v;
  double cddvmod = sqrt(cd * sqrt(tau2));
  if (cddvmod != 0.) { g.ua[k] = u / cddvmod; g.va[k] = v / cddvmod; }
  else { g.ua[k] = 0.; g.va[k] = 0.; }
}

// max over the compute domain of sst*msk (I:5340) and "any calving != 0"; one block
// per 256 cells, result via atomics on an ordered-int encoding
__device__ __forceinline__ unsigned long long enc_f64(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dec_f64(unsigned long long u) {
  unsigned long long r = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)r);
#else
  double d; memcpy(&d, &r, 8); return d;
#endif
}
__global__ void k_sst_max(const __grid_constant__ DevGrid g, const double* __restrict__ sst,
                          const double* __restrict__ calving, unsigned long long* __restrict__ out /*[2]*/) {
  int ni = g.iec - g.isc + 1, nj = g.jec - g.jsc + 1;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = -1e300; int anyc = 0;
  if (k < (long long)ni * nj) {
    size_t c = gidx(g, g.isc + (int)(k % ni), g.jsc + (int)(k / ni));
    double t = sst[k] * g.msk[c];
    if (t > v) v = t;
    if (calving && calving[k] != 0.) anyc = 1;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { v = fmax(v, shfl_down_d(v, d)); anyc |= __shfl_down_sync(0xffffffffu, anyc, d); }
  if ((threadIdx.x & 31) == 0) { atomicMax(&out[0], enc_f64(v)); if (anyc) atomicOr(&out[1], 1ull); }
}
__global__ void k_sst_in(const __grid_constant__ DevGrid g, const double* __restrict__ sst,
                         const unsigned long long* __restrict__ mx) {
  int ni = g.iec - g.isc + 1, nj = g.jec - g.jsc + 1;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)ni * nj) return;
  size_t c = gidx(g, g.isc + (int)(k % ni), g.jsc + (int)(k / ni));
  double max_SST = dec_f64(mx[0]);
  g.sst[c] = (max_SST > 120.0) ? sst[k] - 273.15 : sst[k];
}

// I:5364-5383
__global__ void k_scrub(const __grid_constant__ DevGrid g, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  double* f[10] = {g.ua, g.va, g.uo, g.vo, g.ui, g.vi, g.sst, g.sss, g.cn, g.hi};
  bool land = g.msk[k] < 0.5;
#pragma unroll
  for (int q = 0; q < 10; q++) {
    double v = f[q][k];
    if (land) v = 0.0;
    if (v != v) v = 0.;
    f[q][k] = v;
  }
}

__global__ void k_pack_lonlat(const __grid_constant__ DevGrid g, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  LonLat r; r.lon = g.lon[k]; r.lat = g.lat[k];
  g.lonlat[k] = r;
}
// after k_pack_lonlat: the rectangle records of pos_within_cell
__global__ void k_pack_rect(const __grid_constant__ DevGrid g, const __grid_constant__ DevParams p, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  int i = g.isd + (int)(k % g.nid), j = g.jsd + (int)(k / g.nid);
  g.rect[k] = make_rect(g, p, i, j);
}

// corner + cell records from the ingested fields; ddx_ssh/ddy_ssh I:4903-4926
__global__ void k_pack_forcing(const __grid_constant__ DevGrid g, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  CornerRec r;
  r.uo = g.uo[k]; r.vo = g.vo[k]; r.ui = g.ui[k]; r.vi = g.vi[k]; r.ua = g.ua[k]; r.va = g.va[k];
  r.cosr = g.cosr[k]; r.sinr = g.sinr[k];
  g.corner[k] = r;
  int i = g.isd + (int)(k % g.nid), j = g.jsd + (int)(k / g.nid);
  size_t nid = g.nid;
  CellRec c;
  c.sst = g.sst[k]; c.sss = g.sss[k]; c.cn = g.cn[k]; c.hi = g.hi[k];
  c.od = g.ocean_depth[k] + g.ssh[k];
  { double a_ = g.area[k]; c.rarea = (a_ != 0.) ? 1. / a_ : 0.; }
  double ddx = 0., ddy = 0.;
  if (i + 1 <= g.ied && j - 1 >= g.jsd) {
    double dxp = 0.5 * (g.dx[k + 1] + g.dx[k + 1 - nid]);
    double dx0 = 0.5 * (g.dx[k] + g.dx[k - nid]);
    ddx = 2. * (g.ssh[k + 1] - g.ssh[k]) / (dx0 + dxp) * g.msk[k + 1] * g.msk[k];
  }
  if (j + 1 <= g.jed && i - 1 >= g.isd) {
    double dyp = 0.5 * (g.dy[k + nid] + g.dy[k - 1 + nid]);
    double dy0 = 0.5 * (g.dy[k] + g.dy[k - 1]);
    ddy = 2. * (g.ssh[k + nid] - g.ssh[k]) / (dy0 + dyp) * g.msk[k + nid] * g.msk[k];
  }
  c.ddx = ddx; c.ddy = ddy;
  g.cell[k] = c;
}

// I:5654-5679: what icebergs_run hands back
__global__ void k_outputs(const __grid_constant__ DevGrid g, double* __restrict__ calving_out,
                          double* __restrict__ hflx_out) {
  int ni = g.iec - g.isc + 1, nj = g.jec - g.jsc + 1;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)ni * nj) return;
  size_t c = gidx(g, g.isc + (int)(k % ni), g.jsc + (int)(k / ni));
  double a = g.area[c];
  calving_out[k] = (a > 0.) ? g.calving[c] / a + g.floating_melt[c] : 0.;
  hflx_out[k] = g.calving_hflx[c];
}

// sum_mass F:6606-6633 over the owned bergs and sum(stored_ice) over the compute domain (icebergs_stock_pe I:8102)
__global__ void k_stock(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, long long n_slots, long long n2,
                        double* __restrict__ out /* [2]: berg mass, stored ice */) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0., st = 0.;
  if (s < n_slots) {
    uint8_t f = b.flags[s];
    if ((f & BF_ALIVE) && !(f & (BF_HALO | BF_LEAVER)))
      m = (b.f64[C_MASS][s] + b.f64[C_MASS_OF_BITS][s] + b.f64[C_MASS_OF_FL_BITS][s] + b.f64[C_MASS_OF_FL_BERGY_BITS][s]) * b.f64[C_MASS_SCALING][s];
  }
  int ni = g.iec - g.isc + 1, nj = g.jec - g.jsc + 1;
  if (s < (long long)ni * nj) {
    int c = gidx(g, g.isc + (int)(s % ni), g.jsc + (int)(s / ni));
    for (int k = 0; k < KID_NCLASSES; k++) st += g.stored_ice[c + n2 * k];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { m += shfl_down_d(m, d); st += shfl_down_d(st, d); }
  if ((threadIdx.x & 31) == 0) { if (m != 0.) atomicAdd(&out[0], m); if (st != 0.) atomicAdd(&out[1], st); }
}

// compute-domain slice of a data-domain field (what icebergs_run hands back, I:5663-5678)
__global__ void k_copy_out(const __grid_constant__ DevGrid g, const double* __restrict__ src, double* __restrict__ dst) {
  int ni = g.iec - g.isc + 1, nj = g.jec - g.jsc + 1;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)ni * nj) return;
  dst[k] = src[gidx(g, g.isc + (int)(k % ni), g.jsc + (int)(k / ni))];
}

// ------------------------------------------------------------- calving
struct CalvingTables {
  double initial_mass_s[KID_NCLASSES], distribution_s[KID_NCLASSES], mass_scaling_s[KID_NCLASSES],
      initial_thickness_s[KID_NCLASSES];
  double initial_mass_n[KID_NCLASSES], distribution_n[KID_NCLASSES], mass_scaling_n[KID_NCLASSES],
      initial_thickness_n[KID_NCLASSES];
  double LoW_ratio, rho_bergs;
};

// accumulate_calving I:6153-6222 (first_call: I:6175-6188)
__global__ void k_accumulate_calving(const __grid_constant__ DevGrid g, const __grid_constant__ CalvingTables t,
                                     double dt, int first_call, long long n2) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  int i = g.isd + (int)(k % g.nid), j = g.jsd + (int)(k / g.nid);
  bool south = g.lat[k] < 0.;
  double calving = g.calving[k];
  if (first_call && i >= g.isc && i <= g.iec && j >= g.jsc && j <= g.jec && calving != 0.) {
    double s = 0.;
    for (int q = 0; q < KID_NCLASSES; q++) s += g.stored_ice[k + n2 * q];
    g.stored_heat[k] = s * g.calving_hflx[k] * g.area[k] / calving;
  }
  double rd_s = 1., rd_n = 1.;
  for (int q = 0; q < KID_NCLASSES; q++) {
    double dist = south ? t.distribution_s[q] : t.distribution_n[q];
    g.stored_ice[k + n2 * q] = g.stored_ice[k + n2 * q] + dt * calving * dist;
    rd_s = rd_s - t.distribution_s[q];
    rd_n = rd_n - t.distribution_n[q];
  }
  double rd = south ? rd_s : rd_n;
  g.calving[k] = calving * rd;
  double hf = g.calving_hflx[k];
  double tmp = dt * hf * g.area[k] * (1. - rd);
  g.tmp[k] = tmp;
  g.stored_heat[k] = g.stored_heat[k] + tmp;
  g.calving_hflx[k] = hf * rd;
}

// calve_icebergs I:6225-6402: one thread per compute cell walks the classes in the
// reference order, so ids (per-cell counter, F:4165-4177) and start_day offsets
// (I:6381) are the reference's.
__global__ void k_calve(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                        const __grid_constant__ DevParams p, const __grid_constant__ CalvingTables t,
                        DevCounters* __restrict__ cnt, long long n2) {
  int ni = g.iec - g.isc + 1, nj = g.jec - g.jsc + 1;
  long long kk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= (long long)ni * nj) return;
  int i = g.isc + (int)(kk % ni), j = g.jsc + (int)(kk / ni);
  size_t c = gidx(g, i, j);
  bool south = g.lat[c] < 0.;
  double calved_sum = 0., heat_sum = 0.;
  unsigned long long ncalved = 0;
  for (int k = 0; k < KID_NCLASSES; k++) {
    g.real_calving[c + n2 * k] = 0.;
    double initial_mass = south ? t.initial_mass_s[k] : t.initial_mass_n[k];
    double mass_scaling = south ? t.mass_scaling_s[k] : t.mass_scaling_n[k];
    double initial_thickness = south ? t.initial_thickness_s[k] : t.initial_thickness_n[k];
    double* si = &g.stored_ice[c + n2 * k];
    if (!(*si >= initial_mass * mass_scaling)) continue;
    double initial_width = sqrt(initial_mass / (t.LoW_ratio * t.rho_bergs * initial_thickness));   // F:1540
    double initial_length = t.LoW_ratio * initial_width;                                              // F:1541
    double ddt = 0.;
    while (*si >= initial_mass * mass_scaling) {
      Quad q = load_quad(g, i, j);
      double lon = 0.25 * ((q.x3 + q.x1) + (q.x4 + q.x2));
      double lat = 0.25 * ((q.y3 + q.y1) + (q.y4 + q.y2));
      double xi, yj;
      bool lret = pos_within_cell(g, p, lon, lat, i, j, &xi, &yj, &cnt->error_flags);
      if (!lret) { atomicOr(&cnt->error_flags, 128u); return; }
      unsigned long long s = atomicAdd(&cnt->n_slots, 1ull);
      if ((long long)s >= b.capacity) { atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_CAPACITY); return; }
      for (int col = 0; col < C_NCOLS; col++) if (b.f64[col]) b.f64[col][s] = 0.;
      b.f64[C_LON][s] = lon; b.f64[C_LAT][s] = lat;
      b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
      if (b.f64[C_LON_OLD]) { b.f64[C_LON_OLD][s] = lon; b.f64[C_LAT_OLD][s] = lat; }
      b.f64[C_MASS][s] = initial_mass; b.f64[C_THICKNESS][s] = initial_thickness;
      b.f64[C_WIDTH][s] = initial_width; b.f64[C_LENGTH][s] = initial_length;
      b.f64[C_START_LON][s] = lon; b.f64[C_START_LAT][s] = lat;
      b.f64[C_START_DAY][s] = p.current_yearday + ddt / 86400.;
      b.f64[C_START_MASS][s] = initial_mass;
      b.f64[C_MASS_SCALING][s] = mass_scaling;
      double heat_density = g.stored_heat[c] / (*si);
      b.f64[C_HEAT_DENSITY][s] = heat_density;
      b.start_year[s] = p.current_year;
      b.ine[s] = i; b.jne[s] = j;
      int32_t counter = g.iceberg_counter_grd[c] + 1;            // generate_id F:4165-4179
      g.iceberg_counter_grd[c] = counter;
      b.id[s] = (int64_t)counter * ((int64_t)1 << 32) + (int64_t)(i + g.gni * (j - 1));
      b.halo_code[s] = 0;
      b.flags[s] = BF_ALIVE;
      double calved_to_berg = initial_mass * mass_scaling;
      double heat_to_berg = calved_to_berg * heat_density;
      g.stored_heat[c] = g.stored_heat[c] - heat_to_berg;
      heat_sum += heat_to_berg;
      *si = *si - calved_to_berg;
      calved_sum += calved_to_berg;
      g.real_calving[c + n2 * k] += calved_to_berg / p.dt;
      ddt = ddt - p.dt * 2. / 17.;
      ncalved++;
    }
  }
  if (ncalved) {
    atomicAdd(&cnt->nbergs_calved, ncalved);
    atomicAdd(&cnt->net_calving_to_bergs, calved_sum);
    atomicAdd(&cnt->net_heat_to_bergs, heat_sum);
  }
}

// ------------------------------------------------------- restart ingest
// read_restart_bergs (fmsio:606-975): locate the cell unless ine/jne came with the
// file, keep only bergs of this rank's compute domain, recompute xi,yj, *_old = current.
__global__ void k_locate(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                         const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt, long long s0,
                         long long s1, int have_ij) {
  long long s = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= s1) return;
  double lon = b.f64[C_LON][s], lat = b.f64[C_LAT][s];
  int i, j;
  bool found;
  if (have_ij) { i = b.ine[s]; j = b.jne[s]; found = cell_on_pe(g, i, j); }
  else found = find_cell(g, p, lon, lat, &i, &j, &cnt->error_flags);
  if (!found || i < g.isc || i > g.iec || j < g.jsc || j > g.jec) { b.flags[s] = 0; return; }
  double xi, yj;
  pos_within_cell(g, p, lon, lat, i, j, &xi, &yj, &cnt->error_flags);
  b.ine[s] = i; b.jne[s] = j;
  b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
}

// ids for restart files that carry none (legacy iceberg_num): generate_id in file
// order (fmsio:841-845) -- inherently serial, one thread
__global__ void k_generate_ids(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, long long s0,
                               long long s1) {
  if (blockIdx.x || threadIdx.x) return;
  for (long long s = s0; s < s1; s++) {
    if (!(b.flags[s] & BF_ALIVE)) continue;
    int i = b.ine[s], j = b.jne[s];
    size_t c = gidx(g, i, j);
    int32_t counter = g.iceberg_counter_grd[c] + 1;
    g.iceberg_counter_grd[c] = counter;
    b.id[s] = (int64_t)counter * ((int64_t)1 << 32) + (int64_t)(i + g.gni * (j - 1));
  }
}

__global__ void k_count_alive(const uint8_t* __restrict__ flags, long long n_slots, int include_halo,
                              unsigned long long* __restrict__ out) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool a = false;
  if (s < n_slots) { uint8_t f = flags[s]; a = (f & BF_ALIVE) && !(f & BF_LEAVER) && (include_halo || !(f & BF_HALO)); }
  warp_count_add(out, a);
}

}  // namespace kid
