// kid_interact.cuh -- berg-berg interactions of the single-time-step (STS) scheme.
//
//   calculate_force     I:611-804   spring + damping projectors between two elements
//   interactive_force   I:480-607   3x3 neighbour sweep over the cell-sorted store (+ bond loop)
//   k_ia_velocity       first sweep of evolve_icebergs (I:7106-7175) with interactions on:
//                       verlet_stepping/accel for every owned berg, positions untouched
//   (the second sweep -- update_verlet_position, *_old refresh I:7182-7197 -- send_bergs and
//    thermodynamics run in k_step<.., SPLIT=true>)
//   k_clear_halo / k_ghost_* / k_update_latlon   update_halo_icebergs F:1800, update_latlon F:5128
//
// The per-cell linked lists of the reference are the cell_start/cell_count tables of the counting
// sort: in interactive runs the store is re-sorted every step, so the bergs of a cell are a
// contiguous slot range.  In-cell order is ascending previous slot, not the reference's
// (start_year, start_day, ...) order (F:4318): only the rounding of the force sums depends on it.
#pragma once
#include "kid_physics.cuh"

namespace kid {

struct CellTable { const int32_t* start; const int32_t* count; };

// calculate_force I:611-804 once both bergs are in registers: positions, the other berg's velocity, masses and
// interaction radii.  Returns whether the pair exerted a force (its decision depends on the positions only).
// crit: 0 = no c_crit_dist argument, 1 = c_crit_dist=.true. (contact inside a conglomerate: radii only, the
// bond spring constant, I:716-719)
__device__ __forceinline__ bool calculate_force_core(const DevParams& p, double lon1, double lat1, double lon2, double lat2,
                                                     double u2, double v2, double M1, double M2, double R1, double R2, IAcc& A,
                                                     double u0, double v0, double u1, double v1, bool bonded, int crit) {
  double dlon = lon1 - lon2, dlat = lat1 - lat2;
  double lat_ref = 0.5 * (lat1 + lat2), dx_dlon, dy_dlat;
  convert_from_grid_to_meters(p, lat_ref, dx_dlon, dy_dlat);
  double r_dist_x = dlon * dx_dlon, r_dist_y = dlat * dy_dlat;
  double r_dist = sqrt((r_dist_x * r_dist_x) + (r_dist_y * r_dist_y));
  double M_min = M1 < M2 ? M1 : M2;
  double crit_dist, spring_coef;
  if (bonded) { crit_dist = R1 + R2; spring_coef = p.spring_coef; }
  else if (crit == 1) { crit_dist = R1 + R2; spring_coef = p.spring_coef; }
  else { spring_coef = p.contact_spring_coef; crit_dist = fmax(R1 + R2, p.contact_distance); }
  double radial_damping_coef = p.radial_damping_coef, tangental_damping_coef = p.tangental_damping_coef;
  if (p.critical_interaction_damping_on) {
    radial_damping_coef = 2. * sqrt(spring_coef);
    if (p.tang_crit_int_damp_on) tangental_damping_coef = (2. * sqrt(spring_coef)) / 4;
  }
  bool tbonded = bonded;
  // STS with contact_distance = 0 and one spring constant: a bond only pulls (I:741-748)
  if (bonded && !(p.mts || (p.contact_distance > 0.) || (p.contact_spring_coef != p.spring_coef)))
    if (!(r_dist > crit_dist)) tbonded = false;
  if (!((r_dist > 0.) && (tbonded || (r_dist < crit_dist && !bonded)))) return false;
  {
    double accel_spring = spring_coef * (M_min / M1) * (crit_dist - r_dist);
    A.IA_x = A.IA_x + (accel_spring * (r_dist_x / r_dist));
    A.IA_y = A.IA_y + (accel_spring * (r_dist_y / r_dist));
    double r2 = r_dist * r_dist;
    double P_11 = (r_dist_x * r_dist_x) / r2, P_12 = (r_dist_x * r_dist_y) / r2, P_22 = (r_dist_y * r_dist_y) / r2;
    double P_21 = P_12;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
      double p_ia_coef = (pass == 0 ? radial_damping_coef : tangental_damping_coef) * (M_min / M1);
      if (p.scale_damping_by_pmag) {
        double a1 = ((P_11 * (u2 - u1)) + (P_12 * (v2 - v1))), a2 = ((P_12 * (u2 - u1)) + (P_22 * (v2 - v1)));
        double b1 = ((P_11 * (u2 - u0)) + (P_12 * (v2 - v0))), b2 = ((P_12 * (u2 - u0)) + (P_22 * (v2 - v0)));
        p_ia_coef = p_ia_coef * (0.5 * (sqrt((a1 * a1) + (a2 * a2)) + sqrt((b1 * b1) + (b2 * b2))));
      }
      A.P11 = A.P11 + p_ia_coef * P_11; A.P12 = A.P12 + p_ia_coef * P_12;
      A.P21 = A.P21 + p_ia_coef * P_21; A.P22 = A.P22 + p_ia_coef * P_22;
      A.Pu_x = A.Pu_x + (p_ia_coef * ((P_11 * u2) + (P_12 * v2)));
      A.Pu_y = A.Pu_y + (p_ia_coef * ((P_12 * u2) + (P_22 * v2)));
      P_11 = 1 - P_11; P_12 = -P_12; P_21 = -P_21; P_22 = 1 - P_22;      // normal -> tangential projector
    }
  }
  return true;
}
// the interaction radius of a berg of area A = L*W, I:700-714
__device__ __forceinline__ double ia_radius_of(const DevParams& p, double A) {
  if (p.hexagonal_icebergs) return sqrt(A / (2. * sqrt(3.)));
  if (p.iceberg_bonds_on) return 0.5 * sqrt(A);
  return sqrt(A / p.pi);
}

// I:611-804.  s = primary berg, o = other berg (slots).
__device__ __forceinline__ void calculate_force(const DevBergs& b, const DevParams& p, long long s, long long o,
                                                IAcc& A, double u0, double v0, double u1, double v1, bool bonded,
                                                int crit = 0) {
  if (b.id[s] == b.id[o]) return;
  if (b.f64[C_FL_K][s] == -1. || b.f64[C_FL_K][o] == -1.) return;
  double lon1 = b.f64[C_LON_OLD][s], lat1 = b.f64[C_LAT_OLD][s];
  double lon2 = b.f64[C_LON_OLD][o], lat2 = b.f64[C_LAT_OLD][o];
  double u2 = b.f64[C_UVEL_OLD][o], v2 = b.f64[C_VVEL_OLD][o];
  double M1 = b.f64[C_MASS][s], A1 = b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s];
  double M2 = b.f64[C_MASS][o], A2 = b.f64[C_LENGTH][o] * b.f64[C_WIDTH][o];
  calculate_force_core(p, lon1, lat1, lon2, lat2, u2, v2, M1, M2, ia_radius_of(p, A1), ia_radius_of(p, A2), A, u0, v0, u1, v1,
                       bonded, crit);
}

// ---- the plain branch of interactive_force (STS, contact_distance = 0, one spring constant: every berg of the 3x3 cells
// is a candidate, I:577-592) for large unbonded populations.  What calculate_force reads of the OTHER berg is gathered
// once per step into one 64-byte record per berg (k_ia_prepare) -- one line per candidate instead of eight column
// gathers --; a candidate further away in latitude alone, or in longitude alone, than the two radii is dropped before any
// other arithmetic, on a 32-byte key (r_dist >= |r_dist_y| and >= |r_dist_x|: the reference's own test r_dist < crit_dist
// cannot pass; without this the full pair evaluation runs with one or two lanes of a warp active -- ncu: 12.9 of 32 threads
// per executed instruction); and the pairs that did exert a force are
// remembered, so the corrector evaluation of accel (I:2217) walks those few instead of the 3x3 cells again (whether a pair
// acts depends on the *_old positions only).  Same arithmetic, same order of accumulation.
struct __align__(16) IaRec { double lon, lat, u, v, M, R; long long id; long long no_ia; };
struct __align__(16) IaKey { double lon, lat, R, pad; };   // what the pre-tests read: 32 bytes per candidate
#define KID_IA_MAXHIT 12
struct IaHits { int32_t o[KID_IA_MAXHIT]; int32_t n; bool overflow; };

__device__ __forceinline__ const IaKey* ia_keys(const IaRec* rec, long long capacity) { return (const IaKey*)(rec + capacity); }
__global__ void k_ia_prepare(const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p, IaRec* __restrict__ rec,
                             long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  IaRec r;
  r.lon = b.f64[C_LON_OLD][s]; r.lat = b.f64[C_LAT_OLD][s]; r.u = b.f64[C_UVEL_OLD][s]; r.v = b.f64[C_VVEL_OLD][s];
  r.M = b.f64[C_MASS][s]; r.R = ia_radius_of(p, b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s]);
  r.id = b.id[s]; r.no_ia = (b.f64[C_FL_K][s] == -1.) ? 1 : 0;
  rec[s] = r;
  IaKey k;
  k.lon = r.lon; k.lat = r.lat; k.R = r.R; k.pad = 0.;
  ((IaKey*)(rec + b.capacity))[s] = k;       // the keys follow the capacity records
}

__device__ __forceinline__ void interactive_force_plain(const DevGrid& g, const DevBergs& b, const DevParams& p,
                                                        const CellTable& ct, const IaRec* __restrict__ rec, long long s, int i, int j,
                                                        IAcc& A, double u0, double v0, double u1, double v1, IaHits& hits, bool replay) {
  A.IA_x = A.IA_y = A.P11 = A.P12 = A.P21 = A.P22 = A.Pu_x = A.Pu_y = 0.;
  const IaRec me = rec[s];
  if (me.no_ia) return;
  if (replay && !hits.overflow) {
    for (int k = 0; k < hits.n; k++) {
      const IaRec o = rec[hits.o[k]];
      calculate_force_core(p, me.lon, me.lat, o.lon, o.lat, o.u, o.v, me.M, o.M, me.R, o.R, A, u0, v0, u1, v1, false, 0);
    }
  } else {
    const double dy_dlat = p.grid_is_latlon ? (p.pi / 180.) * p.Rearth : 1.;
    // second pre-test, on the longitude difference: r_dist >= |r_dist_x| = |dlon| dx_dlon with dx_dlon = (pi/180) Rearth
    // cos(lat_ref) >= dx_lo, the value at |lat1| + 0.1 degrees -- valid for pairs whose latitudes differ by less than that,
    // which the first pre-test has established whenever the two radii span less than 0.1 degrees of latitude
    const double lat_span = 0.1;
    double dx_lo = 1.;
    if (p.grid_is_latlon) {
      const double a = fabs(me.lat) + lat_span;
      dx_lo = (a < 90.) ? ((p.pi / 180.) * p.Rearth * cos(a * (p.pi / 180.))) * (1. - 1e-9) : 0.;
    }
    // (R1+R2)(1+1e-9) below span_ok: |dlat| < lat_span for what passed test 1; on a Cartesian grid dx_dlon = 1 always
    const double span_ok = p.grid_is_latlon ? lat_span * dy_dlat : 1e300;
    const IaKey* __restrict__ key = ia_keys(rec, b.capacity);
    if (!replay) { hits.n = 0; hits.overflow = false; }
    for (int grdj = j - 1; grdj <= j + 1; grdj++)
      for (int grdi = i - 1; grdi <= i + 1; grdi++) {
        if (grdi < g.isd || grdi > g.ied || grdj < g.jsd || grdj > g.jed) continue;
        int c = gidx(g, grdi, grdj);
        int n = ct.count[c];
        long long o0 = ct.start[c];
        for (int k = 0; k < n; k++) {
          const IaKey ok = key[o0 + k];
          const double crit_hi = (me.R + ok.R) * (1. + 1e-9);
          if (fabs((me.lat - ok.lat) * dy_dlat) > crit_hi) continue;
          if (crit_hi < span_ok && fabs(me.lon - ok.lon) * dx_lo > crit_hi) continue;
          const IaRec o = rec[o0 + k];
          if (o.id == me.id || o.no_ia) continue;
          bool hit = calculate_force_core(p, me.lon, me.lat, o.lon, o.lat, o.u, o.v, me.M, o.M, me.R, o.R, A, u0, v0, u1, v1, false, 0);
          if (hit && !replay) { if (hits.n < KID_IA_MAXHIT) hits.o[hits.n++] = (int32_t)(o0 + k); else hits.overflow = true; }
        }
      }
  }
  if (p.iceberg_bonds_on) {
    // bonds were formed at the front of the list (F:4818): newest first
    for (int k = b.max_bonds - 1; k >= 0; k--) {
      long long slot = (long long)k * b.capacity + s;
      if (b.bond_other_id[slot] == 0) continue;
      int32_t o = b.bond_other_slot[slot];
      if (o < 0) continue;
      calculate_force(b, p, s, o, A, u0, v0, u1, v1, true);
    }
  }
}

// I:480-607, the branch for STS with contact_distance = 0 and one spring constant (I:577-605)
__device__ __forceinline__ void interactive_force(const DevGrid& g, const DevBergs& b, const DevParams& p,
                                                  const CellTable& ct, long long s, int i, int j, IAcc& A,
                                                  double u0, double v0, double u1, double v1, int mts_part = 0) {
  A.IA_x = A.IA_y = A.P11 = A.P12 = A.P21 = A.P22 = A.Pu_x = A.Pu_y = 0.;
  if (b.f64[C_FL_K][s] == -1.) return;
  if (p.mts || (p.contact_distance > 0.) || (p.contact_spring_coef != p.spring_coef)) {
    // I:512-576 (STS): bonded partners through the bonds, the rest of the berg's own conglomerate by radius
    // contact (partners excluded -- the reference marks them by negating their id), other conglomerates
    // within contact_cells with the contact spring
    // with the MTS scheme the bonded half runs in the fast sub-steps only (mts_part 3), collisions between
    // conglomerates in the long step only (mts_part 1), I:522 / I:563
    const int32_t my_cong = b.conglom_id[s];
    if (p.iceberg_bonds_on && (!p.mts || mts_part == 3)) {
      for (int k = b.max_bonds - 1; k >= 0; k--) {
        long long slot = (long long)k * b.capacity + s;
        if (b.bond_other_id[slot] == 0) continue;
        int32_t o = b.bond_other_slot[slot];
        if (o >= 0) calculate_force(b, p, s, o, A, u0, v0, u1, v1, true);
      }
      for (int grdj = max(j - 2, g.jsd + 1); grdj <= min(j + 2, g.jed); grdj++)
        for (int grdi = max(i - 2, g.isd + 1); grdi <= min(i + 2, g.ied); grdi++) {
          int c = gidx(g, grdi, grdj);
          int n = ct.count[c];
          long long o0 = ct.start[c];
          for (int k = 0; k < n; k++) {
            long long o = o0 + k;
            if (b.conglom_id[o] != my_cong) continue;
            bool partner = false;
            for (int q = 0; q < b.max_bonds; q++) {
              long long slot = (long long)q * b.capacity + s;
              if (b.bond_other_id[slot] != 0 && b.bond_other_slot[slot] == (int32_t)o) partner = true;
            }
            if (!partner) calculate_force(b, p, s, o, A, u0, v0, u1, v1, false, 1);
          }
        }
    }
    if (p.mts && mts_part == 3) return;
    const int nc_x = p.contact_cells_lon, nc_y = p.contact_cells_lat;
    for (int grdj = max(j - nc_y, g.jsd); grdj <= min(j + nc_y, g.jed); grdj++)
      for (int grdi = max(i - nc_x, g.isd); grdi <= min(i + nc_x, g.ied); grdi++) {
        int c = gidx(g, grdi, grdj);
        int n = ct.count[c];
        long long o0 = ct.start[c];
        for (int k = 0; k < n; k++)
          if (b.conglom_id[o0 + k] != my_cong) calculate_force(b, p, s, o0 + k, A, u0, v0, u1, v1, false);
      }
    return;
  }
  for (int grdj = j - 1; grdj <= j + 1; grdj++)
    for (int grdi = i - 1; grdi <= i + 1; grdi++) {
      if (grdi < g.isd || grdi > g.ied || grdj < g.jsd || grdj > g.jed) continue;
      int c = gidx(g, grdi, grdj);
      int n = ct.count[c];
      long long o0 = ct.start[c];
      for (int k = 0; k < n; k++) calculate_force(b, p, s, o0 + k, A, u0, v0, u1, v1, false);
    }
  if (p.iceberg_bonds_on) {
    // bonds were formed at the front of the list (F:4818): newest first
    for (int k = b.max_bonds - 1; k >= 0; k--) {
      long long slot = (long long)k * b.capacity + s;
      if (b.bond_other_id[slot] == 0) continue;
      int32_t o = b.bond_other_slot[slot];
      if (o < 0) continue;        // unmatched halo bond (connect_all_bonds F:5059: not an error for halo bergs)
      calculate_force(b, p, s, o, A, u0, v0, u1, v1, true);
    }
  }
}

// first sweep of evolve_icebergs with interactions on: I:7106-7175 (verlet_stepping I:7203, accel I:1950)
__global__ void __launch_bounds__(KID_BLOCK)
k_ia_velocity(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
              const __grid_constant__ DevParams p, const CellTable ct, DevCounters* __restrict__ cnt, long long n_slots,
              const IaRec* __restrict__ rec /* nullptr: the general interactive_force */) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t flags = b.flags[s];
  if (!(flags & BF_ALIVE) || (flags & (BF_HALO | BF_STATIC))) return;
  const double dt = p.dt, dt_2 = 0.5 * dt;
  int i = b.ine[s], j = b.jne[s];
  double lon = b.f64[C_LON][s], lat = b.f64[C_LAT][s], uvel = b.f64[C_UVEL][s], vvel = b.f64[C_VVEL][s];
  double axn = b.f64[C_AXN][s], ayn = b.f64[C_AYN][s], bxn = b.f64[C_BXN][s], byn = b.f64[C_BYN][s];
  double xi = b.f64[C_XI][s], yj = b.f64[C_YJ][s];
  double M = b.f64[C_MASS][s], T = b.f64[C_THICKNESS][s], W = b.f64[C_WIDTH][s], L = b.f64[C_LENGTH][s];
  double uvel_prev = uvel - dt_2 * bxn, vvel_prev = vvel - dt_2 * byn;
  double uvel3 = uvel + (dt_2 * axn), vvel3 = vvel + (dt_2 * ayn);
  Env e;
  if (!interp_flds(g, p, i, j, xi, yj, e)) atomicOr(&cnt->error_flags, 64u);
  double sin_lat = 0., cos_lat = 1.;
  if (p.grid_is_latlon) sincos_halfpi(p.pi_180 * lat, &sin_lat, &cos_lat);
  double f_cori = (p.grid_is_latlon && !p.use_f_plane) ? p.omega2 * sin_lat : p.f_cori_plane;
  double dragfrac = 1.0;
  if (p.iceberg_bonds_on && p.internal_bergs_for_drag) {       // I:2104-2120
    double N_bonds = 0., N_max = p.hexagonal_icebergs ? 6.0 : 4.0;
    for (int k = 0; k < b.max_bonds; k++) if (b.bond_other_id[(long long)k * b.capacity + s] != 0) N_bonds += 1.0;
    dragfrac = ((N_max - N_bonds) / N_max);
  }
  IAcc ia;
  IaHits hits;
  hits.n = 0; hits.overflow = true;
  if (rec) interactive_force_plain(g, b, p, ct, rec, s, i, j, ia, uvel, vvel, uvel, vvel, hits, false);
  else interactive_force(g, b, p, ct, s, i, j, ia, uvel, vvel, uvel, vvel);      // I:2153
  double ax1, ay1, un_l, vn_l;
  const double u0 = uvel, v0 = vvel;
  accel_core<true, false>(p, M, T, W, L, f_cori, uvel, vvel, dt, e, dragfrac, ia,
                   [&](double us, double vs, IAcc& q) {                                                                    // I:2217
                     if (rec) interactive_force_plain(g, b, p, ct, rec, s, i, j, q, u0, v0, us, vs, hits, true);
                     else interactive_force(g, b, p, ct, s, i, j, q, u0, v0, us, vs);
                   },
                   ax1, ay1, axn, ayn, bxn, byn, un_l, vn_l);
  if ((p.speed_limit > 0.) || (p.speed_limit == -1.)) {
    double speed = sqrt(un_l * un_l + vn_l * vn_l);
    if (speed > 0.) {
      int c = gidx(g, i, j);
      double loc_dx = fmin(0.5 * (g.dx[c] + g.dx[c - g.nid]), 0.5 * (g.dy[c] + g.dy[c - 1]));
      double new_speed = loc_dx / dt * p.speed_limit;
      if (new_speed < speed && p.speed_limit > 0.) atomicAdd(&cnt->nspeeding, 1ull);
    }
  }
  bool tang = (lat > 89.) && p.grid_is_latlon;
  double uveln, vveln;
  if (tang) tang_velocity(p, lon, uvel3, vvel3, ax1, ay1, dt, uveln, vveln);
  else { uveln = uvel3 + (dt * ax1); vveln = vvel3 + (dt * ay1); }
  if (p.override_iceberg_velocities) { uveln = p.u_override; vveln = p.v_override; }
  b.f64[C_UVEL_PREV][s] = uvel_prev; b.f64[C_VVEL_PREV][s] = vvel_prev;
  b.f64[C_AXN][s] = axn; b.f64[C_AYN][s] = ayn; b.f64[C_BXN][s] = bxn; b.f64[C_BYN][s] = byn;
  b.f64[C_UVEL][s] = uveln; b.f64[C_VVEL][s] = vveln;
}

// ---------------------------------------------------------------- ghosts
// update_halo_icebergs F:1800: step 1, clear the halo copies
__global__ void k_clear_halo(uint8_t* __restrict__ flags, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  if (flags[s] & BF_HALO) flags[s] = 0;
}

// Which of the 8 neighbours get a copy of a berg in cell (i,j): the strips of F:1896-1912 (E/W)
// and F:1976-2006 (N/S, over the data-domain columns so that corner copies travel too).
struct GhostPlan {
  int32_t nbr[9];            // rank per direction, -1 none
  int32_t hw, isc, iec, jsc, jec;
  int32_t gni, cyclic_x;
};
__device__ __forceinline__ bool ghost_goes(const GhostPlan& gp, int dir, int i, int j) {
  int dx = dir % 3 - 1, dy = dir / 3 - 1;
  if (gp.nbr[dir] < 0) return false;
  if (dx > 0 && !(i >= gp.iec - gp.hw + 2)) return false;
  if (dx < 0 && !(i <= gp.isc + gp.hw - 1)) return false;
  if (dy > 0 && !(j >= gp.jec - gp.hw + 2)) return false;
  if (dy < 0 && !(j <= gp.jsc + gp.hw - 1)) return false;
  return true;
}

__global__ void k_ghost_count(const __grid_constant__ GhostPlan gp, const uint8_t* __restrict__ flags,
                              const int32_t* __restrict__ ine, const int32_t* __restrict__ jne, long long n_slots,
                              int32_t* __restrict__ counts /* [9] */) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t f = flags[s];
  if (!(f & BF_ALIVE) || (f & (BF_HALO | BF_LEAVER))) return;
  int i = ine[s], j = jne[s];
  for (int dir = 0; dir < 9; dir++)
    if (dir != 4 && ghost_goes(gp, dir, i, j)) atomicAdd(&counts[dir], 1);
}

// pack_berg_into_buffer2 with halo_berg = 1 (F:1902-1905); bonds travel as (other_id, other ine/jne, length)
__global__ void k_ghost_pack(const __grid_constant__ GhostPlan gp, const __grid_constant__ DevBergs b, long long n_slots,
                             const int32_t* __restrict__ offsets /* [9] */, int32_t* __restrict__ cursor /* [9] */,
                             double* __restrict__ sendbuf, const __grid_constant__ RecLayout RL) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t f = b.flags[s];
  if (!(f & BF_ALIVE) || (f & (BF_HALO | BF_LEAVER))) return;
  int i = b.ine[s], j = b.jne[s];
  for (int dir = 0; dir < 9; dir++) {
    if (dir == 4 || !ghost_goes(gp, dir, i, j)) continue;
    int pos = offsets[dir] + atomicAdd(&cursor[dir], 1);
    double* rec = sendbuf + (size_t)pos * RL.w;
    pack_berg(b, s, rec, RL);
    // the copy's cell on the receiving side: one period away when the message crosses the cyclic seam
    int dx = dir % 3 - 1, ci = i;
    if (gp.cyclic_x) { if (dx > 0 && gp.iec == gp.gni) ci = i - gp.gni; else if (dx < 0 && gp.isc == 1) ci = i + gp.gni; }
    rec[PK_INE_JNE] = __longlong_as_double(((long long)(unsigned)ci << 32) | (unsigned)j);
    rec[PK_YEAR_FLAGS] = __longlong_as_double(((long long)(unsigned)b.start_year[s] << 32) | (unsigned)f);
  }
}

// unpack_berg_from_buffer2 F:3468 for halo copies: cell re-found, xi/yj recomputed, *_old = current
__global__ void k_ghost_unpack(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                               const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt,
                               const double* __restrict__ recvbuf, long long n_recv, long long s0,
                               const __grid_constant__ RecLayout RL) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_recv) return;
  long long s = s0 + k;
  const double* rec = recvbuf + (size_t)k * RL.w;
  unpack_berg(b, s, rec, RL);
  double lon = rec[PK_F64_0 + C_LON], lat = rec[PK_F64_0 + C_LAT];
  long long ij = __double_as_longlong(rec[PK_INE_JNE]), yf = __double_as_longlong(rec[PK_YEAR_FLAGS]);
  int i = (int)(ij >> 32), j = (int)(ij & 0xffffffffll);
  b.start_year[s] = (int32_t)(yf >> 32);
  uint8_t f = (uint8_t)(yf & 0xff);
  // check_and_find_cell F:5973: the cell the sender named (already the periodic image when the copy
  // crossed the seam, k_ghost_pack), then its other images, then the scan
  bool found = false;
  int oi = i, oj = j;
  if (cell_on_pe(g, oi, oj)) found = is_point_in_cell(g, p, lon, lat, oi, oj, &cnt->error_flags);
  if (!found && g.cyclic_x) {
    for (int sh = -1; sh <= 1 && !found; sh += 2) {
      oi = i + sh * g.gni;
      if (cell_on_pe(g, oi, oj)) found = is_point_in_cell(g, p, lon, lat, oi, oj, &cnt->error_flags);
    }
  }
  if (!found) found = find_cell_wide(g, p, lon, lat, &oi, &oj, &cnt->error_flags);
  if (!found) { b.flags[s] = 0; return; }      // a copy nobody needs (F:3668 is FATAL for owned bergs only in practice)
  double xi, yj;
  pos_within_cell(g, p, lon, lat, oi, oj, &xi, &yj, &cnt->error_flags);
  b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
  b.ine[s] = oi; b.jne[s] = oj;
  b.halo_code[s] = 1;
  b.flags[s] = (uint8_t)((f | BF_ALIVE | BF_HALO) & ~(BF_LEAVER | BF_ARRIVAL));
}

// first sweep of evolve_icebergs with interactions on and Runge_not_Verlet (the namelist default, F:733):
// Runge_Kutta_stepping I:7331-7679 with interactive_force inside every accel call.  Position, velocity and cell are
// stored; *_old, send_bergs and thermodynamics follow in k_step<.., SPLIT=true>.
__global__ void __launch_bounds__(KID_BLOCK)
k_step_rk_ia(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
             const CellTable ct, DevCounters* __restrict__ cnt, long long n_slots,
             const IaRec* __restrict__ rec /* nullptr: the general interactive_force */) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t flags = (s < n_slots) ? b.flags[s] : (uint8_t)0;
  const bool active = (flags & BF_ALIVE) && !(flags & (BF_HALO | BF_STATIC));
  bool any_bounce = false, speeding = false;
  if (active) {
    int i = b.ine[s], j = b.jne[s];
    const int i0 = i, j0 = j;
    double xi = b.f64[C_XI][s], yj = b.f64[C_YJ][s];
    double lon = b.f64[C_LON][s], lat = b.f64[C_LAT][s], uvel = b.f64[C_UVEL][s], vvel = b.f64[C_VVEL][s];
    const double M = b.f64[C_MASS][s], T = b.f64[C_THICKNESS][s], W = b.f64[C_WIDTH][s], L = b.f64[C_LENGTH][s];
    double dragfrac = 1.0;
    if (p.iceberg_bonds_on && p.internal_bergs_for_drag) {       // I:2104-2120
      double N_bonds = 0., N_max = p.hexagonal_icebergs ? 6.0 : 4.0;
      for (int k = 0; k < b.max_bonds; k++) if (b.bond_other_id[(long long)k * b.capacity + s] != 0) N_bonds += 1.0;
      dragfrac = ((N_max - N_bonds) / N_max);
    }
    // every evaluation of the four stages works on the *_old positions (I:637-640): the pairs that act are found once
    IaHits hits;
    hits.n = 0; hits.overflow = true;
    bool built = false;
    rk_stepping<true>(g, b, p, cnt, s, i, j, xi, yj, lon, lat, uvel, vvel, M, T, W, L, dragfrac,
                      [&](double u0, double v0, double u1, double v1, IAcc& a) {
                        if (rec) { interactive_force_plain(g, b, p, ct, rec, s, i0, j0, a, u0, v0, u1, v1, hits, built); built = true; }
                        else interactive_force(g, b, p, ct, s, i0, j0, a, u0, v0, u1, v1);
                      },
                      any_bounce, speeding);
    b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj; b.ine[s] = i; b.jne[s] = j;
  }
  warp_count_add(&cnt->n_bounced, any_bounce);
  warp_count_add(&cnt->nspeeding, speeding);
}

// update_latlon F:5128-5169: positions re-derived from (cell, xi, yj) so that copies across the
// periodic seam carry the coordinates of the halo cell they sit in
__global__ void k_update_latlon(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                                const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t f = b.flags[s];
  if (!(f & BF_ALIVE) || (f & BF_LEAVER)) return;
  int i = b.ine[s], j = b.jne[s];
  if (!cell_on_pe(g, i, j)) return;
  if (p.mts && b.halo_code[s] >= 2) return;      // copies beyond the halo keep their coordinates (F:4992)
  double lon = b.f64[C_LON][s], lat = b.f64[C_LAT][s];
  double dlon = lon - b.f64[C_LON_OLD][s], dlat = lat - b.f64[C_LAT_OLD][s];
  double xi = b.f64[C_XI][s], yj = b.f64[C_YJ][s];
  bilin_lonlat(g, p, i, j, xi, yj, &lon, &lat);
  b.f64[C_LON][s] = lon; b.f64[C_LAT][s] = lat;
  b.f64[C_LON_OLD][s] = lon - dlon; b.f64[C_LAT_OLD][s] = lat - dlat;
  pos_within_cell(g, p, lon, lat, i, j, &xi, &yj, &cnt->error_flags);
  b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
}

}  // namespace kid

namespace kid {

// ------------------------------------------------------------------ bonds
// bond_address_update F:4887-4915: the (ine,jne) hint of every connected bond follows its partner
__global__ void k_bond_address_update(const __grid_constant__ DevBergs b, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE)) return;
  for (int k = 0; k < b.max_bonds; k++) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] == 0) continue;
    int32_t o = b.bond_other_slot[slot];
    if (o < 0) continue;
    b.bond_other_ine[slot] = b.ine[o]; b.bond_other_jne[slot] = b.jne[o];
  }
}

__device__ __forceinline__ int find_id_in_cell(const DevGrid& g, const DevBergs& b, const CellTable& ct, int i, int j,
                                               int64_t id) {
  if (!((i > g.isd - 1) && (i < g.ied + 1) && (j > g.jsd - 1) && (j < g.jed + 1))) return -1;
  int c = gidx(g, i, j);
  int n = ct.count[c], o0 = ct.start[c];
  for (int k = 0; k < n; k++) if (b.id[o0 + k] == id) return o0 + k;
  return -1;
}

// connect_all_bonds F:4963-5125 on the freshly sorted store: every bond looks its partner up in the
// hinted cell, then around the hint, then in the 5x5 cells around the berg (updating the hint).
// half_period > 0 (cyclic x): the store may hold several periodic images of one berg (the copies of
// transfer_mts_bergs); a candidate more than half a period away in x is an image of the partner, not the partner.
__device__ __forceinline__ int find_partner_in_cell(const DevGrid& g, const DevBergs& b, const CellTable& ct, int i, int j,
                                                    int64_t id, double lon, double half_period) {
  if (!((i > g.isd - 1) && (i < g.ied + 1) && (j > g.jsd - 1) && (j < g.jed + 1))) return -1;
  int c = gidx(g, i, j);
  int n = ct.count[c], o0 = ct.start[c];
  for (int k = 0; k < n; k++)
    if (b.id[o0 + k] == id && !(half_period > 0. && fabs(b.f64[C_LON][o0 + k] - lon) > half_period)) return o0 + k;
  return -1;
}
__global__ void k_connect_bonds(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                                const CellTable ct, DevCounters* __restrict__ cnt, long long n_slots, double half_period) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE)) return;
  const double lon = b.f64[C_LON][s];
  for (int k = 0; k < b.max_bonds; k++) {
    long long slot = (long long)k * b.capacity + s;
    int64_t oid = b.bond_other_id[slot];
    if (oid == 0) continue;
    int i = b.bond_other_ine[slot], j = b.bond_other_jne[slot];
    const int bi = b.ine[s], bj = b.jne[s];
    // Bonded elements sit within a cell or two of each other.  A hint further away names the partner's copy on the
    // other side of the cyclic seam (the berg, or the partner, has just wrapped): with several PEs along x that cell is
    // off-PE and the reference falls through to the search around the berg, which finds the near (halo) copy; on one
    // rank the far copy IS in the hinted cell, so the search around the berg goes first there.
    const bool far = (abs(i - bi) > 2) || (abs(j - bj) > 2);
    int o = -1;
    for (int pass = 0; pass < 2 && o < 0; pass++) {
      if ((pass == 0) == !far) {            // near hint: hinted cell and its ring first; far hint: last
        o = find_partner_in_cell(g, b, ct, i, j, oid, lon, half_period);
        for (int jj = j - 1; jj <= j + 1 && o < 0; jj++)
          for (int ii = i - 1; ii <= i + 1 && o < 0; ii++)
            if (ii != i || jj != j) o = find_partner_in_cell(g, b, ct, ii, jj, oid, lon, half_period);
      } else {
        for (int jj = bj - 2; jj <= bj + 2 && o < 0; jj++)
          for (int ii = bi - 2; ii <= bi + 2 && o < 0; ii++) {
            o = find_partner_in_cell(g, b, ct, ii, jj, oid, lon, half_period);
            if (o >= 0) { b.bond_other_ine[slot] = ii; b.bond_other_jne[slot] = jj; }
          }
      }
    }
    b.bond_other_slot[slot] = o;
    if (o < 0 && !(b.flags[s] & BF_HALO)) atomicOr(&cnt->error_flags, 256u);   // 'A non-halo bond is missing!!!' F:5063
  }
}

// set_conglom_ids F:2601-2646: connected components of the bond graph by label propagation (the reference
// flood-fills recursively, F:2649; only equality of labels is ever tested, so any labelling of the same
// components is equivalent).  Owned bergs start from slot+1; copies that no owned berg reaches stay 0.
__global__ void k_conglom_init(const __grid_constant__ DevBergs b, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t f = b.flags[s];
  b.conglom_id[s] = ((f & BF_ALIVE) && !(f & BF_HALO)) ? (int32_t)(s + 1) : 0;
}
__global__ void k_conglom_sweep(const __grid_constant__ DevBergs b, long long n_slots, int* __restrict__ changed) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE)) return;
  int32_t best = b.conglom_id[s];
  for (int k = 0; k < b.max_bonds; k++) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] == 0) continue;
    if (b.bond_broken && b.bond_broken[slot] == 1) continue;      // dem: a broken bond no longer joins (F:2661)
    int32_t o = b.bond_other_slot[slot];
    if (o < 0) continue;
    int32_t l = b.conglom_id[o];
    if (l != 0 && (best == 0 || l < best)) best = l;
  }
  if (best != b.conglom_id[s]) { b.conglom_id[s] = best; *changed = 1; }
}

// the same propagation to its fixed point inside ONE CTA (stores of a few thousand slots: bonded runs): no launch and no
// host round trip per sweep.  The fixed point -- the smallest seed label of the component -- does not depend on the
// order of the updates, so the labels are those of the sweep kernel.
__global__ void __launch_bounds__(1024) k_conglom_label_one_cta(const __grid_constant__ DevBergs b, long long n_slots) {
  __shared__ int changed;
  const int n = (int)n_slots, nt = blockDim.x, tid = threadIdx.x;
  for (int it = 0; it < 1000000; it++) {
    __syncthreads();
    if (tid == 0) changed = 0;
    __syncthreads();
    for (int s = tid; s < n; s += nt) {
      if (!(b.flags[s] & BF_ALIVE)) continue;
      int32_t best = b.conglom_id[s];
      for (int k = 0; k < b.max_bonds; k++) {
        long long slot = (long long)k * b.capacity + s;
        if (b.bond_other_id[slot] == 0) continue;
        if (b.bond_broken && b.bond_broken[slot] == 1) continue;
        int32_t o = b.bond_other_slot[slot];
        if (o < 0) continue;
        int32_t l = ((volatile int32_t*)b.conglom_id)[o];
        if (l != 0 && (best == 0 || l < best)) best = l;
      }
      if (best != b.conglom_id[s]) { ((volatile int32_t*)b.conglom_id)[s] = best; changed = 1; }
    }
    __syncthreads();
    if (!changed) break;
  }
}

// initialize_iceberg_bonds I:356-441: O(N^2) distance test over every berg of the data domain, in the
// reference's outer/inner order (cells j-major, bergs in list order) so that each berg's bond list
// has the reference's order.  One thread per berg; the inner loop walks the sorted store.
__global__ void k_init_bonds(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                             const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt, long long n_slots,
                             double bond_length_crit, int from_radii) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE)) return;
  double lon1 = b.f64[C_LON][s], lat1 = b.f64[C_LAT][s];
  double rdenom = p.hexagonal_icebergs ? 1. / (2. * sqrt(3.)) : 1. / 4.;
  for (long long o = 0; o < n_slots; o++) {
    if (!(b.flags[o] & BF_ALIVE) || b.id[o] == b.id[s]) continue;
    bool already = false;
    int nfree = -1;
    for (int k = 0; k < b.max_bonds; k++) {
      int64_t oid = b.bond_other_id[(long long)k * b.capacity + s];
      if (oid == b.id[o]) already = true;
      if (oid == 0 && nfree < 0) nfree = k;
    }
    if (already) continue;
    double lon2 = b.f64[C_LON][o], lat2 = b.f64[C_LAT][o];
    double dlon = lon1 - lon2, dlat = lat1 - lat2, dx_dlon, dy_dlat;
    convert_from_grid_to_meters(p, 0.5 * (lat1 + lat2), dx_dlon, dy_dlat);
    double rx = dlon * dx_dlon, ry = dlat * dy_dlat;
    double r_dist = sqrt((rx * rx) + (ry * ry));
    bool bond;
    if (from_radii) {
      double radius1 = sqrt(b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s] * rdenom);
      double radius2 = sqrt(b.f64[C_LENGTH][o] * b.f64[C_WIDTH][o] * rdenom);
      bond = r_dist < 1.25 * (radius1 + radius2);
    } else bond = r_dist < bond_length_crit;
    if (!bond) continue;
    if (nfree < 0) { atomicOr(&cnt->error_flags, 512u); continue; }      // more than max_bonds partners
    long long slot = (long long)nfree * b.capacity + s;
    b.bond_other_id[slot] = b.id[o];
    b.bond_other_slot[slot] = (int32_t)o;
    b.bond_other_ine[slot] = b.ine[o]; b.bond_other_jne[slot] = b.jne[o];
    b.bond_length[slot] = 0.;
  }
}

// orig_bond_length F:4589-4614 (first visit, I:5420): bond%length = current separation of the pair
// in grid units (degrees or metres, whatever lon/lat are)
__global__ void k_orig_bond_length(const __grid_constant__ DevBergs b, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE)) return;
  for (int k = 0; k < b.max_bonds; k++) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] == 0) continue;
    int32_t o = b.bond_other_slot[slot];
    if (o < 0) continue;
    double dl = b.f64[C_LON][s] - b.f64[C_LON][o], dp = b.f64[C_LAT][s] - b.f64[C_LAT][o];
    b.bond_length[slot] = sqrt((dl * dl) + (dp * dp));
  }
}

}  // namespace kid

namespace kid {

// ----------------------------------------------------------------- footloose
// footloose_calving I:2503-2734 + calve_fl_icebergs I:6405-6569 (displace_fl_bergs off: the
// displacement draws from the FMS random stream).  One thread per compute cell walks the cell's bergs
// in store order -- the store is sorted with the reference's list key (k_cell_order keyed) -- so the
// per-cell id counter (generate_id F:4165) hands out the reference's ids.
struct FlConsts { double lfootparam, l_c, lw_c, B_c; };

__device__ __forceinline__ long long fl_new_child(const DevGrid& g, const DevBergs& b, const DevParams& p,
                                                  DevCounters* __restrict__ cnt, long long ps, double k, double l_b,
                                                  bool from_bits) {
  unsigned long long s = atomicAdd(&cnt->n_slots, 1ull);
  if ((long long)s >= b.capacity) { atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_CAPACITY); return -1; }
  for (int c = 0; c < C_NCOLS; c++) if (b.f64[c]) b.f64[c][s] = b.f64[c][ps];      // lon, lat, xi, yj, velocities, accelerations, ...
  double len, wid, thk, mass, ms, mob = 0.;
  if (from_bits) {
    // fl_bits_dimensions I:3370-3387
    const double l_c = p.pi / (2. * sqrt(2.)), lw_c = 1. / (KID_GRAVITY * KID_RHO_SEAWATER);
    const double B_c = 1. / (12. * (1. - pow(0.3, 2.)));
    double l_w = pow(lw_c * p.fl_youngs * B_c * pow(b.f64[C_THICKNESS][ps], 3.), 0.25);
    double lb = l_c * l_w;
    len = 3. * lb; wid = lb; thk = b.f64[C_THICKNESS][ps];
    rolling(p, thk, wid, len);
    mass = thk * len * wid * p.rho_bergs;
    ms = k * p.new_berg_from_fl_bits_mass_thres / mass;
    double pms = b.f64[C_MASS_SCALING][ps];
    double percent_fl = (mass * ms) / (b.f64[C_MASS_OF_FL_BITS][ps] * pms);
    mob = (percent_fl * b.f64[C_MASS_OF_FL_BERGY_BITS][ps] * pms) / ms;
    b.f64[C_MASS_OF_FL_BERGY_BITS][ps] = (1 - percent_fl) * b.f64[C_MASS_OF_FL_BERGY_BITS][ps];
    b.f64[C_MASS_OF_FL_BITS][ps] = b.f64[C_MASS_OF_FL_BITS][ps] - k * p.new_berg_from_fl_bits_mass_thres / pms;
  } else {
    len = l_b * 3.; wid = l_b; thk = b.f64[C_THICKNESS][ps];
    mass = wid * len * thk * p.rho_bergs;
    ms = b.f64[C_MASS_SCALING][ps] * k;
  }
  b.f64[C_LENGTH][s] = len; b.f64[C_WIDTH][s] = wid; b.f64[C_THICKNESS][s] = thk; b.f64[C_MASS][s] = mass;
  b.f64[C_MASS_SCALING][s] = ms; b.f64[C_MASS_OF_BITS][s] = mob;
  b.f64[C_START_LON][s] = b.f64[C_LON][ps]; b.f64[C_START_LAT][s] = b.f64[C_LAT][ps];
  b.f64[C_START_DAY][s] = p.current_yearday;
  b.f64[C_MASS_OF_FL_BITS][s] = 0.; b.f64[C_MASS_OF_FL_BERGY_BITS][s] = 0.;
  b.f64[C_FL_K][s] = -1.0;
  b.start_year[s] = p.current_year;
  int i = b.ine[ps], j = b.jne[ps];
  b.ine[s] = i; b.jne[s] = j;
  int c = gidx(g, i, j);
  int32_t counter = g.iceberg_counter_grd[c] + 1;            // generate_id F:4165-4179, the parent's cell
  g.iceberg_counter_grd[c] = counter;
  b.id[s] = (int64_t)counter * ((int64_t)1 << 32) + (int64_t)(i + g.gni * (j - 1));
  b.halo_code[s] = 0;
  b.flags[s] = (uint8_t)(BF_ALIVE | (b.flags[ps] & BF_STATIC));
  for (int q = 0; q < b.max_bonds; q++) { b.bond_other_id[(long long)q * b.capacity + s] = 0; b.bond_other_slot[(long long)q * b.capacity + s] = -1; }
  return (long long)s;
}

__global__ void k_footloose(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                            const __grid_constant__ DevParams p, const __grid_constant__ FlConsts fc, const CellTable ct,
                            DevCounters* __restrict__ cnt, int fl_style_fl_bits, double mass_thres) {
  int ni = g.iec - g.isc + 1, nj = g.jec - g.jsc + 1;
  long long kk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= (long long)ni * nj) return;
  int grdi = g.isc + (int)(kk % ni), grdj = g.jsc + (int)(kk / ni);
  int cell = gidx(g, grdi, grdj);
  int n = ct.count[cell], s0 = ct.start[cell];
  double l_b = 0., c = 0.;
  unsigned long long ncalved = 0;
  double area = g.area[cell];
  for (int q = 0; q < n; q++) {
    long long s = s0 + q;
    uint8_t f = b.flags[s];
    if (!(f & BF_ALIVE) || (f & (BF_HALO | BF_LEAVER))) continue;
    double fl_k = b.f64[C_FL_K][s];
    double ms = b.f64[C_MASS_SCALING][s];
    if (!((f & BF_STATIC) || fl_k < 0)) {
      double T = b.f64[C_THICKNESS][s], W = b.f64[C_WIDTH][s], L = b.f64[C_LENGTH][s];
      double N_bonds = 0;
      for (int k = 0; k < b.max_bonds; k++) if (b.bond_other_id[(long long)k * b.capacity + s] != 0) N_bonds += 1.0;
      if (N_bonds > 0) { atomicOr(&cnt->error_flags, 1024u); return; }      // 'Bonded footloose calving not yet fully implemented!' I:2567
      double l_w = pow(fc.lw_c * fc.B_c * pow(T, 3.), 0.25);
      l_b = fc.l_c * l_w;
      double l_b3 = 3 * l_b, Lmin, Wmin, max_k, k, foot_area;
      c = ceil((L - l_b3) / l_b3); Lmin = L - c * l_b3;
      c = ceil((W - l_b3) / l_b3); Wmin = W - c * l_b3;
      max_k = fmax(floor((L * W - Lmin * Wmin) / (l_b3 * l_b)), 0.);
      if (max_k == 0) k = 0;
      else {
        double foot_l = fc.lfootparam * T / l_w;
        foot_area = foot_l * l_b3;
        k = floor(fl_k / foot_area);
        if (k > max_k) k = max_k;
        fl_k = fl_k - k * foot_area;
        b.f64[C_FL_K][s] = fl_k;
      }
      if (k > 0) {
        double ds, Ln, Wn;
        if (c > 0) {
          ds = 0.5 * ((L + W) - sqrt(pow(L + W, 2.) - 4. * (l_b3 * l_b * k)));
          Ln = L - ds; Wn = W - ds;
          if (Wn < Wmin) { Ln = Ln * (1 - (Wmin - Wn) / Wmin); Wn = Wmin; }
        } else {
          ds = k * 3. * pow(l_b, 2.) / W;
          Ln = L - ds; Wn = W;
        }
        double dA = L * W - Ln * Wn;
        if (!fl_style_fl_bits) {
          fl_new_child(g, b, p, cnt, s, k, l_b, false);
          ncalved++;
        } else {
          double dM_fl_bits = p.rho_bergs * T * dA;
          b.f64[C_MASS_OF_FL_BITS][s] = b.f64[C_MASS_OF_FL_BITS][s] + dM_fl_bits;
          if (area != 0.) g.fl_bits_src[cell] = g.fl_bits_src[cell] + dM_fl_bits / (p.dt * area) * ms;
        }
        if (Ln <= 0 || Wn <= 0) {
          atomicOr(&cnt->error_flags, 2048u);       // 'non-edge element has fully calved from footloose mechanism' I:2648
          return;
        }
        if (p.allow_bergs_to_roll) rolling(p, T, Wn, Ln);
        b.f64[C_THICKNESS][s] = T; b.f64[C_WIDTH][s] = Wn; b.f64[C_LENGTH][s] = Ln;
        b.f64[C_MASS][s] = Ln * Wn * T * p.rho_bergs;
      }
    }
    if (b.f64[C_MASS_OF_FL_BITS][s] * ms > mass_thres) {
      double k = floor(b.f64[C_MASS_OF_FL_BITS][s] * ms / mass_thres);
      fl_new_child(g, b, p, cnt, s, k, l_b, true);
      ncalved++;
      if (area != 0.) g.fl_bits_src[cell] = g.fl_bits_src[cell] - k * mass_thres / (p.dt * area);
    }
  }
  if (ncalved) atomicAdd(&cnt->nbergs_calved_fl, ncalved);
}

// adjust_fl_berg_interactivity I:2765-2841: a child (fl_k = -1) becomes interactive (-2) once no
// other berg is within contact range
__global__ void k_fl_interactivity(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                                   const __grid_constant__ DevParams p, const CellTable ct, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE) || b.f64[C_FL_K][s] != -1.) return;
  int nc_x = p.contact_cells_lon, nc_y = p.contact_cells_lat;
  bool radial = (nc_x == 1 && nc_y == 1);
  double rdenom = p.hexagonal_icebergs ? 1. / (2. * sqrt(3.)) : (p.iceberg_bonds_on ? 1. / 4. : 1. / p.pi);
  double crit_dist = p.contact_distance * p.contact_distance;
  double lat1 = b.f64[C_LAT][s], lon1 = b.f64[C_LON][s];
  double R1 = radial ? sqrt(b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s] * rdenom) : 0.;
  int gi = b.ine[s], gj = b.jne[s];
  bool contact = false;
  for (int j2 = max(gj - nc_y, g.jsd + 1); j2 <= min(gj + nc_y, g.jed) && !contact; j2++)
    for (int i2 = max(gi - nc_x, g.isd + 1); i2 <= min(gi + nc_x, g.ied) && !contact; i2++) {
      int c = gidx(g, i2, j2);
      int n = ct.count[c], o0 = ct.start[c];
      for (int k = 0; k < n; k++) {
        long long o = o0 + k;
        if (b.id[o] == b.id[s]) continue;
        double lat2 = b.f64[C_LAT][o], lon2 = b.f64[C_LON][o], dlon = lon2 - lon1, dlat = lat2 - lat1, r_dist;
        if (radial) {
          double R2 = sqrt(b.f64[C_LENGTH][o] * b.f64[C_WIDTH][o] * rdenom);
          crit_dist = pow(fmax(R1 + R2, p.contact_distance), 2.);
        }
        if (p.grid_is_latlon) {
          double lat_ref = 0.5 * (lat1 + lat2);
          double dx_dlon = p.pi_180 * p.Rearth * cos(lat_ref * p.pi_180), dy_dlat = p.pi_180 * p.Rearth;
          r_dist = pow(dlon * dx_dlon, 2.) + pow(dlat * dy_dlat, 2.);
        } else r_dist = dlon * dlon + dlat * dlat;
        if (r_dist < crit_dist) { contact = true; break; }
      }
    }
  if (!contact) b.f64[C_FL_K][s] = -2.;
}

}  // namespace kid
