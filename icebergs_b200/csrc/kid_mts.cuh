// kid_mts.cuh -- the multiple-time-step (MTS) velocity Verlet scheme, SURVEY 8(a) rows a15-a16.
//
//   evolve_icebergs_mts        I:6576-7078   the driver: one kernel per sweep over the bergs of the reference
//   accel_mts                  I:1278-1706   long-step forces (part 1) / implicit fast step (part 3)
//   accel_explicit_inner_mts   I:1710-1947   explicit fast step, bonded elements (non-DEM branch)
//   interp_gridded_fields_to_bergs I:4673    the per-berg environment cache (13 columns) the MTS scheme steps with
//
// Every sweep of the reference reads only the *_old state of other bergs and writes its own berg, so a sweep is one
// kernel with one thread per berg; the sweeps of a sub-step are launched back to back on the handle's stream (the
// convergence norms of force_convergence are the only host round trips).  Plain IEEE arithmetic: this is the
// bonded-conglomerate path (hundreds to 1e5 sub-steps of a few thousand elements), launch-latency bound, not the
// HBM-bound free-drift kernel.  One rank; no copies through the cyclic seam (kid_init refuses other layouts).
#pragma once
#include "kid_interact.cuh"

namespace kid {

struct MtsParams {
  double dt_fast, constant_length, constant_width, constant_area, constant_radius;
  int32_t force_convergence, explicit_inner_mts, short_step_mts_grounding, radius_based_drag;
  int32_t constant_interaction_LW, use_grounding_torque, pad0, pad1;
};

// convergence sums of one sweep (usum, usum1, usum2 of I:6600) + had_collision
struct MtsSums { double usum, usum1, usum2; unsigned int had_collision, pad; };

__device__ __forceinline__ double mts_ground_fraction(const DevParams& p, double od, double D) {   // I:1382-1392
  double gf;
  if (p.h_to_init_grounding > 0.0) {
    gf = 1.0 - (od - D) / p.h_to_init_grounding;
    gf = fmax(gf, 0.0); gf = fmin(gf, 1.0);
  } else gf = (D > od) ? 1.0 : 0.0;
  return gf;
}

__device__ __forceinline__ double mts_ia_radius(const DevParams& p, double A) {   // I:684-695
  if (p.hexagonal_icebergs) return sqrt(A / (2. * sqrt(3.)));
  if (p.iceberg_bonds_on) return 0.5 * sqrt(A);
  return sqrt(A / p.pi);
}

__device__ __forceinline__ void mts_speed_ticket(const DevGrid& g, const DevParams& p, DevCounters* cnt, int i, int j,
                                                 double uveln, double vveln, double dt) {
  if ((p.speed_limit > 0.) || (p.speed_limit == -1.)) {
    double speed = sqrt(uveln * uveln + vveln * vveln);
    if (speed > 0.) {
      int c = gidx(g, i, j);
      double loc_dx = fmin(0.5 * (g.dx[c] + g.dx[c - g.nid]), 0.5 * (g.dy[c] + g.dy[c - 1]));
      if (loc_dx / dt * p.speed_limit < speed && p.speed_limit > 0.) atomicAdd(&cnt->nspeeding, 1ull);
    }
  }
}

// quad_interp_from_agrid F:7163-7252 of ocean_depth + ssh (rev_mind = F)
__device__ __noinline__ double mts_quad_interp_od(const DevGrid& g, const DevParams& p, double x, double y, int i, int j,
                                                  double xi, double yj, unsigned int* err) {
  int is, ie, js, je;
  if ((i % 2) == 1) { if (xi >= 0.5) { is = i; ie = i + 2; } else { is = i - 2; ie = i; } } else { is = i - 1; ie = i + 1; }
  if ((j % 2) == 1) { if (yj >= 0.5) { js = j; je = j + 2; } else { js = j - 2; je = j; } } else { js = j - 1; je = j + 1; }
  if (is < g.isd || ie > g.ied || js < g.jsd || je > g.jed) { atomicOr(err, (unsigned)KID_DEVERR_OFF_PE); return 0.; }
  double x1 = g.lonc[gidx(g, is, js)], y1 = g.latc[gidx(g, is, js)];
  double x2 = g.lonc[gidx(g, ie, js)], y2 = g.latc[gidx(g, ie, js)];
  double x3 = g.lonc[gidx(g, ie, je)], y3 = g.latc[gidx(g, ie, je)];
  double x4 = g.lonc[gidx(g, is, je)], y4 = g.latc[gidx(g, is, je)];
  double xloc, yloc;
  if ((!p.grid_is_latlon) && p.grid_is_regular) {
    double dx = fabs(x3 - x4), dy = fabs(y3 - y2);
    x1 = x3 - (dx / 2); y1 = y3 - (dy / 2);
    double Delta_x = amap(x, x1, p.Lx) - x1;
    xloc = ((Delta_x) / dx) + 0.5; yloc = ((y - y1) / dy) + 0.5;
  } else if ((fmax(fmax(y1, y2), fmax(y3, y4)) < 89.999) || (!p.grid_is_latlon)) {
    calc_xiyj(x1, x2, x3, x4, y1, y2, y3, y4, x, y, &xloc, &yloc, p.Lx, err);
  } else {
    double pi_180 = p.pi / 180.;
    double xx = (90. - y) * cos(x * pi_180), yy = (90. - y) * sin(x * pi_180);
    double l1 = g.lon[gidx(g, is, js)], l2 = g.lon[gidx(g, ie, je)], l4 = g.lon[gidx(g, is, je)];
    x1 = (90. - y1) * cos(l1 * pi_180); y1 = (90. - y1) * sin(l1 * pi_180);
    x2 = (90. - y2) * cos(l2 * pi_180); y2 = (90. - y2) * sin(l2 * pi_180);
    x3 = (90. - y3) * cos(l2 * pi_180); y3 = (90. - y3) * sin(l2 * pi_180);
    x4 = (90. - y4) * cos(l4 * pi_180); y4 = (90. - y4) * sin(l4 * pi_180);
    calc_xiyj(x1, x2, x3, x4, y1, y2, y3, y4, xx, yy, &xloc, &yloc, p.Lx, err);
  }
  xloc = xloc * 2 - 1; yloc = yloc * 2 - 1;
  double xb[3], yb[3];
  xb[0] = 0.5 * xloc * (xloc - 1); yb[0] = 0.5 * yloc * (yloc - 1);
  xb[1] = (1 + xloc) * (1 - xloc); yb[1] = (1 + yloc) * (1 - yloc);
  xb[2] = 0.5 * xloc * (xloc + 1); yb[2] = 0.5 * yloc * (yloc + 1);
  double s = 0.;
  for (int bb = 0; bb < 3; bb++)
    for (int a = 0; a < 3; a++) {
      int c = gidx(g, is + a, js + bb);
      s += xb[a] * yb[bb] * (g.ocean_depth[c] + g.ssh[c]);
    }
  return s;
}

// interp_gridded_fields_to_bergs I:4673-4715 (MTS: compute domain only, halo copies keep what they carry)
__global__ void k_mts_env_cache(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                                const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t flags = b.flags[s];
  if (!(flags & BF_ALIVE) || (flags & BF_HALO)) return;
  int i = b.ine[s], j = b.jne[s];
  double xi = b.f64[C_XI][s], yj = b.f64[C_YJ][s];
  Env e;
  if (!interp_flds(g, p, i, j, xi, yj, e)) atomicOr(&cnt->error_flags, 64u);
  e.od = mts_quad_interp_od(g, p, b.f64[C_LON][s], b.f64[C_LAT][s], i, j, xi, yj, &cnt->error_flags);
  b.f64[C_UO][s] = e.uo; b.f64[C_VO][s] = e.vo; b.f64[C_UI][s] = e.ui; b.f64[C_VI][s] = e.vi;
  b.f64[C_UA][s] = e.ua; b.f64[C_VA][s] = e.va; b.f64[C_SSH_X][s] = e.ssh_x; b.f64[C_SSH_Y][s] = e.ssh_y;
  b.f64[C_SST][s] = e.sst; b.f64[C_SSS][s] = e.sss; b.f64[C_CN][s] = e.cn; b.f64[C_HI][s] = e.hi; b.f64[C_OD][s] = e.od;
}

// accel_mts I:1278-1706 for the berg in slot s.  mts_part 1: long step, all forces, collisions between
// conglomerates; mts_part 3: fast step, only_interactive_forces, bonded interactions.
__device__ __noinline__ void accel_mts(const DevGrid& g, const DevBergs& b, const DevParams& p, const MtsParams& mp,
                                       const CellTable& ct, DevCounters* cnt, long long s, int i, int j, double lat,
                                       double uvel0, double vvel0, double dt, int mts_part, double& ax, double& ay,
                                       double& axn, double& ayn, double& bxn, double& byn, double& Fdc_x, double& Fdc_y) {
  const double Cr0 = 0.06, scaling = 0.5;
  const double rho_seawater = KID_RHO_SEAWATER, gravity = KID_GRAVITY;
  double u_star = uvel0 + (axn * (dt / 2.)), v_star = vvel0 + (ayn * (dt / 2.));
  if (mts_part == 1) { u_star = uvel0; v_star = vvel0; }         // uvel0 = berg%uvel at the call (I:1339-1346)
  axn = 0.; ayn = 0.; bxn = 0.; byn = 0.;
  const bool only_ia = (mts_part == 3) || p.only_interactive_forces;
  const bool interactive = p.interactive_icebergs_on;
  double uo = 0, vo = 0, ui = 0, vi = 0, ua = 0, va = 0, ssh_x = 0, ssh_y = 0, f_cori = 0, wave_rad = 0, uwave = 0, vwave = 0;
  double c_ocn = 0, c_atm = 0, c_ice = 0, c_gnd = 0;
  IAcc IA;
  IA.IA_x = IA.IA_y = IA.P11 = IA.P12 = IA.P21 = IA.P22 = IA.Pu_x = IA.Pu_y = 0.;
  if (!only_ia) {
    uo = b.f64[C_UO][s]; vo = b.f64[C_VO][s]; ua = b.f64[C_UA][s]; va = b.f64[C_VA][s]; ui = b.f64[C_UI][s]; vi = b.f64[C_VI][s];
    ssh_x = b.f64[C_SSH_X][s]; ssh_y = b.f64[C_SSH_Y][s];
    double hi = b.f64[C_HI][s], od = b.f64[C_OD][s], pi_180 = p.pi / 180.;
    if (p.grid_is_latlon && !p.use_f_plane) f_cori = (2. * p.omega) * sin(pi_180 * lat);
    else f_cori = (2. * p.omega) * sin(pi_180 * p.lat_ref);
    double M = b.f64[C_MASS][s], T = b.f64[C_THICKNESS][s], D = (p.rho_bergs / rho_seawater) * T, F = T - D;
    double W = b.f64[C_WIDTH][s], L = b.f64[C_LENGTH][s], L2, W2;
    hi = fmin(hi, D);
    double D_hi = fmax(0., D - hi);
    if (p.dem && p.hexagonal_icebergs && mp.radius_based_drag) { L2 = 2. * sqrt(L * W / (2. * sqrt(3.))); W2 = L2; }
    else { L2 = L; W2 = W; }
    double groundfrac = mts_ground_fraction(p, od, D);
    c_gnd = (groundfrac > 0.0) ? (p.cdrag_grounding * W * L * groundfrac) / M : 0.0;
    if (mp.short_step_mts_grounding) c_gnd = 0.;
    uwave = ua - uo; vwave = va - vo;
    double wmod = uwave * uwave + vwave * vwave, ampl = 0.5 * 0.02025 * wmod, Lwavelength = 0.32 * wmod;
    double Lcutoff = 0.125 * Lwavelength, Ltop = 0.25 * Lwavelength;
    double Cr = Cr0 * fmin(fmax(0., (L2 - Lcutoff) / ((Ltop - Lcutoff) + 1.e-30)), 1.);
    wave_rad = 0.5 * rho_seawater / M * Cr * gravity * ampl * fmin(ampl, F) * (2. * W2 * L2) / (W2 + L2);
    wmod = sqrt(ua * ua + va * va);
    if (wmod != 0.) { uwave = ua / wmod; vwave = va / wmod; } else { uwave = 0.; vwave = 0.; wave_rad = 0.; }
    double dragfrac = 1.0;
    if (p.iceberg_bonds_on && p.internal_bergs_for_drag) {
      double N_bonds = 0., N_max = p.hexagonal_icebergs ? 6.0 : 4.0;
      for (int k = 0; k < b.max_bonds; k++) {
        long long slot = (long long)k * b.capacity + s;
        if (b.bond_other_id[slot] != 0 && !(p.dem && b.bond_broken && b.bond_broken[slot] == 1)) N_bonds += 1.0;
      }
      dragfrac = ((N_max - N_bonds) / N_max);
    }
    c_ocn = rho_seawater / M * p.ocean_drag_scale * (0.5 * KID_CD_WV * dragfrac * W2 * (D_hi) + KID_CD_WH * W * L);
    c_atm = KID_RHO_AIR / M * (0.5 * KID_CD_AV * dragfrac * W2 * F + KID_CD_AH * W * L);
    c_ice = (fabs(hi) == 0.) ? 0. : KID_RHO_ICE / M * (0.5 * KID_CD_IV * dragfrac * W2 * hi);
    if (fabs(ui) + fabs(vi) == 0.) c_ice = 0.;
    axn = -gravity * ssh_x + wave_rad * uwave; ayn = -gravity * ssh_y + wave_rad * vwave;
    if (interactive) { interactive_force(g, b, p, ct, s, i, j, IA, uvel0, vvel0, uvel0, vvel0, mts_part); axn += IA.IA_x; ayn += IA.IA_y; }
    axn = axn + f_cori * v_star; ayn = ayn - f_cori * u_star;
  } else if (interactive) interactive_force(g, b, p, ct, s, i, j, IA, uvel0, vvel0, uvel0, vvel0, mts_part);
  double uveln = uvel0, vveln = vvel0;
  double RHS_x = 0, RHS_y = 0, A11 = 1, A12 = 0, A21 = 0, A22 = 1;
  for (int itloop = 1; itloop <= 2; itloop++) {
    double us = (itloop == 2) ? uveln : uvel0, vs = (itloop == 2) ? vveln : vvel0;
    if (only_ia) {
      if (interactive) {
        if (itloop > 1) interactive_force(g, b, p, ct, s, i, j, IA, uvel0, vvel0, us, vs, mts_part);
        RHS_x = (IA.IA_x / 2) - scaling * (((IA.P11 * u_star) + (IA.P12 * v_star)) - IA.Pu_x);
        RHS_y = (IA.IA_y / 2) - scaling * (((IA.P21 * u_star) + (IA.P22 * v_star)) - IA.Pu_y);
        A11 = 1 + (scaling * dt * IA.P11); A22 = 1 + (scaling * dt * IA.P22);
        A12 = (scaling * dt * IA.P12); A21 = (scaling * dt * IA.P21);
      }
    } else {
#define KSQ(x) ((x) * (x))
      double drag_ocn = c_ocn * 0.5 * (sqrt(KSQ(uveln - uo) + KSQ(vveln - vo)) + sqrt(KSQ(uvel0 - uo) + KSQ(vvel0 - vo)));
      double drag_atm = c_atm * 0.5 * (sqrt(KSQ(uveln - ua) + KSQ(vveln - va)) + sqrt(KSQ(uvel0 - ua) + KSQ(vvel0 - va)));
      double drag_ice = c_ice * 0.5 * (sqrt(KSQ(uveln - ui) + KSQ(vveln - vi)) + sqrt(KSQ(uvel0 - ui) + KSQ(vvel0 - vi)));
#undef KSQ
      double drag_gnd = c_gnd;
      RHS_x = (axn / 2) + scaling * (-drag_ocn * (u_star - uo) - drag_atm * (u_star - ua) - drag_ice * (u_star - ui) - drag_gnd * u_star);
      RHS_y = (ayn / 2) + scaling * (-drag_ocn * (v_star - vo) - drag_atm * (v_star - va) - drag_ice * (v_star - vi) - drag_gnd * v_star);
      if (interactive) {
        if (itloop > 1) interactive_force(g, b, p, ct, s, i, j, IA, uvel0, vvel0, us, vs, mts_part);
        RHS_x = RHS_x - scaling * (((IA.P11 * u_star) + (IA.P12 * v_star)) - IA.Pu_x);
        RHS_y = RHS_y - scaling * (((IA.P21 * u_star) + (IA.P22 * v_star)) - IA.Pu_y);
      }
      double lambda = drag_ocn + drag_atm + drag_ice + drag_gnd;
      A11 = 1. + scaling * dt * lambda; A22 = 1. + scaling * dt * lambda;
      A12 = -scaling * dt * f_cori; A21 = scaling * dt * f_cori;
      A12 = A12 / 2.; A21 = A21 / 2.;
      if (interactive) {
        A11 = A11 + (scaling * dt * IA.P11); A22 = A22 + (scaling * dt * IA.P22);
        A12 = A12 + (scaling * dt * IA.P12); A21 = A21 + (scaling * dt * IA.P21);
      }
    }
    double detA = 1. / ((A11 * A22) - (A12 * A21));
    ax = detA * (A22 * RHS_x - A12 * RHS_y); ay = detA * (A11 * RHS_y - A21 * RHS_x);
    uveln = u_star + dt * ax; vveln = v_star + dt * ay;
  }
  if (only_ia) { axn = IA.IA_x; ayn = IA.IA_y; }
  else {
    axn = -gravity * ssh_x + wave_rad * uwave; ayn = -gravity * ssh_y + wave_rad * vwave;
    if (interactive) { axn += IA.IA_x; ayn += IA.IA_y; }
    axn = axn + f_cori * vveln; ayn = ayn - f_cori * uveln;
  }
  bxn = 2 * ax - axn; byn = 2 * ay - ayn;
  if (mts_part == 1) {
    double M = b.f64[C_MASS][s];
    Fdc_x = M * (IA.Pu_x - (IA.P11 * uveln + IA.P12 * vveln));
    Fdc_y = M * (IA.Pu_y - (IA.P21 * uveln + IA.P22 * vveln));
  }
  mts_speed_ticket(g, p, cnt, i, j, uveln, vveln, dt);
  if (p.override_iceberg_velocities) { ax = 0.0; ay = 0.0; axn = 0.0; ayn = 0.0; bxn = 0.0; byn = 0.0; }
}

// accel_explicit_inner_mts I:1710-1947, bonded elements without the DEM forces
__device__ __noinline__ void accel_explicit_inner_mts(const DevGrid& g, const DevBergs& b, const DevParams& p,
                                                      const CellTable& ct, DevCounters* cnt, long long s, int i, int j,
                                                      double uvel0, double vvel0, double dt, double& ax, double& ay,
                                                      double& axn, double& ayn) {
  double u_star = uvel0 + (axn * (dt / 2.)), v_star = vvel0 + (ayn * (dt / 2.));
  double IA_x = 0., IA_y = 0., IAd_x = 0., IAd_y = 0.;
  const double uo_s = b.f64[C_UVEL_OLD][s], vo_s = b.f64[C_VVEL_OLD][s];
  if (p.iceberg_bonds_on) {
    for (int k = b.max_bonds - 1; k >= 0; k--) {          // list order: newest bond first (F:4818)
      long long slot = (long long)k * b.capacity + s;
      if (b.bond_other_id[slot] == 0) continue;
      int32_t o = b.bond_other_slot[slot];
      if (o < 0) { atomicOr(&cnt->error_flags, 256u); continue; }
      IAcc A; A.IA_x = IA_x; A.IA_y = IA_y; A.P11 = A.P12 = A.P21 = A.P22 = A.Pu_x = A.Pu_y = 0.;
      calculate_force(b, p, s, o, A, uvel0, vvel0, uvel0, vvel0, true);
      IA_x = A.IA_x; IA_y = A.IA_y;
      IAd_x += A.P11 * (b.f64[C_UVEL_OLD][o] - uo_s) + A.P12 * (b.f64[C_VVEL_OLD][o] - vo_s);
      IAd_y += A.P12 * (b.f64[C_UVEL_OLD][o] - uo_s) + A.P22 * (b.f64[C_VVEL_OLD][o] - vo_s);
    }
    const int32_t my_cong = b.conglom_id[s];
    for (int grdj = max(j - 1, g.jsd + 1); grdj <= min(j + 1, g.jed); grdj++)
      for (int grdi = max(i - 1, g.isd + 1); grdi <= min(i + 1, g.ied); grdi++) {
        int c = gidx(g, grdi, grdj);
        int n = ct.count[c];
        long long o0 = ct.start[c];
        for (int k = 0; k < n; k++) {
          long long o = o0 + k;
          if (b.conglom_id[o] != my_cong || !(b.n_bonds[o] < b.max_bonds)) continue;
          bool partner = false;
          for (int q = 0; q < b.max_bonds; q++) {
            long long slot = (long long)q * b.capacity + s;
            if (b.bond_other_id[slot] != 0 && b.bond_other_slot[slot] == (int32_t)o) partner = true;
          }
          if (partner) continue;
          IAcc A; A.IA_x = IA_x; A.IA_y = IA_y; A.P11 = A.P12 = A.P21 = A.P22 = A.Pu_x = A.Pu_y = 0.;
          calculate_force(b, p, s, o, A, uvel0, vvel0, uvel0, vvel0, false, 1);
          IA_x = A.IA_x; IA_y = A.IA_y;
          IAd_x += A.P11 * (b.f64[C_UVEL_OLD][o] - uo_s) + A.P12 * (b.f64[C_VVEL_OLD][o] - vo_s);
          IAd_y += A.P12 * (b.f64[C_UVEL_OLD][o] - uo_s) + A.P22 * (b.f64[C_VVEL_OLD][o] - vo_s);
        }
      }
  }
  axn = IA_x + IAd_x; ayn = IA_y + IAd_y;
  ax = 0.5 * (axn + 0.0); ay = 0.5 * (ayn + 0.0);
  double uveln = u_star + dt * ax, vveln = v_star + dt * ay;
  mts_speed_ticket(g, p, cnt, i, j, uveln, vveln, dt);
  if (p.override_iceberg_velocities) { ax = 0.0; ay = 0.0; axn = 0.0; ayn = 0.0; }
}

__device__ __forceinline__ void mts_block_sums(MtsSums* out, double a, double b_, double c) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { a += shfl_down_d(a, d); b_ += shfl_down_d(b_, d); c += shfl_down_d(c, d); }
  if ((threadIdx.x & 31) == 0) {
    if (a != 0.) atomicAdd(&out->usum, a);
    if (b_ != 0.) atomicAdd(&out->usum1, b_);
    if (c != 0.) atomicAdd(&out->usum2, c);
  }
}

// the bergs the sub-steps evolve: static_berg < 0.5 and conglom_id /= 0 (I:6756, I:6792, ...)
__device__ __forceinline__ bool mts_active(const DevBergs& b, long long s, uint8_t flags) {
  return (flags & BF_ALIVE) && !(flags & BF_STATIC) && b.conglom_id[s] != 0;
}

// part 1, one pass of the convergence loop I:6660-6706
__global__ void __launch_bounds__(128)
k_mts_part1(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
            const __grid_constant__ MtsParams mp, const CellTable ct, DevCounters* __restrict__ cnt,
            MtsSums* __restrict__ sums, long long n_slots, int ii) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double su = 0., su1 = 0., su2 = 0.;
  uint8_t flags = (s < n_slots) ? b.flags[s] : (uint8_t)0;
  const bool fc = mp.force_convergence;
  bool act = (flags & BF_ALIVE) && !(flags & BF_STATIC) && (fc || b.conglom_id[s] != 0) && (ii == 1 || (flags & BF_COLLIDED));
  if (act) {
    const double dt = p.dt;
    double uvel = b.f64[C_UVEL][s], vvel = b.f64[C_VVEL][s];
    double ax1, ay1, axn = 0., ayn = 0., bxn = 0., byn = 0., Fdc_x = 0., Fdc_y = 0.;
    accel_mts(g, b, p, mp, ct, cnt, s, b.ine[s], b.jne[s], b.f64[C_LAT][s], uvel, vvel, dt, 1, ax1, ay1, axn, ayn, bxn, byn, Fdc_x, Fdc_y);
    if (Fdc_x != 0. || (Fdc_y != 0. && fc)) { b.flags[s] = flags | BF_COLLIDED; sums->had_collision = 1u; }
    b.f64[C_AXN][s] = axn; b.f64[C_AYN][s] = ayn; b.f64[C_BXN][s] = bxn; b.f64[C_BYN][s] = byn;
    if (fc) {
      double up = uvel + (dt * ax1), vp = vvel + (dt * ay1), uold = b.f64[C_UVEL_OLD][s], vold = b.f64[C_VVEL_OLD][s];
      b.f64[C_UVEL_PREV][s] = up; b.f64[C_VVEL_PREV][s] = vp;
      if (ii == 1) su = uold * uold + vold * vold;
      su1 = up * up + vp * vp;
      su2 = (up - uold) * (up - uold) + (vp - vold) * (vp - vold);
    } else {
      uvel = uvel + (dt * ax1); vvel = vvel + (dt * ay1);
      b.f64[C_UVEL][s] = uvel; b.f64[C_VVEL][s] = vvel;
      b.f64[C_UVEL_PREV][s] = uvel; b.f64[C_VVEL_PREV][s] = vvel;
    }
  }
  if (fc) mts_block_sums(sums, su, su1, su2);
}

// I:6709-6720: the velocities other bergs damp against follow the iterate
__global__ void k_mts_part1_old(const __grid_constant__ DevBergs b, long long n_slots, int last_iter) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t flags = b.flags[s];
  if (!(flags & BF_ALIVE) || (flags & BF_STATIC)) return;
  b.f64[C_UVEL_OLD][s] = b.f64[C_UVEL_PREV][s]; b.f64[C_VVEL_OLD][s] = b.f64[C_VVEL_PREV][s];
  if (last_iter && (flags & BF_COLLIDED)) b.flags[s] = flags & ~BF_COLLIDED;
}

// part 2 I:6744-6764
__global__ void k_mts_part2(const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p, long long n_slots, int fc) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  if (!mts_active(b, s, b.flags[s])) return;
  const double dt_2 = 0.5 * p.dt;
  double u = b.f64[C_UVEL_PREV][s], v = b.f64[C_VVEL_PREV][s];
  u = u + dt_2 * (b.f64[C_AXN][s] + b.f64[C_BXN][s]); v = v + dt_2 * (b.f64[C_AYN][s] + b.f64[C_BYN][s]);
  b.f64[C_UVEL][s] = u; b.f64[C_VVEL][s] = v; b.f64[C_UVEL_OLD][s] = u; b.f64[C_VVEL_OLD][s] = v;
  if (fc) {
    b.f64[C_AXN][s] = b.f64[C_AXN_FAST][s]; b.f64[C_AYN][s] = b.f64[C_AYN_FAST][s];
    b.f64[C_BXN][s] = b.f64[C_BXN_FAST][s]; b.f64[C_BYN][s] = b.f64[C_BYN_FAST][s];
  }
}

// sub-step position update I:6790-6831
__global__ void k_mts_pos(const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p, long long n_slots, double dt) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  if (!mts_active(b, s, b.flags[s])) return;
  const double dt_2 = 0.5 * dt;
  double lon1 = b.f64[C_LON][s], lat1 = b.f64[C_LAT][s], uvel1 = b.f64[C_UVEL][s], vvel1 = b.f64[C_VVEL][s];
  double axn = b.f64[C_AXN_FAST][s], ayn = b.f64[C_AYN_FAST][s], bxn = b.f64[C_BXN_FAST][s], byn = b.f64[C_BYN_FAST][s];
  const bool tang = (lat1 > 89.) && p.grid_is_latlon;
  double dxdl1, dydl, lonn, latn;
  convert_from_meters_to_grid(p, lat1, dxdl1, dydl);
  double uvel2 = uvel1 + (dt_2 * axn) + (dt_2 * bxn), vvel2 = vvel1 + (dt_2 * ayn) + (dt_2 * byn);
  if (tang) {
    double x1, y1, xdot2, ydot2;
    rotpos_to_tang(p, lon1, lat1, x1, y1);
    rotvec_to_tang(p, lon1, uvel2, vvel2, xdot2, ydot2);
    rotpos_from_tang(p, x1 + (dt * xdot2), y1 + (dt * ydot2), lonn, latn);
  } else { lonn = lon1 + (dt * (uvel2 * dxdl1)); latn = lat1 + (dt * (vvel2 * dydl)); }
  b.f64[C_LON][s] = lonn; b.f64[C_LAT][s] = latn; b.f64[C_LON_OLD][s] = lonn; b.f64[C_LAT_OLD][s] = latn;
  b.f64[C_UVEL_OLD][s] = uvel1 + dt_2 * (axn + bxn);
  b.f64[C_VVEL_OLD][s] = vvel1 + dt_2 * (ayn + bxn);          // sic: bxn_fast, I:6827
}

// sub-step velocity update, one pass of I:6846-6934
__global__ void __launch_bounds__(128)
k_mts_vel(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
          const __grid_constant__ MtsParams mp, const CellTable ct, DevCounters* __restrict__ cnt,
          MtsSums* __restrict__ sums, long long n_slots, double dt, int jj) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double su = 0., su1 = 0., su2 = 0.;
  uint8_t flags = (s < n_slots) ? b.flags[s] : (uint8_t)0;
  const bool iterate = mp.force_convergence && !mp.explicit_inner_mts;
  if (mts_active(b, (s < n_slots) ? s : 0, flags)) {
    const double dt_2 = 0.5 * dt;
    double latn = b.f64[C_LAT][s], lonn = b.f64[C_LON][s];
    double bxn = b.f64[C_BXN_FAST][s], byn = b.f64[C_BYN_FAST][s];
    double axn = b.f64[C_AXN_FAST][s] + bxn, ayn = b.f64[C_AYN_FAST][s] + byn;
    double uvel1 = b.f64[C_UVEL][s], vvel1 = b.f64[C_VVEL][s], ax1, ay1;
    double uvel3 = uvel1 + (dt_2 * axn), vvel3 = vvel1 + (dt_2 * ayn);
    int i = b.ine[s], j = b.jne[s];
    if (mp.explicit_inner_mts) {
      accel_explicit_inner_mts(g, b, p, ct, cnt, s, i, j, uvel1, vvel1, dt, ax1, ay1, axn, ayn);
      bxn = 0.; byn = 0.;
      if (mp.short_step_mts_grounding) {
        double T = b.f64[C_THICKNESS][s], D = (p.rho_bergs / KID_RHO_SEAWATER) * T;
        double groundfrac = mts_ground_fraction(p, b.f64[C_OD][s], D), gdrag = 0.;
        if (groundfrac > 0.0) {
          double MM, AA;
          if (mp.constant_interaction_LW) { MM = mp.constant_length * mp.constant_width * T * p.rho_bergs; AA = mp.constant_width * mp.constant_length; }
          else { MM = b.f64[C_MASS][s]; AA = b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s]; }
          gdrag = -p.cdrag_grounding * groundfrac * AA / MM;
        }
        axn = axn + uvel1 * gdrag; ayn = ayn + vvel1 * gdrag;
        ax1 = 0.5 * axn; ay1 = 0.5 * ayn;
      }
    } else {
      double f1, f2;
      accel_mts(g, b, p, mp, ct, cnt, s, i, j, latn, uvel1, vvel1, dt, 3, ax1, ay1, axn, ayn, bxn, byn, f1, f2);
    }
    double uveln, vveln;
    if ((latn > 89.) && p.grid_is_latlon) {
      double xdot3, ydot3, xddot1, yddot1;
      rotvec_to_tang(p, lonn, uvel3, vvel3, xdot3, ydot3);
      rotvec_to_tang(p, lonn, ax1, ay1, xddot1, yddot1);
      rotvec_from_tang(p, lonn, xdot3 + (dt * xddot1), ydot3 + (dt * yddot1), uveln, vveln);
    } else { uveln = uvel3 + (dt * ax1); vveln = vvel3 + (dt * ay1); }
    if (iterate) {
      double uold = b.f64[C_UVEL_OLD][s], vold = b.f64[C_VVEL_OLD][s];
      if (jj == 1) su = uold * uold + vold * vold;
      su1 = uveln * uveln + vveln * vveln;
      su2 = (uveln - uold) * (uveln - uold) + (vveln - vold) * (vveln - vold);
    }
    b.f64[C_AXN_FAST][s] = axn; b.f64[C_AYN_FAST][s] = ayn; b.f64[C_BXN_FAST][s] = bxn; b.f64[C_BYN_FAST][s] = byn;
    b.f64[C_UVEL][s] = uveln; b.f64[C_VVEL][s] = vveln;
  }
  if (iterate) mts_block_sums(sums, su, su1, su2);
}

// the pass did not converge: step back and iterate, I:6951-6968
__global__ void k_mts_vel_retry(const __grid_constant__ DevBergs b, long long n_slots, double dt) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  if (!mts_active(b, s, b.flags[s])) return;
  const double dt_2 = 0.5 * dt;
  double u = b.f64[C_UVEL][s], v = b.f64[C_VVEL][s];
  b.f64[C_UVEL_OLD][s] = u; b.f64[C_VVEL_OLD][s] = v;
  double axn = b.f64[C_AXN][s], ayn = b.f64[C_AYN][s], bxn = b.f64[C_BXN][s], byn = b.f64[C_BYN][s];
  b.f64[C_UVEL][s] = u - dt_2 * (b.f64[C_AXN_FAST][s] + b.f64[C_BXN_FAST][s]) - dt_2 * (axn + bxn);
  b.f64[C_VVEL][s] = v - dt_2 * (b.f64[C_AYN_FAST][s] + b.f64[C_BYN_FAST][s]) - dt_2 * (ayn + byn);
  b.f64[C_AXN_FAST][s] = axn; b.f64[C_AYN_FAST][s] = ayn; b.f64[C_BXN_FAST][s] = bxn; b.f64[C_BYN_FAST][s] = byn;
}

// end of a sub-step I:6974-7041 (without the DEM rotation)
__global__ void k_mts_sub_end(const __grid_constant__ DevBergs b, long long n_slots, int fc) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  if (!mts_active(b, s, b.flags[s])) return;
  b.f64[C_UVEL_OLD][s] = b.f64[C_UVEL][s]; b.f64[C_VVEL_OLD][s] = b.f64[C_VVEL][s];
  if (fc) {
    b.f64[C_AXN][s] = b.f64[C_AXN_FAST][s]; b.f64[C_AYN][s] = b.f64[C_AYN_FAST][s];
    b.f64[C_BXN][s] = b.f64[C_BXN_FAST][s]; b.f64[C_BYN][s] = b.f64[C_BYN_FAST][s];
  }
}

// explicit fast scheme without convergence passes: position + velocity + end-of-sub-step of ONE sub-step need two
// grid-wide dependencies (positions of all bergs before any force, *_old of all bergs before the next positions), so a
// sub-step is k_mts_pos, k_mts_vel, k_mts_sub_end; when the whole population fits one CTA the sub-step loop runs
// inside a single kernel with __syncthreads() between the sweeps (the bonded-conglomerate tests have 1e1..1e3 elements
// and 60..1e5 sub-steps: launch latency, not bandwidth, is what they cost).
__global__ void __launch_bounds__(1024)
k_mts_substeps_one_cta(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
                       const __grid_constant__ MtsParams mp, const CellTable ct, DevCounters* __restrict__ cnt,
                       long long n_slots, double dt, int nsub) {
  const double dt_2 = 0.5 * dt;
  const int per = (int)((n_slots + blockDim.x - 1) / blockDim.x);
  for (int k = 0; k < nsub; k++) {
    for (int q = 0; q < per; q++) {
      long long s = (long long)q * blockDim.x + threadIdx.x;
      if (s >= n_slots || !mts_active(b, s, b.flags[s])) continue;
      double lon1 = b.f64[C_LON][s], lat1 = b.f64[C_LAT][s], uvel1 = b.f64[C_UVEL][s], vvel1 = b.f64[C_VVEL][s];
      double axn = b.f64[C_AXN_FAST][s], ayn = b.f64[C_AYN_FAST][s], bxn = b.f64[C_BXN_FAST][s], byn = b.f64[C_BYN_FAST][s];
      const bool tang = (lat1 > 89.) && p.grid_is_latlon;
      double dxdl1, dydl, lonn, latn;
      convert_from_meters_to_grid(p, lat1, dxdl1, dydl);
      double uvel2 = uvel1 + (dt_2 * axn) + (dt_2 * bxn), vvel2 = vvel1 + (dt_2 * ayn) + (dt_2 * byn);
      if (tang) {
        double x1, y1, xdot2, ydot2;
        rotpos_to_tang(p, lon1, lat1, x1, y1);
        rotvec_to_tang(p, lon1, uvel2, vvel2, xdot2, ydot2);
        rotpos_from_tang(p, x1 + (dt * xdot2), y1 + (dt * ydot2), lonn, latn);
      } else { lonn = lon1 + (dt * (uvel2 * dxdl1)); latn = lat1 + (dt * (vvel2 * dydl)); }
      b.f64[C_LON][s] = lonn; b.f64[C_LAT][s] = latn; b.f64[C_LON_OLD][s] = lonn; b.f64[C_LAT_OLD][s] = latn;
      b.f64[C_UVEL_OLD][s] = uvel1 + dt_2 * (axn + bxn);
      b.f64[C_VVEL_OLD][s] = vvel1 + dt_2 * (ayn + bxn);
    }
    __syncthreads();
    for (int q = 0; q < per; q++) {
      long long s = (long long)q * blockDim.x + threadIdx.x;
      if (s >= n_slots || !mts_active(b, s, b.flags[s])) continue;
      double latn = b.f64[C_LAT][s], lonn = b.f64[C_LON][s];
      double axn = b.f64[C_AXN_FAST][s] + b.f64[C_BXN_FAST][s], ayn = b.f64[C_AYN_FAST][s] + b.f64[C_BYN_FAST][s];
      double uvel1 = b.f64[C_UVEL][s], vvel1 = b.f64[C_VVEL][s], ax1, ay1;
      double uvel3 = uvel1 + (dt_2 * axn), vvel3 = vvel1 + (dt_2 * ayn);
      accel_explicit_inner_mts(g, b, p, ct, cnt, s, b.ine[s], b.jne[s], uvel1, vvel1, dt, ax1, ay1, axn, ayn);
      double uveln, vveln;
      if ((latn > 89.) && p.grid_is_latlon) {
        double xdot3, ydot3, xddot1, yddot1;
        rotvec_to_tang(p, lonn, uvel3, vvel3, xdot3, ydot3);
        rotvec_to_tang(p, lonn, ax1, ay1, xddot1, yddot1);
        rotvec_from_tang(p, lonn, xdot3 + (dt * xddot1), ydot3 + (dt * yddot1), uveln, vveln);
      } else { uveln = uvel3 + (dt * ax1); vveln = vvel3 + (dt * ay1); }
      // the new velocity is parked in *_PREV until every thread has read the old *_OLD of its neighbours
      b.f64[C_AXN_FAST][s] = axn; b.f64[C_AYN_FAST][s] = ayn; b.f64[C_BXN_FAST][s] = 0.; b.f64[C_BYN_FAST][s] = 0.;
      b.f64[C_UVEL][s] = uveln; b.f64[C_VVEL][s] = vveln;
    }
    __syncthreads();
    for (int q = 0; q < per; q++) {
      long long s = (long long)q * blockDim.x + threadIdx.x;
      if (s >= n_slots || !mts_active(b, s, b.flags[s])) continue;
      b.f64[C_UVEL_OLD][s] = b.f64[C_UVEL][s]; b.f64[C_VVEL_OLD][s] = b.f64[C_VVEL][s];
      if (mp.force_convergence) {
        b.f64[C_AXN][s] = b.f64[C_AXN_FAST][s]; b.f64[C_AYN][s] = b.f64[C_AYN_FAST][s];
        b.f64[C_BXN][s] = 0.; b.f64[C_BYN][s] = 0.;
      }
    }
    __syncthreads();
  }
}

// end of evolve_icebergs_mts I:7050-7075: the cell of the new position, grounding
__global__ void k_mts_finish(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                             const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t flags = b.flags[s];
  if (!(flags & BF_ALIVE) || (flags & (BF_STATIC | BF_HALO))) return;
  b.f64[C_UVEL_OLD][s] = b.f64[C_UVEL][s]; b.f64[C_VVEL_OLD][s] = b.f64[C_VVEL][s];
  double lonn = b.f64[C_LON][s], latn = b.f64[C_LAT][s], xi = b.f64[C_XI][s], yj = b.f64[C_YJ][s];
  int i = b.ine[s], j = b.jne[s];
  const int i0 = i, j0 = j;
  if (adjust_index_and_ground(g, p, lonn, latn, i, j, xi, yj, &cnt->error_flags, &cnt->warn_adjust)) atomicAdd(&cnt->n_bounced, 1ull);
  b.f64[C_LON][s] = lonn; b.f64[C_LAT][s] = latn; b.f64[C_LON_OLD][s] = lonn; b.f64[C_LAT_OLD][s] = latn;
  // send_bergs_to_other_pes F:2997 on one rank: through the cyclic seam back into the tile, or out of the model
  if (i > g.iec || i < g.isc || j > g.jec || j < g.jsc) {
    int route = route_berg(g, p, lonn, latn, i, j, xi, yj, &cnt->error_flags, &cnt->n_wrapped);
    if (route != 0) { b.flags[s] = 0; return; }
  }
  b.ine[s] = i; b.jne[s] = j; b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
  if (i != i0 || j != j0) atomicAdd(&cnt->n_cell_moves, 1ull);
}

// assign_n_bonds F:4617-4637
__global__ void k_assign_n_bonds(const __grid_constant__ DevBergs b, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  int n = 0;
  for (int k = 0; k < b.max_bonds; k++) if (b.bond_other_id[(long long)k * b.capacity + s] != 0) n++;
  b.n_bonds[s] = n;
}

}  // namespace kid
