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
// HBM-bound free-drift kernel.  Conglomerates that reach several ranks, or the cyclic seam: complete copies per rank, see the end of this file.
#pragma once
#include <cooperative_groups.h>
#include "kid_interact.cuh"

namespace kid {

struct MtsParams {
  double dt_fast, constant_length, constant_width, constant_area, constant_radius;
  double dem_spring_coef, dem_damping_coef, dem_K_damp, poisson, frac_thres_n, frac_thres_t;
  double dem_tests_start_lon, dem_tests_end_lon;
  int32_t force_convergence, explicit_inner_mts, short_step_mts_grounding, radius_based_drag;
  int32_t constant_interaction_LW, use_grounding_torque;
  int32_t ignore_tangential_force, orig_dem_moment_of_inertia, break_bonds_on_sub_steps, fracture_criterion_stress;
  int32_t use_broken_bonds_for_substep_contact, dem_beam_test, no_frac_first_ts, pad0;
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


// the bergs the sub-steps evolve: static_berg < 0.5 and conglom_id /= 0 (I:6756, I:6792, ...)
__device__ __forceinline__ bool mts_active(const DevBergs& b, long long s, uint8_t flags) {
  return (flags & BF_ALIVE) && !(flags & BF_STATIC) && b.conglom_id[s] != 0;
}

// ------------------------------------------------------------------ DEM (dem=.true.)
// calculate_unbonded_same_conglom_dem_force I:807-955
__device__ __noinline__ void dem_unbonded_force(const DevBergs& b, const DevParams& p, const MtsParams& mp, long long s,
                                                long long o, double& IA_x, double& IA_y, double& IAd_x, double& IAd_y,
                                                double u0, double v0, double u1, double v1) {
  if (b.id[s] == b.id[o] || b.f64[C_FL_K][s] == -1. || b.f64[C_FL_K][o] == -1.) return;
  double dlon = b.f64[C_LON_OLD][s] - b.f64[C_LON_OLD][o], dlat = b.f64[C_LAT_OLD][s] - b.f64[C_LAT_OLD][o], dx_dlon, dy_dlat;
  convert_from_grid_to_meters(p, 0.5 * (b.f64[C_LAT_OLD][s] + b.f64[C_LAT_OLD][o]), dx_dlon, dy_dlat);
  double rx = dlon * dx_dlon, ry = dlat * dy_dlat, r2 = (rx * rx) + (ry * ry), R1, R2, M1, M2;
  if (mp.constant_interaction_LW) {
    if ((2 * mp.constant_radius) * (2 * mp.constant_radius) <= r2) return;
    R1 = mp.constant_radius; R2 = R1;
    M1 = mp.constant_area * b.f64[C_THICKNESS][s] * p.rho_bergs; M2 = mp.constant_area * b.f64[C_THICKNESS][o] * p.rho_bergs;
  } else {
    // length and width do not change during the sub-steps: the radii are the same numbers whether cached or not
    if (b.ia_radius) { R1 = b.ia_radius[s]; R2 = b.ia_radius[o]; }
    else { R1 = mts_ia_radius(p, b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s]); R2 = mts_ia_radius(p, b.f64[C_LENGTH][o] * b.f64[C_WIDTH][o]); }
    if ((R1 + R2) * (R1 + R2) <= r2) return;
    M1 = b.f64[C_MASS][s]; M2 = b.f64[C_MASS][o];
  }
  double r_dist = sqrt(r2), u2 = b.f64[C_UVEL_OLD][o], v2 = b.f64[C_VVEL_OLD][o], M_min = fmin(M1, M2), crit_dist = R1 + R2;
  double spring_coef = p.spring_coef, radial = p.radial_damping_coef, tang = p.tangental_damping_coef;
  if (p.critical_interaction_damping_on) {
    radial = 2. * sqrt(spring_coef);
    if (p.tang_crit_int_damp_on) tang = (2. * sqrt(spring_coef)) / 4;
  }
  if (!((r_dist > 0.) && (r_dist < crit_dist))) return;
  double accel_spring = spring_coef * (M_min / M1) * (crit_dist - r_dist);
  IA_x += (accel_spring * (rx / r_dist)); IA_y += (accel_spring * (ry / r_dist));
  double rr = r_dist * r_dist, P_11 = (rx * rx) / rr, P_12 = (rx * ry) / rr, P_22 = (ry * ry) / rr;
  double Pia_11 = 0., Pia_12 = 0., Pia_22 = 0.;
  for (int pass = 0; pass < 2; pass++) {
    double coef = (pass == 0 ? radial : tang) * (M_min / M1);
    if (p.scale_damping_by_pmag) {
      double a1 = ((P_11 * (u2 - u1)) + (P_12 * (v2 - v1))), a2 = ((P_12 * (u2 - u1)) + (P_22 * (v2 - v1)));
      double b1 = ((P_11 * (u2 - u0)) + (P_12 * (v2 - v0))), b2 = ((P_12 * (u2 - u0)) + (P_22 * (v2 - v0)));
      coef = coef * (0.5 * (sqrt((a1 * a1) + (a2 * a2)) + sqrt((b1 * b1) + (b2 * b2))));
    }
    Pia_11 += coef * P_11; Pia_12 += coef * P_12; Pia_22 += coef * P_22;
    P_11 = 1 - P_11; P_12 = -P_12; P_22 = 1 - P_22;
  }
  double du = b.f64[C_UVEL_OLD][o] - b.f64[C_UVEL_OLD][s], dv = b.f64[C_VVEL_OLD][o] - b.f64[C_VVEL_OLD][s];
  IAd_x += Pia_11 * du + Pia_12 * dv;
  IAd_y += Pia_12 * du + Pia_22 * dv;
}

// index of the half-bond of berg o that points back at berg s (current_bond%other_bond, F:5094-5123); -1 if none
__device__ __forceinline__ long long dem_reverse_bond(const DevBergs& b, long long s, long long o) {
  const int64_t my_id = b.id[s];
  for (int q = 0; q < b.max_bonds; q++) {
    long long slot = (long long)q * b.capacity + o;
    if (b.bond_other_id[slot] == my_id) return slot;
  }
  return -1;
}

// calculate_force_dem I:959-1242 for the pair (s, o), evaluated by the berg that comes first in the reference's
// traversal (the lower slot of the cell-sorted store); results are stored on both half-bonds (save_bond_forces F:53)
// and summed per berg afterwards (k_dem_sum).  Returns true when the bond broke in this sub-step.
__device__ __noinline__ bool dem_bond_force(const DevBergs& b, const DevParams& p, const MtsParams& mp, DevCounters* cnt,
                                            long long s, long long o, long long bs /* my half-bond */, double dt) {
  if (b.id[s] == b.id[o] || b.f64[C_FL_K][s] == -1. || b.f64[C_FL_K][o] == -1.) return false;
  const double hexdenom = 1. / (2. * sqrt(3.));
  const double T1 = b.f64[C_THICKNESS][s], T2 = b.f64[C_THICKNESS][o];
  double M1, M2, R1, R2, Rmin, l0, T_Rmin;
  if (mp.constant_interaction_LW) {
    M1 = mp.constant_area * T1 * p.rho_bergs; M2 = mp.constant_area * T2 * p.rho_bergs;
    R1 = mp.constant_radius; R2 = R1; Rmin = R1; l0 = 2 * R1; T_Rmin = T2;
  } else {
    M1 = b.f64[C_MASS][s]; M2 = b.f64[C_MASS][o];
    double A1 = b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s], A2 = b.f64[C_LENGTH][o] * b.f64[C_WIDTH][o];
    if (p.hexagonal_icebergs) { R1 = sqrt(A1 * hexdenom); R2 = sqrt(A2 * hexdenom); } else { R1 = 0.5 * sqrt(A1); R2 = 0.5 * sqrt(A2); }
    if (R1 < R2) { Rmin = R1; T_Rmin = T1; } else { Rmin = R2; T_Rmin = T2; }
    l0 = R1 + R2;
  }
  double dx_dlon, dy_dlat;
  convert_from_grid_to_meters(p, 0.5 * (b.f64[C_LAT_OLD][s] + b.f64[C_LAT_OLD][o]), dx_dlon, dy_dlat);
  double rx = (b.f64[C_LON_OLD][s] - b.f64[C_LON_OLD][o]) * dx_dlon, ry = (b.f64[C_LAT_OLD][s] - b.f64[C_LAT_OLD][o]) * dy_dlat;
  double len = sqrt((rx * rx) + (ry * ry));
  b.bond_length[bs] = len;
  if (len == 0) { atomicOr(&cnt->error_flags, 8192u); return false; }
  const long long bo = dem_reverse_bond(b, s, o);
  double n1 = rx / len, n2 = ry / len, half_delta = 0.5 * (l0 - len);
  double RR1 = R1 - half_delta, RR2 = R2 - half_delta;
  double RR1x = RR1 * n1, RR1y = RR1 * n2, RR2x = RR2 * n1, RR2y = RR2 * n2;
  double L = 2.0 * (Rmin + (Rmin - half_delta) * fabs(R1 - R2) / len);
  double Thick = T_Rmin + (Rmin - half_delta) * fabs(T1 - T2) / len;
  double Fn_x = mp.dem_spring_coef * Thick * 2. * half_delta * L / l0, Fn_y = Fn_x * n2; Fn_x = Fn_x * n1;
  double ur = b.f64[C_UVEL_OLD][s] - b.f64[C_UVEL_OLD][o], vr = b.f64[C_VVEL_OLD][s] - b.f64[C_VVEL_OLD][o];
  const double av1 = b.f64[C_ANG_VEL][s], av2 = b.f64[C_ANG_VEL][o];
  double tangd1 = b.bond_dem[BD_TANGD1][bs], tangd2 = b.bond_dem[BD_TANGD2][bs];
  {
    double tmag = tangd1 * tangd1 + tangd2 * tangd2, tangdotnt = tangd1 * n1 + tangd2 * n2;
    double t1p = tangd1 - tangdotnt * n1, t2p = tangd2 - tangdotnt * n2, tmagp = t1p * t1p + t2p * t2p;
    if (tmagp > 0.) { double t_rat = sqrt(tmag / tmagp); t1p = t_rat * t1p; t2p = t_rat * t2p; } else { t1p = 0.; t2p = 0.; }
    double rotu = RR1y * av1 + RR2y * av2, rotv = -(RR1x * av1 + RR2x * av2);
    double ur2 = ur + rotu, vr2 = vr + rotv, up = ur2 * n1 + vr2 * n2, vp = up * n2; up = up * n1;
    tangd1 = t1p + (ur2 - up) * dt; tangd2 = t2p + (vr2 - vp) * dt;
  }
  double ss_factor = -L * Thick * mp.dem_spring_coef / (l0 * 2.0 * (1.0 + mp.poisson));
  if (mp.ignore_tangential_force) ss_factor = 0.;
  double Fs_x = ss_factor * tangd1, Fs_y = ss_factor * tangd2;
  double sstress = sqrt(Fs_x * Fs_x + Fs_y * Fs_y) / (L * Thick);
  double Ts = -(RR1x * Fs_y - RR1y * Fs_x);
  double rel_rot = b.bond_dem[BD_REL_ROT][bs] + (av1 - av2) * dt;
  double theta, Tr;
  const double drot = b.f64[C_ROT][s] - b.f64[C_ROT][o];
  if (!mp.orig_dem_moment_of_inertia) { theta = sin(drot); Tr = -mp.dem_spring_coef * pow(L, 3.) * Thick * theta / (12. * l0); }
  else { theta = drot; Tr = -(mp.dem_spring_coef / l0) * (2. / 3.) * pow(0.5 * L, 3.) * Thick * theta; }
  double nstress = (mp.dem_spring_coef / l0) * (-2 * half_delta + fabs(theta * 0.5 * L));
  double damping_coef = mp.dem_damping_coef * sqrt(mp.dem_K_damp * M1 * M2 / (M1 + M2));
  double F_x, F_y, Fd_x, Fd_y, T, T_d, T_other;
  bool broke = false;
  if (mp.break_bonds_on_sub_steps && (nstress > mp.frac_thres_n || sstress > mp.frac_thres_t)) {
    broke = true;
    T = 0.; T_d = 0.; T_other = 0.;
    if (nstress < 0) { F_x = Fn_x; F_y = Fn_y; Fd_x = -damping_coef * ur; Fd_y = -damping_coef * vr; }   // sheared under compression
    else { F_x = 0.; F_y = 0.; Fd_x = 0.; Fd_y = 0.; }
    if (b.cand_dirty) *b.cand_dirty = 1;
    b.bond_broken[bs] = 2;          // 2 = broke in this sub-step: its stored forces still count once (I:1158-1196),
    if (bo >= 0) b.bond_broken[bo] = 2;   //     mts_phase_end turns it into 1
    if (mp.use_broken_bonds_for_substep_contact) { atomicSub(&b.n_bonds[s], 1); atomicSub(&b.n_bonds[o], 1); }
  } else {
    F_x = Fn_x + Fs_x; F_y = Fn_y + Fs_y; Fd_x = -damping_coef * ur; Fd_y = -damping_coef * vr;
    T = Ts + Tr; T_d = -damping_coef * (av1 - av2); T_other = Ts - Tr;
  }
  b.bond_dem[BD_TANGD1][bs] = tangd1; b.bond_dem[BD_TANGD2][bs] = tangd2; b.bond_dem[BD_REL_ROT][bs] = rel_rot;
  b.bond_dem[BD_NSTRESS][bs] = nstress; b.bond_dem[BD_SSTRESS][bs] = sstress;
  b.bond_dem[BD_FX][bs] = F_x; b.bond_dem[BD_FY][bs] = F_y; b.bond_dem[BD_FDX][bs] = Fd_x; b.bond_dem[BD_FDY][bs] = Fd_y;
  b.bond_dem[BD_T][bs] = T; b.bond_dem[BD_TD][bs] = T_d;
  if (bo >= 0) {
    b.bond_length[bo] = len;
    b.bond_dem[BD_TANGD1][bo] = -tangd1; b.bond_dem[BD_TANGD2][bo] = -tangd2; b.bond_dem[BD_REL_ROT][bo] = -rel_rot;
    b.bond_dem[BD_NSTRESS][bo] = nstress; b.bond_dem[BD_SSTRESS][bo] = sstress;
    b.bond_dem[BD_FX][bo] = -F_x; b.bond_dem[BD_FY][bo] = -F_y; b.bond_dem[BD_FDX][bo] = -Fd_x; b.bond_dem[BD_FDY][bo] = -Fd_y;
    b.bond_dem[BD_T][bo] = T_other; b.bond_dem[BD_TD][bo] = -T_d;
  }
  return broke;
}

// first half of the explicit DEM sub-step: every intact bond once, by its first berg
__device__ __forceinline__ void dem_pair_phase(const DevBergs& b, const DevParams& p, const MtsParams& mp, DevCounters* cnt,
                                               long long s, double dt) {
  for (int k = b.max_bonds - 1; k >= 0; k--) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] == 0) continue;
    int32_t o = b.bond_other_slot[slot];
    if (o < 0) { atomicOr(&cnt->error_flags, 256u); continue; }
    if (b.bond_broken[slot] != 0) continue;
    // the pair belongs to the berg the reference's sweep reaches first (lower slot); a partner the sweep skips
    // (static) leaves it to this berg
    if ((long long)o < s && mts_active(b, o, b.flags[o])) continue;
    dem_bond_force(b, p, mp, cnt, s, o, slot, dt);
  }
}

// The part of dem_sum_phase's contact search that does not change from sub-step to sub-step (cells are fixed until
// k_mts_finish; conglomerate labels, bond partners and bond counts change only when a bond breaks): the candidates of
// element s in the order the search visits them.
__device__ __forceinline__ void dem_build_candidates(const DevGrid& g, const DevBergs& b, const CellTable& ct, long long s) {
  const int i = b.ine[s], j = b.jne[s], mb = b.max_bonds;
  const int32_t my_cong = b.conglom_id[s];
  uint8_t* cl = b.cand + (size_t)s * b.cand_stride;
  int nc = 0;
  for (int grdj = max(j - 1, g.jsd + 1); grdj <= min(j + 1, g.jed); grdj++)
    for (int grdi = max(i - 1, g.isd + 1); grdi <= min(i + 1, g.ied); grdi++) {
      const int c = gidx(g, grdi, grdj), n = ct.count[c], o0 = ct.start[c];
      for (int k = 0; k < n; k++) {
        const int o = o0 + k;
        if (b.conglom_id[o] != my_cong || !(b.n_bonds[o] < mb)) continue;
        bool is_partner = false;
        for (int q = 0; q < mb; q++) {
          const long long slot = (long long)q * b.capacity + s;
          if (b.bond_other_id[slot] != 0 && b.bond_other_slot[slot] == o) is_partner = true;
        }
        if (!is_partner) cl[nc++] = (uint8_t)o;
      }
    }
  b.cand_n[s] = nc;
}

// second half: the berg's sums (accel_explicit_inner_mts I:1756-1915 with dem)
__device__ __noinline__ void dem_sum_phase(const DevGrid& g, const DevBergs& b, const DevParams& p, const MtsParams& mp,
                                           const CellTable& ct, DevCounters* cnt, long long s, int i, int j, double uvel0,
                                           double vvel0, double dt, double& ax, double& ay, double& axn, double& ayn) {
  double u_star = uvel0 + (axn * (dt / 2.)), v_star = vvel0 + (ayn * (dt / 2.));
  double IA_x = 0., IA_y = 0., IAd_x = 0., IAd_y = 0., F_x = 0., F_y = 0., T = 0., Fd_x = 0., Fd_y = 0., T_d = 0.;
  for (int k = b.max_bonds - 1; k >= 0; k--) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] == 0) continue;
    int32_t o = b.bond_other_slot[slot];
    if (o < 0) continue;
    // a bond broken before this sub-step acts as contact between unbonded elements of one conglomerate (I:1789)
    if (b.bond_broken[slot] == 1) { dem_unbonded_force(b, p, mp, s, o, IA_x, IA_y, IAd_x, IAd_y, uvel0, vvel0, uvel0, vvel0); continue; }
    F_x += b.bond_dem[BD_FX][slot]; F_y += b.bond_dem[BD_FY][slot]; Fd_x += b.bond_dem[BD_FDX][slot]; Fd_y += b.bond_dem[BD_FDY][slot];
    T += b.bond_dem[BD_T][slot]; T_d += b.bond_dem[BD_TD][slot];
  }
  const bool run_contact = !((b.n_bonds[s] == b.max_bonds) || mp.use_broken_bonds_for_substep_contact);
  if (run_contact) {
    // The sweep over the 3x3 cells visits every element of the conglomerate that sits there (all 29 of a beam, every
    // sub-step): the berg's own partners are looked up once, and on Cartesian grids with cached radii the range test
    // of dem_unbonded_force is made here first -- the same arithmetic (x * 1.0 == x), so the same decision.
    const int32_t my_cong = b.conglom_id[s];
    const int mb = b.max_bonds;
    int32_t partner[12];                            // max_bonds <= 12 (kid_init)
#pragma unroll
    for (int q = 0; q < 12; q++) {
      partner[q] = -2;
      if (q < mb) {
        long long slot = (long long)q * b.capacity + s;
        if (b.bond_other_id[slot] != 0) partner[q] = b.bond_other_slot[slot];
      }
    }
    const int32_t* __restrict__ cong = b.conglom_id;
    const int32_t* __restrict__ nbd = b.n_bonds;
    const double* __restrict__ lon_old = b.f64[C_LON_OLD];
    const double* __restrict__ lat_old = b.f64[C_LAT_OLD];
    const double* __restrict__ rad = b.ia_radius;
    const bool quick = !p.grid_is_latlon && rad != nullptr && !mp.constant_interaction_LW;
    const double lon_s = lon_old[s], lat_s = lat_old[s], R_s = quick ? rad[s] : 0.;
    if (b.cand) {
      // the static part of the search (conglomerate, outer layer, partners, cells) was done once: dem_build_candidates
      const int nc = b.cand_n[s];
      const uint8_t* __restrict__ cl = b.cand + (size_t)s * b.cand_stride;
      for (int k = 0; k < nc; k++) {
        const int o = cl[k];
        if (quick) {
          double rx = lon_s - lon_old[o], ry = lat_s - lat_old[o], r2 = (rx * rx) + (ry * ry), RR = R_s + rad[o];
          if (RR * RR <= r2) continue;
        }
        dem_unbonded_force(b, p, mp, s, o, IA_x, IA_y, IAd_x, IAd_y, uvel0, vvel0, uvel0, vvel0);
      }
    } else
    for (int grdj = max(j - 1, g.jsd + 1); grdj <= min(j + 1, g.jed); grdj++)
      for (int grdi = max(i - 1, g.isd + 1); grdi <= min(i + 1, g.ied); grdi++) {
        int c = gidx(g, grdi, grdj);
        int n = ct.count[c];
        int o0 = ct.start[c];
        for (int k = 0; k < n; k++) {
          const int o = o0 + k;
          if (cong[o] != my_cong || !(nbd[o] < mb)) continue;
          bool is_partner = false;
#pragma unroll
          for (int q = 0; q < 12; q++) is_partner |= (partner[q] == o);
          if (is_partner) continue;
          if (quick) {
            double rx = lon_s - lon_old[o], ry = lat_s - lat_old[o], r2 = (rx * rx) + (ry * ry), RR = R_s + rad[o];
            if (RR * RR <= r2) continue;
          }
          dem_unbonded_force(b, p, mp, s, o, IA_x, IA_y, IAd_x, IAd_y, uvel0, vvel0, uvel0, vvel0);
        }
      }
  }
  if (mp.dem_beam_test == 1) {                      // simply supported beam, I:1862-1869
    double sl = b.f64[C_START_LON][s];
    if (sl == mp.dem_tests_start_lon || sl == mp.dem_tests_end_lon) { F_y = 0.0; Fd_y = 0.0; }
    else if (sl == 0.5 * (mp.dem_tests_start_lon + mp.dem_tests_end_lon)) F_y = F_y - 1.5e5;
  } else if (mp.dem_beam_test == 2) {               // cantilever, I:1870-1876
    if (b.f64[C_START_LON][s] == mp.dem_tests_end_lon) F_y = F_y - 1.5e10 / 3.;
  }
  double M, R1;
  if (mp.constant_interaction_LW) {
    M = mp.constant_length * mp.constant_width * b.f64[C_THICKNESS][s] * p.rho_bergs;
    R1 = mts_ia_radius(p, mp.constant_length * mp.constant_width);
  } else { M = b.f64[C_MASS][s]; R1 = mts_ia_radius(p, b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s]); }
  IA_x += F_x / M; IA_y += F_y / M; IAd_x += Fd_x / M; IAd_y += Fd_y / M;
  b.f64[C_ANG_ACCEL][s] = (T + T_d) / (0.5 * M * (R1 * R1));
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

// ---- the sweeps of one fast sub-step as per-berg phases (called from the per-sweep kernels or, for small
// populations, from the single-CTA loop below)

// position update I:6790-6831
__device__ __forceinline__ void mts_phase_pos(const DevBergs& b, const DevParams& p, long long s, double dt) {
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

// velocity update, one pass of I:6846-6934; su* = this berg's terms of the convergence norms
__device__ __forceinline__ void mts_phase_vel(const DevGrid& g, const DevBergs& b, const DevParams& p, const MtsParams& mp,
                                              const CellTable& ct, DevCounters* cnt, long long s, double dt, int jj,
                                              bool iterate, double& su, double& su1, double& su2) {
  const double dt_2 = 0.5 * dt;
  double latn = b.f64[C_LAT][s], lonn = b.f64[C_LON][s];
  double bxn = b.f64[C_BXN_FAST][s], byn = b.f64[C_BYN_FAST][s];
  double axn = b.f64[C_AXN_FAST][s] + bxn, ayn = b.f64[C_AYN_FAST][s] + byn;
  double uvel1 = b.f64[C_UVEL][s], vvel1 = b.f64[C_VVEL][s], ax1, ay1;
  double uvel3 = uvel1 + (dt_2 * axn), vvel3 = vvel1 + (dt_2 * ayn);
  int i = b.ine[s], j = b.jne[s];
  if (mp.explicit_inner_mts) {
    if (p.dem) dem_sum_phase(g, b, p, mp, ct, cnt, s, i, j, uvel1, vvel1, dt, ax1, ay1, axn, ayn);
    else accel_explicit_inner_mts(g, b, p, ct, cnt, s, i, j, uvel1, vvel1, dt, ax1, ay1, axn, ayn);
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

// end of a sub-step I:6974-7041
__device__ __forceinline__ void mts_phase_end(const DevBergs& b, const DevParams& p, const MtsParams& mp, long long s, double dt) {
  b.f64[C_UVEL_OLD][s] = b.f64[C_UVEL][s]; b.f64[C_VVEL_OLD][s] = b.f64[C_VVEL][s];
  if (p.dem) {
    double gdrag = 0.;
    if (mp.use_grounding_torque) {
      double T = b.f64[C_THICKNESS][s], D = (p.rho_bergs / KID_RHO_SEAWATER) * T, groundfrac = mts_ground_fraction(p, b.f64[C_OD][s], D);
      if (groundfrac > 0.0) {
        double MM, R1;
        if (mp.constant_interaction_LW) { MM = mp.constant_length * mp.constant_width * T * p.rho_bergs; R1 = mts_ia_radius(p, mp.constant_length * mp.constant_width); }
        else { MM = b.f64[C_MASS][s]; R1 = mts_ia_radius(p, b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s]); }
        gdrag = -p.cdrag_grounding * groundfrac * p.pi * pow(R1, 2.) / MM;
      }
    }
    double av = b.f64[C_ANG_VEL][s] + dt * b.f64[C_ANG_ACCEL][s];
    av = av / (1. - gdrag * dt);
    b.f64[C_ANG_VEL][s] = av;
    b.f64[C_ROT][s] = b.f64[C_ROT][s] + dt * av;
    for (int k = 0; k < b.max_bonds; k++) {
      long long slot = (long long)k * b.capacity + s;
      if (b.bond_other_id[slot] != 0 && b.bond_broken[slot] == 2) b.bond_broken[slot] = 1;
    }
  }
  if (mp.force_convergence) {
    b.f64[C_AXN][s] = b.f64[C_AXN_FAST][s]; b.f64[C_AYN][s] = b.f64[C_AYN_FAST][s];
    b.f64[C_BXN][s] = b.f64[C_BXN_FAST][s]; b.f64[C_BYN][s] = b.f64[C_BYN_FAST][s];
  }
}

// break_bonds_dem F:4713-4799 for the half-bonds of one berg (the stresses of a pair are stored on both halves)
__device__ __forceinline__ void dem_phase_break(const DevBergs& b, const MtsParams& mp, long long s) {
  if (mp.no_frac_first_ts) return;
  double tn = mp.frac_thres_n, tt = mp.frac_thres_t;
  if (tn <= 0.0 && tt <= 0.0) return;
  if (tn <= 0.0) tn = 1.7976931348623157e308;
  if (tt <= 0.0) tt = 1.7976931348623157e308;
  for (int k = 0; k < b.max_bonds; k++) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] == 0) continue;
    if (b.bond_dem[BD_NSTRESS][slot] > tn || b.bond_dem[BD_SSTRESS][slot] > tt) {
      b.bond_other_id[slot] = 0; b.bond_other_slot[slot] = -1;
      b.n_bonds[s] -= 1;
      if (b.cand_dirty) *b.cand_dirty = 1;
    }
  }
}

__global__ void k_mts_pos(const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p, long long n_slots, double dt) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !mts_active(b, s, b.flags[s])) return;
  mts_phase_pos(b, p, s, dt);
}

__global__ void k_dem_pairs(const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
                            const __grid_constant__ MtsParams mp, DevCounters* __restrict__ cnt, long long n_slots, double dt) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !mts_active(b, s, b.flags[s])) return;
  dem_pair_phase(b, p, mp, cnt, s, dt);
}

__global__ void __launch_bounds__(128)
k_mts_vel(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
          const __grid_constant__ MtsParams mp, const CellTable ct, DevCounters* __restrict__ cnt,
          MtsSums* __restrict__ sums, long long n_slots, double dt, int jj) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double su = 0., su1 = 0., su2 = 0.;
  const bool iterate = mp.force_convergence && !mp.explicit_inner_mts;
  if (s < n_slots && mts_active(b, s, b.flags[s])) mts_phase_vel(g, b, p, mp, ct, cnt, s, dt, jj, iterate, su, su1, su2);
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

__global__ void k_mts_sub_end(const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
                              const __grid_constant__ MtsParams mp, long long n_slots, double dt) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !mts_active(b, s, b.flags[s])) return;
  mts_phase_end(b, p, mp, s, dt);
}

// break_bonds_dem sweeps every berg of the data domain
__global__ void k_dem_break_bonds(const __grid_constant__ DevBergs b, const __grid_constant__ MtsParams mp, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE)) return;
  dem_phase_break(b, mp, s);
}

// The explicit fast scheme needs no convergence passes, so a sub-step is a fixed sequence of sweeps separated by
// grid-wide dependencies (all positions before any force, all pair forces before any sum, all new velocities before
// the next *_old).  When the whole population fits one CTA the sub-step loop runs inside a single kernel with
// __syncthreads() between the sweeps: the bonded-conglomerate cases have 1e1..1e3 elements and 60..1e5 sub-steps,
// launch latency is what they would otherwise cost.
// The berg and bond state of a small population also fits the CTA's shared memory (~0.9-1.1 KB per element with
// 4-6 bonds): `in_smem` stages every column the sweeps touch into shared memory once, runs the sub-steps on a DevBergs
// view whose pointers address that copy (the sweeps are the same functions), and writes the state back at the end.  A
// sweep then costs shared-memory latency instead of dependent L2 round trips.
__host__ __device__ inline size_t mts_smem_bytes(const DevBergs& b, long long n, bool dem) {
  auto r16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t t = 0;
  for (int c = 0; c < C_NCOLS; c++) if (b.f64[c]) t += r16(8 * n);
  t += r16(8 * n) + 2 * r16(4 * n) + r16(n) + 2 * r16(4 * n) + r16(8 * n);           // id, ine, jne, flags, conglom_id, n_bonds, ia_radius
  const size_t nb = (size_t)n * b.max_bonds;
  if (nb) { t += r16(8 * nb) + r16(4 * nb) + r16(8 * nb); if (dem) t += r16(4 * nb) + BD_N * r16(8 * nb); }
  if (dem && n <= 128) t += r16((size_t)n * n) + r16(4 * n) + 16;      // contact candidates (dem_build_candidates)
  return t;
}

// MAXT: the CTA size the instance is compiled for.  128 threads leave the compiler 255 registers per thread (168 used); the
// 1024-thread instance has 64 and spills the fp64-heavy DEM sweeps to local memory (measured: the contact search and
// every call of a sweep function then run out of L1).
template <int MAXT>
__global__ void __launch_bounds__(MAXT)
k_mts_substeps_one_cta(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
                       const __grid_constant__ MtsParams mp, const CellTable ct, DevCounters* __restrict__ cnt,
                       long long n_slots, double dt, int nsub, int in_smem) {
  extern __shared__ __align__(16) unsigned char kid_smem[];
  const int n = (int)n_slots, nt = blockDim.x, tid = threadIdx.x;
  const int per = (n + nt - 1) / nt;
  const bool brk = p.dem && mp.break_bonds_on_sub_steps && !mp.use_broken_bonds_for_substep_contact;
  const bool dem = p.dem != 0;
  DevBergs v = b;
  if (in_smem) {
    size_t off = 0;
    auto carve = [&](size_t bytes) { void* q = kid_smem + off; off += (bytes + 15) & ~(size_t)15; return q; };
    const long long nb = (long long)n * b.max_bonds;
    v.capacity = n;
    for (int c = 0; c < C_NCOLS; c++) if (b.f64[c]) v.f64[c] = (double*)carve(8 * (size_t)n);
    v.id = (int64_t*)carve(8 * (size_t)n); v.ine = (int32_t*)carve(4 * (size_t)n); v.jne = (int32_t*)carve(4 * (size_t)n);
    v.flags = (uint8_t*)carve((size_t)n); v.conglom_id = (int32_t*)carve(4 * (size_t)n); v.n_bonds = (int32_t*)carve(4 * (size_t)n);
    v.ia_radius = (double*)carve(8 * (size_t)n);
    if (nb) {
      v.bond_other_id = (int64_t*)carve(8 * (size_t)nb); v.bond_other_slot = (int32_t*)carve(4 * (size_t)nb);
      v.bond_length = (double*)carve(8 * (size_t)nb);
      if (dem) { v.bond_broken = (int32_t*)carve(4 * (size_t)nb); for (int q = 0; q < BD_N; q++) v.bond_dem[q] = (double*)carve(8 * (size_t)nb); }
    }
    if (dem && n <= 128 && mp.explicit_inner_mts) {
      v.cand = (uint8_t*)carve((size_t)n * n); v.cand_n = (int32_t*)carve(4 * (size_t)n); v.cand_dirty = (int32_t*)carve(16);
      v.cand_stride = n;
    }
    for (int s = tid; s < n; s += nt) {
      for (int c = 0; c < C_NCOLS; c++) if (b.f64[c]) v.f64[c][s] = b.f64[c][s];
      v.id[s] = b.id[s]; v.ine[s] = b.ine[s]; v.jne[s] = b.jne[s]; v.flags[s] = b.flags[s];
      v.conglom_id[s] = b.conglom_id[s]; v.n_bonds[s] = b.n_bonds[s];
      v.ia_radius[s] = mts_ia_radius(p, b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s]);
      for (int k = 0; k < b.max_bonds; k++) {
        const long long gs = (long long)k * b.capacity + s, ls = (long long)k * n + s;
        v.bond_other_id[ls] = b.bond_other_id[gs]; v.bond_other_slot[ls] = b.bond_other_slot[gs]; v.bond_length[ls] = b.bond_length[gs];
        if (dem) { v.bond_broken[ls] = b.bond_broken[gs]; for (int q = 0; q < BD_N; q++) v.bond_dem[q][ls] = b.bond_dem[q][gs]; }
      }
    }
    __syncthreads();
    if (v.cand) {
      if (tid == 0) *v.cand_dirty = 0;
      for (int s = tid; s < n; s += nt) dem_build_candidates(g, v, ct, s);
      __syncthreads();
    }
  }
#define KID_EACH_ACTIVE(body)                                                 \
  for (int q = 0; q < per; q++) {                                             \
    long long s = (long long)q * nt + tid;                                    \
    if (s < n && mts_active(v, s, v.flags[s])) { body; }                      \
  }                                                                           \
  __syncthreads();
  for (int k = 0; k < nsub; k++) {
    KID_EACH_ACTIVE(mts_phase_pos(v, p, s, dt))
    if (dem) { KID_EACH_ACTIVE(dem_pair_phase(v, p, mp, cnt, s, dt)) }
    if (v.cand && *v.cand_dirty) {           // a bond broke in the pair phase: the outer layer the sum phase searches changed
      __syncthreads();
      if (tid == 0) *v.cand_dirty = 0;
      for (int s = tid; s < n; s += nt) dem_build_candidates(g, v, ct, s);
      __syncthreads();
    }
    double su, su1, su2;
    KID_EACH_ACTIVE(mts_phase_vel(g, v, p, mp, ct, cnt, s, dt, 1, false, su, su1, su2))
    KID_EACH_ACTIVE(mts_phase_end(v, p, mp, s, dt))
    if (brk) {
      for (int q = 0; q < per; q++) {
        long long s = (long long)q * nt + tid;
        if (s < n && (v.flags[s] & BF_ALIVE)) dem_phase_break(v, mp, s);
      }
      __syncthreads();
    }
    if (v.cand && *v.cand_dirty) {           // a bond broke in this sub-step: partners / outer layer changed
      __syncthreads();
      if (tid == 0) *v.cand_dirty = 0;
      for (int s = tid; s < n; s += nt) dem_build_candidates(g, v, ct, s);
      __syncthreads();
    }
  }
#undef KID_EACH_ACTIVE
  if (in_smem) {
    for (int s = tid; s < n; s += nt) {
      for (int c = 0; c < C_NCOLS; c++) if (b.f64[c]) b.f64[c][s] = v.f64[c][s];
      b.n_bonds[s] = v.n_bonds[s];
      for (int k = 0; k < b.max_bonds; k++) {
        const long long gs = (long long)k * b.capacity + s, ls = (long long)k * n + s;
        b.bond_other_id[gs] = v.bond_other_id[ls]; b.bond_other_slot[gs] = v.bond_other_slot[ls]; b.bond_length[gs] = v.bond_length[ls];
        if (dem) { b.bond_broken[gs] = v.bond_broken[ls]; for (int q = 0; q < BD_N; q++) b.bond_dem[q][gs] = v.bond_dem[q][ls]; }
      }
    }
  }
}

// one half-bond of the pair phase: dem_pair_phase's body for a single bond entry (the tasks of the cluster kernel)
__device__ __forceinline__ void dem_pair_task(const DevBergs& b, const DevParams& p, const MtsParams& mp, DevCounters* cnt,
                                              long long s, int k, double dt) {
  long long slot = (long long)k * b.capacity + s;
  if (b.bond_other_id[slot] == 0) return;
  int32_t o = b.bond_other_slot[slot];
  if (o < 0) { atomicOr(&cnt->error_flags, 256u); return; }
  if (b.bond_broken[slot] != 0) return;
  if ((long long)o < s && mts_active(b, o, b.flags[o])) return;
  dem_bond_force(b, p, mp, cnt, s, o, slot, dt);
}

// The sub-step loop of a population of 1e2..1e4 elements on ONE THREAD-BLOCK CLUSTER: the sweeps of a sub-step are
// separated by grid-wide dependencies, and a hardware cluster barrier (barrier.cluster, ~0.2 us, release/acquire at
// cluster scope: the state lives in global memory / L2) costs a small fraction of a kernel launch.  Against the
// one-CTA loop: 8 SMs' worth of fp64 lanes and registers (256-thread CTAs: no spills), and the pair phase takes one
// (element, half-bond) task per thread instead of one element with all its bonds -- a bond is evaluated once, by its
// first element, so half the tasks return at once and the critical path is ONE calculate_force_dem, not max_bonds.
// The phase functions are the ones the per-sweep kernels call: same arithmetic, same results.
template <int NT>
__global__ void __launch_bounds__(NT)
k_mts_substeps_cluster(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, const __grid_constant__ DevParams p,
                       const __grid_constant__ MtsParams mp, const CellTable ct, DevCounters* __restrict__ cnt,
                       long long n_slots, double dt, int nsub) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const long long nth = (long long)cl.num_threads(), gid = (long long)cl.thread_rank();
  const long long n = n_slots;
  const bool dem = p.dem != 0;
  const bool brk = dem && mp.break_bonds_on_sub_steps && !mp.use_broken_bonds_for_substep_contact;
  const long long npair = dem ? n * b.max_bonds : 0;
  // The end of sub-step k (and break_bonds_dem) and the position update of sub-step k+1 touch the element's own state
  // only (own columns, own half-bonds): one sweep, no barrier between them -- three cluster barriers per sub-step.
  for (long long s = gid; s < n; s += nth) if (mts_active(b, s, b.flags[s])) mts_phase_pos(b, p, s, dt);
  cl.sync();
  for (int k = 0; k < nsub; k++) {
    if (dem) {
      for (long long t = gid; t < npair; t += nth) {
        const long long s = t % n;
        if (mts_active(b, s, b.flags[s])) dem_pair_task(b, p, mp, cnt, s, (int)(t / n), dt);
      }
      cl.sync();
    }
    for (long long s = gid; s < n; s += nth) {
      double su, su1, su2;
      if (mts_active(b, s, b.flags[s])) mts_phase_vel(g, b, p, mp, ct, cnt, s, dt, 1, false, su, su1, su2);
    }
    cl.sync();
    for (long long s = gid; s < n; s += nth) {
      const uint8_t f = b.flags[s];
      const bool act = mts_active(b, s, f);
      if (act) mts_phase_end(b, p, mp, s, dt);
      if (brk && (f & BF_ALIVE)) dem_phase_break(b, mp, s);
      if (act && k + 1 < nsub) mts_phase_pos(b, p, s, dt);
    }
    cl.sync();
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
    if (route == 1) {       // to another rank: the exchange that follows packs it (send_bergs_to_other_pes F:2997)
      b.ine[s] = i; b.jne[s] = j; b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
      b.flags[s] = flags | BF_LEAVER;
      atomicAdd(&cnt->n_leavers, 1ull);
      unsigned long long k = atomicAdd(b.leaver_count, 1ull);
      if ((long long)k < b.leaver_cap) b.leaver_list[k] = (int32_t)s;
      else atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_CAPACITY);
      return;
    }
    if (route != 0) { b.flags[s] = 0; return; }
  }
  b.ine[s] = i; b.jne[s] = j; b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
  if (i != i0 || j != j0) atomicAdd(&cnt->n_cell_moves, 1ull);
}

// assign_n_bonds F:4617-4637; with use_broken_bonds_for_substep_contact a broken bond was already taken off the
// count when it broke (I:1172, I:1193)
__global__ void k_assign_n_bonds(const __grid_constant__ DevBergs b, long long n_slots, int skip_broken) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  int n = 0;
  for (int k = 0; k < b.max_bonds; k++) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] != 0 && !(skip_broken && b.bond_broken && b.bond_broken[slot] == 1)) n++;
  }
  b.n_bonds[s] = n;
}

// remove_broken_bonds_between_congloms F:2692-2733 (use_broken_bonds_for_substep_contact): a broken bond whose two
// elements ended up in different conglomerates is dropped, on both elements
__global__ void k_dem_drop_split_bonds(const __grid_constant__ DevBergs b, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE)) return;
  for (int k = 0; k < b.max_bonds; k++) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] == 0 || b.bond_broken[slot] != 1) continue;
    int32_t o = b.bond_other_slot[slot];
    if (o < 0 || b.conglom_id[o] == b.conglom_id[s]) continue;
    if (!(b.n_bonds[s] < b.max_bonds || b.n_bonds[o] < b.max_bonds)) continue;
    b.bond_other_id[slot] = 0; b.bond_other_slot[slot] = -1;
  }
}

// dem_tests_init F:4685-4710: the start position becomes the current one (the beam loads key on start_lon)
__global__ void k_dem_tests_init(const __grid_constant__ DevBergs b, long long n_slots) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !(b.flags[s] & BF_ALIVE)) return;
  b.f64[C_START_LON][s] = b.f64[C_LON][s]; b.f64[C_START_LAT][s] = b.f64[C_LAT][s];
}

// ----------------------------------------------------------------------------------------------------------------
// transfer_mts_bergs F:2136-2216 (+ mts_pack_in_dir F:2219, mts_mark_and_pack_halo_and_congloms F:2386,
// mts_pack_contact_bergs F:2457, mts_send_and_receive F:2834, mts_remove_unused_bergs F:2736): after an MTS step
// every rank needs, besides the ordinary halo copies, a COMPLETE copy of every conglomerate that reaches its tile
// (the sub-steps then run with no communication) and the bergs within contact distance of those conglomerates.
// The reference finds them by walking bond lists recursively on the sender, direction by direction, twice.  Here:
//   1. every rank packs ALL its owned bergs (k_mts_pack_all; populations of bonded runs are small) and the packs go
//      to every rank (NVSwitch: one group of sends/receives, no relay);
//   2. the receiver unpacks every periodic image of every record that does not fall on its compute domain
//      (k_mts_unpack_images): in the data domain the copy sits in its cell (halo_berg = 1), beyond it in the nearest
//      halo cell with coordinates unchanged (halo_berg = 2, F:3634-3660);
//   3. after connect_all_bonds and set_conglom_ids (labels spread from the owned bergs through the bonds, so a
//      conglomerate this tile owns a part of is labelled as a whole), copies outside the halo that carry no label are
//      dropped unless they are within contact distance of a labelled berg (k_mts_prune = mts_remove_unused_bergs,
//      always applied: it is what turns "everything" into the reference's set).
// The set that survives is the reference's: true halo copies, whole conglomerates, contact copies (halo_berg = 10).
__global__ void k_mts_owned_flags(const uint8_t* __restrict__ flags, long long n_slots, int32_t* __restrict__ out) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t f = flags[s];
  out[s] = ((f & BF_ALIVE) && !(f & (BF_HALO | BF_LEAVER))) ? 1 : 0;
}

__global__ void k_mts_pack_all(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b, long long n_slots,
                               const int32_t* __restrict__ own, const int32_t* __restrict__ idx, double* __restrict__ sendbuf,
                               const __grid_constant__ RecLayout RL, double Lx /* period, 0 = not cyclic */) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !own[s]) return;
  double* rec = sendbuf + (size_t)idx[s] * RL.w;
  pack_berg(b, s, rec, RL);
  const int i = b.ine[s], j = b.jne[s];
  rec[PK_INE_JNE] = __longlong_as_double(((long long)(unsigned)i << 32) | (unsigned)j);
  rec[PK_YEAR_FLAGS] = __longlong_as_double(((long long)(unsigned)b.start_year[s] << 32) | (unsigned)b.flags[s]);
  // a berg that has just wrapped through the seam on this rank still carries the longitude of the far side
  // (update_latlon F:5128 re-derives it after the transfer): the images are taken around the cell it sits in
  if (Lx > 0.) {
    int gi, gj;
    const double lon = rec[PK_F64_0 + C_LON];
    guess_cell(g, lon, rec[PK_F64_0 + C_LAT], &gi, &gj);
    if (gi - i > g.gni / 2) rec[PK_F64_0 + C_LON] = lon - Lx;
    else if (i - gi > g.gni / 2) rec[PK_F64_0 + C_LON] = lon + Lx;
  }
}

// images k = -1, 0, +1 periods (nimg = 3, cyclic x) or the berg itself (nimg = 1) of every record of `src`
__global__ void k_mts_unpack_images(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                                    const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt,
                                    const double* __restrict__ src, long long n_src, long long s0,
                                    const __grid_constant__ RecLayout RL, int nimg, int skip_k0) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_src * nimg) return;
  const long long r = t / nimg;
  const int k = (nimg == 3) ? (int)(t % nimg) - 1 : 0;
  const long long s = s0 + t;
  b.flags[s] = 0;
  b.halo_code[s] = 0;
  if (k == 0 && skip_k0) return;
  const double* rec = src + (size_t)r * RL.w;
  const double lon = rec[PK_F64_0 + C_LON] + (double)k * p.Lx, lat = rec[PK_F64_0 + C_LAT];
  int gi, gj;
  guess_cell(g, lon, lat, &gi, &gj);
  if (gi >= g.isc && gi <= g.iec && gj >= g.jsc && gj <= g.jec) return;      // an owned berg's own position: never a copy
  unpack_berg(b, s, rec, RL);
  b.f64[C_LON][s] = lon;
  if (b.f64[C_LON_OLD]) b.f64[C_LON_OLD][s] = lon;
  long long yf = __double_as_longlong(rec[PK_YEAR_FLAGS]);
  b.start_year[s] = (int32_t)(yf >> 32);
  const uint8_t f = (uint8_t)(yf & 0xff);
  // the cell of the image: the guessed one or a neighbour of it (never the wide search: is_point_in_cell works
  // modulo Lx and would find the ORIGINAL's cell for an image that lies just beyond the data domain)
  int oi = gi, oj = gj;
  bool found = false;
  for (int dj = 0; dj <= 2 && !found; dj++)
    for (int di = 0; di <= 2 && !found; di++) {
      const int ci = gi + (di == 0 ? 0 : di == 1 ? -1 : 1), cj = gj + (dj == 0 ? 0 : dj == 1 ? -1 : 1);
      if (cell_on_pe(g, ci, cj) && is_point_in_cell(g, p, lon, lat, ci, cj, &cnt->error_flags)) { oi = ci; oj = cj; found = true; }
    }
  if (found && oi >= g.isc && oi <= g.iec && oj >= g.jsc && oj <= g.jec) return;
  double xi = 0.5, yj = 0.5;
  if (found) {
    pos_within_cell(g, p, lon, lat, oi, oj, &xi, &yj, &cnt->error_flags);
    b.halo_code[s] = 1;
  } else {            // beyond the data domain: the nearest halo cell, coordinates unchanged (F:3634-3660)
    oi = min(max(gi, g.isd + 1), g.ied); oj = min(max(gj, g.jsd + 1), g.jed);
    b.halo_code[s] = 2;
  }
  b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
  b.ine[s] = oi; b.jne[s] = oj;
  b.flags[s] = (uint8_t)((f | BF_ALIVE | BF_HALO) & ~(BF_LEAVER | BF_ARRIVAL | BF_COLLIDED));
}

// mts_remove_unused_bergs F:2736-2831: a copy beyond the halo that no conglomerate of this tile claims stays only
// if it is within contact distance of a berg that one does claim (then halo_berg = 10)
__global__ void k_mts_prune(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                            const __grid_constant__ DevParams p, const CellTable ct, long long n_slots, int* __restrict__ n_pruned) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const uint8_t f = b.flags[s];
  if (!(f & BF_ALIVE) || !(f & BF_HALO)) return;
  if (b.halo_code[s] < 2 || b.conglom_id[s] > 0) return;
  const int nc_x = p.contact_cells_lon, nc_y = p.contact_cells_lat;
  const bool radial = (nc_x == 1 && nc_y == 1);
  const double rdenom = p.hexagonal_icebergs ? 1. / (2. * sqrt(3.)) : (p.iceberg_bonds_on ? 1. / 4. : 1. / p.pi);
  const double lat1 = b.f64[C_LAT][s], lon1 = b.f64[C_LON][s];
  const double R1 = radial ? sqrt(b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s] * rdenom) : 0.;
  double crit = p.contact_distance * p.contact_distance;
  const int gi = b.ine[s], gj = b.jne[s];
  bool contact = false;
  for (int j2 = max(gj - nc_y, g.jsd + 1); j2 <= min(gj + nc_y, g.jed) && !contact; j2++)
    for (int i2 = max(gi - nc_x, g.isd + 1); i2 <= min(gi + nc_x, g.ied) && !contact; i2++) {
      const int c = gidx(g, i2, j2);
      const int n = ct.count[c], o0 = ct.start[c];
      for (int q = 0; q < n && !contact; q++) {
        const int o = o0 + q;
        if (!(b.flags[o] & BF_ALIVE) || b.conglom_id[o] <= 0) continue;
        const double lat2 = b.f64[C_LAT][o], lon2 = b.f64[C_LON][o];
        const double dlon = lon2 - lon1, dlat = lat2 - lat1;
        if (radial) {
          const double R2 = sqrt(b.f64[C_LENGTH][o] * b.f64[C_WIDTH][o] * rdenom);
          const double m = fmax(R1 + R2, p.contact_distance);
          crit = m * m;
        }
        double r_dist;
        if (p.grid_is_latlon) {
          const double lat_ref = 0.5 * (lat1 + lat2);
          const double dx_dlon = p.pi_180 * p.Rearth * cos(lat_ref * p.pi_180), dy_dlat = p.pi_180 * p.Rearth;
          r_dist = (dlon * dx_dlon) * (dlon * dx_dlon) + (dlat * dy_dlat) * (dlat * dy_dlat);
        } else r_dist = dlon * dlon + dlat * dlat;
        if (r_dist < crit) contact = true;
      }
    }
  if (contact) b.halo_code[s] = 10;
  else { b.flags[s] = 0; atomicAdd(n_pruned, 1); }
}

}  // namespace kid
