// kid_physics.cuh -- per-berg physics on the device.
//
// Follows the reference routine by routine (I: = src/icebergs.F90):
//   interp_flds I:4718 (+ ddx_ssh/ddy_ssh I:4903/4916 precomputed per cell, rotate I:4953),
//   accel I:1950, verlet_stepping I:7203, update_verlet_position I:7684,
//   adjust_index_and_ground I:7819, tangent-plane helpers I:7767-7816 / I:8066,
//   thermodynamics I:2844 (+ rolling I:3307, fl_bits_dimensions I:3370).
// Expression order is the reference's (see SURVEY.md appendix A for the quirks
// that are reproduced on purpose).
#pragma once
#include "kid_geom.cuh"

namespace kid {

struct Env { double uo, vo, ui, vi, ua, va, ssh_x, ssh_y, sst, sss, cn, hi, od; };

// accumulated interaction terms of interactive_force (I:480): IA_x, IA_y, P_ia_*, P_ia_times_u_*
struct IAcc { double IA_x, IA_y, P11, P12, P21, P22, Pu_x, Pu_y; };

// F:7071-7088 on one component of the four corner records
#define KID_BILIN(f) (p.old_bug_bilin \
    ? ((c3.f * (1. - xi) + c4.f * xi) * (1. - yj) + (c2.f * (1. - xi) + c1.f * xi) * yj) \
    : ((c3.f * xi + c4.f * (1. - xi)) * yj + (c2.f * xi + c1.f * (1. - xi)) * (1. - yj)))

__device__ __forceinline__ void rotate(double& u, double& v, double cos_rot, double sin_rot) {
  double u_old = u, v_old = v;
  u = cos_rot * u_old + sin_rot * v_old;
  v = cos_rot * v_old - sin_rot * u_old;
}

// I:4718-4900 (non-MTS ocean depth, I:4897).  Returns false when a NaN survived.
__device__ __forceinline__ bool interp_flds(const DevGrid& g, const DevParams& p, int i, int j, double xi,
                                            double yj, Env& e) {
  const CornerRec* __restrict__ cr = g.corner;
  const CellRec* __restrict__ ce = g.cell;
  size_t ne = gidx(g, i, j);
  size_t nid = (size_t)g.nid;
  const CornerRec c3 = cr[ne], c4 = cr[ne - 1], c2 = cr[ne - nid], c1 = cr[ne - nid - 1];
  double cos_rot = KID_BILIN(cosr);
  double sin_rot = KID_BILIN(sinr);
  double uo = KID_BILIN(uo), vo = KID_BILIN(vo);
  double ui = KID_BILIN(ui), vi = KID_BILIN(vi);
  double ua = KID_BILIN(ua), va = KID_BILIN(va);
  if (p.coastal_drift > 0.) {
    const double* __restrict__ msk = g.msk;
    double cd = p.coastal_drift;
    double m0 = msk[ne], mE = msk[ne + 1], mW = msk[ne - 1], mN = msk[ne + nid], mS = msk[ne - nid];
    uo = uo + cd * (mE - mW) * m0;
    ui = ui + cd * (mE - mW) * m0;
    vo = vo + cd * (mN - mS) * m0;
    vi = vi + cd * (mN - mS) * m0;
  }
  const CellRec c0 = ce[ne];
  e.sst = c0.sst; e.sss = c0.sss; e.cn = c0.cn; e.hi = c0.hi; e.od = c0.od;
  double hxp, hxm;
  if (yj >= 0.5) {
    hxp = (yj - 0.5) * ce[ne + nid].ddx + (1.5 - yj) * c0.ddx;
    hxm = (yj - 0.5) * ce[ne + nid - 1].ddx + (1.5 - yj) * ce[ne - 1].ddx;
  } else {
    hxp = (yj + 0.5) * c0.ddx + (0.5 - yj) * ce[ne - nid].ddx;
    hxm = (yj + 0.5) * ce[ne - 1].ddx + (0.5 - yj) * ce[ne - nid - 1].ddx;
  }
  double ssh_x = xi * hxp + (1. - xi) * hxm;
  if (xi >= 0.5) {
    hxp = (xi - 0.5) * ce[ne + 1].ddy + (1.5 - xi) * c0.ddy;
    hxm = (xi - 0.5) * ce[ne - nid + 1].ddy + (1.5 - xi) * ce[ne - nid].ddy;
  } else {
    hxp = (xi + 0.5) * c0.ddy + (0.5 - xi) * ce[ne - 1].ddy;
    hxm = (xi + 0.5) * ce[ne - nid].ddy + (0.5 - xi) * ce[ne - nid - 1].ddy;
  }
  double ssh_y = yj * hxp + (1. - yj) * hxm;
  rotate(uo, vo, cos_rot, sin_rot);
  rotate(ui, vi, cos_rot, sin_rot);
  rotate(ua, va, cos_rot, sin_rot);
  rotate(ssh_x, ssh_y, cos_rot, sin_rot);
  if (ssh_x != ssh_x) ssh_x = 0.;
  if (ssh_y != ssh_y) ssh_y = 0.;
  e.uo = uo; e.vo = vo; e.ui = ui; e.vi = vi; e.ua = ua; e.va = va; e.ssh_x = ssh_x; e.ssh_y = ssh_y;
  bool bad = (uo != uo) || (vo != vo) || (ui != ui) || (vi != vi) || (ua != ua) || (va != va) ||
             (e.sst != e.sst) || (e.sss != e.sss) || (e.cn != e.cn) || (e.hi != e.hi);
  return !bad;
}

// I:444-477
__device__ __forceinline__ void convert_from_grid_to_meters(const DevParams& p, double lat_ref, double& dx_dlon,
                                                            double& dy_dlat) {
  if (p.grid_is_latlon) {
    dx_dlon = (p.pi / 180.) * p.Rearth * cos((lat_ref) * (p.pi / 180.));
    dy_dlat = (p.pi / 180.) * p.Rearth;
  } else { dx_dlon = 1.; dy_dlat = 1.; }
}
__device__ __forceinline__ void convert_from_meters_to_grid(const DevParams& p, double lat_ref, double& dlon_dx,
                                                            double& dlat_dy) {
  if (p.grid_is_latlon) {
    dlon_dx = (180. / p.pi) / (p.Rearth * cos((lat_ref) * (p.pi / 180.)));
    dlat_dy = (180. / p.pi) / p.Rearth;
  } else { dlon_dx = 1.; dlat_dy = 1.; }
}

// accel I:1950-2442 after the environment is known.  IAF(us, vs, IAcc&) evaluates
// interactive_force with the latest velocity estimate (second call, I:2217); ia is
// the first evaluation (I:2153).  dragfrac: I:2104-2120.
template <bool INTERACTIVE, class IAF>
__device__ __forceinline__ void accel_core(const DevParams& p, double M, double T, double W, double L,
                                           double lat, double uvel, double vvel, double uvel0, double vvel0,
                                           double dt, const Env& e, double dragfrac, IAcc ia, IAF&& iaf,
                                           double& ax, double& ay, double& axn, double& ayn, double& bxn,
                                           double& byn, double& uveln_out, double& vveln_out) {
  // Verlet only: alpha=1, C_N=1, beta=1, use_new_predictive_corrective=T (I:2008-2013)
  const double Cr0 = 0.06;
  double u_star = uvel0 + (axn * (dt / 2.));
  double v_star = vvel0 + (ayn * (dt / 2.));
  double uo = e.uo, vo = e.vo, ui = e.ui, vi = e.vi, ua = e.ua, va = e.va;
  double ssh_x = e.ssh_x, ssh_y = e.ssh_y, hi = e.hi, od = e.od;
  double f_cori;
  if (p.grid_is_latlon && !p.use_f_plane) f_cori = p.omega2 * sin(p.pi_180 * lat);
  else f_cori = p.omega2 * sin(p.pi_180 * p.lat_ref);
  double D = (p.rho_bergs / KID_RHO_SEAWATER) * T;
  double F = T - D;
  hi = fmin(hi, D);
  double D_hi = fmax(0., D - hi);
  double groundfrac, c_gnd;
  if (p.h_to_init_grounding > 0.0) {
    groundfrac = 1.0 - (od - D) / p.h_to_init_grounding;
    groundfrac = fmax(groundfrac, 0.0); groundfrac = fmin(groundfrac, 1.0);
  } else {
    groundfrac = (D > od) ? 1.0 : 0.0;
  }
  if (groundfrac > 0.0) c_gnd = (p.cdrag_grounding * W * L * groundfrac) / M; else c_gnd = 0.0;
  double uwave = ua - uo, vwave = va - vo;
  double wmod = uwave * uwave + vwave * vwave;
  double ampl = 0.5 * 0.02025 * wmod;
  double Lwavelength = 0.32 * wmod;
  double Lcutoff = 0.125 * Lwavelength;
  double Ltop = 0.25 * Lwavelength;
  double Cr = Cr0 * fmin(fmax(0., (L - Lcutoff) / ((Ltop - Lcutoff) + 1.e-30)), 1.);
  double wave_rad = 0.5 * KID_RHO_SEAWATER / M * Cr * KID_GRAVITY * ampl * fmin(ampl, F) * (2. * W * L) / (W + L);
  wmod = sqrt(ua * ua + va * va);
  if (wmod != 0.) { uwave = ua / wmod; vwave = va / wmod; }
  else { uwave = 0.; vwave = 0.; wave_rad = 0.; }
  double c_ocn = KID_RHO_SEAWATER / M * p.ocean_drag_scale * (0.5 * KID_CD_WV * dragfrac * W * (D_hi) + KID_CD_WH * W * L);
  double c_atm = KID_RHO_AIR / M * (0.5 * KID_CD_AV * dragfrac * W * F + KID_CD_AH * W * L);
  double c_ice;
  if (fabs(hi) == 0.) c_ice = 0.; else c_ice = KID_RHO_ICE / M * (0.5 * KID_CD_IV * dragfrac * W * hi);
  if (fabs(ui) + fabs(vi) == 0.) c_ice = 0.;
  axn = -KID_GRAVITY * ssh_x + wave_rad * uwave;
  ayn = -KID_GRAVITY * ssh_y + wave_rad * vwave;
  bxn = 0.; byn = 0.;
  if (INTERACTIVE) { axn = axn + ia.IA_x; ayn = ayn + ia.IA_y; }
  axn = axn + f_cori * v_star;
  ayn = ayn - f_cori * u_star;
  double uveln = uvel0, vveln = vvel0;
  double us = uvel0, vs = vvel0;
  // the velocity-at-start halves of the drag magnitudes do not change between the two iterations
  double d0_ocn = sqrt((uvel0 - uo) * (uvel0 - uo) + (vvel0 - vo) * (vvel0 - vo));
  double d0_atm = sqrt((uvel0 - ua) * (uvel0 - ua) + (vvel0 - va) * (vvel0 - va));
  double d0_ice = sqrt((uvel0 - ui) * (uvel0 - ui) + (vvel0 - vi) * (vvel0 - vi));
#pragma unroll
  for (int itloop = 1; itloop <= 2; itloop++) {
    if (itloop == 2) { us = uveln; vs = vveln; }
    double drag_ocn = c_ocn * 0.5 * (sqrt((uveln - uo) * (uveln - uo) + (vveln - vo) * (vveln - vo)) + d0_ocn);
    double drag_atm = c_atm * 0.5 * (sqrt((uveln - ua) * (uveln - ua) + (vveln - va) * (vveln - va)) + d0_atm);
    double drag_ice = c_ice * 0.5 * (sqrt((uveln - ui) * (uveln - ui) + (vveln - vi) * (vveln - vi)) + d0_ice);
    double drag_gnd = c_gnd;
    double RHS_x = (axn / 2) + bxn;
    double RHS_y = (ayn / 2) + byn;
    RHS_x = RHS_x - drag_ocn * (u_star - uo) - drag_atm * (u_star - ua) - drag_ice * (u_star - ui) - drag_gnd * u_star;
    RHS_y = RHS_y - drag_ocn * (v_star - vo) - drag_atm * (v_star - va) - drag_ice * (v_star - vi) - drag_gnd * v_star;
    if (INTERACTIVE) {
      if (itloop > 1) iaf(us, vs, ia);
      RHS_x = RHS_x - (((ia.P11 * u_star) + (ia.P12 * v_star)) - ia.Pu_x);
      RHS_y = RHS_y - (((ia.P21 * u_star) + (ia.P22 * v_star)) - ia.Pu_y);
    }
    double A11, A12, A21, A22;
    if (p.only_interactive_forces) {
      RHS_x = (ia.IA_x / 2) - (((ia.P11 * u_star) + (ia.P12 * v_star)) - ia.Pu_x);
      RHS_y = (ia.IA_y / 2) - (((ia.P21 * u_star) + (ia.P22 * v_star)) - ia.Pu_y);
      A11 = 1 + (dt * ia.P11); A12 = (dt * ia.P12); A21 = (dt * ia.P21); A22 = 1 + (dt * ia.P22);
    } else {
      double lambda = drag_ocn + drag_atm + drag_ice + drag_gnd;
      A11 = 1. + 1.0 * dt * lambda;
      A22 = 1. + 1.0 * dt * lambda;
      A12 = -1.0 * dt * f_cori;
      A21 = 1.0 * dt * f_cori;
      A12 = A12 / 2.; A21 = A21 / 2.;
      if (INTERACTIVE) {
        A11 = A11 + (dt * ia.P11); A12 = A12 + (dt * ia.P12);
        A21 = A21 + (dt * ia.P21); A22 = A22 + (dt * ia.P22);
      }
    }
    double detA = 1. / ((A11 * A22) - (A12 * A21));
    ax = detA * (A22 * RHS_x - A12 * RHS_y);
    ay = detA * (A11 * RHS_y - A21 * RHS_x);
    uveln = u_star + dt * ax;
    vveln = v_star + dt * ay;
  }
  if (p.only_interactive_forces) {
    axn = ia.IA_x; ayn = ia.IA_y;
  } else {
    axn = -KID_GRAVITY * ssh_x + wave_rad * uwave;
    ayn = -KID_GRAVITY * ssh_y + wave_rad * vwave;
    if (INTERACTIVE) { axn = axn + ia.IA_x; ayn = ayn + ia.IA_y; }
    axn = axn + f_cori * vveln;
    ayn = ayn - f_cori * uveln;
  }
  bxn = ax - (axn / 2); byn = ay - (ayn / 2);
  uveln_out = uveln; vveln_out = vveln;
  if (p.override_iceberg_velocities) { ax = 0.0; ay = 0.0; axn = 0.0; ayn = 0.0; bxn = 0.0; byn = 0.0; }
}

// tangent plane helpers I:7767-7816, I:8066-8099 (lat > 89 only)
__device__ __noinline__ void tang_velocity(const DevParams& p, double lonn, double uvel3, double vvel3,
                                           double ax1, double ay1, double dt, double& uveln, double& vveln) {
  double clon = cos(lonn * p.pi_180), slon = sin(lonn * p.pi_180);
  double xdot3 = -slon * uvel3 - clon * vvel3, ydot3 = clon * uvel3 - slon * vvel3;
  double xddot1 = -slon * ax1 - clon * ay1, yddot1 = clon * ax1 - slon * ay1;
  double xdotn = xdot3 + (dt * xddot1), ydotn = ydot3 + (dt * yddot1);
  uveln = -slon * xdotn + clon * ydotn;
  vveln = -clon * xdotn - slon * ydotn;
}
__device__ __noinline__ void tang_position(const DevParams& p, double lon1, double lat1, double uvel2,
                                           double vvel2, double dt, double& lonn, double& latn) {
  double r180_pi = 180. / p.pi;
  double colat = 90. - lat1;
  double r = p.Rearth * (colat * p.pi_180);
  double clon = cos(lon1 * p.pi_180), slon = sin(lon1 * p.pi_180);
  double x1 = r * clon, y1 = r * slon;
  double xdot2 = -slon * uvel2 - clon * vvel2, ydot2 = clon * uvel2 - slon * vvel2;
  double xn = x1 + (dt * xdot2), yn = y1 + (dt * ydot2);
  double rn = sqrt(xn * xn + yn * yn);
  latn = 90. - (r180_pi * rn / p.Rearth);
  lonn = r180_pi * acos(xn / rn) * f_sign1(yn);
}

// I:7819-8063.  Returns bounced.  warn: count of the WARNING-level events.
__device__ __forceinline__ bool adjust_index_and_ground(const DevGrid& g, const DevParams& p, double& lon,
                                                        double& lat, int& i, int& j, double& xi, double& yj,
                                                        unsigned int* err, unsigned int* warn) {
  const double posn_eps = 0.05;
  bool bounced = false;
  int i0 = i, j0 = j;
  bool lret = pos_within_cell(g, p, lon, lat, i, j, &xi, &yj, err);
  if (lret) return false;
  // the debug-only inm/jnm search (I:7903-7936) is inactive; the repeat of pos_within_cell at
  // I:7940 has the same arguments as the call above
  const double* __restrict__ msk = g.msk;
  int icount = 0;
  while (!lret && icount < 4) {
    icount++;
    if (xi < 0.) {
      if (i > g.isd) {
        if (msk[gidx(g, i - 1, j)] > 0.) { if (i > g.isd + 1) i = i - 1; }
        else bounced = true;
      }
    } else if (xi >= 1.) {
      if (i < g.ied) {
        if (msk[gidx(g, i + 1, j)] > 0.) { if (i < g.ied) i = i + 1; }
        else bounced = true;
      }
    }
    if (yj < 0.) {
      if (j > g.jsd) {
        if (msk[gidx(g, i, j - 1)] > 0.) { if (j > g.jsd + 1) j = j - 1; }
        else bounced = true;
      }
    } else if (yj >= 1.) {
      if (j < g.jed) {
        if (msk[gidx(g, i, j + 1)] > 0.) { if (j < g.jed) j = j + 1; }
        else bounced = true;
      }
    }
    if (bounced) {
      if (xi >= 1.) xi = 1. - posn_eps;
      if (xi < 0.) xi = posn_eps;
      if (yj >= 1.) yj = 1. - posn_eps;
      if (yj < 0.) yj = posn_eps;
      bilin_lonlat(g, p, i, j, xi, yj, &lon, &lat);
    }
    lret = pos_within_cell(g, p, lon, lat, i, j, &xi, &yj, err);
  }
  if (!bounced && lret && msk[gidx(g, i, j)] > 0.) return false;
  if (!bounced && !lret) {
    if (abs(i - i0) + abs(j - j0) == 0) {
      if (p.use_roundoff_fix) {
        xi = (xi - 0.5) * (1. - posn_eps) + 0.5;
        yj = (yj - 0.5) * (1. - posn_eps) + 0.5;
      }
      atomicAdd(warn, 1u);
      // the explain call at I:8039 re-evaluates xi,yj for cell (inm,jnm)=(i0,j0)
      pos_within_cell(g, p, lon, lat, i0, j0, &xi, &yj, err);
    } else {
      atomicAdd(warn, 1u);
    }
  }
  if (xi >= 1.) xi = 1. - posn_eps;
  if (xi < 0.) xi = posn_eps;
  if (yj > 1.) yj = 1. - posn_eps;
  if (yj <= 0.) yj = posn_eps;
  bilin_lonlat(g, p, i, j, xi, yj, &lon, &lat);
  lret = pos_within_cell(g, p, lon, lat, i, j, &xi, &yj, err);
  if (!lret) atomicAdd(warn, 1u);
  return bounced;
}

// I:3307-3364
__device__ __forceinline__ void swap_d(double& x, double& y) { double t = x; x = y; y = t; }
__device__ __forceinline__ void rolling(const DevParams& p, double& Tn, double& Wn, double& Ln) {
  const double Delta = 6.0;
  double Dn = (p.rho_bergs / KID_RHO_SEAWATER) * Tn;
  if (Dn > 0.) {
    if ((!p.use_updated_rolling_scheme) && (p.tip_parameter < 999.)) {
      if (fmax(Wn, Ln) < sqrt(0.92 * (Dn * Dn) + 58.32 * Dn)) {
        swap_d(Tn, Wn);
        if (Wn > Ln) swap_d(Wn, Ln);
      }
    } else {
      if (Wn > Ln) swap_d(Ln, Wn);
      if ((!p.use_updated_rolling_scheme) && (p.tip_parameter >= 999.)) {
        double q = p.rho_bergs / KID_RHO_SEAWATER;
        if (Wn < sqrt((6.0 * q * (1 - q) * (Tn * Tn)) - (12 * Delta * q * Tn))) {
          swap_d(Tn, Wn);
          if (Wn > Ln) swap_d(Wn, Ln);
        }
      }
      if (p.use_updated_rolling_scheme) {
        double tip_parameter;
        if (p.tip_parameter > 0.) tip_parameter = p.tip_parameter;
        else tip_parameter = sqrt(6 * (p.rho_bergs / KID_RHO_SEAWATER) * (1 - (p.rho_bergs / KID_RHO_SEAWATER)));
        if ((tip_parameter * Tn) > Wn) {
          swap_d(Tn, Wn);
          if (Wn > Ln) swap_d(Wn, Ln);
        }
      }
    }
  }
}

// per-berg state thermodynamics reads and writes
struct ThermoState {
  double mass, thickness, width, length, mass_scaling, mass_of_bits, mass_of_fl_bits, mass_of_fl_bergy_bits;
  double fl_k, heat_density, start_day;
  int start_year;
};
// what one berg adds to the grid (already divided by area and scaled), I:3116-3199
struct ThermoFlux {
  double floating_melt, calving_hflx, berg_melt, bergy_src, bergy_melt, fl_bits_melt;
  double fl_parent_melt, fl_child_melt, melt_buoy, melt_eros, melt_conv, melt_buoy_fl, melt_eros_fl, melt_conv_fl;
  double fl_bits_src, net_heat;
};

enum { TH_KEEP = 0, TH_DELETE = 1, TH_BECAME_FL = 2 };

// the footloose-bits part of thermodynamics (I:3031-3064) -- out of line, only
// bergs that carry FL bits take it
struct FlBits { double Lfl, Wfl, Tfl, Lnfl, Wnfl, Tnfl, Mnew_fl, dMfl, dMb_fl, dMv_fl, dMe_fl; };
__device__ __noinline__ void thermo_fl_bits(const DevParams& p, double thickness, double mass_of_fl_bits,
                                            double dvo08, double SST, double Mv_fl, double Me_fl, FlBits& f) {
  const double perday = 1. / 86400.;
  const double l_c = p.pi / (2. * sqrt(2.)), lw_c = 1. / (KID_GRAVITY * KID_RHO_SEAWATER);
  const double B_c = 1. / (12. * (1. - pow(0.3, 2.)));
  double dt = p.dt;
  // fl_bits_dimensions I:3370-3387
  double l_w = pow(lw_c * p.fl_youngs * B_c * pow(thickness, 3.), 0.25);
  double l_b = l_c * l_w;
  f.Lfl = 3. * l_b; f.Wfl = l_b; f.Tfl = thickness;
  rolling(p, f.Tfl, f.Wfl, f.Lfl);
  double Lfl = f.Lfl, Wfl = f.Wfl, Tfl = f.Tfl;
  double Mfl = mass_of_fl_bits;
  double Volfl = Lfl * Wfl * Tfl;
  double Mb_fl = fmax(0.58 * dvo08 * (SST + 4.0) / pow(Lfl, 0.2), 0.) * perday;
  double Tnfl = fmax(Tfl - Mb_fl * dt, 0.);
  double Lnfl, Wnfl, nVolfl, Mnew_fl;
  if (p.use_operator_splitting) {
    nVolfl = Tnfl * Wfl * Lfl;
    double Mnew1_fl = (nVolfl / Volfl) * Mfl;
    f.dMb_fl = Mfl - Mnew1_fl;
    Lnfl = fmax(Lfl - Mv_fl * dt, 0.);
    Wnfl = fmax(Wfl - Mv_fl * dt, 0.);
    nVolfl = Tnfl * Wnfl * Lnfl;
    double Mnew2_fl = (nVolfl / Volfl) * Mfl;
    f.dMv_fl = Mnew1_fl - Mnew2_fl;
    Lnfl = fmax(Lnfl - Me_fl * dt, 0.);
    Wnfl = fmax(Wnfl - Me_fl * dt, 0.);
    nVolfl = Tnfl * Wnfl * Lnfl;
    Mnew_fl = (nVolfl / Volfl) * Mfl;
    f.dMe_fl = Mnew2_fl - Mnew_fl;
  } else {
    Lnfl = fmax(Lfl - (Mv_fl + Me_fl) * dt, 0.);
    Wnfl = fmax(Wfl - (Mv_fl + Me_fl) * dt, 0.);
    nVolfl = Tnfl * Wnfl * Lnfl;
    Mnew_fl = (nVolfl / Volfl) * Mfl;
    f.dMb_fl = (Mfl / Volfl) * (Wfl * Lfl) * Mb_fl * dt;
    f.dMe_fl = (Mfl / Volfl) * (Tfl * (Wfl + Lfl)) * Me_fl * dt;
    f.dMv_fl = (Mfl / Volfl) * (Tfl * (Wfl + Lfl)) * Mv_fl * dt;
  }
  f.Lnfl = Lnfl; f.Wnfl = Wnfl; f.Tnfl = Tnfl; f.Mnew_fl = Mnew_fl;
  f.dMfl = Mfl - Mnew_fl;
}

// thermodynamics of one berg, I:2896-3296.  `area` = grd%area(i,j) (caller has
// checked != 0), n_bonds_eff = N_bonds of I:2928-2944.
__device__ __forceinline__ int thermo_berg(const DevParams& p, const Env& e, double uvel, double vvel,
                                           double area, double N_bonds, ThermoState& s, ThermoFlux& fx) {
  const double perday = 1. / 86400.;
  double dt = p.dt;
  double SST = e.sst;
  double IC = fmin(1., e.cn + p.sicn_shift);
  double M = s.mass, T = s.thickness, W = s.width, L = s.length;
  double Vol = T * W * L;
  double dvo = sqrt((uvel - e.uo) * (uvel - e.uo) + (vvel - e.vo) * (vvel - e.vo));
  double dva = sqrt((e.ua - e.uo) * (e.ua - e.uo) + (e.va - e.vo) * (e.va - e.vo));
  double Ss = 1.5 * sqrt(dva) + 0.1 * dva;            // dva**0.5
  double dvo08 = pow(dvo, 0.8);
  double Mv = fmax(7.62e-3 * SST + 1.29e-3 * (SST * SST), 0.) * perday;
  double Mb = fmax(0.58 * dvo08 * (SST + 4.0) / pow(L, 0.2), 0.) * perday;
  double Me = fmax(1. / 12. * (SST + 2.) * Ss * (1 + cos(p.pi * (IC * IC * IC))), 0.) * perday;
  double Mv_fl = 0., Me_fl = 0.;
  if (s.mass_of_fl_bits > 0.) { Mv_fl = Mv; Me_fl = Me; }
  if (p.set_melt_rates_to_zero) { Mv = 0.0; Mb = 0.0; Me = 0.0; }
  double Tn, nVol, Mnew1, Mnew2, Mnew, dMb, dMv, dMe, dM, Ln1 = 0, Wn1 = 0, Ln, Wn;
  if (p.use_operator_splitting) {
    Tn = fmax(T - Mb * dt, 0.);
    nVol = Tn * W * L;
    Mnew1 = (nVol / Vol) * M;
    dMb = M - Mnew1;
    Ln1 = fmax(L - Mv * dt, 0.);
    Wn1 = fmax(W - Mv * dt, 0.);
    nVol = Tn * Wn1 * Ln1;
    Mnew2 = (nVol / Vol) * M;
    dMv = Mnew1 - Mnew2;
    Ln = fmax(Ln1 - Me * dt, 0.);
    Wn = fmax(Wn1 - Me * dt, 0.);
    nVol = Tn * Wn * Ln;
    Mnew = (nVol / Vol) * M;
    dMe = Mnew2 - Mnew;
    dM = M - Mnew;
  } else {
    Ln = fmax(L - (Mv + Me) * (dt), 0.);
    Wn = fmax(W - (Mv + Me) * (dt), 0.);
    Tn = fmax(T - Mb * (dt), 0.);
    nVol = Tn * Wn * Ln;
    Mnew = (nVol / Vol) * M;
    dM = M - Mnew;
    dMb = (M / Vol) * (W * L) * Mb * dt;
    dMe = (M / Vol) * (T * (W + L)) * Me * dt;
    dMv = (M / Vol) * (T * (W + L)) * Mv * dt;
  }
  if (p.footloose) {
    if (s.fl_k >= 0) {
      const double l_c = p.pi / (2. * sqrt(2.)), lw_c = 1. / (KID_GRAVITY * KID_RHO_SEAWATER);
      const double B_c = 1. / (12. * (1. - pow(0.3, 2.)));
      double l_b3 = 3. * l_c * pow(lw_c * p.fl_youngs * B_c * pow(Tn, 3.), 0.25);
      if (L > l_b3) {
        double fb = Tn * (1. - p.rho_bergs / KID_RHO_SEAWATER);
        double kd = Tn - fb;
        if (W > l_b3) {
          s.fl_k = s.fl_k + (dMe / fb - dMv / kd) / p.rho_bergs;
          if (s.fl_k < 0) s.fl_k = 0;
        } else {
          double dMv_l = dMv * (Wn1 + W) / (2. * (Ln1 + W));
          double dMe_l = dMe * (Wn + Wn1) / (2. * (Ln + Wn1));
          s.fl_k = s.fl_k + (dMe_l / fb - dMv_l / kd) / p.rho_bergs;
          if (s.fl_k < 0) s.fl_k = 0;
        }
      }
    }
  }
  FlBits fl;
  fl.Lfl = fl.Wfl = fl.Tfl = fl.Lnfl = fl.Wnfl = fl.Tnfl = 0.;
  bool has_fl = s.mass_of_fl_bits > 0.;
  if (has_fl) {
    thermo_fl_bits(p, s.thickness, s.mass_of_fl_bits, dvo08, SST, Mv_fl, Me_fl, fl);
  } else {
    fl.dMfl = 0.; fl.dMb_fl = 0.; fl.dMv_fl = 0.; fl.dMe_fl = 0.;
    fl.Mnew_fl = s.mass_of_fl_bits;
  }
  double dMfl = fl.dMfl, Mnew_fl = fl.Mnew_fl;
  double dMbitsE, dMbitsM, nMbits, dMbitsE_fl, dMbitsM_fl, nMbits_fl;
  if (p.bergy_bit_erosion_fraction > 0.) {
    double Mbits = s.mass_of_bits;
    dMbitsE = p.bergy_bit_erosion_fraction * dMe;
    nMbits = Mbits + dMbitsE;
    double Lbits = fmin(fmin(L, W), fmin(T, 40.));
    double Abits = (Mbits / p.rho_bergs) / Lbits;
    double Mbb = fmax(0.58 * dvo08 * (SST + 2.0) / pow(Lbits, 0.2), 0.) * perday;
    Mbb = p.rho_bergs * Abits * Mbb;
    dMbitsM = fmin(Mbb * dt, nMbits);
    nMbits = nMbits - dMbitsM;
    if (Mnew == 0.) { dMbitsM = dMbitsM + nMbits; nMbits = 0.; }
    if (has_fl) {
      double Mbits_fl = s.mass_of_fl_bergy_bits;
      dMbitsE_fl = p.bergy_bit_erosion_fraction * fl.dMe_fl;
      nMbits_fl = Mbits_fl + dMbitsE_fl;
      double Lbits_fl = fmin(fmin(fl.Lfl, fl.Wfl), fmin(fl.Tfl, 40.));
      double Abits_fl = (Mbits_fl / p.rho_bergs) / Lbits_fl;
      double Mbb_fl = fmax(0.58 * dvo08 * (SST + 2.0) / pow(Lbits_fl, 0.2), 0.) * perday;
      Mbb_fl = p.rho_bergs * Abits_fl * Mbb_fl;
      dMbitsM_fl = fmin(Mbb_fl * dt, nMbits_fl);
      nMbits_fl = nMbits_fl - dMbitsM_fl;
      if (Mnew_fl == 0.) { dMbitsM_fl = dMbitsM_fl + nMbits_fl; nMbits_fl = 0.; }
    } else {
      dMbitsE_fl = 0.; dMbitsM_fl = 0.; nMbits_fl = 0.;
    }
  } else {
    dMbitsE = 0.; dMbitsM = 0.; nMbits = s.mass_of_bits;
    dMbitsE_fl = 0.; dMbitsM_fl = 0.; nMbits_fl = s.mass_of_fl_bergy_bits;
  }
  // grid contributions I:3116-3199
  {
    double ms = s.mass_scaling;
    double melt = (dM - (dMbitsE - dMbitsM) + dMfl - (dMbitsE_fl - dMbitsM_fl)) / dt;
    fx.floating_melt = melt / area * ms;
    melt = melt * s.heat_density;
    fx.calving_hflx = melt / area * ms;
    fx.net_heat = melt * ms * dt;
    melt = dM / dt;
    fx.berg_melt = melt / area * ms;
    melt = (dMbitsE + dMbitsE_fl) / dt;
    fx.bergy_src = melt / area * ms;
    melt = (dMbitsM + dMbitsM_fl) / dt;
    fx.bergy_melt = melt / area * ms;
    melt = dMfl / dt;
    fx.fl_bits_melt = melt / area * ms;
    fx.fl_parent_melt = fx.fl_child_melt = fx.melt_buoy = fx.melt_eros = fx.melt_conv = 0.;
    fx.melt_buoy_fl = fx.melt_eros_fl = fx.melt_conv_fl = 0.;
    if (p.melt_diagnostics) {
      if (s.fl_k >= 0) {
        melt = (dM - (dMbitsE - dMbitsM)) / dt; fx.fl_parent_melt = melt / area * ms;
        melt = (dMfl - (dMbitsE_fl - dMbitsM_fl)) / dt; fx.fl_child_melt = melt / area * ms;
        melt = dMb / dt; fx.melt_buoy = melt / area * ms;
        melt = dMe / dt; fx.melt_eros = melt / area * ms;
        melt = dMv / dt; fx.melt_conv = melt / area * ms;
        if (dMfl > 0) {
          melt = fl.dMb_fl / dt; fx.melt_buoy_fl = melt / area * ms;
          melt = fl.dMe_fl / dt; fx.melt_eros_fl = melt / area * ms;
          melt = fl.dMv_fl / dt; fx.melt_conv_fl = melt / area * ms;
        }
      } else {
        melt = (dM - (dMbitsE - dMbitsM)) / dt; fx.fl_child_melt = melt / area * ms;
        melt = dMb / dt; fx.melt_buoy_fl = melt / area * ms;
        melt = dMe / dt; fx.melt_eros_fl = melt / area * ms;
        melt = dMv / dt; fx.melt_conv_fl = melt / area * ms;
      }
    }
  }
  fx.fl_bits_src = 0.;
  if (p.allow_bergs_to_roll && N_bonds == 0.) rolling(p, Tn, Wn, Ln);
  if (p.iceberg_melt_without_decay) {
    Mnew = s.mass; Mnew_fl = s.mass_of_fl_bits; nMbits_fl = s.mass_of_fl_bergy_bits;
  } else {
    s.mass = Mnew;
    s.mass_of_bits = nMbits;
    s.mass_of_fl_bits = Mnew_fl;
    s.mass_of_fl_bergy_bits = nMbits_fl;
    s.thickness = Tn;
    s.width = fmin(Wn, Ln);
    s.length = fmax(Wn, Ln);
  }
  if (Mnew <= 0.) {
    if (Mnew_fl > 0) {
      // the parent melted but its footloose bits did not: the bits become the berg, I:3272-3289
      s.mass = fl.Lnfl * fl.Wnfl * fl.Tnfl * p.rho_bergs;
      s.length = fl.Lnfl; s.width = fl.Wnfl; s.thickness = fl.Tnfl;
      nMbits_fl = nMbits_fl * s.mass_scaling;
      s.mass_scaling = Mnew_fl * s.mass_scaling / s.mass;
      s.mass_of_bits = nMbits_fl / s.mass_scaling;
      s.mass_of_fl_bits = 0.;
      s.mass_of_fl_bergy_bits = 0.;
      s.fl_k = -1.;
      s.start_year = p.current_year;
      s.start_day = p.current_yearday;
      fx.fl_bits_src = -(s.mass * s.mass_scaling / (dt * area));
      return TH_BECAME_FL;
    }
    return TH_DELETE;
  }
  return TH_KEEP;
}

}  // namespace kid
