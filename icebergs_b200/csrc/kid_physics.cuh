// kid_physics.cuh -- per-berg physics on the device.
//
// Follows the reference routine by routine (I: = src/icebergs.F90):
//   interp_flds I:4718 (+ ddx_ssh/ddy_ssh I:4903/4916 precomputed per cell, rotate I:4953),
//   accel I:1950, verlet_stepping I:7203, update_verlet_position I:7684,
//   adjust_index_and_ground I:7819, tangent-plane helpers I:7767-7816 / I:8066,
//   thermodynamics I:2844 (+ rolling I:3307, fl_bits_dimensions I:3370).
// The formulas, branch conditions and quirks are the reference's.
// Floating-point evaluation differs from the CPU path only at rounding level (north_star
// tolerance 1e-10): divisions by a common denominator are multiplications by one reciprocal,
// x**0.8 / x**0.2 are exp(c*log x), sin and cos of the latitude share one argument
// reduction, and the interpolation / momentum sums use explicit fma().  The library is compiled
// with -fmad=false: the compiler never contracts on its own, so expressions whose exact
// cancellation the reference relies on (the operator-split mass differences of thermodynamics,
// the cell-membership cross products) keep the plain IEEE sequence.  Everything that decides an
// integer (cell index, bounce, deletion) branches on the reference's own conditions.
#pragma once
#include "kid_geom.cuh"

namespace kid {

struct Env { double uo, vo, ui, vi, ua, va, ssh_x, ssh_y, sst, sss, cn, hi, od; };

// LEAN instances of the hot functions take the namelist switches below as compile-time constants
// (the values of the default free-drift configuration: lat-lon grid aligned with the axes, no f-plane,
// no coastal drift / speed limit / grounding drag / velocity override, operator-split melt, default
// rolling scheme, no footloose, no melt diagnostics); kid_init picks them when the parameters match
// (lean_config()), otherwise the generic instances read every switch at run time.
#define PF(field, leanval) (LEAN ? (leanval) : (p.field))

// Reciprocal and square root for the drag/melt-rate arithmetic: hardware seed (MUFU.RCP64H /
// MUFU.RSQ64H, ~20 bits) refined by Newton / Goldschmidt steps to ~1 ulp, without the IEEE
// corner-case paths of `/` and sqrt() (operands here are masses, lengths, speeds: positive,
// normal range).  Geometry and the mass-difference sequence keep IEEE `/`.
// max/min of ordinary numbers: one compare + select (fmax/fmin also order NaNs and signed zeros,
// which costs several more instructions per call; the operands here are never NaN)
__device__ __forceinline__ double kmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double kmin(double a, double b) { return a < b ? a : b; }

#ifndef KID_IEEE_DIVSQRT
__device__ __forceinline__ double rcp_nr(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);      // seed ~2^-20 -> 2^-40 -> 2^-80
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
// branch-free: x = 0 is nudged to 1e-300 (an exact no-op for any x >= 1e-284), so sqrt_nr(0) = 1e-150
// instead of 0 -- every caller either tests the radicand itself for zero (wind speed, I:2096) or
// only scales terms that vanish with it
__device__ __forceinline__ double sqrt_nr(double x) {
  x = x + 1.e-300;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);          // seed ~2^-20 -> 2^-40, the last line -> 2^-80
  g = fma(g, r, g); h = fma(h, r, h);
  return fma(fma(-g, g, x), h, g);
}
#else
__device__ __forceinline__ double rcp_nr(double x) { return 1. / x; }
__device__ __forceinline__ double sqrt_nr(double x) { return sqrt(x); }
#endif

// sin and cos for |x| <= pi/2 (x = pi/180*lat): Taylor polynomials in x^2 to x^23 / x^24,
// truncation < 1e-20, rounding a few ulp -- no argument reduction, no slow path
__device__ __forceinline__ void sincos_halfpi_nofallback(double x, double* s, double* c);
__device__ __forceinline__ void sincos_halfpi(double x, double* s, double* c) {
  if (!(fabs(x) <= 1.5708)) { sincos(x, s, c); return; }
  sincos_halfpi_nofallback(x, s, c);
}
__device__ __forceinline__ void sincos_halfpi_nofallback(double x, double* s, double* c) {
  const double z = x * x;
  double ps = -1. / 25852016738884976640000.;            // -1/23!
  ps = fma(ps, z, 1. / 51090942171709440000.);            //  1/21!
  ps = fma(ps, z, -1. / 121645100408832000.);             // -1/19!
  ps = fma(ps, z, 1. / 355687428096000.);                 //  1/17!
  ps = fma(ps, z, -1. / 1307674368000.);                  // -1/15!
  ps = fma(ps, z, 1. / 6227020800.);                      //  1/13!
  ps = fma(ps, z, -1. / 39916800.);                       // -1/11!
  ps = fma(ps, z, 1. / 362880.);                          //  1/9!
  ps = fma(ps, z, -1. / 5040.);                           // -1/7!
  ps = fma(ps, z, 1. / 120.);                             //  1/5!
  ps = fma(ps, z, -1. / 6.);                              // -1/3!
  *s = fma(x * z, ps, x);
  double pc = 1. / 620448401733239439360000.;             //  1/24!
  pc = fma(pc, z, -1. / 1124000727777607680000.);         // -1/22!
  pc = fma(pc, z, 1. / 2432902008176640000.);             //  1/20!
  pc = fma(pc, z, -1. / 6402373705728000.);               // -1/18!
  pc = fma(pc, z, 1. / 20922789888000.);                  //  1/16!
  pc = fma(pc, z, -1. / 87178291200.);                    // -1/14!
  pc = fma(pc, z, 1. / 479001600.);                       //  1/12!
  pc = fma(pc, z, -1. / 3628800.);                        // -1/10!
  pc = fma(pc, z, 1. / 40320.);                           //  1/8!
  pc = fma(pc, z, -1. / 720.);                            // -1/6!
  pc = fma(pc, z, 1. / 24.);                              //  1/4!
  pc = fma(pc, z, -0.5);
  *c = fma(pc, z, 1.);
}

// accumulated interaction terms of interactive_force (I:480): IA_x, IA_y, P_ia_*, P_ia_times_u_*
struct IAcc { double IA_x, IA_y, P11, P12, P21, P22, Pu_x, Pu_y; };

// F:7071-7088 on one component of the four corner records
// (w3,w4,wn,ws) = (xi,1-xi,yj,1-yj), or swapped under old_bug_bilin
#define KID_BILIN(f) fma(fma(c3.f, w3, c4.f * w4), wn, fma(c2.f, w3, c1.f * w4) * ws)
#define KID_BILIN_WEIGHTS                                   \
  double w3 = xi, w4 = 1. - xi, wn = yj, ws = 1. - yj;      \
  if (p.old_bug_bilin) { w3 = 1. - xi; w4 = xi; wn = 1. - yj; ws = yj; }

__device__ __forceinline__ void rotate(double& u, double& v, double cos_rot, double sin_rot) {
  double u_old = u, v_old = v;
  u = fma(cos_rot, u_old, sin_rot * v_old);
  v = fma(cos_rot, v_old, -(sin_rot * u_old));
}

// I:4718-4900 (non-MTS ocean depth, I:4897).  Returns false when a NaN survived.
template <bool LEAN = false>
__device__ __forceinline__ bool interp_flds(const DevGrid& g, const DevParams& p, int i, int j, double xi,
                                            double yj, Env& e) {
  const CornerRec* __restrict__ cr = g.corner;
  const CellRec* __restrict__ ce = g.cell;
  const int ne = gidx(g, i, j);
  const int nid = g.nid;
  const CornerRec c3 = cr[ne], c4 = cr[ne - 1], c2 = cr[ne - nid], c1 = cr[ne - nid - 1];
  KID_BILIN_WEIGHTS
  double cos_rot = 1., sin_rot = 0.;
  if (!PF(no_rotation, 1)) { cos_rot = KID_BILIN(cosr); sin_rot = KID_BILIN(sinr); }
  double uo = KID_BILIN(uo), vo = KID_BILIN(vo);
  double ui = KID_BILIN(ui), vi = KID_BILIN(vi);
  double ua = KID_BILIN(ua), va = KID_BILIN(va);
  if (PF(coastal_drift, 0.) > 0.) {
    const double* __restrict__ msk = g.msk;
    double cd = PF(coastal_drift, 0.);
    double m0 = msk[ne], mE = msk[ne + 1], mW = msk[ne - 1], mN = msk[ne + nid], mS = msk[ne - nid];
    uo = uo + cd * (mE - mW) * m0;
    ui = ui + cd * (mE - mW) * m0;
    vo = vo + cd * (mN - mS) * m0;
    vi = vi + cd * (mN - mS) * m0;
  }
  const CellRec c0 = ce[ne];
  e.sst = c0.sst; e.sss = c0.sss; e.cn = c0.cn; e.hi = c0.hi; e.od = c0.od;
  // SSH slopes, I:4830-4860.  The neighbour picked depends on xi,yj >= 0.5; the loads do not
  // wait for that decision: the address is selected, not the branch.
  const bool yn = yj >= 0.5, xe = xi >= 0.5;
  const int rj = yn ? ne + nid : ne - nid;         // row above or below
  const int ri = xe ? ne + 1 : ne - 1;             // column east or west
  const double ddx_j0 = c0.ddx, ddx_j0w = ce[ne - 1].ddx, ddx_j1 = ce[rj].ddx, ddx_j1w = ce[rj - 1].ddx;
  const double ddy_i0 = c0.ddy, ddy_i0s = ce[ne - nid].ddy, ddy_i1 = ce[ri].ddy, ddy_i1s = ce[ri - nid].ddy;
  double hxp, hxm;
  if (yn) {
    hxp = fma((yj - 0.5), ddx_j1, (1.5 - yj) * ddx_j0);
    hxm = fma((yj - 0.5), ddx_j1w, (1.5 - yj) * ddx_j0w);
  } else {
    hxp = fma((yj + 0.5), ddx_j0, (0.5 - yj) * ddx_j1);
    hxm = fma((yj + 0.5), ddx_j0w, (0.5 - yj) * ddx_j1w);
  }
  double ssh_x = fma(xi, hxp, (1. - xi) * hxm);
  if (xe) {
    hxp = fma((xi - 0.5), ddy_i1, (1.5 - xi) * ddy_i0);
    hxm = fma((xi - 0.5), ddy_i1s, (1.5 - xi) * ddy_i0s);
  } else {
    hxp = fma((xi + 0.5), ddy_i0, (0.5 - xi) * ddy_i1);
    hxm = fma((xi + 0.5), ddy_i0s, (0.5 - xi) * ddy_i1s);
  }
  double ssh_y = fma(yj, hxp, (1. - yj) * hxm);
  if (!PF(no_rotation, 1)) {
    rotate(uo, vo, cos_rot, sin_rot);
    rotate(ui, vi, cos_rot, sin_rot);
    rotate(ua, va, cos_rot, sin_rot);
    rotate(ssh_x, ssh_y, cos_rot, sin_rot);
  }
  if (ssh_x != ssh_x) ssh_x = 0.;
  if (ssh_y != ssh_y) ssh_y = 0.;
  e.uo = uo; e.vo = vo; e.ui = ui; e.vi = vi; e.ua = ua; e.va = va; e.ssh_x = ssh_x; e.ssh_y = ssh_y;
  bool bad = (uo != uo) || (vo != vo) || (ui != ui) || (vi != vi) || (ua != ua) || (va != va) ||
             (e.sst != e.sst) || (e.sss != e.sss) || (e.cn != e.cn) || (e.hi != e.hi);
  return !bad;
}

// The grid records one cell's bergs gather, as k_step_fast caches them in shared memory once per warp and cell
// (56 doubles): [0..31] the corner records c1 (i-1,j-1), c2 (i,j-1), c4 (i-1,j), c3 (i,j); [32..39] the cell record;
// [40..43] the rectangle record; [44..48] ddx_ssh of cells (i-1,j), (i,j+1), (i-1,j+1), (i,j-1), (i-1,j-1);
// [49..53] ddy_ssh of cells (i,j-1), (i+1,j), (i+1,j-1), (i-1,j), (i-1,j-1).
#define KID_CACHE_DOUBLES 56
__device__ __forceinline__ CornerRec cache_corner(const double* __restrict__ c, int k) {
  const double2* q = reinterpret_cast<const double2*>(c + 8 * k);
  double2 a = q[0], b = q[1], d = q[2], e = q[3];
  CornerRec r; r.uo = a.x; r.vo = a.y; r.ui = b.x; r.vi = b.y; r.ua = d.x; r.va = d.y; r.cosr = e.x; r.sinr = e.y;
  return r;
}
// interp_flds I:4718-4900 of the lean configuration (grid aligned with lon/lat: rotate() is the identity; no
// coastal drift) on a cached cell: the statements of interp_flds<true> on the same operands
__device__ __forceinline__ bool interp_flds_cached(const DevParams& p, const double* __restrict__ c, double xi, double yj, Env& e) {
  const CornerRec c1 = cache_corner(c, 0), c2 = cache_corner(c, 1), c4 = cache_corner(c, 2), c3 = cache_corner(c, 3);
  KID_BILIN_WEIGHTS
  double uo = KID_BILIN(uo), vo = KID_BILIN(vo);
  double ui = KID_BILIN(ui), vi = KID_BILIN(vi);
  double ua = KID_BILIN(ua), va = KID_BILIN(va);
  e.sst = c[32]; e.sss = c[33]; e.cn = c[34]; e.hi = c[35]; e.od = c[36];
  const bool yn = yj >= 0.5, xe = xi >= 0.5;
  const double ddx_j0 = c[38], ddx_j0w = c[44], ddx_j1 = yn ? c[45] : c[47], ddx_j1w = yn ? c[46] : c[48];
  const double ddy_i0 = c[39], ddy_i0s = c[49], ddy_i1 = xe ? c[50] : c[52], ddy_i1s = xe ? c[51] : c[53];
  double hxp, hxm;
  if (yn) {
    hxp = fma((yj - 0.5), ddx_j1, (1.5 - yj) * ddx_j0);
    hxm = fma((yj - 0.5), ddx_j1w, (1.5 - yj) * ddx_j0w);
  } else {
    hxp = fma((yj + 0.5), ddx_j0, (0.5 - yj) * ddx_j1);
    hxm = fma((yj + 0.5), ddx_j0w, (0.5 - yj) * ddx_j1w);
  }
  double ssh_x = fma(xi, hxp, (1. - xi) * hxm);
  if (xe) {
    hxp = fma((xi - 0.5), ddy_i1, (1.5 - xi) * ddy_i0);
    hxm = fma((xi - 0.5), ddy_i1s, (1.5 - xi) * ddy_i0s);
  } else {
    hxp = fma((xi + 0.5), ddy_i0, (0.5 - xi) * ddy_i1);
    hxm = fma((xi + 0.5), ddy_i0s, (0.5 - xi) * ddy_i1s);
  }
  double ssh_y = fma(yj, hxp, (1. - yj) * hxm);
  if (ssh_x != ssh_x) ssh_x = 0.;
  if (ssh_y != ssh_y) ssh_y = 0.;
  e.uo = uo; e.vo = vo; e.ui = ui; e.vi = vi; e.ua = ua; e.va = va; e.ssh_x = ssh_x; e.ssh_y = ssh_y;
  bool bad = (uo != uo) || (vo != vo) || (ui != ui) || (vi != vi) || (ua != ua) || (va != va) ||
             (e.sst != e.sst) || (e.sss != e.sss) || (e.cn != e.cn) || (e.hi != e.hi);
  return !bad;
}

// What thermodynamics (I:2896-2920) uses of interp_flds: the rotated ocean and wind velocities
// and the A-grid picks of sst, cn; also hands back 1/area of the cell.
struct EnvThermo { double uo, vo, ua, va, sst, cn, rarea; };
template <bool LEAN = false>
__device__ __forceinline__ void interp_thermo(const DevGrid& g, const DevParams& p, int ne, double xi, double yj,
                                              EnvThermo& e) {
  const CornerRec* __restrict__ cr = g.corner;
  const int nid = g.nid;
  const CornerRec c3 = cr[ne], c4 = cr[ne - 1], c2 = cr[ne - nid], c1 = cr[ne - nid - 1];
  KID_BILIN_WEIGHTS
  double uo = KID_BILIN(uo), vo = KID_BILIN(vo);
  double ua = KID_BILIN(ua), va = KID_BILIN(va);
  if (PF(coastal_drift, 0.) > 0.) {
    const double* __restrict__ msk = g.msk;
    double cd = PF(coastal_drift, 0.);
    double m0 = msk[ne], mE = msk[ne + 1], mW = msk[ne - 1], mN = msk[ne + nid], mS = msk[ne - nid];
    uo = uo + cd * (mE - mW) * m0;
    vo = vo + cd * (mN - mS) * m0;
  }
  if (!PF(no_rotation, 1)) {
    double cos_rot = KID_BILIN(cosr), sin_rot = KID_BILIN(sinr);
    rotate(uo, vo, cos_rot, sin_rot);
    rotate(ua, va, cos_rot, sin_rot);
  }
  const CellRec* __restrict__ ce = g.cell;
  e.sst = ce[ne].sst; e.cn = ce[ne].cn; e.rarea = ce[ne].rarea;
  e.uo = uo; e.vo = vo; e.ua = ua; e.va = va;
}

// interp_thermo of the lean configuration on a cached cell (see interp_flds_cached)
__device__ __forceinline__ void interp_thermo_cached(const DevParams& p, const double* __restrict__ c, double xi, double yj,
                                                     EnvThermo& e) {
  const CornerRec c1 = cache_corner(c, 0), c2 = cache_corner(c, 1), c4 = cache_corner(c, 2), c3 = cache_corner(c, 3);
  KID_BILIN_WEIGHTS
  e.uo = KID_BILIN(uo); e.vo = KID_BILIN(vo);
  e.ua = KID_BILIN(ua); e.va = KID_BILIN(va);
  e.sst = c[32]; e.cn = c[34]; e.rarea = c[37];
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
// the first evaluation (I:2153).  dragfrac: I:2104-2120.  f_cori: I:2043-2047.
template <bool INTERACTIVE, bool LEAN = false, class IAF>
__device__ __forceinline__ void accel_core(const DevParams& p, double M, double T, double W, double L,
                                           double f_cori, double uvel0, double vvel0, double dt, const Env& e,
                                           double dragfrac, IAcc ia, IAF&& iaf, double& ax, double& ay,
                                           double& axn, double& ayn, double& bxn, double& byn, double& uveln_out,
                                           double& vveln_out) {
  // Verlet only: alpha=1, C_N=1, beta=1, use_new_predictive_corrective=T (I:2008-2013)
  const double Cr0 = 0.06;
  double u_star = fma(axn, (dt * 0.5), uvel0);
  double v_star = fma(ayn, (dt * 0.5), vvel0);
  double uo = e.uo, vo = e.vo, ui = e.ui, vi = e.vi, ua = e.ua, va = e.va;
  double ssh_x = e.ssh_x, ssh_y = e.ssh_y, hi = e.hi, od = e.od;
  double rM = rcp_nr(M);
  double D = p.rho_ratio * T;
  double F = T - D;
  hi = kmin(hi, D);
  double D_hi = kmax(0., D - hi);
  double c_gnd = 0.0;
  if (PF(cdrag_grounding, 0.) != 0.) {      // I:2066-2082 (c_gnd is exactly 0 otherwise)
    double groundfrac;
    if (p.h_to_init_grounding > 0.0) {
      groundfrac = 1.0 - (od - D) * p.r_h2ig;
      groundfrac = kmin(kmax(groundfrac, 0.0), 1.0);
    } else {
      groundfrac = (D > od) ? 1.0 : 0.0;
    }
    if (groundfrac > 0.0) c_gnd = (PF(cdrag_grounding, 0.) * W * L * groundfrac) * rM;
  }
  double uwave = ua - uo, vwave = va - vo;
  double wmod = fma(uwave, uwave, vwave * vwave);
  double ampl = 0.5 * 0.02025 * wmod;
  double Lwavelength = 0.32 * wmod;
  double Lcutoff = 0.125 * Lwavelength;
  double Ltop = 0.25 * Lwavelength;
  // Cr0*min(max(0,(L-Lcutoff)/((Ltop-Lcutoff)+1e-30)),1): the quotient only matters strictly inside (0,1)
  double cr_num = L - Lcutoff, cr_den = (Ltop - Lcutoff) + 1.e-30;
  double Cr = (cr_num >= cr_den) ? Cr0 : ((cr_num <= 0.) ? 0. : Cr0 * (cr_num / cr_den));
  double wave_rad = 0.5 * KID_RHO_SEAWATER * rM * Cr * KID_GRAVITY * ampl * kmin(ampl, F) * (2. * W * L) * rcp_nr(W + L);
  double wmod2 = fma(ua, ua, va * va);
  wmod = sqrt_nr(wmod2);
  if (wmod2 != 0.) { double rw = rcp_nr(wmod); uwave = ua * rw; vwave = va * rw; }
  else { uwave = 0.; vwave = 0.; wave_rad = 0.; }
  double WL = W * L;
  double c_ocn = KID_RHO_SEAWATER * rM * p.ocean_drag_scale * (0.5 * KID_CD_WV * dragfrac * W * (D_hi) + KID_CD_WH * WL);
  double c_atm = KID_RHO_AIR * rM * (0.5 * KID_CD_AV * dragfrac * W * F + KID_CD_AH * WL);
  double c_ice;
  if (fabs(hi) == 0.) c_ice = 0.; else c_ice = KID_RHO_ICE * rM * (0.5 * KID_CD_IV * dragfrac * W * hi);
  if (fabs(ui) + fabs(vi) == 0.) c_ice = 0.;
  double ax_expl = fma(-KID_GRAVITY, ssh_x, wave_rad * uwave);     // I:2143-2144 and again I:2288-2289
  double ay_expl = fma(-KID_GRAVITY, ssh_y, wave_rad * vwave);
  if (INTERACTIVE) { ax_expl = ax_expl + ia.IA_x; ay_expl = ay_expl + ia.IA_y; }
  axn = fma(f_cori, v_star, ax_expl);
  ayn = fma(-f_cori, u_star, ay_expl);
  bxn = 0.; byn = 0.;
  double uveln = uvel0, vveln = vvel0;
  double us = uvel0, vs = vvel0;
  // the velocity-at-start halves of the drag magnitudes are the same in both iterations, and in
  // the first one uveln = uvel0 so that 0.5*(d0+d0) = d0 exactly
#define KID_HYPOT(a, b) sqrt_nr(fma((a), (a), (b) * (b)))
  double d0_ocn = KID_HYPOT(uvel0 - uo, vvel0 - vo);
  double d0_atm = KID_HYPOT(uvel0 - ua, vvel0 - va);
  double d0_ice = KID_HYPOT(uvel0 - ui, vvel0 - vi);
  double half_cori = 0.5 * (dt * f_cori);                        // A21 = -A12 = alpha*dt*f_cori/2
#pragma unroll
  for (int itloop = 1; itloop <= 2; itloop++) {
    double drag_ocn, drag_atm, drag_ice;
    if (itloop == 1) {
      drag_ocn = c_ocn * d0_ocn; drag_atm = c_atm * d0_atm; drag_ice = c_ice * d0_ice;
    } else {
      us = uveln; vs = vveln;
      drag_ocn = c_ocn * 0.5 * (KID_HYPOT(uveln - uo, vveln - vo) + d0_ocn);
      drag_atm = c_atm * 0.5 * (KID_HYPOT(uveln - ua, vveln - va) + d0_atm);
      drag_ice = c_ice * 0.5 * (KID_HYPOT(uveln - ui, vveln - vi) + d0_ice);
    }
    double drag_gnd = c_gnd;
    double RHS_x = fma(axn, 0.5, bxn);
    double RHS_y = fma(ayn, 0.5, byn);
    RHS_x = fma(-drag_gnd, u_star, fma(-drag_ice, (u_star - ui), fma(-drag_atm, (u_star - ua), fma(-drag_ocn, (u_star - uo), RHS_x))));
    RHS_y = fma(-drag_gnd, v_star, fma(-drag_ice, (v_star - vi), fma(-drag_atm, (v_star - va), fma(-drag_ocn, (v_star - vo), RHS_y))));
    if (INTERACTIVE) {
      if (itloop > 1) iaf(us, vs, ia);
      RHS_x = RHS_x - (((ia.P11 * u_star) + (ia.P12 * v_star)) - ia.Pu_x);
      RHS_y = RHS_y - (((ia.P21 * u_star) + (ia.P22 * v_star)) - ia.Pu_y);
    }
    double A11, A12, A21, A22;
    if (INTERACTIVE && PF(only_interactive_forces, 0)) {
      RHS_x = (ia.IA_x * 0.5) - (((ia.P11 * u_star) + (ia.P12 * v_star)) - ia.Pu_x);
      RHS_y = (ia.IA_y * 0.5) - (((ia.P21 * u_star) + (ia.P22 * v_star)) - ia.Pu_y);
      A11 = 1 + (dt * ia.P11); A12 = (dt * ia.P12); A21 = (dt * ia.P21); A22 = 1 + (dt * ia.P22);
    } else {
      double lambda = drag_ocn + drag_atm + drag_ice + drag_gnd;
      A11 = fma(dt, lambda, 1.);
      A22 = A11;
      A12 = -half_cori;
      A21 = half_cori;
      if (INTERACTIVE) {
        A11 = A11 + (dt * ia.P11); A12 = A12 + (dt * ia.P12);
        A21 = A21 + (dt * ia.P21); A22 = A22 + (dt * ia.P22);
      }
    }
    double detA = rcp_nr(fma(A11, A22, -(A12 * A21)));
    ax = detA * fma(A22, RHS_x, -(A12 * RHS_y));
    ay = detA * fma(A11, RHS_y, -(A21 * RHS_x));
    uveln = fma(dt, ax, u_star);
    vveln = fma(dt, ay, v_star);
  }
  if (INTERACTIVE && PF(only_interactive_forces, 0)) {
    axn = ia.IA_x; ayn = ia.IA_y;
  } else {
    axn = fma(f_cori, vveln, ax_expl);
    ayn = fma(-f_cori, uveln, ay_expl);
  }
  bxn = fma(axn, -0.5, ax); byn = fma(ayn, -0.5, ay);
  uveln_out = uveln; vveln_out = vveln;
  if (PF(override_iceberg_velocities, 0)) { ax = 0.0; ay = 0.0; axn = 0.0; ayn = 0.0; bxn = 0.0; byn = 0.0; }
}

// accel I:1950-2442 as Runge_Kutta_stepping calls it (Runge_not_Verlet=.true.): alpha = 0, beta = 1,
// C_N = 0 (I:2002-2005) -- explicit Coriolis at the stage velocity, implicit drag -- and the namelist's
// use_new_predictive_corrective.  Free bergs only (interactions with RK are refused at init).  The
// stepping default of the reference; kept out of line and in plain IEEE arithmetic.
// INTERACTIVE: with interactive_force (I:2152-2161, I:2216-2227, I:2252-2257); ia = its first evaluation at the
// step's starting velocity, iaf(us, vs, IAcc&) the re-evaluation inside the drag iteration.  dragfrac: I:2104-2120.
template <bool INTERACTIVE, class IAF>
__device__ __noinline__ void accel_rk(const DevParams& p, double M, double T, double W, double L, double lat,
                                      double uvel, double vvel, double uvel0, double vvel0, double dt, const Env& e,
                                      double loc_dx, double dragfrac, IAcc ia, IAF&& iaf, double& ax, double& ay, double& axn,
                                      double& ayn, double& bxn, double& byn, bool& speeding) {
  const double Cr0 = 0.06;
  const bool use_new_pc = p.use_new_predictive_corrective != 0;
  double u_star = uvel0 + (axn * (dt / 2.)), v_star = vvel0 + (ayn * (dt / 2.));
  double uo = e.uo, vo = e.vo, ui = e.ui, vi = e.vi, ua = e.ua, va = e.va, ssh_x = e.ssh_x, ssh_y = e.ssh_y, hi = e.hi, od = e.od;
  double f_cori = (p.grid_is_latlon && !p.use_f_plane) ? p.omega2 * sin(p.pi_180 * lat) : p.f_cori_plane;
  double D = (p.rho_bergs / KID_RHO_SEAWATER) * T, F = T - D;
  hi = fmin(hi, D);
  double D_hi = fmax(0., D - hi);
  double groundfrac, c_gnd;
  if (p.h_to_init_grounding > 0.0) { groundfrac = 1.0 - (od - D) / p.h_to_init_grounding; groundfrac = fmin(fmax(groundfrac, 0.0), 1.0); }
  else groundfrac = (D > od) ? 1.0 : 0.0;
  c_gnd = (groundfrac > 0.0) ? (p.cdrag_grounding * W * L * groundfrac) / M : 0.0;
  double uwave = ua - uo, vwave = va - vo;
  double wmod = uwave * uwave + vwave * vwave;
  double ampl = 0.5 * 0.02025 * wmod, Lwavelength = 0.32 * wmod, Lcutoff = 0.125 * Lwavelength, Ltop = 0.25 * Lwavelength;
  double Cr = Cr0 * fmin(fmax(0., (L - Lcutoff) / ((Ltop - Lcutoff) + 1.e-30)), 1.);
  double wave_rad = 0.5 * KID_RHO_SEAWATER / M * Cr * KID_GRAVITY * ampl * fmin(ampl, F) * (2. * W * L) / (W + L);
  wmod = sqrt(ua * ua + va * va);
  if (wmod != 0.) { uwave = ua / wmod; vwave = va / wmod; } else { uwave = 0.; vwave = 0.; wave_rad = 0.; }
  double c_ocn = KID_RHO_SEAWATER / M * p.ocean_drag_scale * (0.5 * KID_CD_WV * dragfrac * W * (D_hi) + KID_CD_WH * W * L);
  double c_atm = KID_RHO_AIR / M * (0.5 * KID_CD_AV * dragfrac * W * F + KID_CD_AH * W * L);
  double c_ice = (fabs(hi) == 0.) ? 0. : KID_RHO_ICE / M * (0.5 * KID_CD_IV * dragfrac * W * hi);
  if (fabs(ui) + fabs(vi) == 0.) c_ice = 0.;
  axn = 0.; ayn = 0.;
  bxn = -KID_GRAVITY * ssh_x + wave_rad * uwave;
  byn = -KID_GRAVITY * ssh_y + wave_rad * vwave;
  if (INTERACTIVE) { bxn = bxn + ia.IA_x; byn = byn + ia.IA_y; }   // I:2152-2161 (Runge_not_Verlet: into bxn, byn)
  bxn = bxn + f_cori * vvel; byn = byn - f_cori * uvel;            // alpha = 0: explicit Coriolis, I:2173-2174
  double uveln, vveln;
  if (use_new_pc) { uveln = uvel0; vveln = vvel0; } else { uveln = uvel; vveln = vvel; }
  double us_ia = uvel0, vs_ia = vvel0;                             // I:2182-2186
  for (int itloop = 1; itloop <= 2; itloop++) {
    if (itloop == 2) { us_ia = uveln; vs_ia = vveln; }
    double drag_ocn, drag_atm, drag_ice, drag_gnd = c_gnd;
    if (use_new_pc) {
      drag_ocn = c_ocn * 0.5 * (sqrt((uveln - uo) * (uveln - uo) + (vveln - vo) * (vveln - vo)) + sqrt((uvel0 - uo) * (uvel0 - uo) + (vvel0 - vo) * (vvel0 - vo)));
      drag_atm = c_atm * 0.5 * (sqrt((uveln - ua) * (uveln - ua) + (vveln - va) * (vveln - va)) + sqrt((uvel0 - ua) * (uvel0 - ua) + (vvel0 - va) * (vvel0 - va)));
      drag_ice = c_ice * 0.5 * (sqrt((uveln - ui) * (uveln - ui) + (vveln - vi) * (vveln - vi)) + sqrt((uvel0 - ui) * (uvel0 - ui) + (vvel0 - vi) * (vvel0 - vi)));
    } else {
      double us = 0.5 * (uveln + uvel), vs = 0.5 * (vveln + vvel);
      us_ia = us; vs_ia = vs;                                      // (the original scheme overwrites us, vs: I:2199)
      drag_ocn = c_ocn * sqrt((us - uo) * (us - uo) + (vs - vo) * (vs - vo));
      drag_atm = c_atm * sqrt((us - ua) * (us - ua) + (vs - va) * (vs - va));
      drag_ice = c_ice * sqrt((us - ui) * (us - ui) + (vs - vi) * (vs - vi));
    }
    double RHS_x = (axn / 2) + bxn, RHS_y = (ayn / 2) + byn;
    RHS_x = RHS_x - drag_ocn * (u_star - uo) - drag_atm * (u_star - ua) - drag_ice * (u_star - ui) - drag_gnd * u_star;
    RHS_y = RHS_y - drag_ocn * (v_star - vo) - drag_atm * (v_star - va) - drag_ice * (v_star - vi) - drag_gnd * v_star;
    if (INTERACTIVE) {                                             // I:2216-2227
      if (itloop > 1) iaf(us_ia, vs_ia, ia);
      RHS_x = RHS_x - (((ia.P11 * u_star) + (ia.P12 * v_star)) - ia.Pu_x);
      RHS_y = RHS_y - (((ia.P21 * u_star) + (ia.P22 * v_star)) - ia.Pu_y);
    }
    double lambda = drag_ocn + drag_atm + drag_ice + drag_gnd;
    double A11 = 1. + dt * lambda, A22 = A11, A12 = -0.0 * dt * f_cori, A21 = 0.0 * dt * f_cori;
    if (INTERACTIVE) {
      if (p.only_interactive_forces) {                             // I:2236-2242
        RHS_x = (ia.IA_x / 2) - (((ia.P11 * u_star) + (ia.P12 * v_star)) - ia.Pu_x);
        RHS_y = (ia.IA_y / 2) - (((ia.P21 * u_star) + (ia.P22 * v_star)) - ia.Pu_y);
        A11 = 1 + (dt * ia.P11); A12 = (dt * ia.P12); A21 = (dt * ia.P21); A22 = 1 + (dt * ia.P22);
      } else {
        A11 = A11 + (dt * ia.P11); A12 = A12 + (dt * ia.P12); A21 = A21 + (dt * ia.P21); A22 = A22 + (dt * ia.P22);
      }
    }
    double detA = 1. / ((A11 * A22) - (A12 * A21));
    ax = detA * (A22 * RHS_x - A12 * RHS_y);
    ay = detA * (A11 * RHS_y - A21 * RHS_x);
    uveln = u_star + dt * ax; vveln = v_star + dt * ay;
  }
  axn = 0.; ayn = 0.;                                              // I:2283-2296 with Runge_not_Verlet, C_N = 0
  if (INTERACTIVE && p.only_interactive_forces) { axn = ia.IA_x; ayn = ia.IA_y; }
  bxn = ax - (axn / 2); byn = ay - (ayn / 2);
  if ((p.speed_limit > 0.) || (p.speed_limit == -1.)) {
    double speed = sqrt(uveln * uveln + vveln * vveln);
    if (speed > 0.) { double new_speed = loc_dx / dt * p.speed_limit; if (new_speed < speed && p.speed_limit > 0.) speeding = true; }
  }
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

// rotpos_to_tang / rotpos_from_tang / rotvec_to_tang / rotvec_from_tang I:7767-7816, I:8066-8099
__device__ __noinline__ void rotpos_to_tang(const DevParams& p, double lon, double lat, double& x, double& y) {
  double colat = 90. - lat, r = p.Rearth * (colat * p.pi_180);
  x = r * cos(lon * p.pi_180); y = r * sin(lon * p.pi_180);
}
__device__ __noinline__ void rotpos_from_tang(const DevParams& p, double x, double y, double& lon, double& lat) {
  double r = sqrt(x * x + y * y), r180_pi = 180. / p.pi;
  lat = 90. - (r180_pi * r / p.Rearth);
  lon = r180_pi * acos(x / r) * f_sign1(y);
}
__device__ __noinline__ void rotvec_to_tang(const DevParams& p, double lon, double uvel, double vvel, double& xdot, double& ydot) {
  double clon = cos(lon * p.pi_180), slon = sin(lon * p.pi_180);
  xdot = -slon * uvel - clon * vvel; ydot = clon * uvel - slon * vvel;
}
__device__ __noinline__ void rotvec_from_tang(const DevParams& p, double lon, double xdot, double ydot, double& uvel, double& vvel) {
  double clon = cos(lon * p.pi_180), slon = sin(lon * p.pi_180);
  uvel = -slon * xdot + clon * ydot; vvel = -clon * xdot - slon * ydot;
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
template <bool LEAN = false>
__device__ __forceinline__ void rolling(const DevParams& p, double& Tn, double& Wn, double& Ln) {
  const double Delta = 6.0;
  double Dn = p.rho_ratio * Tn;
  if (Dn > 0.) {
    if ((!PF(use_updated_rolling_scheme, 0)) && (LEAN || p.tip_parameter < 999.)) {
      // max(W,L) < sqrt(X): decided on the squares unless the two sides agree to 1e-12, where the
      // reference's own sqrt-then-compare is evaluated
      double mx = kmax(Wn, Ln), X = 0.92 * (Dn * Dn) + 58.32 * Dn, m2 = mx * mx;
      bool tip = m2 < X * (1. - 1.e-12);
      if (!tip && m2 <= X * (1. + 1.e-12)) tip = mx < sqrt(X);
      if (tip) {
        swap_d(Tn, Wn);
        if (Wn > Ln) swap_d(Wn, Ln);
      }
    } else {
      if (Wn > Ln) swap_d(Ln, Wn);
      if ((!PF(use_updated_rolling_scheme, 0)) && (!LEAN && p.tip_parameter >= 999.)) {
        double q = p.rho_bergs / KID_RHO_SEAWATER;
        if (Wn < sqrt((6.0 * q * (1 - q) * (Tn * Tn)) - (12 * Delta * q * Tn))) {
          swap_d(Tn, Wn);
          if (Wn > Ln) swap_d(Wn, Ln);
        }
      }
      if (PF(use_updated_rolling_scheme, 0)) {
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

// find_basal_melt I:3492-3826 (+ calculate_TFreeze I:3790, calculate_density I:3805): ice-shelf style
// two-/three-equation basal melt.  Out of line: only the use_mixed_melting / melt_icebergs_as_ice_shelf
// namelists reach it.
__device__ __noinline__ double find_basal_melt(const DevParams& p, double dvo, double lat, double salt, double temp,
                                               double thickness) {
  const double VK = 0.40, ZETA_N = 0.052, RC = 0.20, c2_3 = 2.0 / 3.0;
  const double dR0_dT = -0.038357, dR0_dS = 0.805876, RHO_T0_S0 = 999.910681, Salin_Ice = 0.0;
  const double kd_molec_salt = 8.02e-10, kd_molec_temp = 1.41e-7, kv_molec = 1.95e-6;
  const double Cp_ml = 3974.0, LF = 3.335e5, p_atm = 101325;
  const double dTFr_dp = -7.53E-08, dTFr_dS = -0.0573, TFr_S0_P0 = 0.0832;
  double density_ice = p.rho_bergs, Rho0 = KID_RHO_SEAWATER, Hml = 10.;
  double p_int = p_atm + (KID_GRAVITY * thickness * density_ice);
  double Rhoml = RHO_T0_S0 + dR0_dT * temp + dR0_dS * salt;
  double I_ZETA_N = 1.0 / ZETA_N, I_LF = 1.0 / LF;
  double SC = kv_molec / kd_molec_salt, PR = kv_molec / kd_molec_temp, I_VK = 1.0 / VK;
  double RhoCp = Rho0 * Cp_ml;
  double Gam_mol_t = 12.5 * pow(PR, c2_3) - 6, Gam_mol_s = 12.5 * pow(SC, c2_3) - 6;
  double ustar = sqrt(p.cdrag_icebergs * (dvo * dvo + p.utide_icebergs * p.utide_icebergs));
  double ustar_h = fmax(p.ustar_icebergs_bg, ustar);
  double pi_180 = p.pi / 180., f_cori;
  if (p.grid_is_latlon && !p.use_f_plane) f_cori = (2. * p.omega) * sin(pi_180 * lat);
  else f_cori = (2. * p.omega) * sin(pi_180 * p.lat_ref);
  double absf = fabs(f_cori), hBL_neut;
  if ((absf * Hml <= VK * ustar_h) || (absf == 0.)) hBL_neut = Hml; else hBL_neut = (VK * ustar_h) / absf;
  double hBL_neut_h_molec = ZETA_N * ((hBL_neut * ustar_h) / (5.0 * kv_molec));
  double ln_neut = 0.0; if (hBL_neut_h_molec > 1.0) ln_neut = log(hBL_neut_h_molec);
  double tfreeze, Gam_turb, I_Gam_T = 0., I_Gam_S = 0., wT_flux, t_flux, lprec = 0.;
  bool out_of_bounds = false;
  if (p.use_three_equation_model) {
    double Sbdry = salt, Sb_max = 0., Sb_min = 0.;
    bool Sb_max_set = false, Sb_min_set = false;
    double dB_dS = (KID_GRAVITY / Rhoml) * dR0_dS, dB_dT = (KID_GRAVITY / Rhoml) * dR0_dT;
    for (int it1 = 1; it1 <= 20; it1++) {
      tfreeze = (TFr_S0_P0 + dTFr_dS * Sbdry) + dTFr_dp * p_int;
      double dT_ustar = (temp - tfreeze) * ustar_h, dS_ustar = (salt - Sbdry) * ustar_h;
      if (p.const_gamma) { I_Gam_T = p.gamma_t_3eq; I_Gam_S = p.gamma_t_3eq / 35.; }
      else {
        Gam_turb = I_VK * (ln_neut + (0.5 * I_ZETA_N - 1.0));
        I_Gam_T = 1.0 / (Gam_mol_t + Gam_turb); I_Gam_S = 1.0 / (Gam_mol_s + Gam_turb);
      }
      wT_flux = dT_ustar * I_Gam_T;
      double wB_flux = dB_dS * (dS_ustar * I_Gam_S) + dB_dT * wT_flux;
      if (wB_flux > 0.0) {
        double n_star_term = (ZETA_N / RC) * (hBL_neut * VK) / pow(ustar_h, 3.);
        // the reference's inner loop (it3 = 1,30) never feeds its Newton estimate back into wB_flux:
        // every pass recomputes the same numbers, so one pass gives its result
        double I_n_star = sqrt(1.0 + n_star_term * wB_flux);
        if (hBL_neut_h_molec > I_n_star * I_n_star) Gam_turb = I_VK * ((ln_neut - 2.0 * log(I_n_star)) + (0.5 * I_ZETA_N * I_n_star - 1.0));
        else Gam_turb = I_VK * (0.5 * I_ZETA_N * I_n_star - 1.0);
        if (p.const_gamma) { I_Gam_T = p.gamma_t_3eq; I_Gam_S = p.gamma_t_3eq / 35.; }
        else { I_Gam_T = 1.0 / (Gam_mol_t + Gam_turb); I_Gam_S = 1.0 / (Gam_mol_s + Gam_turb); }
        wT_flux = dT_ustar * I_Gam_T;
      }
      t_flux = RhoCp * wT_flux;
      double exch_vel_s = ustar_h * I_Gam_S;
      lprec = I_LF * t_flux;
      double mass_exch = exch_vel_s * Rho0;
      double Sbdry_it = (salt * mass_exch + Salin_Ice * lprec) / (mass_exch + lprec);
      double dS_it = Sbdry_it - Sbdry;
      if (fabs(dS_it) < 1e-4 * (0.5 * (salt + Sbdry + 1.e-10))) break;
      if (dS_it < 0.0) {
        if (Sb_max_set && (Sbdry > Sb_max)) { out_of_bounds = true; break; }
        Sb_max = Sbdry; Sb_max_set = true;
      } else {
        if (Sb_min_set && (Sbdry < Sb_min)) { out_of_bounds = true; break; }
        Sb_min = Sbdry; Sb_min_set = true;
      }
      Sbdry = Sbdry_it;                                  // I:3758 overwrites the false-position estimate
    }
  }
  if ((!p.use_three_equation_model) || out_of_bounds) {
    tfreeze = (TFr_S0_P0 + dTFr_dS * salt) + dTFr_dp * p_int;
    Gam_turb = I_VK * (ln_neut + (0.5 * I_ZETA_N - 1.0));
    I_Gam_T = 1.0 / (Gam_mol_t + Gam_turb);
    double exch_vel_t = ustar_h * I_Gam_T;
    wT_flux = exch_vel_t * (temp - tfreeze);
    t_flux = RhoCp * wT_flux;
    lprec = I_LF * t_flux;
  }
  return lprec / density_ice;
}

// what the ice-shelf melt options additionally need of the berg and its cell (I:2945-2956)
struct ShelfIn { double lat, sss, ocean_depth; };

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
  double Mb_fl = kmax(0.58 * dvo08 * (SST + 4.0) / pow(Lfl, 0.2), 0.) * perday;
  double Tnfl = kmax(Tfl - Mb_fl * dt, 0.);
  double Lnfl, Wnfl, nVolfl, Mnew_fl;
  if (p.use_operator_splitting) {
    nVolfl = Tnfl * Wfl * Lfl;
    double Mnew1_fl = (nVolfl / Volfl) * Mfl;
    f.dMb_fl = Mfl - Mnew1_fl;
    Lnfl = kmax(Lfl - Mv_fl * dt, 0.);
    Wnfl = kmax(Wfl - Mv_fl * dt, 0.);
    nVolfl = Tnfl * Wnfl * Lnfl;
    double Mnew2_fl = (nVolfl / Volfl) * Mfl;
    f.dMv_fl = Mnew1_fl - Mnew2_fl;
    Lnfl = kmax(Lnfl - Me_fl * dt, 0.);
    Wnfl = kmax(Wnfl - Me_fl * dt, 0.);
    nVolfl = Tnfl * Wnfl * Lnfl;
    Mnew_fl = (nVolfl / Volfl) * Mfl;
    f.dMe_fl = Mnew2_fl - Mnew_fl;
  } else {
    Lnfl = kmax(Lfl - (Mv_fl + Me_fl) * dt, 0.);
    Wnfl = kmax(Wfl - (Mv_fl + Me_fl) * dt, 0.);
    nVolfl = Tnfl * Wnfl * Lnfl;
    Mnew_fl = (nVolfl / Volfl) * Mfl;
    f.dMb_fl = (Mfl / Volfl) * (Wfl * Lfl) * Mb_fl * dt;
    f.dMe_fl = (Mfl / Volfl) * (Tfl * (Wfl + Lfl)) * Me_fl * dt;
    f.dMv_fl = (Mfl / Volfl) * (Tfl * (Wfl + Lfl)) * Mv_fl * dt;
  }
  f.Lnfl = Lnfl; f.Wnfl = Wnfl; f.Tnfl = Tnfl; f.Mnew_fl = Mnew_fl;
  f.dMfl = Mfl - Mnew_fl;
}

// x**(-0.2) for x > 0 (I:2917 L**0.2, dvo**0.8 = dvo*dvo**(-0.2)): single-precision seed
// (MUFU lg2/ex2, ~1e-6 relative) and two Newton steps on y^-5 = x, y <- y*(6 - x*y^5)/5, which
// converge quadratically to fp64 rounding (~3e-16 relative).
__device__ __forceinline__ double pow_m02(double x) {
#ifdef KID_LIBM_POW
  return exp(-0.2 * log(x));
#else
  float xf = (float)x, lf, ef;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lf) : "f"(xf));
  lf = lf * -0.2f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ef) : "f"(lf));
  double y = (double)ef;
#pragma unroll
  for (int it = 0; it < 2; it++) {
    double y2 = y * y;
    double y5 = y2 * y2 * y;
    y = (y * 0.2) * fma(-x, y5, 6.0);
  }
  return y;
#endif
}
// x**0.8 for x >= 0
__device__ __forceinline__ double pow_08(double x) { return (x > 1.e-30) ? x * pow_m02(x) : 0.; }

// thermodynamics of one berg, I:2896-3296.  e.rarea = 1/grd%area(i,j) (caller has checked the
// cell is not dry), N_bonds per I:2928-2944.
template <bool LEAN = false>
__device__ __forceinline__ int thermo_berg(const DevParams& p, const EnvThermo& e, double uvel, double vvel,
                                           double N_bonds, ThermoState& s, ThermoFlux& fx, const ShelfIn& sh) {
  const double perday = 1. / 86400.;
  double dt = p.dt;
  double SST = e.sst;
  double IC = kmin(1., e.cn + p.sicn_shift);
  double M = s.mass, T = s.thickness, W = s.width, L = s.length;
  double Vol = T * W * L;
  double M_Vol = M / Vol;
  double dvo = KID_HYPOT(uvel - e.uo, vvel - e.vo);
  double dva = KID_HYPOT(e.ua - e.uo, e.va - e.vo);
  double Ss = fma(1.5, sqrt_nr(dva), 0.1 * dva);           // dva**0.5
  double dvo08 = pow_08(dvo);
  double Mv = kmax(7.62e-3 * SST + 1.29e-3 * (SST * SST), 0.) * perday;
  double Mb = kmax(0.58 * dvo08 * (SST + 4.0) * pow_m02(L), 0.) * perday;
  double IC3 = IC * IC * IC;
  double wave_ic = (IC3 == 0.) ? 2. : (1 + cos(p.pi * IC3));   // cos(0) = 1
  double Me = kmax(1. / 12. * (SST + 2.) * Ss * wave_ic, 0.) * perday;
  double Mv_fl = 0., Me_fl = 0.;
  if (s.mass_of_fl_bits > 0.) { Mv_fl = Mv; Me_fl = Me; }
  if (!LEAN && (p.melt_icebergs_as_ice_shelf || p.use_mixed_melting)) {      // I:2945-2968
    double SSS = p.use_mixed_layer_salinity_for_thermo ? sh.sss : 35.0;
    double Ms = kmax(find_basal_melt(p, dvo, sh.lat, SSS, SST, T), 0.);
    if ((p.melt_cutoff >= 0.) && p.apply_thickness_cutoff_to_bergs_melt) {
      double Dn = (p.rho_bergs / KID_RHO_SEAWATER) * T;
      if ((sh.ocean_depth - Dn) < p.melt_cutoff) Ms = 0.;
    }
    if (p.use_mixed_melting) {
      double N_max = p.hexagonal_icebergs ? 6.0 : 4.0;
      Me = ((N_max - N_bonds) / N_max) * (Mv + Me);
      Mv = 0.0;
      Mb = (((N_max - N_bonds) / N_max) * (Mb)) + (N_bonds / N_max) * Ms;
    } else { Mv = 0.0; Me = 0.0; Mb = Ms; }
  }
  if (PF(set_melt_rates_to_zero, 0)) { Mv = 0.0; Mb = 0.0; Me = 0.0; }
  double Tn, Mnew1, Mnew2, Mnew, dMb, dMv, dMe, dM, Ln1 = 0, Wn1 = 0, Ln, Wn;
  if (PF(use_operator_splitting, 1)) {
    // the mass differences below cancel to ~1e-6 of the mass: each (nVol/Vol)*M keeps the
    // reference's own rounding sequence so that dMe (hence the bergy bits) agrees to 1e-10
    Tn = kmax(T - Mb * dt, 0.);
    Mnew1 = ((Tn * W * L) / Vol) * M;
    dMb = M - Mnew1;
    Ln1 = kmax(L - Mv * dt, 0.);
    Wn1 = kmax(W - Mv * dt, 0.);
    Mnew2 = ((Tn * Wn1 * Ln1) / Vol) * M;
    dMv = Mnew1 - Mnew2;
    Ln = kmax(Ln1 - Me * dt, 0.);
    Wn = kmax(Wn1 - Me * dt, 0.);
    Mnew = ((Tn * Wn * Ln) / Vol) * M;
    dMe = Mnew2 - Mnew;
    dM = M - Mnew;
  } else {
    Ln = kmax(L - (Mv + Me) * (dt), 0.);
    Wn = kmax(W - (Mv + Me) * (dt), 0.);
    Tn = kmax(T - Mb * (dt), 0.);
    Mnew = ((Tn * Wn * Ln) / Vol) * M;
    dM = M - Mnew;
    dMb = M_Vol * (W * L) * Mb * dt;
    dMe = M_Vol * (T * (W + L)) * Me * dt;
    dMv = M_Vol * (T * (W + L)) * Mv * dt;
  }
  if (PF(footloose, 0)) {
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
    double Lbits = kmin(kmin(L, W), kmin(T, 40.));
    // Abits=(Mbits/rho)/Lbits; Mbb=rho*Abits*rate: the bergy-bit melt in kg/s, rate ~ Lbits**-0.2
    double rpow, rLbits;
    if (Lbits == 40.) { rpow = p.rpow40_02; rLbits = 1. / 40.; }
    else { rpow = pow_m02(Lbits); rLbits = 1. / Lbits; }
    double Abits = (Mbits * p.r_rho_bergs) * rLbits;
    double Mbb = kmax(0.58 * dvo08 * (SST + 2.0) * rpow, 0.) * perday;
    Mbb = p.rho_bergs * Abits * Mbb;
    dMbitsM = kmin(Mbb * dt, nMbits);
    nMbits = nMbits - dMbitsM;
    if (Mnew == 0.) { dMbitsM = dMbitsM + nMbits; nMbits = 0.; }
    if (has_fl) {
      double Mbits_fl = s.mass_of_fl_bergy_bits;
      dMbitsE_fl = p.bergy_bit_erosion_fraction * fl.dMe_fl;
      nMbits_fl = Mbits_fl + dMbitsE_fl;
      double Lbits_fl = kmin(kmin(fl.Lfl, fl.Wfl), kmin(fl.Tfl, 40.));
      double Abits_fl = (Mbits_fl / p.rho_bergs) / Lbits_fl;
      double Mbb_fl = kmax(0.58 * dvo08 * (SST + 2.0) / pow(Lbits_fl, 0.2), 0.) * perday;
      Mbb_fl = p.rho_bergs * Abits_fl * Mbb_fl;
      dMbitsM_fl = kmin(Mbb_fl * dt, nMbits_fl);
      nMbits_fl = nMbits_fl - dMbitsM_fl;
      if (Mnew_fl == 0.) { dMbitsM_fl = dMbitsM_fl + nMbits_fl; nMbits_fl = 0.; }
    } else {
      dMbitsE_fl = 0.; dMbitsM_fl = 0.; nMbits_fl = 0.;
    }
  } else {
    dMbitsE = 0.; dMbitsM = 0.; nMbits = s.mass_of_bits;
    dMbitsE_fl = 0.; dMbitsM_fl = 0.; nMbits_fl = s.mass_of_fl_bergy_bits;
  }
  // grid contributions I:3116-3199: X/dt/area*mass_scaling = X*k
  {
    double ms = s.mass_scaling;
    double k = p.rdt * e.rarea * ms;
    double melt = (dM - (dMbitsE - dMbitsM) + dMfl - (dMbitsE_fl - dMbitsM_fl));
    fx.floating_melt = melt * k;
    double hk = s.heat_density * k;
    fx.calving_hflx = melt * hk;
    fx.net_heat = melt * s.heat_density * ms;              // melt/dt*heat_density*ms*dt
    fx.berg_melt = dM * k;
    fx.bergy_src = (dMbitsE + dMbitsE_fl) * k;
    fx.bergy_melt = (dMbitsM + dMbitsM_fl) * k;
    fx.fl_bits_melt = dMfl * k;
    fx.fl_parent_melt = fx.fl_child_melt = fx.melt_buoy = fx.melt_eros = fx.melt_conv = 0.;
    fx.melt_buoy_fl = fx.melt_eros_fl = fx.melt_conv_fl = 0.;
    if (PF(melt_diagnostics, 0)) {
      if (s.fl_k >= 0) {
        fx.fl_parent_melt = (dM - (dMbitsE - dMbitsM)) * k;
        fx.fl_child_melt = (dMfl - (dMbitsE_fl - dMbitsM_fl)) * k;
        fx.melt_buoy = dMb * k; fx.melt_eros = dMe * k; fx.melt_conv = dMv * k;
        if (dMfl > 0) { fx.melt_buoy_fl = fl.dMb_fl * k; fx.melt_eros_fl = fl.dMe_fl * k; fx.melt_conv_fl = fl.dMv_fl * k; }
      } else {
        fx.fl_child_melt = (dM - (dMbitsE - dMbitsM)) * k;
        fx.melt_buoy_fl = dMb * k; fx.melt_eros_fl = dMe * k; fx.melt_conv_fl = dMv * k;
      }
    }
  }
  fx.fl_bits_src = 0.;
  if (PF(allow_bergs_to_roll, 1) && N_bonds == 0.) rolling<LEAN>(p, Tn, Wn, Ln);
  if (PF(iceberg_melt_without_decay, 0)) {
    Mnew = s.mass; Mnew_fl = s.mass_of_fl_bits; nMbits_fl = s.mass_of_fl_bergy_bits;
  } else {
    s.mass = Mnew;
    s.mass_of_bits = nMbits;
    s.mass_of_fl_bits = Mnew_fl;
    s.mass_of_fl_bergy_bits = nMbits_fl;
    s.thickness = Tn;
    s.width = kmin(Wn, Ln);
    s.length = kmax(Wn, Ln);
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
      fx.fl_bits_src = -(s.mass * s.mass_scaling * p.rdt * e.rarea);
      return TH_BECAME_FL;
    }
    return TH_DELETE;
  }
  return TH_KEEP;
}

}  // namespace kid
