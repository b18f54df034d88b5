// kid_geom.cuh -- cell geometry on the device.
//
// Follows the reference routine by routine (F: = src/icebergs_framework.F90):
//   apply_modulo_around_point F:6558, sum_sign_dot_prod4/5 F:6163/F:6231,
//   is_point_in_cell F:6076, calc_xiyj F:6439, is_point_within_xi_yj_bounds F:6540,
//   pos_within_cell F:6299, bilin F:7071, find_cell(_wide) F:6011/F:6044.
// The operation order of every expression that feeds an integer decision (cell
// membership, xi/yj clamps) is the reference's, and those expressions are spelled
// with the __d*_rn intrinsics, which the compiler never contracts into FMAs: the
// branch outcomes are the ones the plain IEEE sequence of the CPU path takes even
// though the rest of the library is compiled with FMA contraction on.
#pragma once
#include "kid_device.cuh"

namespace kid {

// linear index of cell (i,j) in the data domain (32-bit: kid_init refuses grids of 2^31 cells)
__device__ __forceinline__ int gidx(const DevGrid& g, int i, int j) {
  return (i - g.isd) + (j - g.jsd) * g.nid;
}

__device__ __forceinline__ double f_sign1(double b) { return signbit(b) ? -1. : 1.; }

// never-contracted arithmetic
#define KMUL(a, b) __dmul_rn((a), (b))
#define KADD(a, b) __dadd_rn((a), (b))
#define KSUB(a, b) __dsub_rn((a), (b))

// Fortran MODULO(a,p) for p>0.  The in-range case (every berg that is within half
// a period of the reference point) is exact and costs two compares.
__device__ __noinline__ double f_modulo_slow(double a, double p) {
  double r = fmod(a, p);
  if (r != 0. && r < 0.) r += p;
  return r;
}
__device__ __forceinline__ double f_modulo(double a, double p) {
  if (a >= 0. && a < p) return a;
  return f_modulo_slow(a, p);
}

// F:6558-6573
__device__ __forceinline__ double amap(double x, double y, double Lx) {
  if (Lx > 0.) {
    double Lx_2 = Lx * 0.5;
    double yy = KSUB(y, Lx_2);
    return KADD(f_modulo(KSUB(x, yy), Lx), yy);
  }
  return x;
}

// F:6163-6228.  Inside a counter-clockwise cell all four cross products are negative, so the
// zero-valued sentinels at F:6203-6206 make the S (p0) and W (p3) edges part of the cell and the
// N and E edges not (the reference's comment says "South and East"; the code is followed)
__device__ __forceinline__ bool sum_sign_dot_prod4(double x0, double y0, double x1, double y1,
                                                   double x2, double y2, double x3, double y3,
                                                   double x, double y, double Lx) {
  double xx = amap(x, x0, Lx);
  double xx0 = amap(x0, x0, Lx), xx1 = amap(x1, x0, Lx), xx2 = amap(x2, x0, Lx), xx3 = amap(x3, x0, Lx);
  double l0 = KSUB(KMUL(KSUB(xx, xx0), KSUB(y1, y0)), KMUL(KSUB(y, y0), KSUB(xx1, xx0)));
  double l1 = KSUB(KMUL(KSUB(xx, xx1), KSUB(y2, y1)), KMUL(KSUB(y, y1), KSUB(xx2, xx1)));
  double l2 = KSUB(KMUL(KSUB(xx, xx2), KSUB(y3, y2)), KMUL(KSUB(y, y2), KSUB(xx3, xx2)));
  double l3 = KSUB(KMUL(KSUB(xx, xx3), KSUB(y0, y3)), KMUL(KSUB(y, y3), KSUB(xx0, xx3)));
  double p0 = f_sign1(l0); if (l0 == 0.) p0 = -0.5;
  double p1 = f_sign1(l1); if (l1 == 0.) p1 = 0.5;
  double p2 = f_sign1(l2); if (l2 == 0.) p2 = 0.5;
  double p3 = f_sign1(l3); if (l3 == 0.) p3 = -0.5;
  return (fabs(p0) + fabs(p2)) + (fabs(p1) + fabs(p3)) == fabs((p0 + p2) + (p1 + p3));
}

// F:6231-6296
__device__ __noinline__ bool sum_sign_dot_prod5(double x0, double y0, double x1, double y1, double x2,
                                                double y2, double x3, double y3, double x4, double y4,
                                                double x, double y, double Lx) {
  double xx = amap(x, x0, Lx);
  double xx0 = amap(x0, x0, Lx), xx1 = amap(x1, x0, Lx), xx2 = amap(x2, x0, Lx), xx3 = amap(x3, x0, Lx),
         xx4 = amap(x4, x0, Lx);
  double l0 = KSUB(KMUL(KSUB(xx, xx0), KSUB(y1, y0)), KMUL(KSUB(y, y0), KSUB(xx1, xx0)));
  double l1 = KSUB(KMUL(KSUB(xx, xx1), KSUB(y2, y1)), KMUL(KSUB(y, y1), KSUB(xx2, xx1)));
  double l2 = KSUB(KMUL(KSUB(xx, xx2), KSUB(y3, y2)), KMUL(KSUB(y, y2), KSUB(xx3, xx2)));
  double l3 = KSUB(KMUL(KSUB(xx, xx3), KSUB(y4, y3)), KMUL(KSUB(y, y3), KSUB(xx4, xx3)));
  double l4 = KSUB(KMUL(KSUB(xx, xx4), KSUB(y0, y4)), KMUL(KSUB(y, y4), KSUB(xx0, xx4)));
  double p0 = f_sign1(l0); if (l0 == 0.) p0 = 0.;
  double p1 = f_sign1(l1); if (l1 == 0.) p1 = 0.;
  double p2 = f_sign1(l2); if (l2 == 0.) p2 = 0.;
  double p3 = f_sign1(l3); if (l3 == 0.) p3 = 0.;
  double p4 = f_sign1(l4); if (l4 == 0.) p4 = 0.;
  return ((fabs(p0) + fabs(p2)) + (fabs(p1) + fabs(p3))) + fabs(p4) -
             fabs(((p0 + p2) + (p1 + p3)) + p4) < 0.5;
}

// The four corners of cell (i,j): 1=SW (i-1,j-1), 2=SE (i,j-1), 3=NE (i,j), 4=NW (i-1,j)
struct Quad { double x1, y1, x2, y2, x3, y3, x4, y4; };

__device__ __forceinline__ Quad load_quad(const DevGrid& g, int i, int j) {
  const LonLat* __restrict__ ll = g.lonlat;
  int ne = gidx(g, i, j);
  LonLat c3 = ll[ne], c4 = ll[ne - 1], c2 = ll[ne - g.nid], c1 = ll[ne - g.nid - 1];
  Quad q;
  q.x1 = c1.lon; q.y1 = c1.lat; q.x2 = c2.lon; q.y2 = c2.lat;
  q.x3 = c3.lon; q.y3 = c3.lat; q.x4 = c4.lon; q.y4 = c4.lat;
  return q;
}

__device__ __forceinline__ bool cell_on_pe(const DevGrid& g, int i, int j) {
  return !(i - 1 < g.isd || i > g.ied || j - 1 < g.jsd || j > g.jed);
}

// F:6076-6160 on an already loaded quad
__device__ __forceinline__ bool is_point_in_quad(const Quad& q, const DevParams& p, double x, double y) {
  double Lx = p.Lx;
  double a = amap(q.x1, x, Lx), b = amap(q.x2, x, Lx), c = amap(q.x4, x, Lx), e = amap(q.x3, x, Lx);
  double xlo = fmin(fmin(a, b), fmin(c, e));
  double xhi = fmax(fmax(a, b), fmax(c, e));
  const double tol = 0.1;
  if (x < (xlo - tol) || x > (xhi + tol)) return false;
  double ylo = fmin(fmin(q.y1, q.y2), fmin(q.y4, q.y3));
  double yhi = fmax(fmax(q.y1, q.y2), fmax(q.y4, q.y3));
  if (y < ylo || y > yhi) return false;
  if (p.grid_is_latlon) {
    if (q.y3 > 89.999) return sum_sign_dot_prod5(q.x1, q.y1, q.x2, q.y2, q.x2, q.y3, q.x4, q.y3, q.x4, q.y4, x, y, Lx);
    else if (q.y4 > 89.999) return sum_sign_dot_prod5(q.x1, q.y1, q.x2, q.y2, q.x3, q.y3, q.x3, q.y4, q.x1, q.y4, x, y, Lx);
    else if (q.y1 > 89.999) return sum_sign_dot_prod5(q.x4, q.y1, q.x2, q.y1, q.x2, q.y2, q.x3, q.y3, q.x4, q.y4, x, y, Lx);
    else if (q.y2 > 89.999) return sum_sign_dot_prod5(q.x1, q.y1, q.x1, q.y2, q.x3, q.y2, q.x3, q.y3, q.x4, q.y4, x, y, Lx);
  }
  return sum_sign_dot_prod4(q.x1, q.y1, q.x2, q.y2, q.x3, q.y3, q.x4, q.y4, x, y, Lx);
}

__device__ __forceinline__ bool is_point_in_cell(const DevGrid& g, const DevParams& p, double x, double y,
                                                 int i, int j, unsigned int* err) {
  if (!cell_on_pe(g, i, j)) { atomicOr(err, (unsigned)KID_DEVERR_OFF_PE); return false; }
  Quad q = load_quad(g, i, j);
  return is_point_in_quad(q, p, x, y);
}

// F:6439-6534
__device__ __forceinline__ void calc_xiyj(double x1, double x2, double x3, double x4, double y1, double y2,
                                          double y3, double y4, double x, double y, double* xi, double* yj,
                                          double Lx, unsigned int* err) {
  double alpha = KSUB(x2, x1), delta = KSUB(y2, y1), beta = KSUB(x4, x1), epsilon = KSUB(y4, y1);
  double gamma = KSUB(KSUB(x3, x1), KADD(alpha, beta));
  double kappa = KSUB(KSUB(y3, y1), KADD(delta, epsilon));
  double a = KSUB(KMUL(kappa, beta), KMUL(gamma, epsilon));
  double dx = KSUB(amap(x, x1, Lx), x1);
  double dy = KSUB(y, y1);
  double b = KSUB(KSUB(KMUL(delta, beta), KMUL(alpha, epsilon)), KSUB(KMUL(kappa, dx), KMUL(gamma, dy)));
  double c = KSUB(KMUL(alpha, dy), KMUL(delta, dx));
  double yy;
  if (fabs(a) > 1.e-12) {
    double d = KSUB(KMUL(0.25, KMUL(b, b)), KMUL(a, c));
    if (d >= 0.) {
      double sq = sqrt(d);
      double yy1 = -KADD(KMUL(0.5, b), sq) / a;
      double yy2 = -KSUB(KMUL(0.5, b), sq) / a;
      yy = (fabs(KSUB(yy1, 0.5)) < fabs(KSUB(yy2, 0.5))) ? yy1 : yy2;
    } else {
      atomicOr(err, (unsigned)KID_DEVERR_COMPLEX_ROOTS);
      yy = 0.;
    }
  } else {
    yy = (b != 0.) ? -c / b : 0.;
  }
  a = KADD(alpha, KMUL(gamma, yy));
  b = KADD(delta, KMUL(kappa, yy));
  double xx;
  if (a != 0.) {
    xx = KSUB(dx, KMUL(beta, yy)) / a;
  } else if (b != 0.) {
    xx = KSUB(dy, KMUL(epsilon, yy)) / b;
  } else {
    c = KADD(KSUB(KMUL(epsilon, alpha), KMUL(beta, delta)), KMUL(KSUB(KMUL(epsilon, gamma), KMUL(beta, kappa)), yy));
    if (c != 0.) {
      xx = KSUB(KMUL(epsilon, dx), KMUL(beta, dy)) / c;
    } else {
      atomicOr(err, (unsigned)KID_DEVERR_NOT_INVERTIBLE);
      xx = 0.;
    }
  }
  *xi = xx; *yj = yy;
}

// F:6540-6552
__device__ __forceinline__ bool within_xi_yj_bounds(double xi, double yj) {
  return (xi >= 0 && xi < 1) && (yj >= 0 && yj < 1);
}

// polar tangent-plane branch of pos_within_cell, F:6359-6405 (kept out of line:
// only cells that touch the pole take it)
__device__ __noinline__ void polar_xiyj(const Quad& q, const DevParams& p, double x, double y, double* xi,
                                        double* yj, bool inside, unsigned int* err) {
  double pi_180 = p.pi_180;
  double xx = (90. - y) * cos(x * pi_180), yy = (90. - y) * sin(x * pi_180);
  double tx1 = (90. - q.y1) * cos(q.x1 * pi_180), ty1 = (90. - q.y1) * sin(q.x1 * pi_180);
  double tx2 = (90. - q.y2) * cos(q.x2 * pi_180), ty2 = (90. - q.y2) * sin(q.x2 * pi_180);
  double tx3 = (90. - q.y3) * cos(q.x3 * pi_180), ty3 = (90. - q.y3) * sin(q.x3 * pi_180);
  double tx4 = (90. - q.y4) * cos(q.x4 * pi_180), ty4 = (90. - q.y4) * sin(q.x4 * pi_180);
  calc_xiyj(tx1, tx2, tx3, tx4, ty1, ty2, ty3, ty4, xx, yy, xi, yj, p.Lx, err);
  if (inside) {
    if (!within_xi_yj_bounds(*xi, *yj)) {
      double fac = 2.1 * fmax(fabs(*xi - 0.5), fabs(*yj - 0.5));
      fac = fmax(1., fac);
      *xi = 0.5 + (*xi - 0.5) / fac;
      *yj = 0.5 + (*yj - 0.5) / fac;
    }
  }
}

// exact membership test kept out of line: pos_within_cell only needs it for points within
// KID_EDGE_BAND of a cell edge (see below)
__device__ __noinline__ bool is_point_in_quad_ool(const Quad& q, const DevParams& p, double x, double y) {
  return is_point_in_quad(q, p, x, y);
}

#define KID_EDGE_BAND 1.e-6

// F:6299-6436.  Returns is_point_in_cell; xi,yj = -999 when (i,j) is off the PE.
// Membership (F:6076) is decided from (xi,yj) when the point is farther than KID_EDGE_BAND
// (in cell units) from every edge -- there the four cross products of sum_sign_dot_prod4 have
// the same sign (inside) or the crude bounds / a cross product rejects the point (outside) by a
// margin ~1e10 times the rounding error; only inside the band, and always in polar cells, is
// the reference's sign test evaluated.
__device__ __noinline__ bool pos_within_cell_general(const DevGrid& g, const DevParams& p, double x, double y,
                                                     int i, int j, double* xi, double* yj, unsigned int* err) {
  *xi = -999.; *yj = -999.;
  if (!cell_on_pe(g, i, j)) return false;
  Quad q = load_quad(g, i, j);
  double a, b;
  if ((!p.grid_is_latlon) && p.grid_is_regular) {
    double dx = fabs(KSUB(q.x3, q.x4));
    double dy = fabs(KSUB(q.y3, q.y2));
    double x1 = KSUB(q.x3, KMUL(dx, 0.5));
    double y1 = KSUB(q.y3, KMUL(dy, 0.5));
    double Delta_x = KSUB(amap(x, x1, p.Lx), x1);
    a = KADD(((Delta_x) / dx), 0.5);
    b = KADD((KSUB(y, y1) / dy), 0.5);
  } else if ((fmax(fmax(q.y1, q.y2), fmax(q.y3, q.y4)) < 89.999) || (!p.grid_is_latlon)) {
    calc_xiyj(q.x1, q.x2, q.x3, q.x4, q.y1, q.y2, q.y3, q.y4, x, y, &a, &b, p.Lx, err);
  } else {
    bool inside = is_point_in_quad_ool(q, p, x, y);
    polar_xiyj(q, p, x, y, xi, yj, inside, err);
    return inside;
  }
  *xi = a; *yj = b;
  const double lo = KID_EDGE_BAND, hi = 1. - KID_EDGE_BAND;
  if (a > lo && a < hi && b > lo && b < hi) return true;
  if (a < -lo || a > 1. + lo || b < -lo || b > 1. + lo) return false;
  return is_point_in_quad_ool(q, p, x, y);
}

// rectangle record of cell k (linear index) from the corner positions; launched once at init
__device__ __forceinline__ RectCell make_rect(const DevGrid& g, const DevParams& p, int i, int j) {
  RectCell r;
  r.x1 = 0.; r.y1 = 0.; r.reps = 0.;
  r.ralpha = __longlong_as_double(0x7ff8000000000000ll);
  if (!cell_on_pe(g, i, j)) return r;
  Quad q = load_quad(g, i, j);
  if ((!p.grid_is_latlon) && p.grid_is_regular) {
    double dx = fabs(KSUB(q.x3, q.x4)), dy = fabs(KSUB(q.y3, q.y2));
    if (!(dx > 0.) || !(dy > 0.)) return r;
    r.x1 = KSUB(q.x3, KMUL(dx, 0.5)); r.y1 = KSUB(q.y3, KMUL(dy, 0.5));
    r.ralpha = 1. / dx; r.reps = 1. / dy;
    return r;
  }
  if (!((fmax(fmax(q.y1, q.y2), fmax(q.y3, q.y4)) < 89.999) || (!p.grid_is_latlon))) return r;   // polar: F:6359
  double alpha = KSUB(q.x2, q.x1), delta = KSUB(q.y2, q.y1), beta = KSUB(q.x4, q.x1), epsilon = KSUB(q.y4, q.y1);
  double gamma = KSUB(KSUB(q.x3, q.x1), KADD(alpha, beta));
  double kappa = KSUB(KSUB(q.y3, q.y1), KADD(delta, epsilon));
  if (delta != 0. || beta != 0. || gamma != 0. || kappa != 0. || !(alpha > 0.) || !(epsilon > 0.)) return r;
  r.x1 = q.x1; r.y1 = q.y1;
  r.ralpha = 1. / alpha; r.reps = 1. / epsilon;
  return r;
}

// pos_within_cell with the rectangle shortcut in line and everything else out of line
__device__ __forceinline__ bool pos_within_cell(const DevGrid& g, const DevParams& p, double x, double y,
                                                int i, int j, double* xi, double* yj, unsigned int* err) {
  if (cell_on_pe(g, i, j)) {
    const RectCell rc = g.rect[gidx(g, i, j)];
    if (rc.ralpha == rc.ralpha) {
      double a = KADD(KMUL(KSUB(amap(x, rc.x1, p.Lx), rc.x1), rc.ralpha), p.rect_add);
      double b = KADD(KMUL(KSUB(y, rc.y1), rc.reps), p.rect_add);
      const double lo = KID_EDGE_BAND, hi = 1. - KID_EDGE_BAND;
      if (a > lo && a < hi && b > lo && b < hi) { *xi = a; *yj = b; return true; }
      if (a < -lo || a > 1. + lo || b < -lo || b > 1. + lo) { *xi = a; *yj = b; return false; }
    }
  }
  return pos_within_cell_general(g, p, x, y, i, j, xi, yj, err);
}

// F:7071-7088 on the corner positions
__device__ __forceinline__ void bilin_lonlat(const DevGrid& g, const DevParams& p, int i, int j, double xi,
                                             double yj, double* lon, double* lat) {
  Quad q = load_quad(g, i, j);
#define KB1(A, B, C, D, w1, w2, w3, w4) KADD(KMUL(KADD(KMUL(A, w1), KMUL(B, w2)), w3), KMUL(KADD(KMUL(C, w1), KMUL(D, w2)), w4))
  double xm = KSUB(1., xi), ym = KSUB(1., yj);
  if (p.old_bug_bilin) {
    *lon = KB1(q.x3, q.x4, q.x2, q.x1, xm, xi, ym, yj);
    *lat = KB1(q.y3, q.y4, q.y2, q.y1, xm, xi, ym, yj);
  } else {
    *lon = KB1(q.x3, q.x4, q.x2, q.x1, xi, xm, yj, ym);
    *lat = KB1(q.y3, q.y4, q.y2, q.y1, xi, xm, yj, ym);
  }
#undef KB1
}

// structured-grid guess shared by find_cell / find_cell_wide, F:6025-6026
__device__ __forceinline__ void guess_cell(const DevGrid& g, double x, double y, int* oi, int* oj) {
  double lon0 = g.lon[0], lat0 = g.lat[0];
  double lon1 = g.lon[(size_t)g.nid + 1], lat1 = g.lat[(size_t)g.nid + 1];
  double fi = floor((x - lon0) / (lon1 - lon0)), fj = floor((y - lat0) / (lat1 - lat0));
  // clamp before the int conversion (the reference's int() of a huge value is undefined)
  fi = fmin(fmax(fi, -1.e9), 1.e9); fj = fmin(fmax(fj, -1.e9), 1.e9);
  *oi = (int)fi + g.isd + 1;
  *oj = (int)fj + g.jsd + 1;
}

// F:6011-6041 (compute domain).  The fall-back scan is O(cells): callers use it for
// restart ingest / unpack only.
__device__ __noinline__ bool find_cell(const DevGrid& g, const DevParams& p, double x, double y, int* oi,
                                       int* oj, unsigned int* err) {
  guess_cell(g, x, y, oi, oj);
  if (*oi > g.isc - 1 && *oi < g.iec + 1 && *oj > g.jsc - 1 && *oj < g.jec + 1)
    if (is_point_in_cell(g, p, x, y, *oi, *oj, err)) return true;
  *oi = -999; *oj = -999;
  for (int j = g.jsc; j <= g.jec; j++)
    for (int i = g.isc; i <= g.iec; i++)
      if (is_point_in_cell(g, p, x, y, i, j, err)) { *oi = i; *oj = j; return true; }
  return false;
}

// F:6044-6073 (data domain)
__device__ __noinline__ bool find_cell_wide(const DevGrid& g, const DevParams& p, double x, double y, int* oi,
                                            int* oj, unsigned int* err) {
  guess_cell(g, x, y, oi, oj);
  if (cell_on_pe(g, *oi, *oj))
    if (is_point_in_cell(g, p, x, y, *oi, *oj, err)) return true;
  *oi = -999; *oj = -999;
  for (int j = g.jsd + 1; j <= g.jed; j++)
    for (int i = g.isd + 1; i <= g.ied; i++)
      if (is_point_in_cell(g, p, x, y, i, j, err)) { *oi = i; *oj = j; return true; }
  return false;
}

}  // namespace kid
