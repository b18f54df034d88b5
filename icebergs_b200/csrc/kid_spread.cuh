// kid_spread.cuh -- mass / area / momentum of the bergs spread onto the ocean grid:
// spread_mass_across_ocean_cells I:3895-4100 with the triangle / hexagon quadrant geometry I:4136-4670,
// sum_up_spread_fields I:6077-6150, the ustar and thickness-cutoff tail of create_gridded_icebergs_fields
// I:3461-3489.  The sign tests that decide on which side of an axis a corner lies use non-contracted
// arithmetic (the library is compiled with -fmad=false anyway).
#pragma once
#include "kid_physics.cuh"

namespace kid {

#define KH_MAX(a, b) ((a) > (b) ? (a) : (b))
#define KH_MIN(a, b) ((a) < (b) ? (a) : (b))
#define KH_MUL(a, b) __dmul_rn((a), (b))
#define KH_SUB(a, b) __dsub_rn((a), (b))
// Area_of_triangle I:4136-4145
__device__ __noinline__ double kh_area_of_triangle(double Ax, double Ay, double Bx, double By, double Cx, double Cy) {
  return fabs(0.5 * ((Ax * (By - Cy)) + (Bx * (Cy - Ay)) + (Cx * (Ay - By))));
}
// point_in_interval I:4158-4172
__device__ __noinline__ int kh_point_in_interval(double Ax, double Ay, double Bx, double By, double px, double py) {
  if ((px <= KH_MAX(Ax, Bx)) && (px >= KH_MIN(Ax, Bx)))
    if ((py <= KH_MAX(Ay, By)) && (py >= KH_MIN(Ay, By))) return 1;
  return 0;
}
// point_is_on_the_line I:4175-4197 (tol = 0)
__device__ __noinline__ int kh_point_is_on_the_line(double Ax, double Ay, double Bx, double By, double qx, double qy) {
  double dxc = qx - Ax, dyc = qy - Ay, dxl = Bx - Ax, dyl = By - Ay;
  double cross = KH_SUB(KH_MUL(dxc, dyl), KH_MUL(dyc, dxl));
  return fabs(cross) <= 0.0;
}
// point_in_triangle I:4203-4236
__device__ __noinline__ int kh_point_in_triangle(double Ax, double Ay, double Bx, double By, double Cx, double Cy, double qx, double qy) {
  if ((Ax == qx && Ay == qy) || (Bx == qx && By == qy) || (Cx == qx && Cy == qy)) return 0;
  if (kh_point_is_on_the_line(Ax, Ay, Bx, By, qx, qy) || kh_point_is_on_the_line(Ax, Ay, Cx, Cy, qx, qy) ||
      kh_point_is_on_the_line(Bx, By, Cx, Cy, qx, qy)) return 0;
  double l0 = KH_SUB(KH_MUL(qx - Ax, By - Ay), KH_MUL(qy - Ay, Bx - Ax));
  double l1 = KH_SUB(KH_MUL(qx - Bx, Cy - By), KH_MUL(qy - By, Cx - Bx));
  double l2 = KH_SUB(KH_MUL(qx - Cx, Ay - Cy), KH_MUL(qy - Cy, Ax - Cx));
  double p0 = (l0 == 0.) ? 0. : (signbit(l0) ? -1. : 1.);
  double p1 = (l1 == 0.) ? 0. : (signbit(l1) ? -1. : 1.);
  double p2 = (l2 == 0.) ? 0. : (signbit(l2) ? -1. : 1.);
  return ((fabs(p0) + fabs(p2)) + (fabs(p1)) == fabs((p0 + p2) + (p1)));
}
// intercept_of_a_line I:4282-4311; axis: 0 = 'x', 1 = 'y'
__device__ __noinline__ void kh_intercept_of_a_line(double Ax, double Ay, double Bx, double By, int axis, double* x0, double* y0) {
  const double No_intercept_val = 100000000000.;
  *x0 = No_intercept_val; *y0 = No_intercept_val;
  if (axis == 0) { if (Ay != By) { *x0 = Ax - (((Ax - Bx) / (Ay - By)) * Ay); *y0 = 0.; } }
  else { if (Ax != Bx) { *x0 = 0.; *y0 = -(((Ay - By) / (Ax - Bx)) * Ax) + Ay; } }
}
// Area_of_triangle_across_axes I:4244-4276
__device__ __noinline__ void kh_area_across_axes(double Ax, double Ay, double Bx, double By, double Cx, double Cy, int axis,
                              double* Area_positive, double* Area_negative) {
  double pABx, pABy, pACx, pACy;
  double A_triangle = kh_area_of_triangle(Ax, Ay, Bx, By, Cx, Cy);
  kh_intercept_of_a_line(Ax, Ay, Bx, By, axis, &pABx, &pABy);
  kh_intercept_of_a_line(Ax, Ay, Cx, Cy, axis, &pACx, &pACy);
  double A0 = (axis == 0) ? Ay : Ax;
  double A_half_triangle = kh_area_of_triangle(Ax, Ay, pABx, pABy, pACx, pACy);
  if (A0 >= 0.) { *Area_positive = A_half_triangle; *Area_negative = A_triangle - A_half_triangle; }
  else { *Area_positive = A_triangle - A_half_triangle; *Area_negative = A_half_triangle; }
}
// divding_triangle_across_axes I:4318-4394; returns 0 on the reference's 'logical error' FATALs
__device__ __noinline__ int kh_dividing_triangle(double Ax, double Ay, double Bx, double By, double Cx, double Cy, int axis,
                              double* Area_positive, double* Area_negative) {
  double A0, B0, C0;
  if (axis == 0) { A0 = Ay; B0 = By; C0 = Cy; } else { A0 = Ax; B0 = Bx; C0 = Cx; }
  double A_triangle = kh_area_of_triangle(Ax, Ay, Bx, By, Cx, Cy);
  if ((B0 * C0) > 0.) {
    if ((A0 * B0) >= 0.) {
      if ((A0 > 0.) || ((A0 == 0.) && (B0 > 0.))) { *Area_positive = A_triangle; *Area_negative = 0.; }
      else { *Area_positive = 0.; *Area_negative = A_triangle; }
    } else kh_area_across_axes(Ax, Ay, Bx, By, Cx, Cy, axis, Area_positive, Area_negative);
  } else if ((B0 * C0) < 0.) {
    if ((A0 * B0) >= 0.) kh_area_across_axes(Cx, Cy, Bx, By, Ax, Ay, axis, Area_positive, Area_negative);
    else kh_area_across_axes(Bx, By, Cx, Cy, Ax, Ay, axis, Area_positive, Area_negative);
  } else {
    if (((A0 == 0.) && (B0 == 0.)) && (C0 == 0.)) { *Area_positive = 0.; *Area_negative = 0.; }
    else if ((A0 * B0 < 0.) || (A0 * C0 < 0.)) kh_area_across_axes(Ax, Ay, Bx, By, Cx, Cy, axis, Area_positive, Area_negative);
    else if (((A0 * B0 > 0.) || (A0 * C0 > 0.)) || (((fabs(A0) > 0.) && (B0 == 0.)) && (C0 == 0.))) {
      if (A0 > 0.) { *Area_positive = A_triangle; *Area_negative = 0.; }
      else { *Area_positive = 0.; *Area_negative = A_triangle; }
    } else if (A0 == 0.) {
      if ((B0 > 0.) || (C0 > 0.)) { *Area_positive = A_triangle; *Area_negative = 0.; }
      else if ((B0 < 0.) || (C0 < 0.)) { *Area_positive = 0.; *Area_negative = A_triangle; }
      else { *Area_positive = 0.; *Area_negative = 0.; return 0; }
    } else { *Area_positive = 0.; *Area_negative = 0.; return 0; }
  }
  return 1;
}
// Triangle_divided_into_four_quadrants I:4399-4534; returns 0 on the reference's FATAL
__device__ __noinline__ int kh_triangle_into_quadrants(double Ax, double Ay, double Bx, double By, double Cx, double Cy,
                                    double* Area_triangle, double* Q1, double* Q2, double* Q3, double* Q4) {
  double Area_Upper, Area_Lower, Area_Right, Area_Left, px = 0., py = 0., qx = 0., qy = 0., Area_key_quadrant;
  int Key_quadrant = 0, ok = 1;
  *Area_triangle = kh_area_of_triangle(Ax, Ay, Bx, By, Cx, Cy);
  ok &= kh_dividing_triangle(Ax, Ay, Bx, By, Cx, Cy, 0, &Area_Upper, &Area_Lower);
  ok &= kh_dividing_triangle(Ax, Ay, Bx, By, Cx, Cy, 1, &Area_Right, &Area_Left);
  if (kh_point_in_triangle(Ax, Ay, Bx, By, Cx, Cy, 0., 0.)) {
    kh_intercept_of_a_line(Ax, Ay, Bx, By, 0, &px, &py);
    kh_intercept_of_a_line(Ax, Ay, Bx, By, 1, &qx, &qy);
    if (!(kh_point_in_interval(Ax, Ay, Bx, By, px, py) && kh_point_in_interval(Ax, Ay, Bx, By, qx, qy))) {
      kh_intercept_of_a_line(Ax, Ay, Cx, Cy, 0, &px, &py);
      kh_intercept_of_a_line(Ax, Ay, Cx, Cy, 1, &qx, &qy);
      if (!(kh_point_in_interval(Ax, Ay, Cx, Cy, px, py) && kh_point_in_interval(Ax, Ay, Cx, Cy, qx, qy))) {
        kh_intercept_of_a_line(Bx, By, Cx, Cy, 0, &px, &py);
        kh_intercept_of_a_line(Bx, By, Cx, Cy, 1, &qx, &qy);
        if (!(kh_point_in_interval(Bx, By, Cx, Cy, px, py) && kh_point_in_interval(Bx, By, Cx, Cy, qx, qy))) ok = 0;
      }
    }
    Area_key_quadrant = kh_area_of_triangle(px, py, qx, qy, 0., 0.);
    if ((px >= 0.) && (qy >= 0.)) Key_quadrant = 1;
    else if ((px < 0.) && (qy >= 0.)) Key_quadrant = 2;
    else if ((px < 0.) && (qy < 0.)) Key_quadrant = 3;
    else if ((px >= 0.) && (qy < 0.)) Key_quadrant = 4;
  } else {
    Area_key_quadrant = 0;
    if ((!((((Ax > 0.) && (Ay > 0.)) || ((Bx > 0.) && (By > 0.))) || ((Cx > 0.) && (Cy > 0.)))) && ((Area_Upper + Area_Right) <= *Area_triangle)) Key_quadrant = 1;
    else if ((!((((Ax < 0.) && (Ay > 0)) || ((Bx < 0.) && (By > 0.))) || ((Cx < 0.) && (Cy > 0.)))) && ((Area_Upper + Area_Left) <= *Area_triangle)) Key_quadrant = 2;
    else if ((!((((Ax < 0.) && (Ay < 0.)) || ((Bx < 0.) && (By < 0.))) || ((Cx < 0.) && (Cy < 0.)))) && ((Area_Lower + Area_Left) <= *Area_triangle)) Key_quadrant = 3;
    else Key_quadrant = 4;
  }
  double A1, A2, A3, A4;
  if (Key_quadrant == 1) { A1 = Area_key_quadrant; A2 = Area_Upper - A1; A4 = Area_Right - A1; A3 = *Area_triangle - (A1 + A2 + A4); }
  else if (Key_quadrant == 2) { A2 = Area_key_quadrant; A1 = Area_Upper - A2; A4 = Area_Right - A1; A3 = *Area_triangle - (A1 + A2 + A4); }
  else if (Key_quadrant == 3) { A3 = Area_key_quadrant; A2 = Area_Left - A3; A1 = Area_Upper - A2; A4 = *Area_triangle - (A1 + A2 + A3); }
  else if (Key_quadrant == 4) { A4 = Area_key_quadrant; A1 = Area_Right - A4; A2 = Area_Upper - A1; A3 = *Area_triangle - (A1 + A2 + A4); }
  else { A1 = A2 = A3 = A4 = 0.; ok = 0; }
  *Q1 = KH_MAX(A1, 0.); *Q2 = KH_MAX(A2, 0.); *Q3 = KH_MAX(A3, 0.); *Q4 = KH_MAX(A4, 0.);
  return ok;
}
// rotate_and_translate I:4537-4554 (theta in degrees)
__device__ __noinline__ void kh_rotate_and_translate(double* px, double* py, double theta, double x0, double y0, double pi) {
  double c = cos(theta * pi / 180), s = sin(theta * pi / 180);
  double px_temp = (c * *px) + (s * *py);
  double py_temp = (-s * *px) + (c * *py);
  *px = px_temp + x0; *py = py_temp + y0;
}
// Hexagon_into_quadrants_using_triangles I:4562-4670; returns 0 on a FATAL of the triangle routines
__device__ __noinline__ int kh_hexagon_into_quadrants(double x0, double y0, double H, double theta, double pi, double* Area_hex,
                                   double* Area_Q1, double* Area_Q2, double* Area_Q3, double* Area_Q4) {
  double S = (2 / sqrt(3.)) * H;
  double Cx[6] = {S, H / sqrt(3.), -H / sqrt(3.), -S, -H / sqrt(3.), H / sqrt(3.)};
  double Cy[6] = {0., H, H, 0., -H, -H};
  int ok = 1;
  for (int k = 0; k < 6; k++) kh_rotate_and_translate(&Cx[k], &Cy[k], theta, x0, y0, pi);
  double TA[6], T1[6], T2[6], T3[6], T4[6];
  for (int k = 0; k < 6; k++) {
    int n = (k + 1) % 6;
    ok &= kh_triangle_into_quadrants(x0, y0, Cx[k], Cy[k], Cx[n], Cy[n], &TA[k], &T1[k], &T2[k], &T3[k], &T4[k]);
  }
  *Area_hex = TA[0] + TA[1] + TA[2] + TA[3] + TA[4] + TA[5];
  double Q1 = T1[0] + T1[1] + T1[2] + T1[3] + T1[4] + T1[5];
  double Q2 = T2[0] + T2[1] + T2[2] + T2[3] + T2[4] + T2[5];
  double Q3 = T3[0] + T3[1] + T3[2] + T3[3] + T3[4] + T3[5];
  double Q4 = T4[0] + T4[1] + T4[2] + T4[3] + T4[4] + T4[5];
  Q1 = KH_MAX(Q1, 0.); Q2 = KH_MAX(Q2, 0.); Q3 = KH_MAX(Q3, 0.); Q4 = KH_MAX(Q4, 0.);
  double Error = *Area_hex - (Q1 + Q2 + Q3 + Q4);
  if (((Q1 >= Q2) && (Q1 >= Q3)) && (Q1 >= Q4)) Q1 = Q1 + Error;
  else if (((Q2 >= Q1) && (Q2 >= Q3)) && (Q2 >= Q4)) Q2 = Q2 + Error;
  else if (((Q3 >= Q1) && (Q3 >= Q2)) && (Q3 >= Q4)) Q3 = Q3 + Error;
  else if (((Q4 >= Q1) && (Q4 >= Q2)) && (Q4 >= Q3)) Q4 = Q4 + Error;
  *Area_Q1 = Q1; *Area_Q2 = Q2; *Area_Q3 = Q3; *Area_Q4 = Q4;
  return ok;
}

// the nine weights of spread_mass_across_ocean_cells I:3949-4082 (order: yDxL, yDxC, yDxR, yCxL, yCxC,
// yCxR, yUxL, yUxC, yUxR) and the inverse of fraction_used.  msk9 = msk of the 3x3 cells in the same order.
// Returns 0 when the reference would stop with 'All the mass is not being used!!!'.
__device__ __noinline__ int kh_spread_weights(int hexagonal, int use_old_spreading, double x, double y, double Area, double cell_area,
                           const double msk9[9], double orientation, double pi, int is_static, double w[9],
                           double* I_fraction_used) {
  double yDxL = 0., yDxC = 0., yDxR = 0., yCxL = 0., yCxR = 0., yUxL = 0., yUxC = 0., yUxR = 0., yCxC = 1., fraction_used;
  int ok = 1;
  if (!hexagonal) {
    double L, xL, xC, xR, yD, yC, yU;
    if (cell_area > 0) L = KH_MIN(sqrt(Area / cell_area), 1.0); else L = 1.;
    if (use_old_spreading) {
      xL = KH_MIN(0.5, KH_MAX(0., 0.5 - x)); xR = KH_MIN(0.5, KH_MAX(0., x - 0.5)); xC = KH_MAX(0., 1. - (xL + xR));
      yD = KH_MIN(0.5, KH_MAX(0., 0.5 - y)); yU = KH_MIN(0.5, KH_MAX(0., y - 0.5)); yC = KH_MAX(0., 1. - (yD + yU));
    } else {
      xL = KH_MIN(0.5, KH_MAX(0., 0.5 - (x / L))); xR = KH_MIN(0.5, KH_MAX(0., (x / L) + (0.5 - (1 / L)))); xC = KH_MAX(0., 1. - (xL + xR));
      yD = KH_MIN(0.5, KH_MAX(0., 0.5 - (y / L))); yU = KH_MIN(0.5, KH_MAX(0., (y / L) + (0.5 - (1 / L)))); yC = KH_MAX(0., 1. - (yD + yU));
    }
    yDxL = yD * xL * msk9[0]; yDxC = yD * xC * msk9[1]; yDxR = yD * xR * msk9[2];
    yCxL = yC * xL * msk9[3]; yCxR = yC * xR * msk9[5];
    yUxL = yU * xL * msk9[6]; yUxC = yU * xC * msk9[7]; yUxR = yU * xR * msk9[8];
    yCxC = 1. - (((yDxL + yUxR) + (yDxR + yUxL)) + ((yCxL + yCxR) + (yDxC + yUxC)));
    fraction_used = 1.;
  } else {
    double H, S, origin_x = 1., origin_y = 1., x0, y0, Area_hex, Q1, Q2, Q3, Q4;
    if (cell_area > 0) H = KH_MIN(((sqrt(Area / (2. * sqrt(3.))) / sqrt(cell_area))), 1.);
    else H = (sqrt(3.) / 2) * (0.49);
    S = (2 / sqrt(3.)) * H; (void)S;
    if (x < 0.5) origin_x = 0.;
    if (y < 0.5) origin_y = 0.;
    x0 = (x - origin_x); y0 = (y - origin_y);
    ok &= kh_hexagon_into_quadrants(x0, y0, H, orientation, pi, &Area_hex, &Q1, &Q2, &Q3, &Q4);
    Q1 = Q1 / Area_hex; Q2 = Q2 / Area_hex; Q3 = Q3 / Area_hex; Q4 = Q4 / Area_hex;
    if ((x >= 0.5) && (y >= 0.5)) { yUxR = Q1; yUxC = Q2; yCxC = Q3; yCxR = Q4; }
    else if ((x < 0.5) && (y >= 0.5)) { yUxC = Q1; yUxL = Q2; yCxL = Q3; yCxC = Q4; }
    else if ((x < 0.5) && (y < 0.5)) { yCxC = Q1; yCxL = Q2; yDxL = Q3; yDxC = Q4; }
    else if ((x >= 0.5) && (y < 0.5)) { yCxR = Q1; yCxC = Q2; yDxC = Q3; yDxR = Q4; }
    if (fabs(yCxC - (1. - (((yDxL + yUxR) + (yDxR + yUxL)) + ((yCxL + yCxR) + (yDxC + yUxC))))) > 1.e-10) ok = 0;
    fraction_used = ((yDxL * msk9[0]) + (yDxC * msk9[1]) + (yDxR * msk9[2]) + (yCxL * msk9[3]) + (yCxR * msk9[5]) +
                     (yUxL * msk9[6]) + (yUxC * msk9[7]) + (yUxR * msk9[8]) + pow(yCxC, msk9[4]));     /* yCxC**msk, I:4081 */
    if (is_static) fraction_used = 1.;
  }
  w[0] = yDxL; w[1] = yDxC; w[2] = yDxR; w[3] = yCxL; w[4] = yCxC; w[5] = yCxR; w[6] = yUxL; w[7] = yUxC; w[8] = yUxR;
  *I_fraction_used = 1. / fraction_used;
  return ok;
}

// find_orientation_using_iceberg_bonds I:3829-3893
__device__ __noinline__ double bond_orientation(const DevGrid& g, const DevBergs& b, const DevParams& p, long long s,
                                                double orientation) {
  int i = b.ine[s], j = b.jne[s];
  if (!(((i > g.isd) && (i < g.ied)) && ((j >= g.jsd) && (j <= g.jed)))) return orientation;
  double bond_count = 0., Average_angle = 0.;
  double lat1 = b.f64[C_LAT][s], lon1 = b.f64[C_LON][s];
  for (int k = b.max_bonds - 1; k >= 0; k--) {
    long long slot = (long long)k * b.capacity + s;
    if (b.bond_other_id[slot] == 0) continue;
    int32_t o = b.bond_other_slot[slot];
    if (o < 0) continue;
    double lat2 = b.f64[C_LAT][o], lon2 = b.f64[C_LON][o], dx_dlon, dy_dlat, angle;
    convert_from_grid_to_meters(p, 0.5 * (lat1 + lat2), dx_dlon, dy_dlat);
    double rx = (lon2 - lon1) * dx_dlon, ry = (lat2 - lat1) * dy_dlat;
    if (rx == 0.) angle = p.pi / 2.;
    else {
      angle = atan(ry / rx);
      angle = ((p.pi / 2.) - (orientation * (p.pi / 180.))) - angle;
      angle = f_modulo_slow(angle, p.pi / 3.);
    }
    bond_count += 1.; Average_angle += angle;
  }
  if (bond_count > 0) Average_angle = Average_angle / bond_count; else Average_angle = 0.;
  return f_modulo_slow(Average_angle, p.pi / 3.);
}

struct SpreadParams {
  double grounding_fraction, clipping_depth, initial_orientation, cdrag_icebergs, utide_icebergs, ustar_icebergs_bg, melt_cutoff;
  int32_t add_weight, use_old_spreading, rotate, diag, apply_cutoff_gridded, bergy;   // bergy: id_bergy_mass > 0 or add_weight_to_ocean, I:5061
};
struct SpreadFields {
  double *mass_on_ocean, *area_on_ocean, *uvel_on_ocean, *vvel_on_ocean;   // [9][n2]
  double *mass, *bergy_mass, *spread_mass, *spread_area, *spread_uvel, *spread_vvel, *ustar_iceberg;
};

// calculate_mass_on_ocean I:4970-5011: every owned berg deposits mass, area and area-weighted velocity
// into the 9 layers of its own cell (halo copies only ever touch halo cells, which mpp_update_domains
// overwrites, I:6105: they are skipped)
__global__ void k_spread_bergs(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                               const __grid_constant__ DevParams p, const __grid_constant__ SpreadParams sp,
                               const __grid_constant__ SpreadFields sf, DevCounters* __restrict__ cnt, long long n_slots,
                               long long n2) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  uint8_t f = b.flags[s];
  if (!(f & BF_ALIVE) || (f & (BF_HALO | BF_LEAVER))) return;
  int i = b.ine[s], j = b.jne[s];
  int c = gidx(g, i, j);
  double area = g.area[c];
  if (!(area > 0.)) return;
  double M = b.f64[C_MASS][s], ms = b.f64[C_MASS_SCALING][s], mob = b.f64[C_MASS_OF_BITS][s];
  double mfl = b.f64[C_MASS_OF_FL_BITS][s], mflb = b.f64[C_MASS_OF_FL_BERGY_BITS][s];
  if (sp.diag) atomicAdd(&sf.mass[c], M / area * ms);                       // id_mass > 0, I:5049
  if (sp.bergy) atomicAdd(&sf.bergy_mass[c], (mob + mflb) / area * ms);      // id_bergy_mass > 0 or add_weight_to_ocean, I:5061
  if (!sp.add_weight) return;
  const double rho_seawater = 1035.;          // the local value of spread_mass_across_ocean_cells, I:3919
  double Tn = b.f64[C_THICKNESS][s], A = b.f64[C_LENGTH][s] * b.f64[C_WIDTH][s];
  double Mass_berg = M, Mfl = mfl;
  if (sp.grounding_fraction > 0.) {
    double Hocean = sp.grounding_fraction * (g.ocean_depth[c] + g.ssh[c]);
    double Dn = (p.rho_bergs / rho_seawater) * Tn;
    if (Dn > Hocean) Mass_berg = Mass_berg * fmin(1., Hocean / Dn);
    if (Mfl > 0.) {
      const double l_c = p.pi / (2. * sqrt(2.)), lw_c = 1. / (KID_GRAVITY * KID_RHO_SEAWATER), B_c = 1. / (12. * (1. - pow(0.3, 2.)));
      double l_b = l_c * pow(lw_c * p.fl_youngs * B_c * pow(Tn, 3.), 0.25);
      double Lfl = 3. * l_b, Wfl = l_b, Tfl = Tn;
      rolling(p, Tfl, Wfl, Lfl);
      Dn = (p.rho_bergs / rho_seawater) * Tfl;
      if (Dn > Hocean) Mfl = Mfl * fmin(1., Hocean / Dn);
    }
  }
  Mass_berg = Mass_berg + Mfl;
  double Mass = (Mass_berg + mob + mflb) * ms;
  if (sp.clipping_depth > 0.) Mass = fmin(Mass, sp.clipping_depth * area * rho_seawater);
  double msk9[9], w[9], Ifu, orientation = sp.initial_orientation;
  for (int dj = -1; dj <= 1; dj++) for (int di = -1; di <= 1; di++) msk9[(dj + 1) * 3 + (di + 1)] = g.msk[c + di + dj * g.nid];
  if (p.hexagonal_icebergs && p.iceberg_bonds_on && sp.rotate) orientation = bond_orientation(g, b, p, s, orientation);
  if (!kh_spread_weights(p.hexagonal_icebergs, sp.use_old_spreading, b.f64[C_XI][s], b.f64[C_YJ][s], A, area, msk9, orientation,
                         p.pi, (f & BF_STATIC) ? 1 : 0, w, &Ifu)) { atomicOr(&cnt->error_flags, 4096u); return; }
  double uv = b.f64[C_UVEL][s], vv = b.f64[C_VVEL][s];
  for (int k = 0; k < 9; k++) {
    if (w[k] == 0.) continue;
    atomicAdd(&sf.mass_on_ocean[c + n2 * k], w[k] * Mass * Ifu);
    atomicAdd(&sf.area_on_ocean[c + n2 * k], w[k] * (A * ms) * Ifu);
    atomicAdd(&sf.uvel_on_ocean[c + n2 * k], w[k] * (uv * A * ms) * Ifu);
    atomicAdd(&sf.vvel_on_ocean[c + n2 * k], w[k] * (vv * A * ms) * Ifu);
  }
}

// sum_up_spread_fields I:6077-6150 for one field (no tripolar fold: parity_x = 1)
__device__ __forceinline__ double sum9(const DevGrid& g, const double* __restrict__ v, int c, long long n2, bool is_area) {
  const int nid = g.nid;
#define V9(di, dj, k) v[c + (di) + (dj) * nid + n2 * ((k) - 1)]
  double dmda = V9(0, 0, 5) + (((V9(-1, -1, 9) + V9(1, 1, 1)) + (V9(1, -1, 7) + V9(-1, 1, 3))) +
                               ((V9(-1, 0, 6) + V9(1, 0, 4)) + (V9(0, -1, 8) + V9(0, 1, 2))));
#undef V9
  double a = g.area[c];
  if (a > 0) dmda = dmda / a * g.msk[c];
  if (is_area) dmda = fmin(dmda, 1.0);
  return dmda;
}

__global__ void k_sum_spread(const __grid_constant__ DevGrid g, const __grid_constant__ SpreadParams sp,
                             const __grid_constant__ SpreadFields sf, long long n2) {
  int ni = g.iec - g.isc + 1, nj = g.jec - g.jsc + 1;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)ni * nj) return;
  int c = gidx(g, g.isc + (int)(k % ni), g.jsc + (int)(k / ni));
  sf.spread_mass[c] = sum9(g, sf.mass_on_ocean, c, n2, false);
  if (sp.diag) {
    double su = sum9(g, sf.uvel_on_ocean, c, n2, false), sv = sum9(g, sf.vvel_on_ocean, c, n2, false);
    double sa = sum9(g, sf.area_on_ocean, c, n2, true);
    sf.spread_uvel[c] = su; sf.spread_vvel[c] = sv; sf.spread_area[c] = sa;
    // I:3461-3473
    double du = su - g.uo[c], dv = sv - g.vo[c];
    double dvo = sqrt(du * du + dv * dv);
    double ustar = sqrt(sp.cdrag_icebergs * (dvo * dvo + sp.utide_icebergs * sp.utide_icebergs));
    double ustar_h = fmax(sp.ustar_icebergs_bg, ustar);
    if (sa == 0.0) ustar_h = 0.;
    sf.ustar_iceberg[c] = ustar_h;
  }
}

// find_melt_using_spread_mass, I:3436-3448: the melt is what the spread mass lost over the step (data domain; the spread
// sums are zero outside the compute domain), and the heat flux follows from it
__global__ void k_melt_from_spread_mass(const __grid_constant__ DevGrid g, const double* __restrict__ spread_mass_old,
                                        const double* __restrict__ spread_mass_new, double dt, double hlf, long long n2) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n2) return;
  double fm = 0.0;
  if (g.area[c] > 0.0) fm = fmax((spread_mass_old[c] - spread_mass_new[c]) / dt, 0.0);
  g.floating_melt[c] = fm;
  int i = g.isd + (int)(c % g.nid), j = g.jsd + (int)(c / g.nid);
  if (i >= g.isc && i <= g.iec && j >= g.jsc && j <= g.jec) g.calving_hflx[c] = fm * hlf;
}

// I:3476-3488: no melt into water shallower than melt_cutoff under the average draught (data domain)
__global__ void k_thickness_cutoff(const __grid_constant__ DevGrid g, const __grid_constant__ DevParams p,
                                   const __grid_constant__ SpreadParams sp, const __grid_constant__ SpreadFields sf, long long n2) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n2) return;
  double sa = sf.spread_area[c];
  if ((sp.melt_cutoff >= 0.) && (sa > 0.)) {
    double ave_thickness = sf.spread_mass[c] / (sa * p.rho_bergs);
    double ave_draft = ave_thickness * (p.rho_bergs / KID_RHO_SEAWATER);
    if ((g.ocean_depth[c] - ave_draft) < sp.melt_cutoff) { g.floating_melt[c] = 0.0; g.calving_hflx[c] = 0.0; }
  }
}

}  // namespace kid
