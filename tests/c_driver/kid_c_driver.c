/*
 * A plain-C caller of the drop-in boundary: what the Fortran ISO_C_BINDING shim does, without
 * Fortran.  Column-major (i fastest) arrays exactly as icebergs_init / icebergs_run receive them (I:92-117,
 * I:5074-5096), built here in C; kid_init / kid_set_bergs / kid_run x3 / kid_get_bergs / kid_end through
 * include/kid_b200.h only, then the same inputs through the CPU oracle (test infrastructure) and a comparison:
 * ids, cell indices and counts bit-exact, positions / velocities / masses to 1e-10 relative.
 * Exit code 0 and "C DRIVER OK" on success.  Built and run by tests/test_c_driver.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kid_b200.h"
#include "kid_oracle.h"

#define NI 72
#define NJ 36
#define NB 3000
#define HALO 4

static double* grid(int ring) { return (double*)calloc((size_t)(NI + 2 * ring) * (NJ + 2 * ring), sizeof(double)); }
/* element (i,j) of an (isc-ring:iec+ring, jsc-ring:jec+ring) Fortran array, i and j 1-based global indices */
#define AT(a, ring, i, j) (a)[(size_t)((i) - 1 + (ring)) + (size_t)((j) - 1 + (ring)) * (NI + 2 * (ring))]

static uint64_t lcg(uint64_t* s) { *s = *s * 6364136223846793005ULL + 1442695040888963407ULL; return *s >> 11; }
static double u01(uint64_t* s) { return (double)lcg(s) / 9007199254740992.0; }

static int cmp_id(const void* a, const void* b) {
  int64_t x = **(const int64_t* const*)a, y = **(const int64_t* const*)b;
  return x < y ? -1 : x > y;
}

int main(void) {
  const double pi = 3.14159265358979323846, rad = pi / 180., Re = 6.36e6, dlon = 360. / NI, dlat = 180. / NJ;
  double *lon = grid(0), *lat = grid(0), *area = grid(0), *depth = grid(0);
  double *wet = grid(1), *dx = grid(1), *dy = grid(1), *cosr = grid(1), *sinr = grid(1);
  double *uo = grid(1), *vo = grid(1), *ui = grid(1), *vi = grid(1), *ssh = grid(1), *cn = grid(1), *hi = grid(1);
  double *taux = grid(0), *tauy = grid(0), *sst = grid(0), *sss = grid(0);
  for (int j = 0; j <= NJ + 1; j++)
    for (int i = 0; i <= NI + 1; i++) {
      double lo = i * dlon, la = -90. + j * dlat, lac = la - 0.5 * dlat;
      AT(wet, 1, i, j) = (fabs(lac) < 80.) ? 1. : 0.;
      AT(dx, 1, i, j) = fabs(Re * cos(la * rad) * dlon * rad);
      AT(dy, 1, i, j) = Re * dlat * rad;
      AT(cosr, 1, i, j) = 1.; AT(sinr, 1, i, j) = 0.;
      AT(uo, 1, i, j) = 0.3 * cos(la * rad) * sin(2 * lo * rad);
      AT(vo, 1, i, j) = 0.2 * sin(lo * rad) * cos(3 * la * rad);
      AT(ui, 1, i, j) = 0.5 * AT(uo, 1, i, j); AT(vi, 1, i, j) = 0.5 * AT(vo, 1, i, j);
      AT(ssh, 1, i, j) = 0.5 * sin(2 * (lo - 0.5 * dlon) * rad) * cos(3 * lac * rad);
      double c = (fabs(lac) - 55.) / 20.; c = c < 0. ? 0. : (c > 1. ? 1. : c);
      AT(cn, 1, i, j) = c; AT(hi, 1, i, j) = 1.5 * c;
      if (i >= 1 && i <= NI && j >= 1 && j <= NJ) {
        AT(lon, 0, i, j) = lo; AT(lat, 0, i, j) = la;
        AT(area, 0, i, j) = fabs(Re * cos(lac * rad) * dlon * rad * Re * dlat * rad);
        AT(depth, 0, i, j) = 4000.;
        AT(taux, 0, i, j) = 10. * cos(2 * la * rad); AT(tauy, 0, i, j) = 3. * sin(3 * lo * rad);
        AT(sst, 0, i, j) = -1.5 + 4. * cos(lac * rad) * cos(lac * rad); AT(sss, 0, i, j) = 34.;
      }
    }
  /* bergs */
  static double blon[NB], blat[NB], mass[NB], thick[NB], width[NB], length[NB], sday[NB], zero[NB], one[NB];
  static int32_t ine[NB], jne[NB], syear[NB];
  static int64_t id[NB];
  static int32_t counter[NJ + 2 * HALO][NI + 2 * HALO];
  uint64_t seed = 20240521;
  for (int k = 0; k < NB; k++) {
    int i, j;
    do { i = 1 + (int)(u01(&seed) * NI); j = 1 + (int)(u01(&seed) * NJ); } while (AT(wet, 1, i, j) < 0.5 || i > NI || j > NJ);
    double xi = 0.05 + 0.9 * u01(&seed), yj = 0.05 + 0.9 * u01(&seed);
    blon[k] = (i - 1 + xi) * dlon; blat[k] = -90. + (j - 1 + yj) * dlat;
    mass[k] = 3.3e9; thick[k] = 133.; width[k] = sqrt(mass[k] / (1.5 * 850. * thick[k])); length[k] = 1.5 * width[k];
    sday[k] = (k % 10 + 1) / 17.; zero[k] = 0.; one[k] = 50.;
    ine[k] = i; jne[k] = j; syear[k] = 1;
    int32_t c = ++counter[j - 1 + HALO][i - 1 + HALO];
    id[k] = (int64_t)c * ((int64_t)1 << 32) + (int64_t)(i + NI * (j - 1));      /* generate_id F:4165-4177 */
  }
  KidBergColumns in;
  memset(&in, 0, sizeof(in));
  in.lon = blon; in.lat = blat; in.mass = mass; in.thickness = thick; in.width = width; in.length = length;
  in.uvel = zero; in.vvel = zero; in.axn = zero; in.ayn = zero; in.bxn = zero; in.byn = zero;
  in.start_lon = blon; in.start_lat = blat; in.start_day = sday; in.start_mass = mass; in.mass_scaling = one;
  in.mass_of_bits = zero; in.heat_density = zero; in.start_year = syear; in.ine = ine; in.jne = jne; in.id = id;

  KidParams p;
  kid_default_params(&p);
  p.runge_not_verlet = 0; p.bergy_bit_erosion_fraction = 0.1; p.tau_is_velocity = 1; p.add_weight_to_ocean = 0;
  p.Rearth = Re; p.halo = HALO; p.dt = 3600.;
  KidDomain d;
  kid_single_domain(&d, NI, NJ, HALO, 1, 0, 0);

  kid_t* h = NULL;
  int rc = kid_init(&h, &p, &d, 1, 0.0, 4 * NB, lon, lat, wet, dx, dy, area, cosr, sinr, depth, 0);
  if (rc) { fprintf(stderr, "kid_init: %d %s\n", rc, kid_last_error(h)); return 2; }
  rc = kid_set_calving_state(h, NULL, NULL, &counter[0][0]);
  if (!rc) rc = kid_set_bergs(h, NB, &in);
  if (rc) { fprintf(stderr, "kid_set_bergs: %d %s\n", rc, kid_last_error(h)); return 2; }
  Oracle* o = oracle_create(&p, &d, 1, 0.0, lon, lat, wet, dx, dy, area, cosr, sinr, depth, 0);
  oracle_set_calving_state(o, NULL, NULL, &counter[0][0]);
  oracle_set_bergs(o, NB, &in);
  double *calv = grid(0), *hflx = grid(0), *calv_o = grid(0), *hflx_o = grid(0);
  for (int step = 0; step < 3; step++) {
    memset(calv, 0, sizeof(double) * NI * NJ); memset(hflx, 0, sizeof(double) * NI * NJ);
    memset(calv_o, 0, sizeof(double) * NI * NJ); memset(hflx_o, 0, sizeof(double) * NI * NJ);
    rc = kid_run(h, 1, step / 24.0, calv, uo, vo, ui, vi, taux, tauy, ssh, sst, hflx, cn, hi, KID_BGRID_NE, KID_BGRID_NE, sss,
                 NULL, NULL, NULL);
    if (rc) { fprintf(stderr, "kid_run: %d %s\n", rc, kid_last_error(h)); return 2; }
    rc = oracle_run(o, 1, step / 24.0, calv_o, uo, vo, ui, vi, taux, tauy, ssh, sst, hflx_o, cn, hi, KID_BGRID_NE, KID_BGRID_NE,
                    sss, NULL, NULL, NULL);
    if (rc) { fprintf(stderr, "oracle_run: %d %s\n", rc, oracle_last_error(o)); return 2; }
  }
  /* results */
  static double g_lon[NB], g_lat[NB], g_u[NB], g_v[NB], g_m[NB], o_lon[NB], o_lat[NB], o_u[NB], o_v[NB], o_m[NB];
  static int32_t g_i[NB], g_j[NB], o_i[NB], o_j[NB];
  static int64_t g_id[NB], o_id[NB];
  KidBergColumns og, oo;
  memset(&og, 0, sizeof(og)); memset(&oo, 0, sizeof(oo));
  og.lon = g_lon; og.lat = g_lat; og.uvel = g_u; og.vvel = g_v; og.mass = g_m; og.ine = g_i; og.jne = g_j; og.id = g_id;
  oo.lon = o_lon; oo.lat = o_lat; oo.uvel = o_u; oo.vvel = o_v; oo.mass = o_m; oo.ine = o_i; oo.jne = o_j; oo.id = o_id;
  int64_t ng = NB, no = NB;
  rc = kid_get_bergs(h, &ng, &og, 0);
  if (rc) { fprintf(stderr, "kid_get_bergs: %d %s\n", rc, kid_last_error(h)); return 2; }
  oracle_get_bergs(o, &no, &oo, 0);
  if (ng != no) { fprintf(stderr, "berg count %lld != %lld\n", (long long)ng, (long long)no); return 1; }
  static const int64_t* pg[NB]; static const int64_t* po[NB];
  for (int k = 0; k < ng; k++) { pg[k] = &g_id[k]; po[k] = &o_id[k]; }
  qsort(pg, (size_t)ng, sizeof(pg[0]), cmp_id); qsort(po, (size_t)no, sizeof(po[0]), cmp_id);
  double worst = 0.;
  for (int k = 0; k < ng; k++) {
    int a = (int)(pg[k] - g_id), b = (int)(po[k] - o_id);
    if (g_id[a] != o_id[b] || g_i[a] != o_i[b] || g_j[a] != o_j[b]) { fprintf(stderr, "id / cell mismatch at %d\n", k); return 1; }
    const double* gv[5] = {g_lon, g_lat, g_u, g_v, g_m}; const double* ov[5] = {o_lon, o_lat, o_u, o_v, o_m};
    for (int q = 0; q < 5; q++) {
      double x = gv[q][a], y = ov[q][b], s = fmax(fmax(fabs(x), fabs(y)), 1e-300), e = fabs(x - y) / s;
      if ((q == 2 || q == 3) && fabs(x - y) < 1e-13) e = 0.;
      if (e > worst) worst = e;
    }
  }
  double fl = 0.;
  for (int k = 0; k < NI * NJ; k++) fl = fmax(fl, fabs(calv[k] - calv_o[k]));
  kid_end(&h);
  oracle_destroy(o);
  printf("bergs %lld worst rel err %.3e max |calving out - oracle| %.3e\n", (long long)ng, worst, fl);
  if (worst > 1e-10 || fl > 1e-12) return 1;
  printf("C DRIVER OK\n");
  return 0;
}
