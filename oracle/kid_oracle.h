/*
 * kid_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * Plain-C restatement of the reference's per-timestep iceberg hot path
 * (NOAA-GFDL/icebergs; I: = src/icebergs.F90, F: = src/icebergs_framework.F90).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (icebergs_b200/) never
 * links, imports or calls anything here.
 *
 * PARITY PINNING: the reference is Fortran + GFDL FMS; no Fortran compiler and
 * no FMS exist in this image, so the reference itself cannot be run.  The oracle
 * is pinned against the reference's own known answers that are portable
 * (tests/test_oracle_golden.py): bilin corner exactness F:7313-7316, the 64-bit
 * id round trip F:7319-7325, the point-in-triangle regression I:234-242 and the
 * 8 hexagon area identities I:261-348 (for the spreading rows), the berg counts
 * of the collision tests (16) and of the footloose test (12 after 69 120 steps),
 * the reference's two DEM beam tests against beam theory (Wang 2020 3.1 / 3.2;
 * tests/dem_ssbeam_test, tests/dem_cbeam_test: calculate_force_dem, the rotation
 * update and the MTS sub-steps), and by review against the cited lines.
 * accel, thermodynamics and calculate_force have no numeric pin in the reference
 * beyond build-specific checksums: for those rows parity is "oracle-vs-kernel,
 * oracle by review" = parity unpinned.
 * The tripolar fold (mpp_update_domains across FOLD_NORTH_EDGE) restates FMS, which is outside the reference tree:
 * unpinned against FMS, pinned geometrically by the analytic continuation of a bipolar cap (tests/test_fold_oracle.py).
 * Cross-checks that do not depend on this file's arithmetic: tests/test_oracle_options.py (the melt found from the
 * spread mass equals the per-berg melt, the running-mean calving follows its closed form).
 */
#ifndef KID_ORACLE_H
#define KID_ORACLE_H

#include "../include/kid_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct Oracle Oracle;

/* same argument meaning as kid_init / kid_run (include/kid_b200.h) */
Oracle* oracle_create(const KidParams* p, const KidDomain* dom, int32_t year, double yearday,
                      const double* lon, const double* lat, const double* wet,
                      const double* dx, const double* dy, const double* area,
                      const double* cos_rot, const double* sin_rot,
                      const double* ocean_depth, int32_t fractional_area);
void    oracle_destroy(Oracle* o);
const char* oracle_last_error(const Oracle* o);

int32_t oracle_set_bergs(Oracle* o, int64_t n, const KidBergColumns* c);
int64_t oracle_count_bergs(const Oracle* o, int32_t include_halo);
/* list order: do j=jsd,jed; do i=isd,ied; walk list (F:1769) */
int32_t oracle_get_bergs(const Oracle* o, int64_t* n, KidBergColumns* c, int32_t include_halo);
/* record_posn F:5328-5498 and the flattened trajectory store (write_trajectory fmsio:1575) */
int32_t oracle_record_posn(Oracle* o);
int64_t oracle_trajectory_count(const Oracle* o);
int32_t oracle_get_trajectory(Oracle* o, int64_t* n, KidTrajColumns* c, int32_t clear);
int32_t oracle_set_bonds(Oracle* o, int64_t nb, const KidBondColumns* c);
int32_t oracle_get_bonds(const Oracle* o, int64_t* nb, KidBondColumns* c);
int32_t oracle_set_calving_state(Oracle* o, const double* stored_ice, const double* stored_heat,
                                 const int32_t* counter);
int32_t oracle_set_calving_rmean(Oracle* o, const double* rmean_calving, const double* rmean_calving_hflx);
int32_t oracle_get_calving_state(const Oracle* o, double* stored_ice, double* stored_heat,
                                 int32_t* counter);

int32_t oracle_run(Oracle* o, int32_t year, double yearday,
                   double* calving, const double* uo, const double* vo,
                   const double* ui, const double* vi,
                   const double* tauxa, const double* tauya,
                   const double* ssh, const double* sst, double* calving_hflx,
                   const double* cn, const double* hi,
                   int32_t stagger, int32_t stress_stagger, const double* sss,
                   double* mass_berg, double* ustar_berg, double* area_berg);
/* re-run the step `nsteps` times with the forcing of the last oracle_run and no
 * return fields; nthreads>1 uses OpenMP over cell rows where the reference loop
 * is order-independent (CPU baseline timing) */
int32_t oracle_step_again(Oracle* o, int32_t nsteps, int32_t year, double yearday, int32_t nthreads);

int32_t oracle_get_grid_field(const Oracle* o, int32_t field_id, double* out);
int32_t oracle_get_counters(const Oracle* o, KidCounters* c);
/* seconds spent in (momentum, thermodynamics, rest) of the last run/step_again call */
void    oracle_last_timing(const Oracle* o, double sec[4]);

/* ---- unit-level entry points (for golden-vector tests) ---- */
double  oracle_bilin(const Oracle* o, int32_t field_id, int32_t i, int32_t j, double xi, double yj);
int32_t oracle_is_point_in_cell(const Oracle* o, double x, double y, int32_t i, int32_t j);
int32_t oracle_pos_within_cell(const Oracle* o, double x, double y, int32_t i, int32_t j,
                               double* xi, double* yj);
int32_t oracle_find_cell(const Oracle* o, double x, double y, int32_t* oi, int32_t* oj);
int32_t oracle_find_cell_wide(const Oracle* o, double x, double y, int32_t* oi, int32_t* oj);
double  oracle_apply_modulo_around_point(double x, double y, double Lx);
int64_t oracle_id_from_2_ints(int32_t counter, int32_t ijhash);
void    oracle_split_id(int64_t id, int32_t* counter, int32_t* ijhash);
double  oracle_yearday(int32_t imon, int32_t iday, int32_t ihr, int32_t imin, int32_t isec);
int32_t oracle_point_in_triangle(double Ax, double Ay, double Bx, double By,
                                 double Cx, double Cy, double qx, double qy);
void    oracle_hexagon_into_quadrants(double x0, double y0, double H, double theta,
                                      double* Area_hex, double* Area_Q1, double* Area_Q2,
                                      double* Area_Q3, double* Area_Q4);
void    oracle_rolling(const KidParams* p, double* Tn, double* Wn, double* Ln);
/* accel() for a free berg with the given environment (no grid): out[0..5] =
 * ax, ay, axn, ayn, bxn, byn  (I:1950-2301) */
void    oracle_accel_free(const KidParams* p, const double berg[8] /* M,T,W,L,lat,uvel,vvel,(unused) */,
                          const double acc_in[4] /* axn,ayn,bxn,byn */,
                          const double env[13] /* uo,vo,ui,vi,ua,va,ssh_x,ssh_y,sst,sss,cn,hi,od */,
                          double out[6]);

/* namelist defaults F:686-822 + FMS constants, a 1x1 layout, and the OpenMP width of this build: what bench.py's
 * CPU legs need so that they never load the product library */
void    oracle_default_params(KidParams* p);
void    oracle_single_domain(KidDomain* d, int32_t gni, int32_t gnj, int32_t halo, int32_t cyclic_x, int32_t cyclic_y);
int32_t oracle_omp_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
