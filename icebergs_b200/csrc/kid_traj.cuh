// kid_traj.cuh -- trajectory sampling, SURVEY 8(f2): record_posn F:5328-5498.
//
// The reference pushes a type(xyt) record onto the berg's own trajectory list (push_posn F:5502), moves the list to
// bergs%trajectories when the berg dies or leaves the PE (move_trajectory F:5611) and flattens everything in
// write_trajectory (fmsio:1575).  On the device the samples go straight into one column-major store (TR_NCOL arrays
// of `cap` doubles, appended with one atomic per sampled berg); the host drains it with kid_get_trajectory.
//
// What a record holds of the berg's environment (uo ... hi) is what thermodynamics / interp_gridded_fields_to_bergs
// left on the berg: the fields interpolated at the berg's CURRENT position from the forcing of the current call
// (I:2890-2894 with old_interp_flds_order, I:5473 / I:5457 otherwise), so the sample re-interpolates there -- or reads
// the environment cache the MTS scheme keeps.  `od` is never stored on the berg in the old order (accel interpolates
// into locals, I:2035): it stays 0.
#pragma once
#include "kid_physics.cuh"

namespace kid {

enum TrajCol : int {
  TR_LON = 0, TR_LAT, TR_DAY, TR_YEAR, TR_ID, TR_MASS, TR_START_MASS, TR_THICKNESS, TR_MASS_OF_BITS, TR_UVEL, TR_VVEL,
  TR_MASS_SCALING, TR_MASS_OF_FL_BITS, TR_MASS_OF_FL_BERGY_BITS, TR_FL_K,
  TR_UVEL_PREV, TR_VVEL_PREV, TR_HEAT_DENSITY, TR_WIDTH, TR_LENGTH,
  TR_UO, TR_VO, TR_UI, TR_VI, TR_UA, TR_VA, TR_SSH_X, TR_SSH_Y, TR_SST, TR_SSS, TR_CN, TR_HI,
  TR_AXN, TR_AYN, TR_BXN, TR_BYN, TR_HALO_BERG, TR_STATIC_BERG, TR_OD,
  TR_AXN_FAST, TR_AYN_FAST, TR_BXN_FAST, TR_BYN_FAST, TR_N_BONDS, TR_ANG_VEL, TR_ANG_ACCEL, TR_ROT,
  TR_NCOL
};

struct TrajParams {
  double area_thres, area_thres2, area_thres3;     // m^2 (F:5362-5364)
  double save_all_traj_year, start_mass_thres_n, start_mass_thres_s, rho_bergs;
  int32_t save_nonfl_traj_by_class, old_interp_flds_order, mts, dem;
};

__global__ void k_record_posn(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                              const __grid_constant__ DevParams p, const __grid_constant__ TrajParams tp,
                              DevCounters* __restrict__ cnt, long long n_slots, double* __restrict__ buf, long long cap,
                              unsigned long long* __restrict__ cursor) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const uint8_t flags = b.flags[s];
  if (!(flags & BF_ALIVE) || (flags & (BF_HALO | BF_LEAVER))) return;          // compute domain only (debug_write = F)
  const double mass = b.f64[C_MASS][s], thickness = b.f64[C_THICKNESS][s], lat = b.f64[C_LAT][s];
  const double fl_k = b.f64[C_FL_K][s], start_mass = b.f64[C_START_MASS][s];
  const double berg_area = mass / (tp.rho_bergs * thickness);
  bool by_class = false;
  if (tp.save_nonfl_traj_by_class && fl_k >= 0. && berg_area > tp.area_thres2)
    by_class = (lat < 0.) ? (start_mass >= tp.start_mass_thres_s) : (start_mass >= tp.start_mass_thres_n);
  const bool save_fl_berg = (fl_k < 0. && berg_area > tp.area_thres3);
  int n_bonds = 0;
  for (int k = 0; k < b.max_bonds; k++) if (b.bond_other_id[(long long)k * b.capacity + s] != 0) n_bonds++;
  if (!((double)p.current_year > tp.save_all_traj_year || by_class || berg_area >= tp.area_thres || n_bonds > 0 || save_fl_berg)) return;
  const unsigned long long k = atomicAdd(cursor, 1ull);
  if ((long long)k >= cap) { atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_CAPACITY); return; }
#define TR(c) buf[(size_t)(c) * cap + k]
  TR(TR_LON) = b.f64[C_LON][s]; TR(TR_LAT) = lat;
  TR(TR_DAY) = p.current_yearday; TR(TR_YEAR) = (double)p.current_year;
  TR(TR_ID) = __longlong_as_double(b.id[s]);
  TR(TR_MASS) = mass; TR(TR_START_MASS) = start_mass; TR(TR_THICKNESS) = thickness;
  TR(TR_MASS_OF_BITS) = b.f64[C_MASS_OF_BITS][s];
  TR(TR_UVEL) = b.f64[C_UVEL][s]; TR(TR_VVEL) = b.f64[C_VVEL][s];
  TR(TR_MASS_SCALING) = b.f64[C_MASS_SCALING][s];
  TR(TR_MASS_OF_FL_BITS) = b.f64[C_MASS_OF_FL_BITS][s];
  TR(TR_MASS_OF_FL_BERGY_BITS) = b.f64[C_MASS_OF_FL_BERGY_BITS][s];
  TR(TR_FL_K) = fl_k;
  TR(TR_UVEL_PREV) = b.f64[C_UVEL_PREV][s]; TR(TR_VVEL_PREV) = b.f64[C_VVEL_PREV][s];
  TR(TR_HEAT_DENSITY) = b.f64[C_HEAT_DENSITY][s];
  TR(TR_WIDTH) = b.f64[C_WIDTH][s]; TR(TR_LENGTH) = b.f64[C_LENGTH][s];
  Env e;
  if (tp.mts) {
    e.uo = b.f64[C_UO][s]; e.vo = b.f64[C_VO][s]; e.ui = b.f64[C_UI][s]; e.vi = b.f64[C_VI][s];
    e.ua = b.f64[C_UA][s]; e.va = b.f64[C_VA][s]; e.ssh_x = b.f64[C_SSH_X][s]; e.ssh_y = b.f64[C_SSH_Y][s];
    e.sst = b.f64[C_SST][s]; e.sss = b.f64[C_SSS][s]; e.cn = b.f64[C_CN][s]; e.hi = b.f64[C_HI][s]; e.od = b.f64[C_OD][s];
  } else {
    interp_flds(g, p, b.ine[s], b.jne[s], b.f64[C_XI][s], b.f64[C_YJ][s], e);
    if (tp.old_interp_flds_order) e.od = 0.;
  }
  TR(TR_UO) = e.uo; TR(TR_VO) = e.vo; TR(TR_UI) = e.ui; TR(TR_VI) = e.vi; TR(TR_UA) = e.ua; TR(TR_VA) = e.va;
  TR(TR_SSH_X) = e.ssh_x; TR(TR_SSH_Y) = e.ssh_y; TR(TR_SST) = e.sst; TR(TR_SSS) = e.sss; TR(TR_CN) = e.cn; TR(TR_HI) = e.hi;
  TR(TR_OD) = e.od;
  TR(TR_AXN) = b.f64[C_AXN][s]; TR(TR_AYN) = b.f64[C_AYN][s]; TR(TR_BXN) = b.f64[C_BXN][s]; TR(TR_BYN) = b.f64[C_BYN][s];
  TR(TR_HALO_BERG) = (double)b.halo_code[s];
  TR(TR_STATIC_BERG) = (flags & BF_STATIC) ? 1. : 0.;
  TR(TR_AXN_FAST) = tp.mts ? b.f64[C_AXN_FAST][s] : 0.; TR(TR_AYN_FAST) = tp.mts ? b.f64[C_AYN_FAST][s] : 0.;
  TR(TR_BXN_FAST) = tp.mts ? b.f64[C_BXN_FAST][s] : 0.; TR(TR_BYN_FAST) = tp.mts ? b.f64[C_BYN_FAST][s] : 0.;
  TR(TR_N_BONDS) = (double)((tp.mts && b.n_bonds) ? b.n_bonds[s] : n_bonds);
  TR(TR_ANG_VEL) = tp.dem ? b.f64[C_ANG_VEL][s] : 0.; TR(TR_ANG_ACCEL) = tp.dem ? b.f64[C_ANG_ACCEL][s] : 0.;
  TR(TR_ROT) = tp.dem ? b.f64[C_ROT][s] : 0.;
#undef TR
}

}  // namespace kid
