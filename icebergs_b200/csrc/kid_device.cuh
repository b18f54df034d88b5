// kid_device.cuh -- device-side data model of the B200 KID hot path.
//
// Storage: the reference keeps heap-allocated type(iceberg) nodes in per-cell
// linked lists (F:290-359, F:416-423).  Here bergs are structure-of-arrays
// columns in HBM indexed by slot, kept (periodically) sorted by cell so that a
// warp's gathers fall in a handful of 32-byte sectors of the packed grid records.
//
// Reference citations: I: = src/icebergs.F90, F: = src/icebergs_framework.F90.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/kid_b200.h"

namespace kid {

// berg flag bits (uint8 column `flags`)
enum : uint8_t {
  BF_ALIVE = 1,      // slot holds a berg
  BF_STATIC = 2,     // static_berg >= 0.5   (F:323)
  BF_LEAVER = 4,     // left the rank's compute domain this step, awaiting migration
  BF_HALO = 8        // halo copy (halo_berg >= 0.5; exact code in column halo_code)
};

// Scalars the kernels need, passed by value (lives in the constant bank).
struct DevParams {
  double dt, pi, pi_180, omega2 /* 2*omega */, Lx, invLx, Rearth;
  double lat_ref, rho_bergs, speed_limit, coastal_drift, ocean_drag_scale;
  double cdrag_grounding, h_to_init_grounding, u_override, v_override;
  double bergy_bit_erosion_fraction, sicn_shift, tip_parameter, melt_cutoff;
  double spring_coef, contact_spring_coef, contact_distance, radial_damping_coef, tangental_damping_coef;
  double fl_youngs, rho_ratio /* rho_bergs/rho_seawater */;
  double initial_mass_s[KID_NCLASSES], initial_mass_n[KID_NCLASSES];
  int32_t grid_is_latlon, grid_is_regular, old_bug_bilin, use_roundoff_fix, use_f_plane;
  int32_t use_new_predictive_corrective, only_interactive_forces, override_iceberg_velocities;
  int32_t old_interp_flds_order, interactive_icebergs_on, iceberg_bonds_on, internal_bergs_for_drag;
  int32_t hexagonal_icebergs, critical_interaction_damping_on, tang_crit_int_damp_on, scale_damping_by_pmag;
  int32_t use_operator_splitting, set_melt_rates_to_zero, allow_bergs_to_roll, use_updated_rolling_scheme;
  int32_t iceberg_melt_without_decay, melt_diagnostics, footloose, mts, dem;
  int32_t contact_cells_lon, contact_cells_lat, max_bonds;
  int32_t current_year; int32_t pad0;
  double current_yearday;
};

// corner record: the 8 B-grid fields interp_flds bilinearly gathers (I:4757-4765)
struct __align__(16) CornerRec { double uo, vo, ui, vi, ua, va, cosr, sinr; };
// cell record: A-grid picks (I:4815-4818), od (I:4897), and what thermodynamics /
// adjust_index_and_ground need from the cell
struct __align__(16) CellRec { double sst, cn, hi, od, area, msk, sss, depth; };
// corner position record
struct __align__(16) LonLat { double lon, lat; };

struct DevGrid {
  int32_t isd, ied, jsd, jed, isc, iec, jsc, jec, nid, njd, gni, gnj;
  int32_t cyclic_x, cyclic_y, wrap_x_local /* cyclic x and this rank spans all of x */, pad;
  // static (data domain)
  double *lon, *lat, *lonc, *latc, *dx, *dy, *area, *msk, *cosr, *sinr, *ocean_depth;
  // forcing (data domain)
  double *uo, *vo, *ui, *vi, *ua, *va, *ssh, *sst, *sss, *cn, *hi;
  double *calving, *calving_hflx;
  // packed per-step records for the hot kernel
  CornerRec* corner;
  CellRec* cell;
  LonLat* lonlat;
  double *ddx, *ddy;   // ddx_ssh / ddy_ssh per cell (I:4903-4926)
  // flux / diagnostic outputs
  double *floating_melt, *berg_melt, *bergy_src, *bergy_melt, *fl_bits_melt, *fl_bits_src;
  double *melt_buoy, *melt_eros, *melt_conv, *melt_buoy_fl, *melt_eros_fl, *melt_conv_fl;
  double *fl_parent_melt, *fl_child_melt;
  double *stored_heat, *stored_ice, *real_calving;
  int32_t* iceberg_counter_grd;
};

// SoA berg store.  Columns follow F:290-359 (Appendix B of SURVEY.md).
struct DevBergs {
  int64_t capacity;
  double *lon, *lat, *uvel, *vvel, *axn, *ayn, *bxn, *byn, *uvel_prev, *vvel_prev;
  double *xi, *yj, *mass, *thickness, *width, *length, *mass_scaling, *mass_of_bits, *heat_density;
  double *start_lon, *start_lat, *start_day, *start_mass;
  double *mass_of_fl_bits, *mass_of_fl_bergy_bits, *fl_k;
  double *uvel_old, *vvel_old, *lon_old, *lat_old;           // interactive only
  double *env;                                                // 13 x capacity env cache (new interp order)
  double *axn_fast, *ayn_fast, *bxn_fast, *byn_fast;          // mts
  double *ang_vel, *ang_accel, *rot;                          // dem
  int64_t* id;
  int32_t *ine, *jne, *start_year, *n_bonds, *conglom_id;
  uint8_t *flags, *halo_code;
};

// device-side counters (one struct in HBM per handle)
struct DevCounters {
  unsigned long long n_alive;        // slots in use (high-water mark = n_slots on host)
  unsigned long long nbergs_melted, nbergs_calved, nbergs_calved_fl, nspeeding, n_bounced;
  unsigned long long n_leavers, n_lost;
  unsigned long long n_slots;        // append cursor
  double net_heat_to_ocean, net_calving_to_bergs, net_heat_to_bergs;
  unsigned int error_flags;
  unsigned int warn_adjust;
};

}  // namespace kid
