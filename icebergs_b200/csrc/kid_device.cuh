// kid_device.cuh -- device-side data model of the B200 KID hot path.
//
// Storage: the reference keeps heap-allocated type(iceberg) nodes in per-cell
// linked lists (F:290-359, F:416-423).  Here bergs are structure-of-arrays
// columns in HBM indexed by slot, kept (periodically) sorted by cell so that a
// warp's gathers fall in a handful of sectors of the packed grid records.
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
  BF_HALO = 8,       // halo copy (halo_berg >= 0.5)
  BF_ARRIVAL = 16,   // arrived by migration this step: thermodynamics still to do
  BF_COLLIDED = 32   // static_berg == 0.1: "had a collision" mark inside the MTS convergence loop (I:6685)
};

// module constants, I:68-80
#define KID_RHO_ICE 916.7
#define KID_RHO_WATER 999.8
#define KID_RHO_AIR 1.1
#define KID_RHO_SEAWATER 1025.
#define KID_GRAVITY 9.8
#define KID_CD_AV 1.3
#define KID_CD_AH 0.0055
#define KID_CD_WV 0.9
#define KID_CD_WH 0.0012
#define KID_CD_IV 0.9

// Scalars the kernels need, passed by value as a __grid_constant__ parameter.
struct DevParams {
  double dt, pi, pi_180, omega2 /* 2*omega */, Lx, Rearth;
  double lat_ref, rho_bergs, speed_limit, coastal_drift, ocean_drag_scale;
  double cdrag_grounding, h_to_init_grounding, u_override, v_override;
  double bergy_bit_erosion_fraction, sicn_shift, tip_parameter, melt_cutoff;
  double spring_coef, contact_spring_coef, contact_distance, radial_damping_coef, tangental_damping_coef;
  double fl_youngs, new_berg_from_fl_bits_mass_thres;
  double gamma_t_3eq, ustar_icebergs_bg, utide_icebergs, cdrag_icebergs, omega;   // find_basal_melt I:3492
  double current_yearday;
  // host-precomputed constants of the step (same libm as the CPU path)
  double rdt;            // 1/dt
  double rho_ratio;      // rho_bergs/rho_seawater
  double r_rho_bergs;    // 1/rho_bergs
  double r_h2ig;         // 1/h_to_init_grounding
  double rpow40_02;      // 1/40**0.2  (bergy bits of bergs whose smallest dimension is >= 40 m, I:3076)
  double dlat_dy;        // (180/pi)/Rearth  I:470
  double r180_pi;        // 180/pi
  double f_cori_plane;   // 2*omega*sin(pi_180*lat_ref)  I:2046
  double rect_add;       // 0.5 on the regular Cartesian path of pos_within_cell (F:6331), else 0
  int32_t grid_is_latlon, grid_is_regular, old_bug_bilin, use_roundoff_fix, use_f_plane;
  int32_t use_new_predictive_corrective, only_interactive_forces, override_iceberg_velocities;
  int32_t old_interp_flds_order, interactive_icebergs_on, iceberg_bonds_on, internal_bergs_for_drag;
  int32_t runge_not_verlet, pad_rk_;
  int32_t hexagonal_icebergs, critical_interaction_damping_on, tang_crit_int_damp_on, scale_damping_by_pmag;
  int32_t use_operator_splitting, set_melt_rates_to_zero, allow_bergs_to_roll, use_updated_rolling_scheme;
  int32_t iceberg_melt_without_decay, melt_diagnostics, footloose, mts, dem;
  int32_t contact_cells_lon, contact_cells_lat, max_bonds;
  int32_t current_year;
  int32_t passive_mode;
  int32_t use_mixed_melting, melt_icebergs_as_ice_shelf, use_three_equation_model, const_gamma;
  int32_t use_mixed_layer_salinity_for_thermo, apply_thickness_cutoff_to_bergs_melt;
  int32_t no_rotation;   // cos_rot == 1 and sin_rot == 0 on the whole data domain: rotate() (I:4953) is the identity
};

// corner record: the 8 B-grid fields interp_flds bilinearly gathers (I:4757-4765)
struct __align__(16) CornerRec { double uo, vo, ui, vi, ua, va, cosr, sinr; };
// cell record: A-grid picks (I:4815-4818), od (I:4897) and what thermodynamics /
// adjust_index_and_ground need from the cell; ddx/ddy = ddx_ssh/ddy_ssh (I:4903-4926)
struct __align__(16) CellRec { double sst, sss, cn, hi, od, rarea /* 1/area, 0 where area==0 */, ddx, ddy; };
// corner position record
struct __align__(16) LonLat { double lon, lat; };
// Cells that are exact rectangles in (lon,lat) (every cell of a regular grid away from the pole):
// calc_xiyj (F:6439) degenerates to xi = dx/alpha, yj = dy/epsilon, the regular Cartesian path
// (F:6325-6332) to xi = dx/|alpha| + 0.5.  x1,y1 = reference corner (SW, or the cell centre for the
// Cartesian path), ralpha/reps = the reciprocals.  ralpha = NaN: not such a cell, take the general path.
struct __align__(16) RectCell { double x1, y1, ralpha, reps; };

struct DevGrid {
  int32_t isd, ied, jsd, jed, isc, iec, jsc, jec, nid, njd, gni, gnj;
  int32_t cyclic_x, cyclic_y;
  int32_t pe_E_self, pe_W_self;     // cyclic x and this rank is its own E/W neighbour
  int32_t has_E, has_W, has_N, has_S;  // a neighbour rank exists in that direction
  int32_t fold_north, pad1;         // this tile touches a folded northern edge (KidDomain.fold_north and jec == gnj)
  // static (data domain)
  double *lon, *lat, *lonc, *latc, *dx, *dy, *area, *msk, *cosr, *sinr, *ocean_depth;
  // forcing (data domain)
  double *uo, *vo, *ui, *vi, *ua, *va, *ssh, *sst, *sss, *cn, *hi;
  double *calving, *calving_hflx;
  // packed records for the hot kernel
  CornerRec* corner;
  CellRec* cell;
  LonLat* lonlat;
  RectCell* rect;
  // flux / diagnostic outputs
  double *floating_melt, *berg_melt, *bergy_src, *bergy_melt, *fl_bits_melt, *fl_bits_src;
  double *melt_buoy, *melt_eros, *melt_conv, *melt_buoy_fl, *melt_eros_fl, *melt_conv_fl;
  double *fl_parent_melt, *fl_child_melt;
  double *stored_heat, *stored_ice, *real_calving, *tmp;
  int32_t* iceberg_counter_grd;
};

// SoA berg store.  Columns follow F:290-359 (Appendix B of SURVEY.md).
// The f64 columns are addressed through `f64[col]` so that generic code (sort
// gather, pack/unpack, get/set) can loop over them.
enum BergCol : int {
  C_LON = 0, C_LAT, C_UVEL, C_VVEL, C_AXN, C_AYN, C_BXN, C_BYN, C_UVEL_PREV, C_VVEL_PREV,
  C_XI, C_YJ, C_MASS, C_THICKNESS, C_WIDTH, C_LENGTH, C_MASS_SCALING, C_MASS_OF_BITS, C_HEAT_DENSITY,
  C_START_LON, C_START_LAT, C_START_DAY, C_START_MASS,
  C_MASS_OF_FL_BITS, C_MASS_OF_FL_BERGY_BITS, C_FL_K,
  C_NBASE,                                       // columns always allocated
  C_UVEL_OLD = C_NBASE, C_VVEL_OLD, C_LON_OLD, C_LAT_OLD,   // interactive
  C_NINTER,
  // MTS (F:349-359): the environment cache of interp_gridded_fields_to_bergs and the fast accelerations
  C_UO = C_NINTER, C_VO, C_UI, C_VI, C_UA, C_VA, C_SSH_X, C_SSH_Y, C_SST, C_SSS, C_CN, C_HI, C_OD,
  C_AXN_FAST, C_AYN_FAST, C_BXN_FAST, C_BYN_FAST,
  C_NMTS,
  C_ANG_VEL = C_NMTS, C_ANG_ACCEL, C_ROT,        // dem (F:355-357)
  C_NDEM,
  C_NCOLS = C_NDEM
};

struct DevBergs {
  int64_t capacity;
  double* f64[C_NCOLS];
  int64_t* id;
  int32_t *ine, *jne, *start_year;
  uint8_t *flags, *halo_code;
  int32_t* leaver_list;       // slots of the bergs that left the tile this step (multi-rank only)
  unsigned long long* leaver_count;   // entries of leaver_list filled this step (two list/count pairs alternate, step_core)
  int64_t leaver_cap;
  // bonds, type(bond) F:362-386: max_bonds half-bonds per berg, entry k of slot s at [k*capacity + s];
  // other_id == 0 marks an empty entry, other_slot is re-resolved after every sort (connect_all_bonds F:4963)
  int32_t max_bonds, pad_;
  int64_t* bond_other_id;
  int32_t* conglom_id;        // connected component of the bond graph (set_conglom_ids F:2601); 0 = halo copy not reached
  int32_t *bond_other_slot, *bond_other_ine, *bond_other_jne;
  double* bond_length;
  int32_t* n_bonds;           // assign_n_bonds F:4617 (MTS contact search skips interior elements)
  int32_t* bond_broken;       // dem only, else nullptr
  // dem bond history and the saved pair forces (type(bond) F:372-386, save_bond_forces F:53), same layout as bond_length
  double* bond_dem[11];
  double* ia_radius;          // scratch: interaction radius per berg, valid inside the shared-memory MTS kernel only
  // scratch of the shared-memory MTS kernel (<= 128 elements): per element the slots the sub-step contact search
  // has to look at (same conglomerate, outer layer, not a bond partner, in the 3x3 cells), in the search's own order;
  // rebuilt when a bond breaks (cand_dirty).  nullptr everywhere else.
  uint8_t* cand;
  int32_t* cand_n;
  int32_t* cand_dirty;
  int32_t cand_stride, pad_cand_;
};
enum BondDem : int { BD_TANGD1 = 0, BD_TANGD2, BD_REL_ROT, BD_NSTRESS, BD_SSTRESS, BD_FX, BD_FY, BD_FDX, BD_FDY, BD_T, BD_TD, BD_N };

// device-side counters (one struct in HBM per handle)
struct DevCounters {
  unsigned long long nbergs_melted, nbergs_calved, nbergs_calved_fl, nspeeding, n_bounced;
  unsigned long long n_leavers, n_lost, n_wrapped;
  unsigned long long n_slots;        // append cursor (high-water mark of used slots)
  unsigned long long n_alive;        // filled by the count kernel
  unsigned long long n_cell_moves;   // bergs whose cell changed this step (sort heuristics)
  unsigned long long n_leaver_list;  // entries of DevBergs::leaver_list filled this step
  double net_heat_to_ocean, net_calving_to_bergs, net_heat_to_bergs;
  unsigned int error_flags;
  unsigned int warn_adjust;
};

}  // namespace kid
