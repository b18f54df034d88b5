/*
 * kid_b200.h -- C ABI of the B200-native KID iceberg hot path.
 *
 * This is the drop-in boundary: the entry points below are exactly what an
 * ISO_C_BINDING Fortran shim named `ice_bergs` binds in place of the reference's
 * own module procedures (reference = NOAA-GFDL/icebergs; I: = src/icebergs.F90,
 * F: = src/icebergs_framework.F90, D: = driver/icebergs_driver.F90):
 *
 *   kid_init            <- icebergs_init        I:92-178  (+ ice_bergs_framework_init F:641-1712)
 *   kid_run             <- icebergs_run         I:5074-5887
 *   kid_end             <- icebergs_end         I:8152-8258
 *   kid_set_bergs       <- read_restart_bergs   (columns of icebergs.res.nc, fmsio:261-337)
 *   kid_get_bergs       <- write_restart_bergs / icebergs_save_restart I:8136
 *   kid_set_bonds       <- read_restart_bonds / initialize_iceberg_bonds I:356-441
 *   kid_set_calving_state <- read_restart_calving (stored_ice, stored_heat, iceberg_counter_grd)
 *   kid_stock           <- icebergs_stock_pe    I:8102
 *   kid_incr_mass       <- icebergs_incr_mass   I:6046
 *
 * Conventions
 *   - every array argument is caller-owned HOST memory, column-major (Fortran)
 *     fp64 unless the type says otherwise; the library owns only the opaque handle.
 *   - every function returns 0 on success; on failure kid_last_error() gives the
 *     text the shim forwards to FMS error_mesg(..., FATAL).
 *   - one CUDA stream per handle; a handle is not thread-safe; distinct handles
 *     (ranks / GPUs) are independent.
 *   - there is no CPU fallback: every entry point that computes fails with
 *     KID_ERR_NO_DEVICE when no sm_100a device is present.
 *
 * Index convention (identical to the reference): cell (i,j) has its NE corner at
 * lon(i,j),lat(i,j); compute domain isc:iec x jsc:jec, data domain = compute +
 * `halo` ring; indices are GLOBAL (1-based) as handed out by mpp_define_domains.
 */
#ifndef KID_B200_H
#define KID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KID_NCLASSES 10          /* F:24  nclasses */
#define KID_ABI_VERSION 3

/* status codes */
enum {
  KID_OK = 0,
  KID_ERR_ARG = 1,
  KID_ERR_NO_DEVICE = 2,
  KID_ERR_CUDA = 3,
  KID_ERR_CAPACITY = 4,
  KID_ERR_UNSUPPORTED = 5,
  KID_ERR_STATE = 6,
  KID_ERR_COMM = 7
};

/* staggering codes of icebergs_run (FMS mpp_parameter_mod values are opaque to
 * us; the shim maps BGRID_NE/CGRID_NE/AGRID to these)  I:5119-5120 */
enum { KID_BGRID_NE = 0, KID_CGRID_NE = 1, KID_AGRID = 2 };

/* KidDomain.comm_kind */
enum { KID_COMM_NCCL = 0, KID_COMM_LOCAL = 1 };

/* stocks, icebergs_stock_pe I:8102 */
enum { KID_ISTOCK_WATER = 0, KID_ISTOCK_HEAT = 1 };

/* ids for kid_get_grid_field (all on the data domain isd:ied x jsd:jed) */
enum {
  KID_FLD_FLOATING_MELT = 0, KID_FLD_BERG_MELT, KID_FLD_BERGY_SRC, KID_FLD_BERGY_MELT,
  KID_FLD_FL_BITS_MELT, KID_FLD_FL_BITS_SRC, KID_FLD_CALVING_HFLX, KID_FLD_CALVING,
  KID_FLD_MELT_BUOY, KID_FLD_MELT_EROS, KID_FLD_MELT_CONV,
  KID_FLD_MELT_BUOY_FL, KID_FLD_MELT_EROS_FL, KID_FLD_MELT_CONV_FL,
  KID_FLD_FL_PARENT_MELT, KID_FLD_FL_CHILD_MELT,
  KID_FLD_UO, KID_FLD_VO, KID_FLD_UI, KID_FLD_VI, KID_FLD_UA, KID_FLD_VA,
  KID_FLD_SSH, KID_FLD_SST, KID_FLD_SSS, KID_FLD_CN, KID_FLD_HI,
  KID_FLD_LON, KID_FLD_LAT, KID_FLD_LONC, KID_FLD_LATC, KID_FLD_DX, KID_FLD_DY,
  KID_FLD_AREA, KID_FLD_MSK, KID_FLD_COS, KID_FLD_SIN, KID_FLD_OCEAN_DEPTH,
  KID_FLD_STORED_HEAT, KID_FLD_MASS, KID_FLD_BERGY_MASS, KID_FLD_SPREAD_MASS,
  KID_FLD_SPREAD_AREA, KID_FLD_USTAR_ICEBERG, KID_FLD_SPREAD_UVEL, KID_FLD_SPREAD_VVEL,
  KID_FLD_RMEAN_CALVING, KID_FLD_RMEAN_CALVING_HFLX,
  KID_FLD_COUNT_
};

/* ----------------------------------------------------------------------------
 * KidParams: the subset of &icebergs_nml (F:825-856, defaults F:686-822) and of
 * the FMS constants that changes hot-path arithmetic.  kid_default_params()
 * fills in the reference defaults.  Logical flags are int32 (0/1).
 * -------------------------------------------------------------------------- */
typedef struct KidParams {
  int32_t abi_version;            /* = KID_ABI_VERSION */
  int32_t halo;                   /* F:686 (4) */
  double dt;                      /* icebergs_init argument dt (I:105) -> bergs%dt F:1328 */
  /* FMS constants_mod (external to the reference tree; GFDL defaults) */
  double pi;                      /* 3.14159265358979323846 */
  double omega;                   /* 7.292e-5  s^-1 */
  double radius;                  /* 6371.0e3  m  (fractional_area only) */
  double hlf;                     /* 3.34e5    J/kg */
  /* frame */
  int32_t grid_is_latlon;         /* F:748 (T) */
  int32_t grid_is_regular;        /* F:749 (T) */
  double Lx;                      /* F:712 (360.) ; <=0 => not periodic */
  double Rearth;                  /* F:44  (6360000.) */
  /* stepping */
  int32_t runge_not_verlet;       /* F:733 (T) -- this library implements Verlet (F) only */
  int32_t old_bug_bilin;          /* F:40  (T) */
  int32_t use_roundoff_fix;       /* F:37  (T) */
  int32_t use_f_plane;            /* F:747 (F) */
  double lat_ref;                 /* F:709 (0.) */
  int32_t use_new_predictive_corrective; /* F:770 (F); forced T under Verlet I:2012 */
  int32_t only_interactive_forces;/* F:757 */
  int32_t override_iceberg_velocities; /* F:746 */
  double u_override, v_override;  /* F:710-711 */
  int32_t static_icebergs;        /* F:756 */
  int32_t old_interp_flds_order;  /* derived F:1483: T unless mts/dem/footloose */
  double rho_bergs;               /* F:694 (850.) */
  double speed_limit;             /* F:726 (0.) */
  double coastal_drift;           /* F:731 */
  double tidal_drift;             /* F:732 (must be 0: FMS RNG is external) */
  double ocean_drag_scale;        /* F:813 (1.) */
  double cdrag_grounding;         /* F:697 (0.) */
  double h_to_init_grounding;     /* F:698 (100.) */
  int32_t tau_is_velocity;        /* F:727 (F) */
  int32_t add_iceberg_thickness_to_ssh; /* F:745 (unsupported: must be 0) */
  /* interactions (I:480-804) */
  int32_t interactive_icebergs_on;/* F:771 */
  int32_t iceberg_bonds_on;       /* F:51  */
  int32_t internal_bergs_for_drag;/* F:735 */
  int32_t hexagonal_icebergs;     /* F:753 */
  int32_t critical_interaction_damping_on; /* F:773 (T) */
  int32_t tang_crit_int_damp_on;  /* F:774 (T) */
  int32_t scale_damping_by_pmag;  /* F:772 (T) */
  int32_t max_bonds;              /* F:693 (6); 0 when bonds are off F:1264 */
  double spring_coef;             /* F:695 (1e-8) */
  double contact_spring_coef;     /* F:696 (<=0 -> spring_coef, F:1312) */
  double contact_distance;        /* F:782 (0.) */
  double radial_damping_coef;     /* F:704 (1e-4) */
  double tangental_damping_coef;  /* F:705 (2e-5) */
  int32_t contact_cells_lon;      /* derived F:1492-1519 */
  int32_t contact_cells_lat;
  int32_t manually_initialize_bonds; /* F:769 */
  int32_t manually_initialize_bonds_from_radii; /* F:725 */
  double length_for_manually_initialize_bonds;  /* F:724 (1000.) */
  /* MTS / DEM (I:1278-1947, I:6576-7078).  This build: one rank, conglomerates clear of the cyclic seam
   * (transfer_mts_bergs F:2144 is not built), save_bond_forces = T, skip_first_outer_mts_step = F. */
  int32_t mts;                    /* F:48 */
  int32_t mts_sub_steps;          /* F:780 (-1 = auto F:1296-1301) */
  int32_t force_convergence;      /* F:783 */
  int32_t explicit_inner_mts;     /* F:784 (forced T when dem F:1433) */
  double convergence_tolerance;   /* F:785 (1e-8) */
  int32_t dem;                    /* F:52 */
  int32_t ignore_tangential_force;/* F:802 */
  int32_t orig_dem_moment_of_inertia; /* F:60 */
  int32_t fracture_criterion_stress;  /* F:800 'stress' => 1, 'none' => 0 */
  int32_t break_bonds_on_sub_steps;   /* F:61 */
  int32_t use_broken_bonds_for_substep_contact; /* F:806 */
  int32_t save_bond_forces;       /* F:53 (T) */
  int32_t constant_interaction_LW;/* F:810 */
  int32_t short_step_mts_grounding; /* F:54 */
  int32_t radius_based_drag;      /* F:55 */
  int32_t use_grounding_torque;   /* F:801 */
  int32_t dem_beam_test;          /* F:808 */
  int32_t skip_first_outer_mts_step; /* F:62 */
  int32_t no_frac_first_ts;       /* F:63 */
  int32_t remove_unused_bergs;    /* F:781 */
  double poisson;                 /* F:803 (0.3) */
  double dem_spring_coef;         /* F:804 */
  double dem_damping_coef;        /* F:805 (0.1) */
  double frac_thres_n, frac_thres_t; /* F:699-700 already scaled by frac_thres_scaling F:1355 */
  double constant_length, constant_width; /* F:811-812 */
  /* thermodynamics (I:2844-3300) */
  int32_t use_operator_splitting; /* F:720 (T) */
  int32_t set_melt_rates_to_zero; /* F:751 */
  int32_t allow_bergs_to_roll;    /* F:752 (T) */
  int32_t use_updated_rolling_scheme; /* F:738 (F) */
  int32_t iceberg_melt_without_decay; /* F:744 */
  int32_t use_mixed_melting;      /* F:734 */
  int32_t melt_icebergs_as_ice_shelf; /* F:743 */
  int32_t apply_thickness_cutoff_to_bergs_melt;   /* F:737 */
  int32_t apply_thickness_cutoff_to_gridded_melt; /* F:736 */
  int32_t melt_diagnostics;       /* 1 => fill melt_buoy/eros/conv(+_fl), fl_parent/child_melt
                                     (reference: diag ids registered, I:3146-3198) */
  int32_t passive_mode;           /* F:722 */
  double bergy_bit_erosion_fraction; /* F:707 (0.) */
  double sicn_shift;              /* F:708 (0.) */
  double tip_parameter;           /* F:729 (0.) */
  double melt_cutoff;             /* F:718 (-1.) */
  /* footloose (I:2503-2841, I:6405-6569) */
  int32_t footloose;              /* F:64 */
  int32_t displace_fl_bergs;      /* F:819 (T) -- needs FMS RNG: must be 0 here */
  int32_t fl_style_fl_bits;       /* F:820: 'fl_bits' => 1, 'new_bergs' => 0 */
  int32_t fl_bits_erosion_to_bergy_bits; /* F:821 (T) */
  double fl_youngs;               /* F:817 (1e7) */
  double fl_strength;             /* F:818 (250.) */
  double new_berg_from_fl_bits_mass_thres; /* F:822 (1e12) */
  /* calving classes (F:787-796, F:1534-1550) */
  double LoW_ratio;               /* F:706 (1.5) */
  double initial_mass_s[KID_NCLASSES], distribution_s[KID_NCLASSES];
  double mass_scaling_s[KID_NCLASSES], initial_thickness_s[KID_NCLASSES];
  double initial_mass_n[KID_NCLASSES], distribution_n[KID_NCLASSES];
  double mass_scaling_n[KID_NCLASSES], initial_thickness_n[KID_NCLASSES];
  /* ice-shelf style basal melt, find_basal_melt I:3492-3826 (used with use_mixed_melting /
   * melt_icebergs_as_ice_shelf) */
  int32_t use_three_equation_model;            /* F:742 (T) */
  int32_t const_gamma;                         /* F:719 (T) */
  int32_t use_mixed_layer_salinity_for_thermo; /* F:740 (F) */
  int32_t pad0_;
  double gamma_t_3eq;                          /* F:717 (0.022) */
  double ustar_icebergs_bg;                    /* F:715 (0.001) */
  double utide_icebergs;                       /* F:714 (0.) */
  double cdrag_icebergs;                       /* F:716 (1.5e-3) */
  /* mass / area / momentum spread onto the ocean grid, I:3895-4100, I:4970-5011, I:6077-6150 */
  int32_t add_weight_to_ocean;                 /* F:721 (T) */
  int32_t time_average_weight;                 /* F:723 (F).  T: the weight is spread inside the stepping stages (I:7264, I:7395..) and
                                                  calculate_mass_on_ocean I:4984-4997 zeroes it again before anything reads it: the
                                                  spread fields of such a run are zero, here as in the reference */
  int32_t use_old_spreading;                   /* F:778 (T) */
  int32_t rotate_icebergs_for_mass_spreading;  /* F:750 (T) */
  int32_t pass_fields_to_ocean_model;          /* F:739 (F): also fill spread_area / spread_uvel / spread_vvel / ustar */
  int32_t old_bug_rotated_weights;             /* F:38 (F): skip the 180-degree turn of the nine weights beyond a folded northern edge, I:6110 */
  double grounding_fraction;                   /* F:730 (0.) */
  double clipping_depth;                       /* F:227 (0.) */
  double initial_orientation;                  /* F:713 (0.) degrees */
  double tau_calving;                          /* F:728 (0.) years: running mean of calving / calving_hflx, get_running_mean_calving I:5999-6038 */
  int32_t find_melt_using_spread_mass;         /* F:741 (F): floating_melt from the spread mass before / after the melt, I:5490-5500,
                                                  I:3436-3448 (free-drifting Verlet bergs; refused with RK4, interactions, footloose,
                                                  MTS or Iceberg_melt_without_decay) */
  /* trajectory sampling, record_posn F:5328-5498 (WHEN to sample is the caller's decision, I:5173-5178: kid_record_posn) */
  int32_t save_short_traj;                     /* F:759 (T): the file holds lon, lat, year, day, id only */
  int32_t save_fl_traj;                        /* F:762 (T): + masses, thickness, velocity (and the footloose state) */
  int32_t save_nonfl_traj_by_class;            /* F:764 (F) */
  double traj_area_thres;                      /* F:687 (0.) km^2: area a non-bonded berg must have to be sampled */
  double traj_area_thres_sntbc;                /* F:688 (0.) km^2, with save_nonfl_traj_by_class */
  double traj_area_thres_fl;                   /* F:689 (1e9) km^2, footloose children (fl_k < 0) */
  double save_all_traj_year;                   /* F:763 (huge): from this year on every berg is sampled */
  double save_traj_by_class_start_mass_thres_n;/* F:765 (0.) */
  double save_traj_by_class_start_mass_thres_s;/* F:766 (0.) */
} KidParams;

/* ----------------------------------------------------------------------------
 * KidDomain: what mpp_define_domains / mpp_get_*_domain / mpp_get_neighbor_pe
 * give the reference (F:915-930).  pe_* = -1 is NULL_PE.  For a single rank that
 * is cyclic in x, pe_E = pe_W = rank.
 * -------------------------------------------------------------------------- */
typedef struct KidDomain {
  int32_t gni, gnj;               /* global cells */
  int32_t isc, iec, jsc, jec;     /* compute domain (global, 1-based) */
  int32_t isd, ied, jsd, jed;     /* data domain = compute +/- halo */
  int32_t cyclic_x, cyclic_y;     /* CYCLIC_GLOBAL_DOMAIN flags */
  int32_t rank, nranks;
  int32_t layout_x, layout_y;     /* ranks in i and j; rank = px + layout_x*py */
  int32_t pe_N, pe_S, pe_E, pe_W;
  int32_t device;                 /* CUDA device ordinal to use */
  int32_t comm_kind;              /* KID_COMM_NCCL (0) or KID_COMM_LOCAL (1): what nccl_comm points to */
  void*   nccl_comm;              /* ncclComm_t, or a kid_local_comm group, or NULL (single rank) */
  int32_t fold_north;             /* FOLD_NORTH_EDGE (tripolar grid, F:649, F:933): row gnj is folded onto itself, cell
                                     (i, gnj+k) is cell (gni+1-i, gnj+1-k) turned by 180 degrees.  Needs cyclic_x.  kid_single_domain /
                                     kid_define_domain leave it 0: the caller sets it from its mpp_define_domains y-flags */
  int32_t pad_;
} KidDomain;

/* ----------------------------------------------------------------------------
 * Berg columns: one entry per variable of icebergs.res.nc (fmsio:261-337) plus
 * the non-persisted ones that cross ranks (F:3268-3302).  Any pointer may be
 * NULL on input (=> reference default: 0, or derived); on output NULL columns
 * are skipped.
 * -------------------------------------------------------------------------- */
typedef struct KidBergColumns {
  double *lon, *lat, *uvel, *vvel;
  double *mass, *thickness, *width, *length;
  double *axn, *ayn, *bxn, *byn;
  double *uvel_prev, *vvel_prev;
  double *uvel_old, *vvel_old, *lon_old, *lat_old;
  double *start_lon, *start_lat, *start_day, *start_mass;
  double *mass_scaling, *mass_of_bits, *mass_of_fl_bits, *mass_of_fl_bergy_bits;
  double *fl_k, *heat_density;
  double *halo_berg, *static_berg;
  double *xi, *yj;
  double *axn_fast, *ayn_fast, *bxn_fast, *byn_fast;   /* mts */
  double *ang_vel, *ang_accel, *rot;                   /* dem */
  double *uo, *vo, *ui, *vi, *ua, *va, *ssh_x, *ssh_y, *sst, *sss, *cn, *hi, *od; /* env cache */
  int32_t *start_year, *ine, *jne;
  int32_t *n_bonds, *conglom_id;
  int64_t *id;
} KidBergColumns;

/* bonds: one row per (berg, partner) half-bond, as bonds_iceberg.res.nc (fmsio:473-493) */
typedef struct KidBondColumns {
  int64_t *first_id, *other_id;
  int32_t *first_ine, *first_jne, *other_ine, *other_jne;
  double *length;                                      /* may be NULL */
  double *tangd1, *tangd2, *nstress, *sstress, *rel_rotation; /* dem; may be NULL */
  int32_t *broken;                                     /* dem; may be NULL */
} KidBondColumns;

/* ----------------------------------------------------------------------------
 * Trajectory samples: one entry per record of iceberg_trajectories.nc (type(xyt) F:257-287, write_trajectory
 * fmsio:1575-2047).  Any pointer may be NULL (column skipped).
 * -------------------------------------------------------------------------- */
typedef struct KidTrajColumns {
  double *lon, *lat, *day;
  int32_t *year;
  int64_t *id;
  double *mass, *start_mass, *thickness, *mass_of_bits, *uvel, *vvel;                   /* save_fl_traj */
  double *mass_scaling, *mass_of_fl_bits, *mass_of_fl_bergy_bits, *fl_k;                /* ... with footloose */
  double *uvel_prev, *vvel_prev, *heat_density, *width, *length;                        /* .not. save_short_traj */
  double *uo, *vo, *ui, *vi, *ua, *va, *ssh_x, *ssh_y, *sst, *sss, *cn, *hi;
  double *axn, *ayn, *bxn, *byn, *halo_berg, *static_berg, *od;
  double *axn_fast, *ayn_fast, *bxn_fast, *byn_fast;                                    /* mts */
  int32_t *n_bonds;                                                                     /* iceberg_bonds_on */
  double *ang_vel, *ang_accel, *rot;                                                    /* dem */
} KidTrajColumns;

/* scalar budget / event counters accumulated by kid_run (reference: I:5685-5776,
 * type(icebergs) F:541-568) */
typedef struct KidCounters {
  int64_t nbergs;                 /* bergs owned by this rank now */
  int64_t nbergs_calved;
  int64_t nbergs_calved_fl;
  int64_t nbergs_melted;
  int64_t nspeeding_tickets;
  int64_t nbonds;
  int64_t n_sent, n_received;     /* last step's migration */
  int64_t n_bounced;              /* last step, diagnostic (not in the reference) */
  double  net_heat_to_ocean;
  double  net_calving_to_bergs, net_heat_to_bergs;
  int32_t error_flags;            /* device-side fatal conditions, see KID_DEVERR_* */
  int32_t reserved;
} KidCounters;

enum {
  KID_DEVERR_GROUNDED = 1,        /* thermodynamics: area(i,j)==0  I:3207 */
  KID_DEVERR_COMPLEX_ROOTS = 2,   /* calc_xiyj F:6502 */
  KID_DEVERR_NOT_INVERTIBLE = 4,  /* calc_xiyj F:6529 */
  KID_DEVERR_OFF_PE = 8,          /* is_point_in_cell F:6099 */
  KID_DEVERR_CAPACITY = 16,       /* append beyond capacity */
  KID_DEVERR_LOST_BERG = 32       /* unpack could not place a berg F:3668 */
};

typedef struct kid_handle kid_t;

/* reference namelist defaults + FMS constants */
void kid_default_params(KidParams* p);

/* single-rank convenience: fills KidDomain for a gni x gnj domain on one rank */
void kid_single_domain(KidDomain* d, int32_t gni, int32_t gnj, int32_t halo,
                       int32_t cyclic_x, int32_t cyclic_y, int32_t device);

/* restatement of mpp_define_layout + mpp_define_domains for nranks (D:157-164,
 * F:915-930); returns 0 or an error code */
int32_t kid_define_domain(KidDomain* d, int32_t gni, int32_t gnj, int32_t halo,
                          int32_t cyclic_x, int32_t cyclic_y,
                          int32_t rank, int32_t nranks, int32_t device);

/*
 * icebergs_init  I:92-117.
 *   lon, lat, area, ocean_depth : (isc:iec, jsc:jec)
 *   wet, dx, dy, cos_rot, sin_rot: (isc-1:iec+1, jsc-1:jec+1)   D:341-344, F:1021-1056
 *   ocean_depth may be NULL.  capacity = max bergs this rank can hold (0 => 1<<20).
 */
int32_t kid_init(kid_t** h, const KidParams* p, const KidDomain* dom,
                 int32_t year, double yearday, int64_t capacity,
                 const double* lon, const double* lat, const double* wet,
                 const double* dx, const double* dy, const double* area,
                 const double* cos_rot, const double* sin_rot,
                 const double* ocean_depth, int32_t fractional_area);

int32_t kid_set_bergs(kid_t* h, int64_t n, const KidBergColumns* cols);
int32_t kid_get_bergs(kid_t* h, int64_t* n, KidBergColumns* cols, int32_t include_halo);
int32_t kid_count_bergs(kid_t* h, int64_t* n);

int32_t kid_set_bonds(kid_t* h, int64_t nb, const KidBondColumns* cols);
int32_t kid_get_bonds(kid_t* h, int64_t* nb, KidBondColumns* cols);

/* stored_ice (isd:ied,jsd:jed,10), stored_heat, iceberg_counter_grd (isd:ied,jsd:jed);
 * any may be NULL (left unchanged) */
int32_t kid_set_calving_state(kid_t* h, const double* stored_ice, const double* stored_heat,
                              const int32_t* iceberg_counter_grd);
int32_t kid_get_calving_state(kid_t* h, double* stored_ice, double* stored_heat,
                              int32_t* iceberg_counter_grd);

/* rmean_calving, rmean_calving_hflx (isd:ied,jsd:jed): the running means of get_running_mean_calving I:5999-6038
 * (tau_calving > 0) as calving.res.nc carries them (fmsio:568-569, read fms2io:1517-1534).  NULL = left alone.  A mean that
 * was set counts as initialised; otherwise the first kid_run starts it from that call's field (I:6010-6017). */
int32_t kid_set_calving_rmean(kid_t* h, const double* rmean_calving, const double* rmean_calving_hflx);
int32_t kid_get_calving_rmean(kid_t* h, double* rmean_calving, double* rmean_calving_hflx);

/*
 * icebergs_run  I:5074-5096.
 *   calving, calving_hflx (inout), tauxa, tauya, sst, sss : (isc:iec, jsc:jec)
 *   uo, vo, ui, vi, ssh, cn, hi                           : (isc-1:iec+1, jsc-1:jec+1)
 *   sss, mass_berg, ustar_berg, area_berg may be NULL.
 */
int32_t kid_run(kid_t* h, int32_t year, double yearday,
                double* calving, const double* uo, const double* vo,
                const double* ui, const double* vi,
                const double* tauxa, const double* tauya,
                const double* ssh, const double* sst, double* calving_hflx,
                const double* cn, const double* hi,
                int32_t stagger, int32_t stress_stagger, const double* sss,
                double* mass_berg, double* ustar_berg, double* area_berg);

/*
 * record_posn F:5328-5498: sample the bergs of the compute domain that pass the trajectory criteria (area thresholds,
 * bonded, save_all_traj_year, ...) at the time of the last kid_run into a device-side trajectory store.  The reference
 * decides when to sample from the model date (sample_traj I:5173-5178, every traj_sample_hrs); that decision stays with
 * the caller, who holds the date.  kid_get_trajectory hands the samples back (n: capacity in, count out; clear /= 0
 * empties the store afterwards), in no particular order -- push_posn / move_trajectory (F:5502, F:5611) keep per-berg
 * lists that write_trajectory flattens; the records, not their order, are the data.
 */
int32_t kid_record_posn(kid_t* h);
int32_t kid_trajectory_count(kid_t* h, int64_t* n);
int32_t kid_get_trajectory(kid_t* h, int64_t* n, KidTrajColumns* c, int32_t clear);

/*
 * Optional: announce the inputs of the NEXT kid_run() early.  Their host-to-device copies are queued on a copy stream
 * and overlap the step that is still computing; the call returns at once.  The arrays (pinned host memory for a
 * real overlap) must stay unchanged until the kid_run() that consumes them has returned; that kid_run() must be given
 * exactly these pointers, otherwise the prefetch is dropped and kid_run() copies as usual.  No reference counterpart:
 * icebergs_run (I:5074) receives its forcing when it is called; a coupler that has the next step's fields ready
 * (the stand-alone driver's forcing is analytic, D:341-420) hides the PCIe transfer this way.
 */
int32_t kid_prefetch_forcing(kid_t* h,
                             const double* calving, const double* uo, const double* vo,
                             const double* ui, const double* vi,
                             const double* tauxa, const double* tauya,
                             const double* ssh, const double* sst, const double* calving_hflx,
                             const double* cn, const double* hi, const double* sss);

/*
 * Device-resident variant used for the HBM-resident throughput figure: the
 * forcing set by the last kid_run()/kid_set_forcing() stays on the device and
 * `nsteps` steps are taken with it (no host<->device traffic, no return fields).
 */
int32_t kid_set_forcing(kid_t* h,
                        const double* calving, const double* uo, const double* vo,
                        const double* ui, const double* vi,
                        const double* tauxa, const double* tauya,
                        const double* ssh, const double* sst, const double* calving_hflx,
                        const double* cn, const double* hi,
                        int32_t stagger, int32_t stress_stagger, const double* sss);
int32_t kid_step_resident(kid_t* h, int32_t nsteps, int32_t year, double yearday);

/* timing of the phases of the last kid_run / kid_step_resident call, in ms,
 * measured with CUDA events on the handle's stream.  Names follow the FMS clocks
 * F:896-903: [0] interface (forcing ingest) [1] calving [2] momentum
 * [3] communication [4] thermodyn [5] sort [6] total */
int32_t kid_last_timing(kid_t* h, double ms[8]);
/* number of kernels launched by this handle so far */
int64_t kid_kernel_launches(kid_t* h);

int32_t kid_get_counters(kid_t* h, KidCounters* c);
int32_t kid_get_grid_field(kid_t* h, int32_t field_id, double* out /* (isd:ied,jsd:jed) */);
int32_t kid_stock(kid_t* h, int32_t index, double* value);
int32_t kid_incr_mass(kid_t* h, double* mass /* (isc:iec,jsc:jec) inout */);

/* force a cell-sort of the berg store now (normally periodic, internal) */
int32_t kid_sort_bergs(kid_t* h);
/* The periodic sort of free-drifting bergs runs every `interval` steps (KID_SORT_INTERVAL, default 32; 0 keeps the
 * current value); steps_since_sort sets where in that period the next step falls (< 0 keeps it).  Measurement
 * control: bench.py places the sorts so that every timed window is charged ceil(steps/interval) of them. */
int32_t kid_set_sort_phase(kid_t* h, int32_t interval, int32_t steps_since_sort);
/* cell sorts done by this handle so far */
int64_t kid_sorts_done(kid_t* h);
/* bergs the fast step kernel handed to the slow-path kernel in the last step (diagnostics; -1 = no fast launch yet) */
int64_t kid_last_slow_count(kid_t* h);
int32_t kid_synchronize(kid_t* h);

int32_t kid_end(kid_t** h);

/* rank that owns global cell (i,j) under d's layout; -1 = outside the model (NULL_PE).  i may be
 * any number of periods off on a cyclic-x domain (send_bergs_to_other_pes F:3024-3050 routes by
 * direction; here a berg goes straight to its owner). */
int32_t kid_owner_rank(const KidDomain* d, int32_t i, int32_t j);

/* ----------------------------------------------------------------------------
 * Multi-rank plumbing (replaces mpp_send/mpp_recv of send_bergs_to_other_pes,
 * F:3053-3194).  The Fortran shim has MPI but no NCCL binding, so the library
 * wraps the three NCCL calls it needs: rank 0 calls kid_nccl_unique_id, the
 * shim broadcasts the 128 bytes (MPI_Bcast / mpp_broadcast), every rank calls
 * kid_nccl_init and stores the result in KidDomain.nccl_comm.  NCCL is loaded
 * at run time (dlopen of libnccl.so.2): single-rank use needs no NCCL at all.
 * -------------------------------------------------------------------------- */
#define KID_NCCL_UNIQUE_ID_BYTES 128
int32_t kid_nccl_unique_id(char* out, int32_t nbytes);
int32_t kid_nccl_init(void** comm, const char* id, int32_t nbytes, int32_t nranks, int32_t rank,
                      int32_t device);
int32_t kid_nccl_destroy(void* comm);
/* In-process group: `nranks` handles that live in ONE process (one host thread per rank; several
 * tiles on one GPU, or GPUs of a host without NCCL) exchange halos and migrating bergs with
 * device-to-device copies after a rendezvous.  Every rank's kid_init / kid_run / kid_step_resident /
 * kid_set_forcing must be called concurrently from its own thread, as MPI ranks would.
 * Store the group in KidDomain.nccl_comm with comm_kind = KID_COMM_LOCAL. */
int32_t kid_local_comm_create(void** group, int32_t nranks);
int32_t kid_local_comm_destroy(void* group);
/* fp64 words one migrating berg occupies in the exchange buffer (reference: buffer_width F:21) */
int32_t kid_pack_width(void);


/* Unit-level entries (parity tests): one evaluation ON THE DEVICE of the mass-spreading geometry.
 *   Hexagon_into_quadrants_using_triangles I:4562-4670: out = Area_hex, Area_Q1..Q4   (hexagon_test I:261-348)
 *   point_in_triangle I:4166-4199 for triangle A,B,C and point q: v = Ax,Ay,Bx,By,Cx,Cy,qx,qy; area = Area_of_triangle */
int32_t kid_unit_hexagon_into_quadrants(int32_t device, double x0, double y0, double H, double theta, double out[5]);
int32_t kid_unit_point_in_triangle(int32_t device, const double v[8], int32_t* inside, double* area);

const char* kid_last_error(const kid_t* h);   /* h may be NULL: last init error */
const char* kid_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KID_B200_H */
