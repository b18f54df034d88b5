"""Deterministic synthetic workload of BASELINE.md / SURVEY.md 8(d): a regular 1/4-degree
lat-lon grid with analytic land mask and forcing, and seeded bergs.  Pure numpy input
generation (no files); used by tests/ and bench.py.  Nothing here is on the hot path.

All 2-D arrays are the reference's column-major (i, j) arrays = numpy shape (nj, ni).
"""
from __future__ import annotations

import numpy as np

REARTH = 6.36e6  # F:44
SEED = 20240521

# F:787-790 (southern-hemisphere table; the reference default for both hemispheres' geometry
# is taken per hemisphere by calve_icebergs, here bergs are seeded from the southern table)
INITIAL_MASS = np.array([8.8e7, 4.1e8, 3.3e9, 1.8e10, 3.8e10, 7.5e10, 1.2e11, 2.2e11, 3.9e11, 7.4e11])
MASS_SCALING = np.array([2000, 200, 50, 20, 10, 5, 2, 1, 1, 1], dtype=np.float64)
INITIAL_THICKNESS = np.array([40., 67., 133., 175., 250., 250., 250., 250., 250., 250.])
LOW_RATIO = 1.5
RHO_BERGS = 850.


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Counter-based RNG: splitmix64 finaliser of (seed + counter)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def u01(stream: int, idx: np.ndarray, seed: int = SEED) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream) * np.uint64(0xD6E8FEB86659FD93)
        r = splitmix64(idx.astype(np.uint64) * np.uint64(0x2545F4914F6CDD1D) + key)
    return (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class Grid:
    """Regular lat-lon grid tile.  Cell (i,j) (global, 1-based) has its NE corner at
    lon = i*dlon, lat = -90 + j*dlat (F:6325-6332)."""

    def __init__(self, gni=1440, gnj=720, isc=1, iec=None, jsc=1, jec=None):
        self.gni, self.gnj = gni, gnj
        self.isc, self.iec = isc, gni if iec is None else iec
        self.jsc, self.jec = jsc, gnj if jec is None else jec
        self.dlon, self.dlat = 360.0 / gni, 180.0 / gnj

    # index ranges
    def _ij(self, ring):
        i = np.arange(self.isc - ring, self.iec + ring + 1)
        j = np.arange(self.jsc - ring, self.jec + ring + 1)
        return np.meshgrid(i, j)  # (nj, ni)

    def corner_lonlat(self, ring=0):
        i, j = self._ij(ring)
        return i * self.dlon, -90.0 + j * self.dlat

    def centre_lonlat(self, ring=0):
        lon, lat = self.corner_lonlat(ring)
        return lon - 0.5 * self.dlon, lat - 0.5 * self.dlat

    def wet(self, ring=1):
        lonc, latc = self.centre_lonlat(ring)
        lonm = np.mod(lonc, 360.0)
        w = (np.abs(latc) < 80.0)
        w &= ~((lonm > 20.0) & (lonm < 60.0) & (latc > -40.0) & (latc < 40.0))
        w &= ~((lonm > 250.0) & (lonm < 300.0) & (latc > -50.0) & (latc < 60.0))
        return w.astype(np.float64)

    def init_args(self):
        """Arguments of icebergs_init (I:92-117) for this tile."""
        lon, lat = self.corner_lonlat(0)
        lonr, latr = self.corner_lonlat(1)
        _, latc = self.centre_lonlat(0)
        rad = np.pi / 180.0
        dx = REARTH * np.cos(latr * rad) * (self.dlon * rad)      # length of the cell's northern edge
        dy = np.full_like(dx, REARTH * self.dlat * rad)
        area = REARTH * np.cos(latc * rad) * (self.dlon * rad) * (REARTH * self.dlat * rad)
        return dict(ice_lon=lon, ice_lat=lat, ice_wet=self.wet(1), ice_dx=np.abs(dx), ice_dy=dy,
                    ice_area=np.abs(area), cos_rot=np.ones_like(dx), sin_rot=np.zeros_like(dx),
                    ocean_depth=np.full_like(lon, 4000.0))

    def forcing(self):
        """Arguments of icebergs_run (I:5074-5096); winds are velocities (tau_is_velocity=T)."""
        rad = np.pi / 180.0
        lam_r, phi_r = [a * rad for a in self.corner_lonlat(1)]
        lam_c, phi_c = [a * rad for a in self.corner_lonlat(0)]
        lamc_r, phic_r = [a * rad for a in self.centre_lonlat(1)]
        _, phic_deg = self.centre_lonlat(0)
        _, phir_deg = self.centre_lonlat(1)
        uo = 0.3 * np.cos(phi_r) * np.sin(2 * lam_r)
        vo = 0.2 * np.sin(lam_r) * np.cos(3 * phi_r)
        f = dict(uo=uo, vo=vo, ui=0.5 * uo, vi=0.5 * vo,
                 tauxa=10.0 * np.cos(2 * phi_c), tauya=3.0 * np.sin(3 * lam_c),
                 ssh=0.5 * np.sin(2 * lamc_r) * np.cos(3 * phic_r),
                 sst=-1.5 + 4.0 * np.cos(phic_deg * rad) ** 2,
                 sss=np.full_like(lam_c, 34.0))
        cn = np.clip((np.abs(phir_deg) - 55.0) / 20.0, 0.0, 1.0)
        f["cn"] = cn
        f["hi"] = 1.5 * cn
        f["calving"] = np.zeros_like(lam_c)
        f["calving_hflx"] = np.zeros_like(lam_c)
        return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in f.items()}

    def seed_bergs(self, n, stream=0, seed=SEED, start=0, counter0=None):
        """n bergs uniformly over this tile's wet cells: (xi,yj) in U(.05,.95)^2,
        class U{1..10}, LoW geometry (F:1540-1541), zero velocity, ids as generate_id F:4165-4177.
        start / counter0: bergs start..start+n-1 of a population seeded in chunks (counter0 = the per-cell
        id counters the earlier chunks returned)."""
        wet = self.wet(0) > 0.5
        jj, ii = np.nonzero(wet)
        ncell = len(ii)
        k = np.arange(start, start + n, dtype=np.uint64)
        base = np.uint64(stream) << np.uint64(40)
        c = np.minimum((u01(1, k + base, seed) * ncell).astype(np.int64), ncell - 1)
        i = (ii[c] + self.isc).astype(np.int32)
        j = (jj[c] + self.jsc).astype(np.int32)
        xi = 0.05 + 0.9 * u01(2, k + base, seed)
        yj = 0.05 + 0.9 * u01(3, k + base, seed)
        cls = np.minimum((u01(4, k + base, seed) * 10).astype(np.int64), 9)
        lon = (i - 1 + xi) * self.dlon
        lat = -90.0 + (j - 1 + yj) * self.dlat
        mass = INITIAL_MASS[cls]
        thick = INITIAL_THICKNESS[cls]
        width = np.sqrt(mass / (LOW_RATIO * RHO_BERGS * thick))
        length = LOW_RATIO * width
        # per-cell counter in generation order -> id = counter*2^32 + (i + gni*(j-1))
        cell = (j.astype(np.int64) - 1) * self.gni + (i.astype(np.int64) - 1)
        order = np.argsort(cell, kind="stable")
        sc = cell[order]
        first = np.r_[0, np.nonzero(np.diff(sc))[0] + 1]
        first_of = np.repeat(first, np.diff(np.r_[first, n]))
        counter = np.empty(n, dtype=np.int64)
        counter[order] = np.arange(n) - first_of + 1
        if counter0 is not None:
            counter += counter0.reshape(-1)[cell]
        ident = counter * (1 << 32) + (i.astype(np.int64) + self.gni * (j.astype(np.int64) - 1))
        z = np.zeros(n)
        return dict(lon=lon, lat=lat, uvel=z.copy(), vvel=z.copy(), mass=mass.copy(), thickness=thick.copy(),
                    width=width, length=length, axn=z.copy(), ayn=z.copy(), bxn=z.copy(), byn=z.copy(),
                    start_lon=lon.copy(), start_lat=lat.copy(), start_day=(cls + 1) / 17.0,
                    start_mass=mass.copy(), mass_scaling=MASS_SCALING[cls].copy(), mass_of_bits=z.copy(),
                    heat_density=z.copy(), start_year=np.ones(n, dtype=np.int32), ine=i, jne=j,
                    id=ident.astype(np.int64)), counter_grid(self, cell, n, counter0)


def counter_grid(grid: Grid, cell, n, counter0=None):
    """iceberg_counter_grd consistent with the seeded ids (global cell -> count)."""
    cnt = np.bincount(cell, minlength=grid.gni * grid.gnj).astype(np.int32).reshape(grid.gnj, grid.gni)
    return cnt if counter0 is None else cnt + counter0


def workload_params(default_params, **over):
    """Physics of the synthetic case: namelist defaults except Verlet stepping, bergy bits
    on, wind passed as velocity (BASELINE.md 'Physics')."""
    kw = dict(runge_not_verlet=0, bergy_bit_erosion_fraction=0.1, tau_is_velocity=1, old_bug_bilin=1,
              Rearth=REARTH,
              # the metric is dyn+thermo: spreading the berg mass onto the ocean grid (8f1) is off
              add_weight_to_ocean=0)
    kw.update(over)
    return default_params(**kw)


class CartesianGrid:
    """The stand-alone driver's test grid (driver/icebergs_driver.F90:274-286): gridres-metre
    square cells, lon(i,j) = gridres*i, lat(i,j) = gridres*j, all ocean, 1000 m deep."""

    def __init__(self, ni=20, nj=20, gridres=1.0e3, isc=1, iec=None, jsc=1, jec=None):
        self.gni, self.gnj, self.res = ni, nj, gridres
        self.isc, self.iec = isc, ni if iec is None else iec
        self.jsc, self.jec = jsc, nj if jec is None else jec

    def _ij(self, ring):
        i = np.arange(self.isc - ring, self.iec + ring + 1)
        j = np.arange(self.jsc - ring, self.jec + ring + 1)
        return np.meshgrid(i, j)

    def init_args(self):
        i0, j0 = self._ij(0)
        i1, _ = self._ij(1)
        one0, one1 = np.ones(i0.shape), np.ones(i1.shape)
        return dict(ice_lon=self.res * i0, ice_lat=self.res * j0, ice_wet=one1.copy(), ice_dx=self.res * one1,
                    ice_dy=self.res * one1, ice_area=self.res * self.res * one0, cos_rot=one1.copy(),
                    sin_rot=0.0 * one1, ocean_depth=1000.0 * one0)

    def forcing(self, ibuo=0.2, ibvo=0.2, sst=-2.0, collision_test=True):
        """D:218-229; collision_test: the meridional current converges on lat = mid (D:313-327)."""
        i0, j0 = self._ij(0)
        i1, j1 = self._ij(1)
        lon1, lat1 = self.res * i1, self.res * j1
        uo = np.full(i1.shape, ibuo)
        vo = np.full(i1.shape, ibvo)
        if collision_test:
            mid = 10.0e3
            vo = np.where((lon1 > mid) | (lon1 <= 0.0) | (lat1 == mid), 0.0, np.where(lat1 > mid, -ibvo, ibvo))
        z0, z1 = np.zeros(i0.shape), np.zeros(i1.shape)
        f = dict(uo=uo, vo=vo, ui=z1.copy(), vi=z1.copy(), tauxa=z0.copy(), tauya=z0.copy(), ssh=z1.copy(),
                 sst=np.full(i0.shape, sst), sss=np.full(i0.shape, 34.0), cn=z1.copy(), hi=z1.copy(),
                 calving=z0.copy(), calving_hflx=z0.copy())
        return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in f.items()}


def collision_bergs(gridres=1.0e3, n=20, h_ice=300.0, rho_ice=918.0):
    """The two 8-element conglomerates of tests/collision_tests (README: 16 bergs): a 1 km-radius
    disc of 300 m thick ice centred on (4.5 km, 4.5 km), hex-packed with elements of radius
    R = sqrt(3)/2 * 0.45 * gridres (initialize_bergs_in_pattern.py:826-836, :1795-1801), mirrored about
    y = n*gridres/2 (:905-914).  Element area 2*sqrt(3)*R^2 (:663-671), width = length = sqrt(area),
    mass = thickness * rho_ice * area with the script's rho_ice = 918 (:196, :638)."""
    R = np.sqrt(3.0) / 2.0 * 0.45 * gridres
    xc = (np.arange(n) + 0.5) * gridres
    X, Y = np.meshgrid(xc, xc)
    ice = ((X - 4500.0) ** 2 + (Y - 4500.0) ** 2) <= 1000.0 ** 2
    pts = []
    for j in range(3 * n):
        for i in range(3 * n):
            x = (2.0 / np.sqrt(3.0)) * R + np.sqrt(3.0) * R * i
            y = R + (i % 2) * R + 2.0 * R * j
            if x >= n * gridres or y >= n * gridres:
                continue
            if ice[int(y // gridres), int(x // gridres)]:
                pts.append((x, y))
    pts = np.array(pts)
    x = np.r_[pts[:, 0], pts[:, 0]]
    y = np.r_[pts[:, 1], n * gridres - pts[:, 1]]
    m = len(x)
    area = 2.0 * np.sqrt(3.0) * R * R
    w = np.full(m, np.sqrt(area))
    z = np.zeros(m)
    return dict(lon=x, lat=y, uvel=z.copy(), vvel=z.copy(), mass=np.full(m, h_ice * rho_ice * area),
                thickness=np.full(m, h_ice), width=w.copy(), length=w.copy(), axn=z.copy(), ayn=z.copy(),
                bxn=z.copy(), byn=z.copy(), start_lon=x.copy(), start_lat=y.copy(), start_day=z.copy(),
                start_mass=np.full(m, h_ice * rho_ice * area), mass_scaling=np.ones(m), mass_of_bits=z.copy(),
                heat_density=z.copy(), start_year=np.zeros(m, dtype=np.int32))


def collision_params(default_params, **over):
    """&icebergs_nml of tests/collision_tests/input_KID.nml (the values that reach the hot path)."""
    kw = dict(halo=3, Lx=20000.0, grid_is_latlon=0, grid_is_regular=1, hexagonal_icebergs=1, rho_bergs=850.0,
              spring_coef=1.0e-5, radial_damping_coef=1.0e-4, tangental_damping_coef=2.0e-5,
              critical_interaction_damping_on=1, LoW_ratio=1.5, bergy_bit_erosion_fraction=0.0, sicn_shift=0.0,
              use_operator_splitting=1, speed_limit=0.0, tip_parameter=0.0, coastal_drift=0.4, tidal_drift=0.0,
              runge_not_verlet=0, allow_bergs_to_roll=1, use_updated_rolling_scheme=1, melt_cutoff=10.0,
              apply_thickness_cutoff_to_gridded_melt=1, apply_thickness_cutoff_to_bergs_melt=1,
              set_melt_rates_to_zero=1, iceberg_bonds_on=1, interactive_icebergs_on=1, only_interactive_forces=0,
              use_new_predictive_corrective=1, max_bonds=6, manually_initialize_bonds=1,
              length_for_manually_initialize_bonds=800.0, use_roundoff_fix=1, old_bug_bilin=0, tau_is_velocity=0,
              add_weight_to_ocean=1, use_old_spreading=0, rotate_icebergs_for_mass_spreading=1, pass_fields_to_ocean_model=0)
    kw.update(over)
    return default_params(**kw)


def footloose_bergs(grdres=5000.0):
    """tests/footloose_tests/makeberg/makeberg.py:245-274: two elements side by side at y = 10000.1."""
    radius = (np.sqrt(3.0) / 2.0) * (0.45 * grdres)
    area = (3.0 * np.sqrt(3.0) / 2.0) * ((4.0 / 3.0) * radius ** 2)
    w = np.sqrt(area)
    x = np.array([10000.1 - radius, 10000.1 + radius])
    y = np.array([10000.1, 10000.1])
    z = np.zeros(2)
    m = np.full(2, 300.0 * 850.0 * area)
    return dict(lon=x, lat=y, uvel=z.copy(), vvel=z.copy(), mass=m, thickness=np.full(2, 300.0), width=np.full(2, w),
                length=np.full(2, w), axn=z.copy(), ayn=z.copy(), bxn=z.copy(), byn=z.copy(), start_lon=x.copy(),
                start_lat=y.copy(), start_day=z.copy(), start_mass=m.copy(), mass_scaling=np.ones(2), mass_of_bits=z.copy(),
                heat_density=z.copy(), start_year=np.zeros(2, dtype=np.int32))


def footloose_params(default_params, **over):
    """&icebergs_nml of tests/footloose_tests/input.nml (values that reach the hot path); the child
    displacement needs the FMS random stream and is off here."""
    kw = dict(halo=3, Lx=20000.0, grid_is_latlon=0, grid_is_regular=1, rho_bergs=850.0, spring_coef=1.0e-5,
              bergy_bit_erosion_fraction=1.0, use_operator_splitting=1, coastal_drift=0.0, runge_not_verlet=0,
              melt_cutoff=10.0, apply_thickness_cutoff_to_gridded_melt=1, apply_thickness_cutoff_to_bergs_melt=1,
              allow_bergs_to_roll=1, use_updated_rolling_scheme=1, tip_parameter=0.0, set_melt_rates_to_zero=0,
              iceberg_bonds_on=0, interactive_icebergs_on=0, use_new_predictive_corrective=1, passive_mode=1,
              old_bug_bilin=0, use_roundoff_fix=1, footloose=1, displace_fl_bergs=0, fl_style_fl_bits=1,
              fl_youngs=1.0e8, fl_strength=250.0, new_berg_from_fl_bits_mass_thres=3.0e11, tau_is_velocity=0,
              add_weight_to_ocean=0)
    kw.update(over)
    return default_params(**kw)


def footloose_forcing(grid, ibuo=1.0, ibvo=0.1, ibua=-1.0, sst=-0.5):
    """driver forcing with fl_test=.true.: vo changes sign east of x = 10 km (D:309-311)."""
    f = grid.forcing(ibuo=ibuo, ibvo=ibvo, sst=sst, collision_test=False)
    i1, _ = grid._ij(1)
    f["vo"] = np.ascontiguousarray(np.where(grid.res * i1 > 10000.0, -ibvo, ibvo), dtype=np.float64)
    f["tauxa"] = np.full_like(f["tauxa"], ibua)
    return f


def beam_bergs(nbergs=29, r=0.25, xs=101.0e3, ys=151.0e3, h=1.0, rho_ice=800.0):
    """tests/dem_ssbeam_test/makeberg/makeberg.py:253-314 (and dem_cbeam_test): a row of square-packed elements of
    radius r, 2r apart, starting at (xs + 0, ys + 2r)."""
    area = (2.0 * r) ** 2
    x = xs + 2.0 * r * np.arange(nbergs)
    y = np.full(nbergs, ys + 2.0 * r)
    z = np.zeros(nbergs)
    w = np.full(nbergs, np.sqrt(area))
    m = np.full(nbergs, h * rho_ice * area)
    return dict(lon=x, lat=y, uvel=z.copy(), vvel=z.copy(), mass=m, thickness=np.full(nbergs, h), width=w.copy(), length=w.copy(),
                axn=z.copy(), ayn=z.copy(), bxn=z.copy(), byn=z.copy(), start_lon=x.copy(), start_lat=y.copy(), start_day=z.copy(),
                start_mass=m.copy(), mass_scaling=np.ones(nbergs), mass_of_bits=z.copy(), heat_density=z.copy(),
                start_year=np.zeros(nbergs, dtype=np.int32))


def beam_params(default_params, **over):
    """&icebergs_nml of tests/dem_ssbeam_test/input.nml (the values that reach the hot path; mass spreading off)."""
    kw = dict(halo=3, Lx=300.0e3, grid_is_latlon=0, grid_is_regular=1, hexagonal_icebergs=0, rho_bergs=800.0,
              dem_beam_test=1, dem=1, poisson=0.3, dem_damping_coef=0.1, dem_spring_coef=1.0e9, mts=1, mts_sub_steps=100000,
              force_convergence=1, convergence_tolerance=1e-8, contact_distance=2000.0, contact_spring_coef=1.0e-8,
              cdrag_grounding=3.16e6, h_to_init_grounding=200.0, spring_coef=1.0e-5, radial_damping_coef=0.0,
              tangental_damping_coef=0.0, scale_damping_by_pmag=0, critical_interaction_damping_on=0, tang_crit_int_damp_on=0,
              LoW_ratio=1.5, bergy_bit_erosion_fraction=0.0, sicn_shift=0.0, use_operator_splitting=1, speed_limit=0.0,
              tip_parameter=0.0, coastal_drift=0.0, tidal_drift=0.0, runge_not_verlet=0, allow_bergs_to_roll=0,
              use_updated_rolling_scheme=0, melt_cutoff=10.0, apply_thickness_cutoff_to_gridded_melt=1,
              apply_thickness_cutoff_to_bergs_melt=1, set_melt_rates_to_zero=1, iceberg_bonds_on=1, interactive_icebergs_on=1,
              only_interactive_forces=1, use_new_predictive_corrective=1, max_bonds=4, internal_bergs_for_drag=1,
              manually_initialize_bonds=1, manually_initialize_bonds_from_radii=1, use_roundoff_fix=1, old_bug_bilin=0,
              tau_is_velocity=0, add_weight_to_ocean=0)
    kw.update(over)
    return default_params(**kw)


def cantilever_bergs(rows=3, per_row=30, r=2500.0, xs=101.0e3, ys=151.0e3, h=1.0, rho_ice=900.0):
    """tests/dem_cbeam_test/makeberg/makeberg.py:253-314: rows x per_row square-packed elements, the first of each
    row static (the clamped end)."""
    area = (2.0 * r) ** 2
    x = np.tile(xs + 2.0 * r * np.arange(per_row), rows)
    y = np.repeat(ys + 2.0 * r * np.arange(rows), per_row)
    n = rows * per_row
    cols = beam_bergs(n, r, xs, ys, h, rho_ice)
    cols.update(lon=x, lat=y, start_lon=x.copy(), start_lat=y.copy())
    cols["static_berg"] = np.where(np.arange(n) % per_row == 0, 1.0, 0.0)
    return cols


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: "tests/a68_test scaled" -- a bonded tabular berg.  The A68a forcing data of the reference
# test is FTP-only (tests/a68_test/get_data.sh:4), so the berg is synthetic: a rectangle of
# square-packed elements of radius r (1.5 km in the reference case, 2r apart, bonds from the radii), on a Cartesian
# grid, with the physics of tests/a68_test/long_run.nml and analytic forcing: a sheared current that turns and
# bends the berg, and a shoal (ocean_depth below the draught) that one corner runs onto, so grounding stress
# builds up and bonds break (fracture_criterion='stress').
def tabular_berg(nx=12, ny=16, r=1500.0, x0=60.0e3, y0=60.0e3, h=200.0, rho_ice=850.0):
    area = (2.0 * r) ** 2
    x = np.tile(x0 + 2.0 * r * np.arange(nx), ny)
    y = np.repeat(y0 + 2.0 * r * np.arange(ny), nx)
    n = nx * ny
    z = np.zeros(n)
    w = np.full(n, np.sqrt(area))
    m = np.full(n, h * rho_ice * area)
    return dict(lon=x, lat=y, uvel=z.copy(), vvel=z.copy(), mass=m, thickness=np.full(n, h), width=w.copy(), length=w.copy(),
                axn=z.copy(), ayn=z.copy(), bxn=z.copy(), byn=z.copy(), start_lon=x.copy(), start_lat=y.copy(), start_day=z.copy(),
                start_mass=m.copy(), mass_scaling=np.ones(n), mass_of_bits=z.copy(), heat_density=z.copy(),
                start_year=np.zeros(n, dtype=np.int32))


def a68_params(default_params, **over):
    """&icebergs_nml of tests/a68_test/long_run.nml (the values that reach the hot path), the template's <..> entries
    filled in: mts_sub_steps=60, ocean_drag_scale=1, cdrag_grounding=1e3 (one of the values the template comments list), frac_thres_n=60e3; a Cartesian grid."""
    kw = dict(halo=3, Lx=300.0e3, grid_is_latlon=0, grid_is_regular=1, hexagonal_icebergs=0, rho_bergs=850.0,
              short_step_mts_grounding=1, tau_is_velocity=1, ocean_drag_scale=1.0, skip_first_outer_mts_step=1,
              constant_interaction_LW=1, break_bonds_on_sub_steps=1, save_bond_forces=1,
              use_broken_bonds_for_substep_contact=1, dem=1, poisson=0.3, dem_damping_coef=1.0, explicit_inner_mts=1,
              dem_spring_coef=5.0e6, spring_coef=0.00065359477124183, mts=1, mts_sub_steps=60, force_convergence=1,
              convergence_tolerance=1e-4, contact_distance=4.0e3, contact_spring_coef=1.0e-7, cdrag_grounding=1.0e3,
              h_to_init_grounding=0.0, fracture_criterion_stress=1, frac_thres_n=60.0e3, frac_thres_t=100.0e3,
              radial_damping_coef=0.0, tangental_damping_coef=0.0, scale_damping_by_pmag=0,
              critical_interaction_damping_on=1, tang_crit_int_damp_on=0, LoW_ratio=1.5, bergy_bit_erosion_fraction=0.0,
              sicn_shift=0.0, use_operator_splitting=1, speed_limit=0.0, tip_parameter=0.0, coastal_drift=0.0,
              tidal_drift=0.0, runge_not_verlet=0, use_mixed_melting=1, apply_thickness_cutoff_to_gridded_melt=1,
              apply_thickness_cutoff_to_bergs_melt=1, melt_cutoff=10.0, allow_bergs_to_roll=1,
              use_updated_rolling_scheme=1, use_three_equation_model=0, const_gamma=0, ustar_icebergs_bg=0.0,
              set_melt_rates_to_zero=1, iceberg_bonds_on=1, interactive_icebergs_on=1, only_interactive_forces=0,
              use_new_predictive_corrective=1, use_old_spreading=0, max_bonds=4, internal_bergs_for_drag=1,
              manually_initialize_bonds=1, manually_initialize_bonds_from_radii=1, use_roundoff_fix=1, old_bug_bilin=0,
              add_weight_to_ocean=0, remove_unused_bergs=1)
    kw.update(over)
    return default_params(**kw)


class TabularGrid(CartesianGrid):
    """60 x 60 cells of 5 km (cyclic in x, Lx = 300 km), 1000 m deep except a shoal north-east of the berg's start."""

    def __init__(self, ni=60, nj=60, gridres=5.0e3, isc=1, iec=None, jsc=1, jec=None):
        super().__init__(ni, nj, gridres, isc, iec, jsc, jec)

    def init_args(self):
        a = super().init_args()
        i0, j0 = self._ij(0)
        x, y = self.res * (i0 - 0.5), self.res * (j0 - 0.5)
        shoal = ((x - 112.5e3) ** 2 + (y - 97.5e3) ** 2) < (12.0e3) ** 2
        a["ocean_depth"] = np.where(shoal, 120.0, 1000.0)       # draught of the 200 m berg: 166 m
        return a

    def forcing(self, u0=0.35, shear=0.25, v0=0.12, ua=6.0):
        """currents (m/s, B-grid corners): eastward, stronger to the north (the berg turns), a weak northward drift;
        winds passed as velocities (tau_is_velocity=T)"""
        i0, j0 = self._ij(0)
        i1, j1 = self._ij(1)
        y1 = self.res * j1 / (self.res * self.gnj)
        z0, z1 = np.zeros(i0.shape), np.zeros(i1.shape)
        f = dict(uo=u0 + shear * (y1 - 0.5), vo=np.full(i1.shape, v0), ui=z1.copy(), vi=z1.copy(),
                 tauxa=np.full(i0.shape, ua), tauya=z0.copy(), ssh=z1.copy(), sst=np.full(i0.shape, -1.0),
                 sss=np.full(i0.shape, 34.0), cn=z1.copy(), hi=z1.copy(), calving=z0.copy(), calving_hflx=z0.copy())
        return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in f.items()}


# ---------------------------------------------------------------------------------------------------------------
# SURVEY 8 f4, tripolar fold: an analytic bipolar cap.  In the polar stereographic plane (X, Y) = 2 tan(colat/2) *
# (cos lon, sin lon) the corners lie on confocal ellipses X = a cosh(mu) cos(nu), Y = a sinh(mu) sin(nu) with
# nu = 2 pi i / gni (cyclic in i) and mu = mu_max (gnj - j) / gnj, so row gnj (mu = 0) is the segment between the two
# grid poles (+-a, 0) walked twice: corner (i, gnj) IS corner (gni - i, gnj) -- FOLD_NORTH_EDGE of a tripolar grid (Murray
# 1996), and cell (i, gnj+k) of the analytic continuation mu < 0 is cell (gni+1-i, gnj+1-k) turned by 180 degrees.
class BipolarCapGrid:
    def __init__(self, gni=90, gnj=24, isc=1, iec=None, jsc=1, jec=None, colat_pole_deg=25.0, mu_max=0.6):
        assert gni % 2 == 0
        self.gni, self.gnj = gni, gnj
        self.isc, self.iec = isc, gni if iec is None else iec
        self.jsc, self.jec = jsc, gnj if jec is None else jec
        self.a = 2.0 * np.tan(0.5 * np.radians(colat_pole_deg))
        self.mu_max = mu_max

    def _ij(self, ring):
        i = np.arange(self.isc - ring, self.iec + ring + 1)
        j = np.arange(self.jsc - ring, self.jec + ring + 1)
        return np.meshgrid(i, j)

    def _munu(self, i, j):
        return self.mu_max * (self.gnj - np.asarray(j, dtype=np.float64)) / self.gnj, 2.0 * np.pi * np.asarray(i, dtype=np.float64) / self.gni

    def plane(self, i, j):
        mu, nu = self._munu(i, j)
        return self.a * np.cosh(mu) * np.cos(nu), self.a * np.sinh(mu) * np.sin(nu)

    @staticmethod
    def plane_to_lonlat(X, Y):
        r = np.hypot(X, Y)
        return np.degrees(np.arctan2(Y, X)), 90.0 - np.degrees(2.0 * np.arctan(0.5 * r))

    def lonlat(self, i, j):
        """lon in the branch that is continuous along a row: [0, 360] from i = 0 to gni (and beyond, periodically)"""
        i = np.asarray(i, dtype=np.float64)
        X, Y = self.plane(i, j)
        lon, lat = self.plane_to_lonlat(X, Y)
        mu, _ = self._munu(i, j)
        # geometric longitude of the point, then the branch: rows south of the fold run 0..360 with i, the continuation
        # beyond the fold is the mirror image (360 - that)
        base = np.mod(i, self.gni) / self.gni * 360.0
        base = np.where(mu < 0, 360.0 - base, base)
        lon = lon + 360.0 * np.round((base - lon) / 360.0)
        return lon + 360.0 * np.floor(i / self.gni) * np.where(mu < 0, -1.0, 1.0), lat

    def corner_lonlat(self, ring=0):
        return self.lonlat(*self._ij(ring))

    @staticmethod
    def _sphere(X, Y):
        r2 = X * X + Y * Y
        return np.stack([4.0 * X / (4.0 + r2), 4.0 * Y / (4.0 + r2), (4.0 - r2) / (4.0 + r2)])

    def _xyz(self, i, j):
        return self._sphere(*self.plane(i, j))

    @staticmethod
    def _dist(p, q):
        return REARTH * np.arccos(np.clip((p * q).sum(axis=0), -1.0, 1.0))

    def grid_angle(self, i, j):
        """(cos, sin) of the angle from local east to the grid's +i direction, counter-clockwise"""
        mu, nu = self._munu(i, j)
        tx, ty = -np.cosh(mu) * np.sin(nu), np.sinh(mu) * np.cos(nu)
        n = np.maximum(np.hypot(tx, ty), 1e-300)
        tx, ty = tx / n, ty / n
        X, Y = self.plane(i, j)
        lam = np.arctan2(Y, X)
        return tx * (-np.sin(lam)) + ty * np.cos(lam), tx * (-np.cos(lam)) + ty * (-np.sin(lam))

    def wet(self, ring=1):
        i, j = self._ij(ring)
        iw = np.mod(i - 1, self.gni) + 1
        jj = np.where(j > self.gnj, 2 * self.gnj + 1 - j, j)
        iw = np.where(j > self.gnj, self.gni + 1 - iw, iw)
        # land around the two grid poles (the degenerate cells at nu = 0, pi) and along the southern edge
        pole = (np.minimum(iw, self.gni + 1 - iw) <= 3) | (np.abs(iw - (self.gni // 2 + 0.5)) <= 3)
        _, latc = self.lonlat(iw - 0.5, jj - 0.5)
        return (~(pole | (jj <= 2) | (latc > 87.0))).astype(np.float64)      # ... and around the geographic pole

    def init_args(self):
        i0, j0 = self._ij(0)
        i1, j1 = self._ij(1)
        lon, lat = self.lonlat(i0, j0)
        dx = self._dist(self._xyz(i1 - 1, j1), self._xyz(i1, j1))           # northern edge of cell (i,j)
        dy = self._dist(self._xyz(i1, j1 - 1), self._xyz(i1, j1))           # eastern edge
        d1 = self._xyz(i0, j0) - self._xyz(i0 - 1, j0 - 1)
        d2 = self._xyz(i0 - 1, j0) - self._xyz(i0, j0 - 1)
        area = 0.5 * REARTH ** 2 * np.sqrt((np.cross(d1, d2, axis=0) ** 2).sum(axis=0))
        ca, sa = self.grid_angle(i1 - 0.5, j1 - 0.5)
        return dict(ice_lon=lon, ice_lat=lat, ice_wet=self.wet(1), ice_dx=dx, ice_dy=dy, ice_area=area,
                    cos_rot=ca, sin_rot=-sa,                         # rotate I:4953: u_geo = cos*u + sin*v
                    ocean_depth=np.full(lon.shape, 4000.0))

    def forcing(self, speed=0.9, direction_deg=80.0):
        """A uniform stream in the stereographic plane (it crosses the fold line Y = 0), given as grid-oriented B-grid
        components; the ice and the wind follow it, scalars are smooth functions of the plane coordinates."""
        i0, j0 = self._ij(0)
        i1, j1 = self._ij(1)
        vx, vy = speed * np.cos(np.radians(direction_deg)), speed * np.sin(np.radians(direction_deg))

        def grid_components(i, j):
            mu, nu = self._munu(i, j)
            tx, ty = -np.cosh(mu) * np.sin(nu), np.sinh(mu) * np.cos(nu)
            n = np.maximum(np.hypot(tx, ty), 1e-300)
            tx, ty = tx / n, ty / n
            return vx * tx + vy * ty, -vx * ty + vy * tx          # along +i, along +j (= t rotated by +90 degrees)

        uo, vo = grid_components(i1, j1)
        ua, va = grid_components(i0, j0)
        Xc, Yc = self.plane(i1 - 0.5, j1 - 0.5)
        X0, Y0 = self.plane(i0 - 0.5, j0 - 0.5)
        f = dict(uo=uo, vo=vo, ui=0.5 * uo, vi=0.5 * vo, tauxa=8.0 * ua, tauya=8.0 * va,
                 ssh=0.3 * np.sin(3.0 * Xc) * np.cos(2.0 * Yc), sst=-1.0 + 2.5 * np.cos(2.0 * X0) ** 2,
                 sss=np.full(i0.shape, 34.0), cn=np.clip(0.6 + 0.5 * np.sin(4.0 * Yc), 0.0, 1.0),
                 hi=1.0 + 0.5 * np.cos(3.0 * Xc), calving=np.zeros(i0.shape), calving_hflx=np.zeros(i0.shape))
        return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in f.items()}

    def seed_bergs(self, n, stream=0, seed=SEED, rows=4):
        """n bergs in the wet cells of the `rows` rows below the fold that this tile owns"""
        i0, j0 = self._ij(0)
        ok = (self.wet(0) > 0.5) & (j0 > self.gnj - rows)
        # keep clear of the geographic pole (it sits in the middle of the fold line)
        lonc, latc = self.lonlat(i0 - 0.5, j0 - 0.5)
        ok &= latc < 86.0
        jj, ii = np.nonzero(ok)
        ncell = len(ii)
        k = np.arange(n, dtype=np.uint64)
        base = np.uint64(stream) << np.uint64(40)
        c = np.minimum((u01(1, k + base, seed) * ncell).astype(np.int64), ncell - 1)
        i = (ii[c] + self.isc).astype(np.int32)
        j = (jj[c] + self.jsc).astype(np.int32)
        xi = 0.05 + 0.9 * u01(2, k + base, seed)
        yj = 0.05 + 0.9 * u01(3, k + base, seed)
        cls = np.minimum((u01(4, k + base, seed) * 10).astype(np.int64), 9)
        cr = {q: self.lonlat(i + di, j + dj) for q, (di, dj) in dict(ne=(0, 0), nw=(-1, 0), se=(0, -1), sw=(-1, -1)).items()}
        pos = []
        for q in (0, 1):       # bilin F:7071 (the non-bug formula)
            pos.append((cr["ne"][q] * xi + cr["nw"][q] * (1 - xi)) * yj + (cr["se"][q] * xi + cr["sw"][q] * (1 - xi)) * (1 - yj))
        lon, lat = pos
        mass = INITIAL_MASS[cls]
        thick = INITIAL_THICKNESS[cls]
        width = np.sqrt(mass / (LOW_RATIO * RHO_BERGS * thick))
        length = LOW_RATIO * width
        cell = (j.astype(np.int64) - 1) * self.gni + (i.astype(np.int64) - 1)
        order = np.argsort(cell, kind="stable")
        sc = cell[order]
        first = np.r_[0, np.nonzero(np.diff(sc))[0] + 1]
        first_of = np.repeat(first, np.diff(np.r_[first, n]))
        counter = np.empty(n, dtype=np.int64)
        counter[order] = np.arange(n) - first_of + 1
        ident = counter * (1 << 32) + (i.astype(np.int64) + self.gni * (j.astype(np.int64) - 1))
        z = np.zeros(n)
        return dict(lon=lon, lat=lat, uvel=z.copy(), vvel=z.copy(), mass=mass.copy(), thickness=thick.copy(),
                    width=width, length=length, axn=z.copy(), ayn=z.copy(), bxn=z.copy(), byn=z.copy(),
                    start_lon=lon.copy(), start_lat=lat.copy(), start_day=(cls + 1) / 17.0,
                    start_mass=mass.copy(), mass_scaling=MASS_SCALING[cls].copy(), mass_of_bits=z.copy(),
                    heat_density=z.copy(), start_year=np.ones(n, dtype=np.int32), ine=i, jne=j,
                    id=ident.astype(np.int64)), counter_grid(self, cell, n)
