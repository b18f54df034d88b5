"""Restart files of the reference as the on-disk interface of the hot path.

``icebergs.res.nc`` (variables of src/icebergs_fmsio.F90:261-337, read back at fmsio:742-790),
``bonds_iceberg.res.nc`` (fmsio:473-493) and ``calving.res.nc`` (fmsio:564-569) are NetCDF-3 classic
files with one unlimited dimension ``i`` -- exactly what the makeberg scripts of the reference tests
write (tests/*/makeberg/*.py).  They are read and written here with ``scipy.io.netcdf_file`` (no FMS,
no netCDF4) and turned into the column dictionaries ``Icebergs.set_bergs / get_bergs / set_bonds /
get_bonds / set_calving_state`` take.  Host-side file handling only: nothing here computes.
"""
from __future__ import annotations

import numpy as np
from scipy.io import netcdf_file

# (name, dtype, long_name, units) in the order of fmsio:261-337
BERG_VARS = [
    ("lon", "d", "longitude", "degrees_E"), ("lat", "d", "latitude", "degrees_N"),
    ("uvel", "d", "zonal velocity", "m/s"), ("vvel", "d", "meridional velocity", "m/s"), ("mass", "d", "mass", "kg"),
    ("axn", "d", "explicit zonal acceleration", "m/s^2"), ("ayn", "d", "explicit meridional acceleration", "m/s^2"),
    ("bxn", "d", "implicit zonal acceleration", "m/s^2"), ("byn", "d", "implicit meridional acceleration", "m/s^2"),
    ("ine", "i", "i index", "none"), ("jne", "i", "j index", "none"),
    ("thickness", "d", "thickness", "m"), ("width", "d", "width", "m"), ("length", "d", "length", "m"),
    ("start_lon", "d", "longitude of calving location", "degrees_E"),
    ("start_lat", "d", "latitude of calving location", "degrees_N"),
    ("start_year", "i", "calendar year of calving event", "years"),
    ("id_cnt", "i", "counter component of iceberg id", "dimensionless"),
    ("id_ij", "i", "position component of iceberg id", "dimensionless"),
    ("start_day", "d", "year day of calving event", "days"), ("start_mass", "d", "initial mass of calving berg", "kg"),
    ("mass_scaling", "d", "scaling factor for mass of calving berg", "none"),
    ("mass_of_bits", "d", "mass of bergy bits", "kg"), ("heat_density", "d", "heat density", "J/kg"),
    ("fl_k", "d", "footloose calving k", "m"), ("mass_of_fl_bits", "d", "mass of footloose bits", "kg"),
    ("mass_of_fl_bergy_bits", "d", "mass of bergy bits associated with footloose bits", "kg"),
    ("static_berg", "d", "static_berg", "dimensionless"),
]
# written only by runs that carry them (mts: fmsio:300-306, dem: fmsio:308-314) and read back when present
MTS_VARS = [("axn_fast", "d", "explicit zonal acceleration for fast dynamics", "m/s^2"),
            ("ayn_fast", "d", "explicit meridional acceleration for fast dynamics", "m/s^2"),
            ("bxn_fast", "d", "implicit zonal acceleration for fast dynamics", "m/s^2"),
            ("byn_fast", "d", "implicit meridional acceleration for fast dynamics", "m/s^2")]
DEM_VARS = [("ang_vel", "d", "angular velocity", "rad/s"), ("ang_accel", "d", "angular acceleration", "rad/s^2"),
            ("rot", "d", "accumulated rotation", "rad")]
BOND_DEM_VARS = [("tangd1", "d", "tangential displacement, x", "m"), ("tangd2", "d", "tangential displacement, y", "m"),
                 ("nstress", "d", "normal stress", "Pa"), ("sstress", "d", "shear stress", "Pa"),
                 ("rel_rotation", "d", "relative rotation", "rad"), ("broken", "i", "bond is broken", "none")]
OPTIONAL_ZERO = ("axn", "ayn", "bxn", "byn", "fl_k", "mass_of_bits", "mass_of_fl_bits", "mass_of_fl_bergy_bits",
                 "heat_density", "static_berg")          # read_real_vector(..., value_if_not_in_file=0.), fmsio:746-781


def split_id(ident):
    """split_id F:7285-7296: id = id_cnt*2^32 + id_ij."""
    ident = np.asarray(ident, dtype=np.int64)
    return (ident >> 32).astype(np.int32), (ident & 0xFFFFFFFF).astype(np.int32)


def id_from_2_ints(cnt, ij):
    """id_from_2_ints F:7276-7282."""
    return np.asarray(cnt, dtype=np.int64) * (1 << 32) + np.asarray(ij, dtype=np.int64)


def _write(path, nrec, variables):
    f = netcdf_file(path, "w", version=1)
    f.createDimension("i", None)
    for name, typ, long_name, units, data in variables:
        v = f.createVariable(name, typ, ("i",))
        v.long_name, v.units = long_name, units
        if nrec:
            v[:nrec] = data
    f.close()


def write_restart_bergs(path, cols):
    """write_restart_bergs fmsio:170-430 from the columns of ``Icebergs.get_bergs()``."""
    n = len(cols["lon"])
    cnt, ij = split_id(cols["id"]) if "id" in cols else (np.zeros(n, np.int32), np.zeros(n, np.int32))
    out = []
    for name, typ, long_name, units in BERG_VARS:
        if name == "id_cnt":
            data = cnt
        elif name == "id_ij":
            data = ij
        elif name in cols:
            data = cols[name]
        elif name in OPTIONAL_ZERO:
            data = np.zeros(n)
        else:
            raise KeyError(f"write_restart_bergs: column {name!r} is missing")
        out.append((name, typ, long_name, units, np.asarray(data, dtype=np.float64 if typ == "d" else np.int32)))
    for name, typ, long_name, units in MTS_VARS + DEM_VARS:
        if name in cols:
            out.append((name, typ, long_name, units, np.asarray(cols[name], dtype=np.float64)))
    _write(path, n, out)


def read_restart_bergs(path, ignore_ij_restart=False):
    """read_restart_bergs fmsio:606-975 -> columns for ``Icebergs.set_bergs``.  Files with the legacy
    ``iceberg_num`` (makeberg scripts) carry no ids: the library then generates them in file order
    (generate_id, fmsio:841-845), as the reference does."""
    f = netcdf_file(path, "r", mmap=False)
    names = set(f.variables)
    n = f.variables["lon"].shape[0] if "lon" in names else 0
    cols = {}
    for name, typ, _, _ in BERG_VARS:
        if name in ("id_cnt", "id_ij"):
            continue
        if name in names:
            cols[name] = np.array(f.variables[name][:], dtype=np.float64 if typ == "d" else np.int32)
        elif name in OPTIONAL_ZERO:
            cols[name] = np.zeros(n)
    for name, _, _, _ in MTS_VARS + DEM_VARS:
        if name in names:
            cols[name] = np.array(f.variables[name][:], dtype=np.float64)
    if "id_cnt" in names and "id_ij" in names:
        cols["id"] = id_from_2_ints(np.array(f.variables["id_cnt"][:]), np.array(f.variables["id_ij"][:]))
    f.close()
    if ignore_ij_restart:            # namelist ignore_ij_restart: find the cell from lon/lat (fmsio:871)
        cols.pop("ine", None)
        cols.pop("jne", None)
    return cols


def write_restart_bonds(path, bonds):
    """write_restart_bonds fmsio:432-560 from ``Icebergs.get_bonds()`` (STS bonds: no DEM history)."""
    n = len(bonds["first_id"])
    fc, fij = split_id(bonds["first_id"])
    oc, oij = split_id(bonds["other_id"])
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    extra = [(nm, t, ln, u, np.asarray(bonds[nm], dtype=np.float64 if t == "d" else np.int32)) for nm, t, ln, u in BOND_DEM_VARS if nm in bonds]
    _write(path, n, extra + [
        ("first_berg_ine", "i", "iceberg ine of first berg in bond", "dimensionless", i32(bonds["first_ine"])),
        ("first_berg_jne", "i", "iceberg jne of first berg in bond", "dimensionless", i32(bonds["first_jne"])),
        ("first_id_cnt", "i", "counter component of iceberg id first berg in bond", "dimensionless", fc),
        ("first_id_ij", "i", "position component of iceberg id first berg in bond", "dimensionless", fij),
        ("other_berg_ine", "i", "iceberg ine of second berg in bond", "dimensionless", i32(bonds["other_ine"])),
        ("other_berg_jne", "i", "iceberg jne of second berg in bond", "dimensionless", i32(bonds["other_jne"])),
        ("other_id_cnt", "i", "counter component of iceberg id second berg in bond", "dimensionless", oc),
        ("other_id_ij", "i", "position component of iceberg id second berg in bond", "dimensionless", oij),
    ])


def read_restart_bonds(path):
    """read_restart_bonds fmsio:1282-1430 -> columns for ``Icebergs.set_bonds``."""
    f = netcdf_file(path, "r", mmap=False)
    g = lambda k: np.array(f.variables[k][:])
    out = dict(first_id=id_from_2_ints(g("first_id_cnt"), g("first_id_ij")), other_id=id_from_2_ints(g("other_id_cnt"), g("other_id_ij")),
               first_ine=g("first_berg_ine").astype(np.int32), first_jne=g("first_berg_jne").astype(np.int32),
               other_ine=g("other_berg_ine").astype(np.int32), other_jne=g("other_berg_jne").astype(np.int32))
    for nm, t, _, _ in BOND_DEM_VARS:
        if nm in f.variables:
            out[nm] = g(nm).astype(np.float64 if t == "d" else np.int32)
    f.close()
    return out


def _calving_window(domain, nj, ni):
    """Where a rank's compute domain sits in a file array of (nj, ni) points: the file holds either the global grid
    (FMS domain-aware restarts are read and written in global index space, fmsio:564-566 / fmsio:1448-1470) or, for a
    distributed per-tile file, just this rank's compute domain."""
    d = domain
    if (nj, ni) == (d.gnj, d.gni):
        return slice(d.jsc - 1, d.jec), slice(d.isc - 1, d.iec)
    if (nj, ni) == (d.njc, d.nic):
        return slice(0, nj), slice(0, ni)
    raise ValueError(f"calving.res.nc holds {nj} x {ni} points: neither the global grid {d.gnj} x {d.gni} nor the "
                     f"compute domain {d.njc} x {d.nic} of this rank")


def write_restart_calving(path, stored_ice, stored_heat, iceberg_counter_grd, domain=None, global_file=False,
                          rmean_calving=None, rmean_calving_hflx=None):
    """calving.res.nc (fmsio:564-566: register_restart_field of stored_ice, stored_heat, iceberg_counter_grd on the
    FMS domain).  FMS writes the COMPUTE domain (halos stripped) with a Time record axis: (Time, zaxis_1, yaxis_1,
    xaxis_1).  ``stored_ice`` etc. are the library's data-domain arrays (``Icebergs.get_calving_state``); with
    ``domain`` the halos are stripped, and with ``global_file`` the compute domain is placed in a global-size array
    (what a single-PE or mpp_io-combined file looks like; the other ranks' points are left zero).  Without ``domain``
    the arrays are written as they are (already halo-free).  ``rmean_calving`` / ``rmean_calving_hflx``
    (``Icebergs.get_calving_rmean``, tau_calving > 0) are written as the reference does (fmsio:568-569)."""
    si, sh, ic = np.asarray(stored_ice), np.asarray(stored_heat), np.asarray(iceberg_counter_grd)
    extra = {k: np.asarray(v) for k, v in (("rmean_calving", rmean_calving), ("rmean_calving_hflx", rmean_calving_hflx)) if v is not None}
    if domain is not None:
        d = domain
        h = d.halo
        si, sh, ic = si[:, h:h + d.njc, h:h + d.nic], sh[h:h + d.njc, h:h + d.nic], ic[h:h + d.njc, h:h + d.nic]
        extra = {k: v[h:h + d.njc, h:h + d.nic] for k, v in extra.items()}
        if global_file:
            for k in list(extra):
                gv = np.zeros((d.gnj, d.gni)); gv[d.jsc - 1:d.jec, d.isc - 1:d.iec] = extra[k]; extra[k] = gv
            gsi = np.zeros((si.shape[0], d.gnj, d.gni)); gsh = np.zeros((d.gnj, d.gni)); gic = np.zeros((d.gnj, d.gni), dtype=np.int32)
            js, is_ = slice(d.jsc - 1, d.jec), slice(d.isc - 1, d.iec)
            gsi[:, js, is_], gsh[js, is_], gic[js, is_] = si, sh, ic
            si, sh, ic = gsi, gsh, gic
    f = netcdf_file(path, "w", version=1)
    nk, nj, ni = si.shape
    f.createDimension("Time", None)        # (scipy's NetCDF-3 writer wants the record dimension first)
    f.createDimension("xaxis_1", ni); f.createDimension("yaxis_1", nj); f.createDimension("zaxis_1", nk)
    for nm, n in (("xaxis_1", ni), ("yaxis_1", nj), ("zaxis_1", nk)):
        v = f.createVariable(nm, "d", (nm,)); v[:] = np.arange(1, n + 1, dtype=np.float64)
    v = f.createVariable("Time", "d", ("Time",)); v[0] = 1.0
    v = f.createVariable("stored_ice", "d", ("Time", "zaxis_1", "yaxis_1", "xaxis_1")); v[0] = si
    v = f.createVariable("stored_heat", "d", ("Time", "yaxis_1", "xaxis_1")); v[0] = sh
    v = f.createVariable("iceberg_counter_grd", "i", ("Time", "yaxis_1", "xaxis_1")); v[0] = ic
    for k, a in extra.items():
        v = f.createVariable(k, "d", ("Time", "yaxis_1", "xaxis_1")); v[0] = a
    f.close()


def read_restart_calving_rmean(path, domain=None):
    """-> (rmean_calving, rmean_calving_hflx) for ``Icebergs.set_calving_rmean``; None for a variable the file does not
    hold (the reference then starts the mean from the first field it sees, fms2io:1517-1534, I:6010-6017)."""
    f = netcdf_file(path, "r", mmap=False)
    out = []
    for name in ("rmean_calving", "rmean_calving_hflx"):
        if name not in f.variables:
            out.append(None)
            continue
        a = np.array(f.variables[name][:], dtype=np.float64)
        if a.ndim == 3:
            a = a[-1]
        if domain is not None:
            d = domain
            js, is_ = _calving_window(d, a.shape[0], a.shape[1])
            o = np.zeros((d.njd, d.nid)); o[d.halo:d.halo + d.njc, d.halo:d.halo + d.nic] = a[js, is_]
            a = o
        out.append(a)
    f.close()
    return tuple(out)


def read_restart_calving(path, domain=None):
    """-> (stored_ice, stored_heat, iceberg_counter_grd) for ``Icebergs.set_calving_state``.  A leading Time axis is
    dropped (last record).  With ``domain`` the rank's compute-domain window of a global (or per-tile) file is cut out
    and padded with zero halos to the data domain the library holds (the reference fills the halos by
    mpp_update_domains on the first icebergs_run, I:5203); without it the arrays come back as stored."""
    f = netcdf_file(path, "r", mmap=False)
    def var(name, nd, dtype):
        a = np.array(f.variables[name][:], dtype=dtype)
        if a.ndim == nd + 1:
            a = a[-1]
        return a
    si, sh = var("stored_ice", 3, np.float64), var("stored_heat", 2, np.float64)
    ic = var("iceberg_counter_grd", 2, np.int32) if "iceberg_counter_grd" in f.variables else np.zeros(sh.shape, dtype=np.int32)
    f.close()
    if domain is None:
        return si, sh, ic
    d = domain
    js, is_ = _calving_window(d, sh.shape[0], sh.shape[1])
    h = d.halo
    osi = np.zeros((si.shape[0], d.njd, d.nid)); osh = np.zeros((d.njd, d.nid)); oic = np.zeros((d.njd, d.nid), dtype=np.int32)
    osi[:, h:h + d.njc, h:h + d.nic] = si[:, js, is_]
    osh[h:h + d.njc, h:h + d.nic] = sh[js, is_]
    oic[h:h + d.njc, h:h + d.nic] = ic[js, is_]
    return osi, osh, oic


# ---------------------------------------------------------------------------------------------------------------
# iceberg_trajectories.nc, write_trajectory fmsio:1575-2047 (the non-FMS-IO branch, fmsio:1893-2040: one unlimited
# dimension "i", the variables below with long_name / units attributes).  Which variables a file holds follows
# save_short_traj / save_fl_traj and the scheme switches exactly as the reference's def_var sequence does.
TRAJ_BASE = [("lon", "d", "longitude", "degrees_E"), ("lat", "d", "latitude", "degrees_N"), ("year", "i", "year", "years"),
             ("day", "d", "year day", "days"), ("id_cnt", "i", "counter component of iceberg id", "dimensionless"),
             ("id_ij", "i", "position component of iceberg id", "dimensionless")]
TRAJ_FL = [("mass", "d", "mass", "kg"), ("start_mass", "d", "start_mass", "kg"), ("thickness", "d", "thickness", "m"),
           ("mass_of_bits", "d", "mass_of_bits", "kg"), ("uvel", "d", "zonal spped", "m/s"), ("vvel", "d", "meridional spped", "m/s")]
TRAJ_FL_FOOTLOOSE = [("mass_scaling", "d", "mass_scaling", "dimensionless"), ("mass_of_fl_bits", "d", "mass_of_fl_bits", "kg"),
                     ("mass_of_fl_bergy_bits", "d", "mass_of_fl_bergy_bits", "kg"), ("fl_k", "d", "footloose calving k", "none")]
# (long names and units as the reference writes them, typos and the 'm' of the accelerations included)
TRAJ_LONG = [("uvel_prev", "d", "zonal speed mts", "m/s"), ("vvel_prev", "d", "meridional speed mts", "m/s"),
             ("uo", "d", "ocean zonal spped", "m/s"), ("vo", "d", "ocean meridional spped", "m/s"),
             ("ui", "d", "ice zonal spped", "m/s"), ("vi", "d", "ice meridional spped", "m/s"),
             ("ua", "d", "atmos zonal spped", "m/s"), ("va", "d", "atmos meridional spped", "m/s"),
             ("heat_density", "d", "heat_density", "J/kg"), ("width", "d", "width", "m"), ("length", "d", "length", "m"),
             ("ssh_x", "d", "sea surface height gradient_x", "non-dim"), ("ssh_y", "d", "sea surface height gradient_y", "non-dim"),
             ("sst", "d", "sea surface temperature", "degrees_C"), ("sss", "d", "sea surface salinity", "psu"),
             ("cn", "d", "sea ice concentration", "none"), ("hi", "d", "sea ice thickness", "m"),
             ("axn", "d", "explicit zonal acceleration", "m"), ("ayn", "d", "explicit meridional acceleration", "m"),
             ("bxn", "d", "implicit zonal acceleration", "m"), ("byn", "d", "implicit meridional acceleration", "m"),
             ("halo_berg", "d", "halo status", "non-dim"), ("od", "d", "ocean_depth", "m")]
TRAJ_MTS = [("axn_fast", "d", "explicit fast step zonal acceleration", "m"), ("ayn_fast", "d", "explicit fast step meridional acceleration", "m"),
            ("bxn_fast", "d", "implicit fast step zonal acceleration", "m"), ("byn_fast", "d", "implicit fast step meridional acceleration", "m")]
TRAJ_BONDS = [("n_bonds", "i", "number of bonds", "dimensionless")]
TRAJ_DEM = [("ang_vel", "d", "angular velocity", "rad/s"), ("ang_accel", "d", "angular acceleration", "rad/s^2"), ("rot", "d", "accumulated rotation", "rad")]


def trajectory_variables(save_short_traj=True, save_fl_traj=True, footloose=False, mts=False, iceberg_bonds_on=False, dem=False):
    """The variables of iceberg_trajectories.nc for a given configuration (fmsio:1925-1990)."""
    v = list(TRAJ_BASE)
    if save_fl_traj:
        v += TRAJ_FL + (TRAJ_FL_FOOTLOOSE if footloose else [])
    if not save_short_traj:
        v += TRAJ_LONG + (TRAJ_MTS if mts else []) + (TRAJ_BONDS if iceberg_bonds_on else []) + (TRAJ_DEM if dem else [])
    return v


def write_trajectory(path, traj, **config):
    """write_trajectory (fmsio:1575) from ``Icebergs.get_trajectory()``; ``config`` as in trajectory_variables.  Records are
    written grouped by berg and in time order, as the reference's per-berg lists come out."""
    order = np.lexsort((traj["day"], traj["year"], traj["id"])) if len(traj["id"]) else np.zeros(0, dtype=np.int64)
    cnt, ij = split_id(traj["id"][order]) if len(order) else (np.zeros(0, np.int32), np.zeros(0, np.int32))
    cols = {k: np.asarray(v)[order] for k, v in traj.items()}
    cols["id_cnt"], cols["id_ij"] = cnt, ij
    f = netcdf_file(path, "w", version=1)
    f.createDimension("i", None)
    for name, t, long_name, units in trajectory_variables(**config):
        v = f.createVariable(name, t, ("i",))
        v.long_name = long_name; v.units = units
        a = cols[name].astype(np.float64 if t == "d" else np.int32)
        if len(a):
            v[:len(a)] = a
    f.close()


def read_trajectory(path):
    f = netcdf_file(path, "r", mmap=False)
    out = {k: np.array(v[:]) for k, v in f.variables.items()}
    f.close()
    if "id_cnt" in out:
        out["id"] = id_from_2_ints(out["id_cnt"], out["id_ij"])
    return out
