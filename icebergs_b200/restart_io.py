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


def write_restart_calving(path, stored_ice, stored_heat, iceberg_counter_grd):
    """calving.res.nc (fmsio:564-566) on the rank's data domain: stored_ice (nclasses, nj, ni)."""
    f = netcdf_file(path, "w", version=1)
    nk, nj, ni = stored_ice.shape
    f.createDimension("xaxis_1", ni); f.createDimension("yaxis_1", nj); f.createDimension("zaxis_1", nk)
    v = f.createVariable("stored_ice", "d", ("zaxis_1", "yaxis_1", "xaxis_1")); v[:] = stored_ice
    v = f.createVariable("stored_heat", "d", ("yaxis_1", "xaxis_1")); v[:] = stored_heat
    v = f.createVariable("iceberg_counter_grd", "i", ("yaxis_1", "xaxis_1")); v[:] = iceberg_counter_grd
    f.close()


def read_restart_calving(path):
    f = netcdf_file(path, "r", mmap=False)
    out = (np.array(f.variables["stored_ice"][:], dtype=np.float64), np.array(f.variables["stored_heat"][:], dtype=np.float64),
           np.array(f.variables["iceberg_counter_grd"][:], dtype=np.int32))
    f.close()
    return out
