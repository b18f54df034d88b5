"""ctypes wrapper of the CPU ORACLE (oracle/libkid_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It offers the same call shapes as icebergs_b200.api so a parity
test can push identical inputs through both.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def load_by_path(name, relpath):
    """A module of the repo loaded from its file, WITHOUT importing the package around it: the struct mirrors
    (icebergs_b200/_cdefs.py, generated from include/kid_b200.h) and the synthetic workload generator
    (icebergs_b200/synthetic.py, pure numpy) are shared with the product, the product itself is never imported here."""
    full = "icebergs_b200." + name
    if full in sys.modules:           # a parity test has the package loaded already: share its classes
        return sys.modules[full]
    key = "_kid_oracle_" + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


D = load_by_path("_cdefs", os.path.join("icebergs_b200", "_cdefs.py"))

# struct pointers are declared void* below: a parity test may hand over structs of the product's own
# icebergs_b200._cdefs classes (same layout, generated from the same header) whatever the import order
LIB_PATH = os.path.join(_HERE, "libkid_oracle.so")
FAST_LIB_PATH = os.path.join(_HERE, "_fast", "libkid_oracle_fast.so")
_vp = C.c_void_p
_LIB = None
_FAST = False


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def use_fast_build():
    """Timing legs of bench.py: the -O3 -march=native -fopenmp build, compiled on this box (oracle/Makefile `fast`).
    Must be called before the first lib() call.  Fails loudly when OpenMP is missing."""
    global _FAST, LIB_PATH
    if _LIB is not None:
        raise RuntimeError("use_fast_build() after the oracle library was loaded")
    subprocess.run(["make", "-s", "-B", "-C", _HERE, "fast"], check=True)   # always: -march=native must match THIS box
    LIB_PATH = FAST_LIB_PATH
    _FAST = True
    if lib().oracle_omp_max_threads() < 1:
        raise RuntimeError("the fast oracle build has no OpenMP")


def default_params(**overrides):
    """Reference namelist defaults (F:686-822) from the oracle itself."""
    p = D.KidParams()
    lib().oracle_default_params(C.byref(p))
    for k, v in overrides.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            for q, x in enumerate(v):
                cur[q] = x
        else:
            setattr(p, k, v)
    return p


class SingleDomain:
    """What api.Domain.single offers the oracle: .c (KidDomain) and the tile sizes."""

    def __init__(self, gni, gnj, halo=4, cyclic_x=True, cyclic_y=False):
        self.c = D.KidDomain()
        lib().oracle_single_domain(C.byref(self.c), gni, gnj, halo, int(cyclic_x), int(cyclic_y))

    def __getattr__(self, k):
        return getattr(self.c, k)

    nic = property(lambda s: s.c.iec - s.c.isc + 1)
    njc = property(lambda s: s.c.jec - s.c.jsc + 1)
    nid = property(lambda s: s.c.ied - s.c.isd + 1)
    njd = property(lambda s: s.c.jed - s.c.jsd + 1)


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_default_params.argtypes = [_vp]
        L.oracle_default_params.restype = None
        L.oracle_single_domain.argtypes = [_vp] + [C.c_int32] * 5
        L.oracle_single_domain.restype = None
        L.oracle_omp_max_threads.restype = C.c_int32
        L.oracle_create.restype = _vp
        L.oracle_create.argtypes = [_vp, _vp, C.c_int32, C.c_double] + [_vp] * 9 + [C.c_int32]
        L.oracle_destroy.argtypes = [_vp]
        L.oracle_last_error.argtypes = [_vp]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_set_bergs.argtypes = [_vp, C.c_int64, _vp]
        L.oracle_count_bergs.argtypes = [_vp, C.c_int32]
        L.oracle_count_bergs.restype = C.c_int64
        L.oracle_get_bergs.argtypes = [_vp, C.POINTER(C.c_int64), _vp, C.c_int32]
        L.oracle_record_posn.argtypes = [_vp]
        L.oracle_trajectory_count.argtypes = [_vp]
        L.oracle_trajectory_count.restype = C.c_int64
        L.oracle_get_trajectory.argtypes = [_vp, C.POINTER(C.c_int64), _vp, C.c_int32]
        L.oracle_set_bonds.argtypes = [_vp, C.c_int64, _vp]
        L.oracle_get_bonds.argtypes = [_vp, C.POINTER(C.c_int64), _vp]
        L.oracle_set_calving_state.argtypes = [_vp, _vp, _vp, _vp]
        L.oracle_get_calving_state.argtypes = [_vp, _vp, _vp, _vp]
        L.oracle_set_calving_rmean.argtypes = [_vp, _vp, _vp]
        L.oracle_run.argtypes = [_vp, C.c_int32, C.c_double] + [_vp] * 12 + [C.c_int32, C.c_int32] + [_vp] * 4
        L.oracle_step_again.argtypes = [_vp, C.c_int32, C.c_int32, C.c_double, C.c_int32]
        L.oracle_get_grid_field.argtypes = [_vp, C.c_int32, _vp]
        L.oracle_get_counters.argtypes = [_vp, _vp]
        L.oracle_last_timing.argtypes = [_vp, C.POINTER(C.c_double)]
        L.oracle_last_timing.restype = None
        L.oracle_bilin.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double]
        L.oracle_bilin.restype = C.c_double
        L.oracle_is_point_in_cell.argtypes = [_vp, C.c_double, C.c_double, C.c_int32, C.c_int32]
        L.oracle_pos_within_cell.argtypes = [_vp, C.c_double, C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_find_cell.argtypes = [_vp, C.c_double, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.oracle_find_cell_wide.argtypes = [_vp, C.c_double, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.oracle_apply_modulo_around_point.argtypes = [C.c_double] * 3
        L.oracle_apply_modulo_around_point.restype = C.c_double
        L.oracle_id_from_2_ints.argtypes = [C.c_int32, C.c_int32]
        L.oracle_id_from_2_ints.restype = C.c_int64
        L.oracle_split_id.argtypes = [C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.oracle_split_id.restype = None
        L.oracle_yearday.argtypes = [C.c_int32] * 5
        L.oracle_yearday.restype = C.c_double
        L.oracle_rolling.argtypes = [_vp] + [C.POINTER(C.c_double)] * 3
        L.oracle_rolling.restype = None
        L.oracle_accel_free.argtypes = [_vp] + [C.POINTER(C.c_double)] * 4
        L.oracle_accel_free.restype = None
        L.oracle_point_in_triangle.argtypes = [C.c_double] * 8
        L.oracle_hexagon_into_quadrants.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_double)] * 5
        L.oracle_hexagon_into_quadrants.restype = None
        _LIB = L
    return _LIB


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def make_columns(n, want=None, cls=None, **arrays):
    cls = cls or D.KidBergColumns
    cols = cls()
    keep = {}
    for name, ctype in cls._fields_:
        a = arrays.get(name)
        base = ctype._type_
        dt = {C.c_double: np.float64, C.c_int32: np.int32, C.c_int64: np.int64}[base]
        if a is None and want and name in want:
            a = np.zeros(n, dtype=dt)
        if a is None:
            continue
        a = np.ascontiguousarray(a, dtype=dt)
        keep[name] = a
        setattr(cols, name, a.ctypes.data_as(ctype))
    return cols, keep


class OracleFatal(RuntimeError):
    pass


class Oracle:
    """CPU restatement of icebergs_init/run/end with the argument meaning of the reference."""

    def __init__(self, gni, gnj, dt, Time, ice_lon, ice_lat, ice_wet, ice_dx, ice_dy, ice_area, cos_rot, sin_rot,
                 ocean_depth=None, fractional_area=False, params=None, domain=None):
        self.params = params
        self.params.dt = dt
        self.domain = domain
        year, yearday = Time
        arrs = [_f64(a) for a in (ice_lon, ice_lat, ice_wet, ice_dx, ice_dy, ice_area, cos_rot, sin_rot, ocean_depth)]
        self._h = lib().oracle_create(C.byref(self.params), C.byref(domain.c), int(year), float(yearday),
                                      *[_ptr(a) for a in arrs], int(fractional_area))
        self._ok()

    def _ok(self, rc=0):
        msg = lib().oracle_last_error(self._h)
        if rc or (msg and len(msg)):
            raise OracleFatal(f"[{rc}] {msg.decode() if msg else ''}")

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bergs(self, **cols):
        n = len(cols["lon"])
        c, keep = make_columns(n, **cols)
        self._ok(lib().oracle_set_bergs(self._h, n, C.byref(c)))

    def count_bergs(self, include_halo=False):
        return lib().oracle_count_bergs(self._h, int(include_halo))

    def get_bergs(self, names, include_halo=False):
        n = self.count_bergs(include_halo)
        cap = max(n, 1)
        c, keep = make_columns(cap, want=set(names))
        m = C.c_int64(cap)
        self._ok(lib().oracle_get_bergs(self._h, C.byref(m), C.byref(c), int(include_halo)))
        return {k: v[: m.value].copy() for k, v in keep.items()}

    def record_posn(self):
        """record_posn F:5328 at the time of the last run"""
        self._ok(lib().oracle_record_posn(self._h))

    def get_trajectory(self, clear=True):
        n = int(lib().oracle_trajectory_count(self._h))
        cap = max(n, 1)
        c, keep = make_columns(cap, want={f[0] for f in D.KidTrajColumns._fields_}, cls=D.KidTrajColumns)
        m = C.c_int64(cap)
        self._ok(lib().oracle_get_trajectory(self._h, C.byref(m), C.byref(c), int(clear)))
        return {k: v[: m.value].copy() for k, v in keep.items()}

    def set_bonds(self, **cols):
        n = len(cols["first_id"]) if cols else 0
        c, keep = make_columns(n, cls=D.KidBondColumns, **cols)
        self._ok(lib().oracle_set_bonds(self._h, n, C.byref(c)))

    def get_bonds(self):
        n = C.c_int64(0)
        lib().oracle_get_bonds(self._h, C.byref(n), None)
        cap = max(n.value, 1)
        names = {"first_id", "other_id", "first_ine", "first_jne", "other_ine", "other_jne", "length", "tangd1", "tangd2",
                 "nstress", "sstress", "rel_rotation", "broken"}
        c, keep = make_columns(cap, want=names, cls=D.KidBondColumns)
        m = C.c_int64(cap)
        self._ok(lib().oracle_get_bonds(self._h, C.byref(m), C.byref(c)))
        return {k: v[: m.value].copy() for k, v in keep.items()}

    def set_calving_state(self, stored_ice=None, stored_heat=None, iceberg_counter_grd=None):
        si, sh = _f64(stored_ice), _f64(stored_heat)
        ic = None if iceberg_counter_grd is None else np.ascontiguousarray(iceberg_counter_grd, dtype=np.int32)
        self._ok(lib().oracle_set_calving_state(self._h, _ptr(si), _ptr(sh), _ptr(ic)))

    def set_calving_rmean(self, rmean_calving=None, rmean_calving_hflx=None):
        a, b = _f64(rmean_calving), _f64(rmean_calving_hflx)
        self._ok(lib().oracle_set_calving_rmean(self._h, _ptr(a), _ptr(b)))

    def get_calving_state(self):
        d = self.domain
        si = np.zeros((D.KID_NCLASSES, d.njd, d.nid))
        sh = np.zeros((d.njd, d.nid))
        ic = np.zeros((d.njd, d.nid), dtype=np.int32)
        lib().oracle_get_calving_state(self._h, _ptr(si), _ptr(sh), _ptr(ic))
        return si, sh, ic

    def run(self, time, calving, uo, vo, ui, vi, tauxa, tauya, ssh, sst, calving_hflx, cn, hi, stagger=0,
            stress_stagger=None, sss=None, mass_berg=None, ustar_berg=None, area_berg=None):
        if stress_stagger is None:
            stress_stagger = stagger
        ins = [_f64(a) for a in (uo, vo, ui, vi, tauxa, tauya, ssh, sst)]
        cn_, hi_, sss_ = _f64(cn), _f64(hi), _f64(sss)
        year, yearday = time
        rc = lib().oracle_run(self._h, int(year), float(yearday), _ptr(calving), *[_ptr(a) for a in ins],
                              _ptr(calving_hflx), _ptr(cn_), _ptr(hi_), stagger, stress_stagger, _ptr(sss_),
                              _ptr(mass_berg), _ptr(ustar_berg), _ptr(area_berg))
        self._ok(rc)

    def step_again(self, nsteps, year=0, yearday=0.0, nthreads=1):
        self._ok(lib().oracle_step_again(self._h, nsteps, year, yearday, nthreads))

    def grid_field(self, field_id):
        d = self.domain
        out = np.zeros((d.njd, d.nid))
        rc = lib().oracle_get_grid_field(self._h, field_id, _ptr(out))
        if rc:
            raise OracleFatal(f"grid field {field_id}")
        return out

    def counters(self):
        c = D.KidCounters()
        lib().oracle_get_counters(self._h, C.byref(c))
        return {n: getattr(c, n) for n, _ in D.KidCounters._fields_}

    def last_timing(self):
        s = (C.c_double * 4)()
        lib().oracle_last_timing(self._h, s)
        return {"momentum": s[0], "thermodyn": s[1], "rest": s[2]}
