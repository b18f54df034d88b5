"""Host-side mirror of the reference's public interface for the hot path.

The reference exposes ``icebergs_init / icebergs_run / icebergs_end`` (module
``ice_bergs``, src/icebergs.F90:65-66, signatures I:92-117, I:5074-5096, I:8152)
around an opaque ``type(icebergs), pointer``.  Here the same three calls (same
argument names, meaning and order; the FMS-only arguments ``layout, io_layout, axes,
dom_x_flags, dom_y_flags`` collapse into a :class:`Domain`) drive the CUDA library
through its C ABI (include/kid_b200.h).  Errors the reference reports with
``error_mesg(..., FATAL)`` raise :class:`KidFatal` carrying the same text.

Array convention: every 2-D array is the reference's column-major ``(i, j)`` array,
i.e. a C-contiguous numpy array of shape ``(nj, ni)``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cdefs as D
from . import _lib

BGRID_NE, CGRID_NE, AGRID = D.KID_BGRID_NE, D.KID_CGRID_NE, D.KID_AGRID

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _lib.load()
    return _LIB


class KidFatal(RuntimeError):
    """error_mesg(..., FATAL) of the reference."""

    def __init__(self, code, msg):
        super().__init__(f"[kid status {code}] {msg}")
        self.code = code
        self.msg = msg


def default_params(**overrides) -> D.KidParams:
    """Reference namelist defaults (F:686-822) + FMS constants; keyword overrides."""
    p = D.KidParams()
    lib().kid_default_params(C.byref(p))
    set_params(p, **overrides)
    return p


def set_params(p, **kw):
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(f"KidParams has no field {k!r}")
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            for q, x in enumerate(v):
                cur[q] = x
        else:
            setattr(p, k, v)
    return p


@dataclass
class Domain:
    """What mpp_define_domains hands the reference (F:915-930)."""
    c: D.KidDomain

    @classmethod
    def single(cls, gni, gnj, halo=4, cyclic_x=True, cyclic_y=False, device=0):
        d = D.KidDomain()
        lib().kid_single_domain(C.byref(d), gni, gnj, halo, int(cyclic_x), int(cyclic_y), device)
        return cls(d)

    @classmethod
    def decomposed(cls, gni, gnj, rank, nranks, halo=4, cyclic_x=True, cyclic_y=False, device=0, comm=None,
                   comm_kind=D.KID_COMM_NCCL):
        d = D.KidDomain()
        rc = lib().kid_define_domain(C.byref(d), gni, gnj, halo, int(cyclic_x), int(cyclic_y), rank, nranks, device)
        if rc:
            raise KidFatal(rc, "kid_define_domain failed")
        if comm is not None:
            d.nccl_comm = comm
            d.comm_kind = comm_kind
        return cls(d)

    def owner_rank(self, i, j) -> int:
        """Rank that owns global cell (i, j) (-1 = outside the model, NULL_PE)."""
        return lib().kid_owner_rank(C.byref(self.c), int(i), int(j))

    def __getattr__(self, k):
        return getattr(self.c, k)

    @property
    def halo(self):
        return self.c.isc - self.c.isd

    @property
    def nic(self):
        return self.c.iec - self.c.isc + 1

    @property
    def njc(self):
        return self.c.jec - self.c.jsc + 1

    @property
    def nid(self):
        return self.c.ied - self.c.isd + 1

    @property
    def njd(self):
        return self.c.jed - self.c.jsd + 1


def _f64(a, shape=None, name="array"):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_columns(n, want=None, cls=None, **arrays):
    """Builds a KidBergColumns over numpy arrays.  Returns (struct, dict of arrays kept alive)."""
    cls = cls or D.KidBergColumns
    cols = cls()
    keep = {}
    for name, ctype in cls._fields_:
        a = arrays.get(name)
        if a is None and want and name in want:
            base = ctype._type_
            a = np.zeros(n, dtype={C.c_double: np.float64, C.c_int32: np.int32, C.c_int64: np.int64}[base])
        if a is None:
            continue
        base = ctype._type_
        dt = {C.c_double: np.float64, C.c_int32: np.int32, C.c_int64: np.int64}[base]
        a = np.ascontiguousarray(a, dtype=dt)
        if a.shape != (n,):
            raise ValueError(f"column {name}: expected {n} entries, got {a.shape}")
        keep[name] = a
        setattr(cols, name, a.ctypes.data_as(ctype))
    return cols, keep


RESTART_COLUMNS = ("lon", "lat", "uvel", "vvel", "mass", "thickness", "width", "length", "axn", "ayn", "bxn",
                   "byn", "uvel_prev", "vvel_prev", "start_lon", "start_lat", "start_day", "start_mass",
                   "mass_scaling", "mass_of_bits", "heat_density", "xi", "yj", "static_berg", "halo_berg",
                   "mass_of_fl_bits", "mass_of_fl_bergy_bits", "fl_k", "start_year", "ine", "jne", "id")


class Icebergs:
    """The opaque ``type(icebergs), pointer :: bergs`` of the reference."""

    def __init__(self, handle, params, domain):
        self._h = handle
        self.params = params
        self.domain = domain

    # -- error plumbing ----------------------------------------------------
    def _check(self, rc):
        if rc != D.KID_OK:
            msg = lib().kid_last_error(self._h)
            raise KidFatal(rc, msg.decode() if msg else "?")

    @property
    def handle(self):
        if not self._h:
            raise KidFatal(D.KID_ERR_STATE, "icebergs handle already ended")
        return self._h

    # -- restart-side accessors (read_restart_bergs / write_restart) --------
    def set_bergs(self, **cols):
        n = len(cols["lon"])
        c, keep = make_columns(n, **cols)
        self._check(lib().kid_set_bergs(self.handle, n, C.byref(c)))

    def count_bergs(self) -> int:
        n = C.c_int64(0)
        self._check(lib().kid_count_bergs(self.handle, C.byref(n)))
        return n.value

    def get_bergs(self, names=RESTART_COLUMNS, include_halo=False) -> dict:
        if include_halo:
            # owned bergs + halo copies: ask the library how many there are (a call with no room reports the count)
            m = C.c_int64(0)
            empty = D.KidBergColumns()
            rc = lib().kid_get_bergs(self.handle, C.byref(m), C.byref(empty), 1)
            if rc not in (D.KID_OK, D.KID_ERR_CAPACITY):
                self._check(rc)
            n = m.value
        else:
            n = self.count_bergs()
        cap = max(n, 1)
        c, keep = make_columns(cap, want=set(names))
        m = C.c_int64(cap)
        self._check(lib().kid_get_bergs(self.handle, C.byref(m), C.byref(c), int(include_halo)))
        return {k: v[: m.value].copy() for k, v in keep.items()}

    def record_posn(self):
        """record_posn F:5328-5498: sample the bergs that pass the trajectory criteria at the time of the last
        icebergs_run (the reference samples every traj_sample_hrs, I:5173-5178; the caller holds the date)."""
        self._check(lib().kid_record_posn(self.handle))

    def get_trajectory(self, clear=True) -> dict:
        """The samples taken so far, one entry per record of iceberg_trajectories.nc (no particular order)."""
        n = C.c_int64(0)
        self._check(lib().kid_trajectory_count(self.handle, C.byref(n)))
        cap = max(n.value, 1)
        c, keep = make_columns(cap, want={f[0] for f in D.KidTrajColumns._fields_}, cls=D.KidTrajColumns)
        m = C.c_int64(cap)
        self._check(lib().kid_get_trajectory(self.handle, C.byref(m), C.byref(c), int(clear)))
        return {k: v[: m.value].copy() for k, v in keep.items()}

    def set_bonds(self, **cols):
        """read_restart_bonds (columns of bonds_iceberg.res.nc, fmsio:473-493); with no columns and
        manually_initialize_bonds the bonds come from initialize_iceberg_bonds (I:356-441)."""
        n = len(cols["first_id"]) if cols else 0
        c, keep = make_columns(n, cls=D.KidBondColumns, **cols)
        self._check(lib().kid_set_bonds(self.handle, n, C.byref(c)))

    def get_bonds(self) -> dict:
        n = C.c_int64(0)
        self._check(lib().kid_get_bonds(self.handle, C.byref(n), None))
        cap = max(n.value, 1)
        names = {"first_id", "other_id", "first_ine", "first_jne", "other_ine", "other_jne", "length", "tangd1", "tangd2",
                 "nstress", "sstress", "rel_rotation", "broken"}
        c, keep = make_columns(cap, want=names, cls=D.KidBondColumns)
        m = C.c_int64(cap)
        self._check(lib().kid_get_bonds(self.handle, C.byref(m), C.byref(c)))
        return {k: v[: m.value].copy() for k, v in keep.items()}

    def stock_pe(self, index) -> float:
        """icebergs_stock_pe, I:8102 (index = KID_ISTOCK_WATER or KID_ISTOCK_HEAT)."""
        v = C.c_double(0.0)
        self._check(lib().kid_stock(self.handle, int(index), C.byref(v)))
        return v.value

    def incr_mass(self, mass):
        """icebergs_incr_mass, I:6046: mass (njc, nic) += spread berg mass (kg/m2)."""
        self._check(lib().kid_incr_mass(self.handle, _ptr(mass)))

    def set_calving_state(self, stored_ice=None, stored_heat=None, iceberg_counter_grd=None):
        d = self.domain
        si = _f64(stored_ice, (D.KID_NCLASSES, d.njd, d.nid), "stored_ice")
        sh = _f64(stored_heat, (d.njd, d.nid), "stored_heat")
        ic = None if iceberg_counter_grd is None else np.ascontiguousarray(iceberg_counter_grd, dtype=np.int32)
        self._check(lib().kid_set_calving_state(self.handle, _ptr(si), _ptr(sh), _ptr(ic)))

    def get_calving_state(self):
        d = self.domain
        si = np.zeros((D.KID_NCLASSES, d.njd, d.nid))
        sh = np.zeros((d.njd, d.nid))
        ic = np.zeros((d.njd, d.nid), dtype=np.int32)
        self._check(lib().kid_get_calving_state(self.handle, _ptr(si), _ptr(sh), _ptr(ic)))
        return si, sh, ic

    def set_calving_rmean(self, rmean_calving=None, rmean_calving_hflx=None):
        """running means of get_running_mean_calving I:5999 (tau_calving > 0), data-domain arrays as calving.res.nc carries them"""
        d = self.domain
        a, b = _f64(rmean_calving, (d.njd, d.nid), "rmean_calving"), _f64(rmean_calving_hflx, (d.njd, d.nid), "rmean_calving_hflx")
        self._check(lib().kid_set_calving_rmean(self.handle, _ptr(a), _ptr(b)))

    def get_calving_rmean(self):
        d = self.domain
        a, b = np.zeros((d.njd, d.nid)), np.zeros((d.njd, d.nid))
        self._check(lib().kid_get_calving_rmean(self.handle, _ptr(a), _ptr(b)))
        return a, b

    def grid_field(self, field_id) -> np.ndarray:
        d = self.domain
        out = np.zeros((d.njd, d.nid))
        self._check(lib().kid_get_grid_field(self.handle, field_id, _ptr(out)))
        return out

    def counters(self) -> dict:
        c = D.KidCounters()
        self._check(lib().kid_get_counters(self.handle, C.byref(c)))
        out = {n: getattr(c, n) for n, _ in D.KidCounters._fields_}
        out["n_slots_hint"] = out["nbergs"]
        return out

    def set_forcing(self, uo, vo, ui, vi, tauxa, tauya, ssh, sst, cn, hi, calving=None, calving_hflx=None,
                    stagger=BGRID_NE, stress_stagger=None, sss=None):
        d = self.domain
        ring, comp = (d.njc + 2, d.nic + 2), (d.njc, d.nic)
        if stress_stagger is None:
            stress_stagger = stagger
        a = [_f64(calving, comp, "calving"), _f64(uo, ring, "uo"), _f64(vo, ring, "vo"), _f64(ui, ring, "ui"),
             _f64(vi, ring, "vi"), _f64(tauxa, comp, "tauxa"), _f64(tauya, comp, "tauya"), _f64(ssh, ring, "ssh"),
             _f64(sst, comp, "sst"), _f64(calving_hflx, comp, "calving_hflx"), _f64(cn, ring, "cn"),
             _f64(hi, ring, "hi")]
        s = _f64(sss, comp, "sss")
        self._check(lib().kid_set_forcing(self.handle, *[_ptr(x) for x in a], stagger, stress_stagger, _ptr(s)))

    def step_resident(self, nsteps, year=0, yearday=0.0):
        self._check(lib().kid_step_resident(self.handle, nsteps, year, yearday))

    def sort(self):
        self._check(lib().kid_sort_bergs(self.handle))

    def set_sort_phase(self, interval=0, steps_since_sort=-1):
        self._check(lib().kid_set_sort_phase(self.handle, int(interval), int(steps_since_sort)))

    def sorts_done(self) -> int:
        return lib().kid_sorts_done(self.handle)

    def slow_fraction(self):
        """share of the bergs the fast step kernel deferred to the slow-path kernel in the last step (None: not used)"""
        n = lib().kid_last_slow_count(self.handle)
        return None if n < 0 else n / max(self.count_bergs(), 1)

    def last_timing(self):
        ms = (C.c_double * 8)()
        self._check(lib().kid_last_timing(self.handle, ms))
        names = ["interface", "calving", "momentum+thermodyn", "communication", "thermodyn_arrivals", "sort", "total", "_"]
        return dict(zip(names, list(ms)))

    def kernel_launches(self) -> int:
        return lib().kid_kernel_launches(self.handle)


def icebergs_init(gni, gnj, dt, Time, ice_lon, ice_lat, ice_wet, ice_dx, ice_dy, ice_area, cos_rot, sin_rot,
                  ocean_depth=None, fractional_area=False, params=None, domain=None, capacity=0) -> Icebergs:
    """icebergs_init, I:92-117.

    ``Time`` is ``(year, yearday)`` (what get_date/yearday F:4431 make of the FMS time).
    ice_lon/ice_lat/ice_area/ocean_depth: (njc, nic); ice_wet/ice_dx/ice_dy/cos_rot/sin_rot:
    (njc+2, nic+2) (D:341-344, F:1021-1056).
    """
    p = params if params is not None else default_params()
    p.dt = dt
    dom = domain if domain is not None else Domain.single(gni, gnj, halo=p.halo, cyclic_x=p.Lx > 0.)
    comp, ring = (dom.njc, dom.nic), (dom.njc + 2, dom.nic + 2)
    arrs = [_f64(ice_lon, comp, "ice_lon"), _f64(ice_lat, comp, "ice_lat"), _f64(ice_wet, ring, "ice_wet"),
            _f64(ice_dx, ring, "ice_dx"), _f64(ice_dy, ring, "ice_dy"), _f64(ice_area, comp, "ice_area"),
            _f64(cos_rot, ring, "cos_rot"), _f64(sin_rot, ring, "sin_rot"), _f64(ocean_depth, comp, "ocean_depth")]
    h = C.c_void_p()
    year, yearday = Time
    rc = lib().kid_init(C.byref(h), C.byref(p), C.byref(dom.c), int(year), float(yearday), int(capacity),
                        *[_ptr(a) for a in arrs], int(fractional_area))
    if rc != D.KID_OK:
        msg = lib().kid_last_error(h if h.value else None)
        if h.value:
            lib().kid_end(C.byref(h))
        raise KidFatal(rc, msg.decode() if msg else "?")
    return Icebergs(h, p, dom)


def icebergs_run(bergs: Icebergs, time, calving, uo, vo, ui, vi, tauxa, tauya, ssh, sst, calving_hflx, cn, hi,
                 stagger=BGRID_NE, stress_stagger=None, sss=None, mass_berg=None, ustar_berg=None,
                 area_berg=None, sample_traj=False):
    """icebergs_run, I:5074-5096.  ``calving`` and ``calving_hflx`` are updated in place
    (intent inout), as are the optional mass_berg/ustar_berg/area_berg outputs."""
    d = bergs.domain
    ring, comp = (d.njc + 2, d.nic + 2), (d.njc, d.nic)
    if stress_stagger is None:
        stress_stagger = stagger
    for name, a in (("calving", calving), ("calving_hflx", calving_hflx)):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.shape == comp):
            raise ValueError(f"{name} must be a C-contiguous float64 array of shape {comp} (it is intent inout)")
    ins = [_f64(uo, ring, "uo"), _f64(vo, ring, "vo"), _f64(ui, ring, "ui"), _f64(vi, ring, "vi"),
           _f64(tauxa, comp, "tauxa"), _f64(tauya, comp, "tauya"), _f64(ssh, ring, "ssh"), _f64(sst, comp, "sst")]
    cn_, hi_, sss_ = _f64(cn, ring, "cn"), _f64(hi, ring, "hi"), _f64(sss, comp, "sss")
    year, yearday = time
    rc = lib().kid_run(bergs.handle, int(year), float(yearday), _ptr(calving), *[_ptr(a) for a in ins],
                       _ptr(calving_hflx), _ptr(cn_), _ptr(hi_), stagger, stress_stagger, _ptr(sss_),
                       _ptr(mass_berg), _ptr(ustar_berg), _ptr(area_berg))
    bergs._check(rc)
    if sample_traj:                       # I:5516: if (sample_traj .or. writeandstop) call record_posn(bergs)
        bergs.record_posn()


def icebergs_prefetch(bergs: Icebergs, calving, uo, vo, ui, vi, tauxa, tauya, ssh, sst, calving_hflx, cn, hi, sss=None):
    """kid_prefetch_forcing: queue the host-to-device copies of the NEXT icebergs_run's inputs now (they overlap the step
    in flight).  The arrays must be C-contiguous float64 of the shapes icebergs_run takes, stay unchanged until that
    icebergs_run has returned, and be passed to it as the same objects."""
    d = bergs.domain
    ring, comp = (d.njc + 2, d.nic + 2), (d.njc, d.nic)
    arrs = [(calving, comp), (uo, ring), (vo, ring), (ui, ring), (vi, ring), (tauxa, comp), (tauya, comp), (ssh, ring),
            (sst, comp), (calving_hflx, comp), (cn, ring), (hi, ring), (sss, comp)]
    for a, shape in arrs:
        if a is not None and not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.shape == shape):
            raise ValueError("icebergs_prefetch: arrays must be C-contiguous float64 of the shapes icebergs_run takes (no copies are made)")
    bergs._check(lib().kid_prefetch_forcing(bergs.handle, *[_ptr(a) for a, _ in arrs]))


def icebergs_end(bergs: Icebergs):
    """icebergs_end, I:8152."""
    if bergs._h:
        lib().kid_end(C.byref(bergs._h))
        bergs._h = None
