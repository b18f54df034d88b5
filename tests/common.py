"""Shared case builders for the parity tests: the same inputs go to the CUDA library
(icebergs_b200.api) and to the CPU oracle (oracle/kid_oracle_py.py)."""
from __future__ import annotations

import numpy as np

import kid_oracle_py as O
from icebergs_b200 import api
from icebergs_b200 import synthetic as S

COMPARE_F64 = ("lon", "lat", "uvel", "vvel", "axn", "ayn", "bxn", "byn", "uvel_prev", "vvel_prev", "xi", "yj",
               "mass", "thickness", "width", "length", "mass_of_bits", "mass_scaling", "heat_density",
               "start_day", "start_mass", "start_lon", "start_lat")
COMPARE_INT = ("ine", "jne", "start_year", "id")
RTOL = 1e-10  # north_star: positions, velocities, masses within 1e-10 relative after one step

RUN_KEYS = ("uo", "vo", "ui", "vi", "tauxa", "tauya", "ssh", "sst", "cn", "hi")


class Case:
    def __init__(self, gni, gnj, n, dt=3600.0, seed=S.SEED, params=None, halo=4, grid=None, capacity=0, **over):
        self.grid = grid or S.Grid(gni, gnj)
        self.gni, self.gnj, self.dt = gni, gnj, dt
        self.over = dict(over)
        self.halo = halo
        self.init = self.grid.init_args()
        self.forcing = self.grid.forcing()
        self.bergs, self.counter = self.grid.seed_bergs(n, seed=seed) if n else (None, None)
        self.capacity = capacity or max(4 * n, 1024)

    def params(self):
        return S.workload_params(api.default_params, halo=self.halo, **self.over)

    def domain(self):
        return api.Domain.single(self.gni, self.gnj, halo=self.halo, cyclic_x=True)

    def counter_dd(self, dom):
        """iceberg_counter_grd on the data domain, consistent with the seeded ids."""
        c = np.zeros((dom.njd, dom.nid), dtype=np.int32)
        h = self.halo
        c[h:h + self.gnj, h:h + self.gni] = self.counter
        return c

    def make_gpu(self):
        p, d = self.params(), self.domain()
        b = api.icebergs_init(self.gni, self.gnj, self.dt, (1, 0.0), params=p, domain=d, capacity=self.capacity,
                              **self.init)
        if self.bergs is not None:
            b.set_calving_state(iceberg_counter_grd=self.counter_dd(d))
            b.set_bergs(**self.bergs)
        return b

    def make_oracle(self):
        p, d = self.params(), self.domain()
        o = O.Oracle(self.gni, self.gnj, self.dt, (1, 0.0), params=p, domain=d, **self.init)
        if self.bergs is not None:
            o.set_calving_state(iceberg_counter_grd=self.counter_dd(d))
            o.set_bergs(**self.bergs)
        return o

    def run_args(self, **replace):
        f = dict(self.forcing)
        f.update(replace)
        calving = np.ascontiguousarray(f["calving"]).copy()
        hflx = np.ascontiguousarray(f["calving_hflx"]).copy()
        return calving, hflx, f


def run_gpu(bergs, case, time=(1, 0.0), **replace):
    calving, hflx, f = case.run_args(**replace)
    api.icebergs_run(bergs, time, calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"],
                     f["sst"], hflx, f["cn"], f["hi"], sss=f["sss"])
    return calving, hflx


def run_oracle(orc, case, time=(1, 0.0), **replace):
    calving, hflx, f = case.run_args(**replace)
    orc.run(time, calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hflx,
            f["cn"], f["hi"], sss=f["sss"])
    return calving, hflx


def by_id(cols):
    order = np.argsort(cols["id"], kind="stable")
    return {k: v[order] for k, v in cols.items()}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    return np.where(a == b, 0.0, np.abs(a - b) / scale)


def assert_bergs_match(got, want, rtol=RTOL, names=COMPARE_F64, context="", acc_floor=0.0):
    """ids, counts, cell indices bit-exact; floats within rtol relative (absolute floor for
    values that are differences of O(1) numbers: accelerations near zero)."""
    assert len(got["id"]) == len(want["id"]), f"{context}: berg count {len(got['id'])} != {len(want['id'])}"
    g, w = by_id(got), by_id(want)
    for k in COMPARE_INT:
        assert np.array_equal(g[k], w[k]), f"{context}: {k} differs at {np.nonzero(g[k] != w[k])[0][:5]}"
    worst = {}
    for k in names:
        if k not in g or k not in w:
            continue
        e = rel_err(g[k], w[k])
        # accelerations and velocities are sums with cancellation: compare against the field's scale
        floor = max(1e-13 * max(np.max(np.abs(w[k])), 1e-300), acc_floor) if k in ("axn", "ayn", "bxn", "byn", "uvel", "vvel", "uvel_prev", "vvel_prev", "uvel_old", "vvel_old", "ang_vel", "ang_accel", "rot") else 0.0
        if k == "rot":          # radians; before any torque acts the rotation is rounding noise of ~1e-12
            floor = max(floor, 1e-10)
        if k in ("xi", "yj"):   # fractions of a cell in [0,1): the tolerance is relative to the cell, not to xi
            floor = rtol
        bad = (e > rtol) & (np.abs(g[k] - w[k]) > floor)
        worst[k] = float(e.max()) if len(e) else 0.0
        assert not bad.any(), f"{context}: {k} rel err {e[bad].max():.3e} at id {g['id'][np.nonzero(bad)[0][:3]]}"
    return worst


def grid_rel(a, b):
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)
