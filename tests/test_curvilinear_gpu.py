"""General (distorted-quadrilateral) cells: calc_xiyj F:6439-6531 through pos_within_cell F:6299 with
grid_is_regular=.false. -- the inverse bilinear map with its quadratic root choice, is_point_in_cell's cross
products on cell edges, the cell walk and the coast bounce on a mesh whose cells are not lon/lat rectangles.
The regular-grid shortcut of the library (RectCell) does not apply to any cell here."""
import numpy as np
import pytest

from common import COMPARE_F64, Case, assert_bergs_match, grid_rel, run_gpu, run_oracle
from icebergs_b200 import _cdefs as D
from icebergs_b200 import api
from icebergs_b200 import synthetic as S

pytestmark = pytest.mark.gpu

NAMES = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]


class SkewGrid(S.Grid):
    """The synthetic lat-lon grid with its corners displaced: lon by a function of lat, lat by a function of lon
    (periodic in lon, so the mesh stays cyclic).  Every cell is a general quadrilateral."""

    def _distort(self, lon, lat):
        rad = np.pi / 180.0
        return (lon + 0.35 * self.dlon * np.sin(3.0 * lat * rad) + 0.1 * self.dlon * np.cos(4.0 * lon * rad),
                lat + 0.30 * self.dlat * np.sin(2.0 * lon * rad))

    def corner_lonlat(self, ring=0):
        return self._distort(*super().corner_lonlat(ring))

    def seed_bergs(self, n, **kw):
        cols, counter = super().seed_bergs(n, **kw)
        i, j = cols["ine"].astype(np.float64), cols["jne"].astype(np.float64)
        xi = cols["lon"] / self.dlon - (i - 1.0)
        yj = (cols["lat"] + 90.0) / self.dlat - (j - 1.0)
        c = {}
        for name, (di, dj) in dict(ne=(0, 0), nw=(-1, 0), se=(0, -1), sw=(-1, -1)).items():
            c[name] = self._distort((i + di) * self.dlon, -90.0 + (j + dj) * self.dlat)
        for q, key in ((0, "lon"), (1, "lat")):      # bilin F:7071 (the non-bug formula)
            cols[key] = (c["ne"][q] * xi + c["nw"][q] * (1 - xi)) * yj + (c["se"][q] * xi + c["sw"][q] * (1 - xi)) * (1 - yj)
        cols["start_lon"], cols["start_lat"] = cols["lon"].copy(), cols["lat"].copy()
        return cols, counter


def _case(n=6000, **over):
    return Case(96, 48, n, grid=SkewGrid(96, 48), grid_is_regular=0, old_bug_bilin=0, **over)


def test_no_cell_is_a_rectangle():
    g = SkewGrid(96, 48)
    lon, lat = g.corner_lonlat(0)
    assert (np.abs(np.diff(lat, axis=1)) > 1e-6).any() and (np.abs(np.diff(lon, axis=0)) > 1e-6).any()


def test_positions_within_general_cells_match_oracle():
    """xi, yj of the restart ingest (pos_within_cell at the given ine/jne) on distorted cells"""
    case = _case()
    b, o = case.make_gpu(), case.make_oracle()
    got, want = b.get_bergs(["xi", "yj", "ine", "jne", "id", "start_year"]), o.get_bergs(["xi", "yj", "ine", "jne", "id", "start_year"])
    assert_bergs_match(got, want, names=("xi", "yj"), context="ingest on the curvilinear grid")
    assert (np.abs(got["xi"] - 0.5) < 0.5).all() and (np.abs(got["yj"] - 0.5) < 0.5).all()
    api.icebergs_end(b)
    o.close()


@pytest.mark.parametrize("dt,steps", [(3600.0, 6), (43200.0, 6)])
def test_steps_on_curvilinear_grid_match_oracle(dt, steps):
    """calc_xiyj on every position update; with half-day steps most bergs change cell every step (the cell walk of
    adjust_index_and_ground on general cells) and many meet the coast"""
    case = _case(dt=dt)
    b, o = case.make_gpu(), case.make_oracle()
    fast = {} if dt < 4000 else {k: np.full_like(case.forcing[k], v) for k, v in dict(uo=0.9, vo=0.25).items()}
    moved = 0
    for step in range(steps):
        before = b.get_bergs(["ine", "jne", "id"])
        run_gpu(b, case, **fast); run_oracle(o, case, **fast)
        got, want = b.get_bergs(NAMES), o.get_bergs(NAMES)
        assert_bergs_match(got, want, rtol=1e-10 if step == 0 else 1e-8, context=f"curvilinear dt={dt} step {step}")
        if len(before["id"]) == len(got["id"]):
            ob, og = np.argsort(before["id"]), np.argsort(got["id"])
            moved += int(((before["ine"][ob] != got["ine"][og]) | (before["jne"][ob] != got["jne"][og])).sum())
        for fid in (D.KID_FLD_FLOATING_MELT, D.KID_FLD_BERG_MELT):
            assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-9
    cg, co = b.counters(), o.counters()
    assert cg["n_bounced"] == co["n_bounced"]
    if dt > 4000:
        assert moved > 2000 and cg["n_bounced"] > 0, (moved, cg["n_bounced"])
    assert cg["error_flags"] == 0
    api.icebergs_end(b)
    o.close()
