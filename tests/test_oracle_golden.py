"""Pins the CPU oracle against the reference's own portable known answers:
bilin corner exactness (icebergs_framework.F90:7313-7316), the 64-bit id round trip
(F:7319-7325), yearday (F:4431-4441), and structural invariants of the restated step."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

import kid_oracle_py as O
from common import Case, run_oracle
from icebergs_b200 import _cdefs as D
from icebergs_b200 import synthetic as S


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KNOWN = json.load(open(os.path.join(GOLDEN, "known_answers.json")))


def test_id_round_trip_F7319():
    L = O.lib()
    k = KNOWN["id_round_trip"]
    i, c1 = k["i"], k["counter"]
    ident = L.oracle_id_from_2_ints(c1, i)
    assert ident == k["id"] == c1 * 2 ** 32 + i
    c2, j = C.c_int32(), C.c_int32()
    L.oracle_split_id(ident, C.byref(c2), C.byref(j))
    assert (c2.value, j.value) == (c1, i)


def test_yearday_F4431():
    L = O.lib()
    for mon, day, hr, mi, sec, want in KNOWN["yearday"]["cases"]:
        assert L.oracle_yearday(mon, day, hr, mi, sec) == want


@pytest.mark.parametrize("old_bug", [0])
def test_bilin_corner_exactness_F7313(old_bug):
    # unit_tests() runs with the corrected bilin (every shipped namelist sets old_bug_bilin=.false.)
    case = Case(48, 24, 0, old_bug_bilin=old_bug)
    o = case.make_oracle()
    L = O.lib()
    d = case.domain()
    lon, lat = o.grid_field(D.KID_FLD_LON), o.grid_field(D.KID_FLD_LAT)
    at = lambda f, i, j: f[j - d.jsd, i - d.isd]
    i, j = d.isc, d.jsc
    assert L.oracle_bilin(o._h, D.KID_FLD_LON, i, j, 0.0, 1.0) == at(lon, i - 1, j)
    assert L.oracle_bilin(o._h, D.KID_FLD_LON, i, j, 1.0, 1.0) == at(lon, i, j)
    assert L.oracle_bilin(o._h, D.KID_FLD_LAT, i, j, 1.0, 0.0) == at(lat, i, j - 1)
    assert L.oracle_bilin(o._h, D.KID_FLD_LAT, i, j, 1.0, 1.0) == at(lat, i, j)


def test_apply_modulo_around_point_F6558():
    L = O.lib()
    f = L.oracle_apply_modulo_around_point
    assert f(370.0, 0.0, 360.0) == 10.0
    assert f(-10.0, 350.0, 360.0) == 350.0
    assert f(5.0, 5.0, -1.0) == 5.0          # non-periodic: identity
    for x, y in [(0.1, 359.9), (359.9, 0.1), (720.5, 10.0)]:
        r = f(x, y, 360.0)
        assert abs(r - y) <= 180.0 and abs(((r - x) / 360.0) - round((r - x) / 360.0)) < 1e-12


def test_cell_membership_edges_F6199():
    # The half-valued sentinels at F:6203-6206 make two of the four edges part of the cell.  The
    # comment above them says "South and East", but with the counter-clockwise corner order the
    # cross products l0..l3 are all NEGATIVE inside, so the edges whose zero maps to a negative
    # sentinel (p0 = S, p3 = W) are the inclusive ones.  The code is followed, not the comment.
    case = Case(48, 24, 0, old_bug_bilin=0)
    o = case.make_oracle()
    L = O.lib()
    i, j = 10, 12
    dlon, dlat = 360.0 / 48, 180.0 / 24
    x0, x1, y0, y1 = (i - 1) * dlon, i * dlon, -90 + (j - 1) * dlat, -90 + j * dlat
    inside = lambda x, y: L.oracle_is_point_in_cell(o._h, x, y, i, j)
    assert inside(0.5 * (x0 + x1), 0.5 * (y0 + y1))
    assert inside(0.5 * (x0 + x1), y0) and not inside(0.5 * (x0 + x1), y1)
    assert inside(x0, 0.5 * (y0 + y1)) and not inside(x1, 0.5 * (y0 + y1))


def test_step_conserves_count_and_ids():
    case = Case(96, 48, 3000)
    o = case.make_oracle()
    before = np.sort(o.get_bergs(["id"])["id"])
    run_oracle(o, case)
    after = o.get_bergs(["id", "ine", "jne", "lon", "lat"])
    assert np.array_equal(np.sort(after["id"]), before)
    # every berg sits in the cell that contains it
    L = O.lib()
    for k in range(0, len(after["id"]), 37):
        assert L.oracle_is_point_in_cell(o._h, after["lon"][k], after["lat"][k], int(after["ine"][k]), int(after["jne"][k]))


def test_melt_budget_closes():
    # mass lost by the bergs (x mass_scaling) = what the grid received, I:3116-3133
    case = Case(96, 48, 2000)
    o = case.make_oracle()
    names = ["id", "mass", "mass_of_bits", "mass_scaling"]
    b0 = o.get_bergs(names)
    run_oracle(o, case)
    b1 = o.get_bergs(names)
    o0 = np.argsort(b0["id"]); o1 = np.argsort(b1["id"])
    lost = np.sum(((b0["mass"] + b0["mass_of_bits"])[o0] - (b1["mass"] + b1["mass_of_bits"])[o1]) * b0["mass_scaling"][o0])
    fm = o.grid_field(D.KID_FLD_FLOATING_MELT)
    area = o.grid_field(D.KID_FLD_AREA)
    assert lost > 0
    assert abs(np.sum(fm * area) * case.dt - lost) / lost < 1e-9


def test_oracle_regression_fixture():
    """tests/golden/oracle_step_48x24.npz (made by tests/golden/make_golden.py with this oracle):
    freezes the restated arithmetic -- not a reference pin."""
    from common import COMPARE_F64, by_id
    fx = np.load(os.path.join(GOLDEN, "oracle_step_48x24.npz"))
    case = Case(48, 24, 400)
    o = case.make_oracle()
    run_oracle(o, case)
    o.step_again(3, 1, 0.0)
    b = by_id(o.get_bergs(list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]))
    for k in fx.files:
        if b[k].dtype.kind == "i":
            assert np.array_equal(b[k], fx[k]), k
        else:
            assert np.allclose(b[k], fx[k], rtol=1e-13, atol=0), k


def test_oracle_fold_regression_fixture():
    """tests/golden/oracle_fold_90x24.npz: freezes the restated tripolar fold (halo rows beyond the fold, bergs handed across
    it, turned spreading weights).  Not a reference pin: the fold's pin is the analytic continuation of the bipolar cap,
    tests/test_fold_oracle.py."""
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    fx = np.load(os.path.join(GOLDEN, "oracle_fold_90x24.npz"))
    got = MG.fold_run()
    assert set(fx.files) == set(got)
    for k in fx.files:
        if got[k].dtype.kind == "i":
            assert np.array_equal(got[k], fx[k]), k
        else:
            assert np.allclose(got[k], fx[k], rtol=1e-12, atol=1e-300), k
    assert np.abs(fx["halo.uo"]).max() > 0 and fx["spread_mass"].max() > 0


def test_oracle_mts_regression_fixture():
    """tests/golden/oracle_mts_20h.npz: freezes the MTS / DEM restatement (not a reference pin; those are the berg
    counts and the beam tests below)."""
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    fx = np.load(os.path.join(GOLDEN, "oracle_mts_20h.npz"))
    for tag, over in (("mts_kid", MG.MTS_NML), ("ikid", MG.IKID_NML)):
        b = MG.mts_run(over)
        for k, v in b.items():
            want = fx[f"{tag}.{k}"]
            if v.dtype.kind == "i":
                assert np.array_equal(v, want), (tag, k)
            else:
                assert np.allclose(v, want, rtol=1e-11, atol=1e-18), (tag, k)


def test_calving_tables_F787():
    p = Case(48, 24, 0).params()
    t = KNOWN["calving_tables"]
    for name in ("initial_mass_s", "distribution_s", "mass_scaling_s", "initial_thickness_s"):
        assert list(getattr(p, name)) == t[name]


def test_footloose_test_ends_with_12_bergs():
    """Known answer of the reference's footloose test (tests/footloose_tests/input.nml:1, restart
    checksum line '#=12'): two 3.6 km elements, 192 h of 10 s steps, fl_style='fl_bits' -- the two
    parents plus ten bergs made from footloose bits.  The child displacement (FMS random stream) is off
    here; it moves children, it does not change how many there are."""
    from icebergs_b200 import api
    from icebergs_b200 import synthetic as S
    g = S.CartesianGrid()
    dom = api.Domain.single(20, 20, halo=3, cyclic_x=True)
    o = O.Oracle(20, 20, 10.0, (1, 0.0), params=S.footloose_params(api.default_params), domain=dom, **g.init_args())
    o.set_bergs(**S.footloose_bergs())
    f = S.footloose_forcing(g)
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"])
    nsteps, done = int(192 * 3600 / 10), 1
    while done < nsteps:
        n = min(8640, nsteps - done)
        o.step_again(n, 1, done * 10.0 / 86400.0)
        done += n
    assert o.count_bergs() == KNOWN["restart_counts"]["footloose"] == 12
    assert o.counters()["nbergs_calved_fl"] == 10


def test_point_in_triangle_I234():
    k = KNOWN["point_in_triangle"]
    L = O.lib()
    assert bool(L.oracle_point_in_triangle(*k["A"], *k["B"], *k["C"], *k["q"])) is k["inside"]


def test_hexagon_quadrant_areas_I261():
    """hexagon_test (I:261-348): the 8 area identities of Hexagon_into_quadrants_using_triangles, tol 1e-10."""
    k = KNOWN["hexagon"]
    L = O.lib()
    out = [C.c_double() for _ in range(5)]
    for case in k["cases"]:
        L.oracle_hexagon_into_quadrants(case["x0"], case["y0"], k["H"], k["theta"], *[C.byref(v) for v in out])
        area, q = out[0].value, [v.value for v in out[1:]]
        assert abs(area - k["area"]) <= k["tol"], case["name"]
        for got, want in zip(q, case["Q"]):
            assert abs(got - want) <= k["tol"], (case["name"], q, case["Q"])


# ------------------------------------------------------------------------------------------------------------------
# DEM known answers: the reference's two beam tests (Wang 2020, sections 3.1 and 3.2).  tests/dem_ssbeam_test and
# tests/dem_cbeam_test are checked by eye in the reference ("the beam should bend into alignment with the plotted
# line", README; the line is beam theory, animate_trajectories.py:143-157 / :149-157); here the same comparison is a
# number.  They pin calculate_force_dem (normal + shear springs, torque from relative rotation), the rotation update
# and the explicit MTS sub-steps of the oracle.
def _beam_run(bergs, nsteps, dt, **over):
    from icebergs_b200 import api
    g = S.CartesianGrid(20, 20, 15000.0)
    dom = api.Domain.single(20, 20, halo=3, cyclic_x=True)
    o = O.Oracle(20, 20, dt, (1, 0.0), params=S.beam_params(api.default_params, **over), domain=dom, **g.init_args())
    o.set_bergs(**bergs)
    o.set_bonds()
    f = g.forcing(ibuo=0.0, ibvo=0.0, collision_test=False)
    for k in range(nsteps):
        c, h = f["calving"].copy(), f["calving_hflx"].copy()
        o.run((1, k * dt / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h,
              f["cn"], f["hi"], sss=f["sss"])
    b = o.get_bergs(["lon", "lat", "vvel", "rot", "start_lon", "start_lat"])
    o.close()
    order = np.lexsort((b["start_lon"], b["start_lat"]))
    return {k: v[order] for k, v in b.items()}


def test_dem_simply_supported_beam_known_answer():
    """dem_ssbeam_test: 29 elements of radius 0.25 m, E = 1e9 Pa, 1.5e5 N at mid-span, ends held vertically
    (I:1862-1869); 1e5 sub-steps per 1 s step.  Euler-Bernoulli mid-span deflection P l^3 / (48 E I) = -0.823 m."""
    b0 = S.beam_bergs()
    b = _beam_run(b0, 8, 1.0)
    xa = b0["lon"] - b0["lon"][0]
    l, P, YM, AI = xa.max(), -1.5e5, 1.0e9, 1.0 * 0.5 ** 3 / 12.0
    w1 = -P * xa * (4.0 * xa * xa - 3.0 * l * l) / (48.0 * YM * AI)
    w2 = P * (xa - l) * (l * l - 8.0 * l * xa + 4.0 * xa * xa) / (48.0 * YM * AI)
    w = np.where(xa > 0.5 * l, w2, w1)                       # animate_trajectories.py:143-157
    d = b["lat"] - b0["lat"]
    assert abs(d[14] / w[14] - 1.0) < 0.03, (d[14], w[14])
    assert np.sqrt(np.mean((d - w) ** 2)) < 0.03 * np.abs(w).max()
    assert np.abs(b["vvel"]).max() < 0.05                    # damped out


def test_dem_cantilever_beam_known_answer():
    """dem_cbeam_test geometry (3 x 30 elements of radius 2.5 km, clamped = static first column, 1.5e10 N shared by the
    three end elements I:1870-1876, orig_dem_moment_of_inertia), stiffened to E = 1e11 Pa so that the linear theory the
    reference plots applies: tip deflection P l^3 / (3 E I), I = thick (3 * 5 km)^3 / 12."""
    b0 = S.cantilever_bergs()
    E = 1.0e11
    b = _beam_run(b0, 200, 100.0, dem_beam_test=2, orig_dem_moment_of_inertia=1, dem_damping_coef=0.7, rho_bergs=900.0,
                  mts_sub_steps=2000, dem_spring_coef=E)
    l, hh, P = 29 * 5000.0, 3.0 * 5000.0, -1.5e10
    wtip = P * l ** 3 / (3.0 * E * hh ** 3 / 12.0)
    tips = [29, 59, 89]
    tip = np.mean(b["lat"][tips] - b0["lat"][tips])
    assert abs(tip / wtip - 1.0) < 0.03, (tip, wtip)          # 1.5 % above: shear deformation of the lattice
    assert np.abs(b["lat"][[0, 30, 60]] - b0["lat"][[0, 30, 60]]).max() == 0.0      # the clamped column is static
    theta = P * l * l / (2.0 * E * hh ** 3 / 12.0)
    assert abs(np.mean(b["rot"][tips]) / theta - 1.0) < 0.03


# ------------------------------------------------------------------------------------------------------------------
# tests/collision_tests in the reference's own configuration: KID with ibdt = 60 s, MTS_KID / iKID with ibdt = 3600 s and
# 60 sub-steps, 48 h (README: '#= 16' bergs at the end of each).  The reference runs them on 4 PEs; on one rank a bonded
# conglomerate that crosses the cyclic seam (after ~24 h) needs the seam-aware partner search of connect_all_bonds
# (oracle/kid_oracle_ext.inc), and the MTS scheme stops short of the seam (no transfer_mts_bergs).
def _collision_run(dt, nsteps, **over):
    from icebergs_b200 import api
    g = S.CartesianGrid()
    dom = api.Domain.single(g.gni, g.gnj, halo=3, cyclic_x=True)
    o = O.Oracle(g.gni, g.gnj, dt, (1, 0.0), params=S.collision_params(api.default_params, **over), domain=dom, **g.init_args())
    o.set_bergs(**S.collision_bergs())
    o.set_bonds()
    f = g.forcing()
    for k in range(nsteps):
        c, h = f["calving"].copy(), f["calving_hflx"].copy()
        o.run((1, k * dt / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h,
              f["cn"], f["hi"], sss=f["sss"])
    b = o.get_bergs(["lon", "lat", "uvel", "vvel", "start_lat"])
    n = o.count_bergs()
    o.close()
    south = b["start_lat"] < 10.0e3
    return n, float(b["lat"][~south].mean() - b["lat"][south].mean()), float(b["lon"].mean()), float(np.abs(np.r_[b["uvel"], b["vvel"]]).max())


MTS_NML = dict(mts=1, mts_sub_steps=60, explicit_inner_mts=1, force_convergence=1, convergence_tolerance=1e-8,
               contact_distance=1.75e3, contact_spring_coef=1.0e-7)


def test_collision_test_kid_48h_berg_count():
    n, sep, x, vmax = _collision_run(60.0, 2880)
    assert n == 16                                            # README:16-17
    assert 1500.0 < sep < 2500.0 and vmax < 0.3, (sep, vmax)  # the conglomerates met, bounced and stay together; nothing blew up
    assert 15.0e3 < x < 17.5e3                                # 34 km of drift: once through the 20 km periodic seam


def test_collision_test_mts_and_ikid_reference_time_steps():
    ref = _collision_run(60.0, 1200)
    for over in (MTS_NML, dict(MTS_NML, dem=1, poisson=0.3, dem_damping_coef=1.0, dem_spring_coef=4471.94)):
        n, sep, x, vmax = _collision_run(3600.0, 20, **over)
        assert n == 16                                        # README:19-22
        assert abs(x / ref[2] - 1.0) < 0.06 and vmax < 0.3, (x, ref[2], vmax)     # same drift as the single-time-step scheme
        assert 2000.0 < sep < 6000.0
