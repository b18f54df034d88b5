"""Berg mass / area / momentum spread onto the ocean grid (calculate_mass_on_ocean
I:4970, spread_mass_across_ocean_cells I:3895 with the hexagon quadrant geometry I:4562,
sum_up_spread_fields I:6077, ustar and the thickness cutoff I:3461-3488) and what icebergs_run
returns in mass_berg / ustar_berg / area_berg (I:5663-5678): CUDA path against the CPU oracle."""
import numpy as np
import pytest

import kid_oracle_py as O
from common import Case, grid_rel
from icebergs_b200 import _cdefs as D
from icebergs_b200 import api
from icebergs_b200 import synthetic as S

pytestmark = pytest.mark.gpu

SPREAD = (D.KID_FLD_SPREAD_MASS, D.KID_FLD_SPREAD_AREA, D.KID_FLD_SPREAD_UVEL, D.KID_FLD_SPREAD_VVEL,
          D.KID_FLD_USTAR_ICEBERG, D.KID_FLD_MASS, D.KID_FLD_BERGY_MASS, D.KID_FLD_FLOATING_MELT)


def outputs(run, who, case, **kw):
    comp = (case.gnj, case.gni)
    m, u, a = np.zeros(comp), np.zeros(comp), np.zeros(comp)
    calving, hflx, f = case.run_args(**kw)
    args = ((1, 0.0), calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hflx, f["cn"], f["hi"])
    if run == "gpu":
        api.icebergs_run(who, *args, sss=f["sss"], mass_berg=m, ustar_berg=u, area_berg=a)
    else:
        who.run(*args, sss=f["sss"], mass_berg=m, ustar_berg=u, area_berg=a)
    return calving, m, u, a


@pytest.mark.parametrize("old_spreading", [1, 0])
def test_rectangular_spreading_latlon(old_spreading):
    """Default rectangular bergs on the lat-lon grid; big scaled bergs so that the weights reach the
    neighbouring cells, dense enough that several bergs share a cell."""
    case = Case(96, 48, 30000, add_weight_to_ocean=1, use_old_spreading=old_spreading, pass_fields_to_ocean_model=1,
                apply_thickness_cutoff_to_gridded_melt=1, melt_cutoff=3995.0)
    b, o = case.make_gpu(), case.make_oracle()
    for step in range(2):
        cg, mg, ug, ag = outputs("gpu", b, case)
        co, mo, uo_, ao = outputs("ora", o, case)
        for fid in SPREAD:
            assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-10, f"field {fid} step {step}"
        assert grid_rel(mg, mo) < 1e-10 and grid_rel(ug, uo_) < 1e-10 and grid_rel(ag, ao) < 1e-10 and grid_rel(cg, co) < 1e-10
    assert mo.max() > 0 and ao.max() > 0 and uo_.max() > 0
    tot = np.sum(o.grid_field(D.KID_FLD_SPREAD_MASS) * o.grid_field(D.KID_FLD_AREA))
    bg = b.get_bergs(["mass", "mass_of_bits", "mass_scaling"])
    assert abs(tot - np.sum((bg["mass"] + bg["mass_of_bits"]) * bg["mass_scaling"])) / tot < 1e-9     # all the mass is on the grid
    mass = np.zeros((case.gnj, case.gni))
    b.incr_mass(mass)
    assert grid_rel(mass, mo) < 1e-12                                                                     # icebergs_incr_mass I:6046
    allm = b.get_bergs(["mass", "mass_of_bits", "mass_scaling"])
    want = np.sum((allm["mass"] + allm["mass_of_bits"]) * allm["mass_scaling"])
    assert abs(b.stock_pe(D.KID_ISTOCK_WATER) - want) / want < 1e-12                                      # icebergs_stock_pe I:8102
    assert abs(b.stock_pe(D.KID_ISTOCK_HEAT) + want * case.params().hlf) / (want * case.params().hlf) < 1e-12
    api.icebergs_end(b)


def test_add_iceberg_thickness_to_ssh():
    """add_iceberg_thickness_to_SSH (I:5330-5337): from the second call on the sea surface height the bergs feel is the
    freeboard-equivalent of the mass spread in the previous step, not the input field."""
    from common import COMPARE_F64, assert_bergs_match
    case = Case(96, 48, 30000, add_weight_to_ocean=1, use_old_spreading=0, add_iceberg_thickness_to_ssh=1)
    b, o = case.make_gpu(), case.make_oracle()
    names = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
    for step in range(3):
        outputs("gpu", b, case); outputs("ora", o, case)
        assert grid_rel(b.grid_field(D.KID_FLD_SSH), o.grid_field(D.KID_FLD_SSH)) < 1e-10, f"ssh, call {step}"
        assert_bergs_match(b.get_bergs(names), o.get_bergs(names), rtol=1e-9, context=f"call {step}")
    ssh = o.grid_field(D.KID_FLD_SSH)
    assert ssh.max() > 0.0 and not np.allclose(ssh[4:-4, 4:-4], case.forcing["ssh"][1:-1, 1:-1])      # it did replace the input
    api.icebergs_end(b)


def test_hexagonal_spreading_with_bond_orientation():
    """tests/collision_tests: hexagonal elements, orientation from the bonds (I:3829), 1 km cells."""
    from test_interactions_gpu import Pair
    params = lambda: S.collision_params(api.default_params, pass_fields_to_ocean_model=1)
    p = Pair(S.collision_bergs(), params)
    for k in range(3):
        p.step(20)
        for fid in SPREAD:
            a, b_ = p.b.grid_field(fid), p.o.grid_field(fid)
            assert np.max(np.abs(a - b_)) <= 1e-9 * max(np.max(np.abs(b_)), 1e-300), f"field {fid} after {20 * (k + 1)} steps"
    sm = p.o.grid_field(D.KID_FLD_SPREAD_MASS)
    assert (sm > 0).sum() > 16            # the 16 elements cover more cells than they sit in
    p.end()


# ---- the reference's own known answers through the CUDA geometry (not only through the oracle's copy of it)
def _known():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))


def test_hexagon_identities_on_the_device():
    """hexagon_test I:261-348: the 8 area identities of Hexagon_into_quadrants_using_triangles, tol 1e-10, evaluated
    by kid_spread.cuh's kh_hexagon_into_quadrants on the GPU."""
    import ctypes as C
    k = _known()["hexagon"]
    out = (C.c_double * 5)()
    assert len(k["cases"]) >= 7
    for case in k["cases"]:
        rc = api.lib().kid_unit_hexagon_into_quadrants(0, case["x0"], case["y0"], k["H"], k["theta"], out)
        assert rc == 0, case["name"]
        assert abs(out[0] - k["area"]) <= k["tol"], case["name"]
        for got, want in zip(list(out)[1:], case["Q"]):
            assert abs(got - want) <= k["tol"], (case["name"], list(out), case["Q"])
        assert abs(sum(list(out)[1:]) - out[0]) <= k["tol"]


def test_point_in_triangle_regression_on_the_device():
    """I:234-242: the point the reference's unit test found misclassified by round-off"""
    import ctypes as C
    k = _known()["point_in_triangle"]
    v = (C.c_double * 8)(*k["A"], *k["B"], *k["C"], *k["q"])
    inside, area = C.c_int32(-1), C.c_double(0.0)
    assert api.lib().kid_unit_point_in_triangle(0, v, C.byref(inside), C.byref(area)) == 0
    assert bool(inside.value) is k["inside"]
    assert area.value > 0.0


@pytest.mark.parametrize("add_weight", [1, 0])
def test_melt_from_the_spread_mass(add_weight):
    """find_melt_using_spread_mass (I:5490-5500, I:3436-3448): the spread mass is summed between the move and the melt and
    again after it; floating_melt is the difference per step, calving_hflx = floating_melt * HLF, and what icebergs_run
    hands back in calving / calving_hflx follows.  The move and the melt run as separate launches for this."""
    from common import COMPARE_F64, assert_bergs_match
    case = Case(96, 48, 20000, add_weight_to_ocean=add_weight, find_melt_using_spread_mass=1, use_old_spreading=0,
                hexagonal_icebergs=1)
    b, o = case.make_gpu(), case.make_oracle()
    names = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
    for step in range(3):
        calving, hflx, f = case.run_args()
        args = (f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"])
        cg, hg = calving.copy(), hflx.copy()
        api.icebergs_run(b, (1, 0.0), cg, *args, hg, f["cn"], f["hi"], sss=f["sss"])
        co, ho = calving.copy(), hflx.copy()
        o.run((1, 0.0), co, *args, ho, f["cn"], f["hi"], sss=f["sss"])
        assert_bergs_match(b.get_bergs(names), o.get_bergs(names), rtol=1e-10 if step == 0 else 1e-9, context=f"fmusm step {step}")
        for fid in (D.KID_FLD_SPREAD_MASS, D.KID_FLD_BERGY_MASS, D.KID_FLD_BERG_MELT):
            assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-10, (fid, step)
        # the melt is a difference of two sums of O(1e12 kg) spread masses: compare against the spread mass per second
        scale = np.abs(o.grid_field(D.KID_FLD_SPREAD_MASS)).max() / case.dt
        for fid in (D.KID_FLD_FLOATING_MELT, D.KID_FLD_CALVING_HFLX):
            got, want = b.grid_field(fid), o.grid_field(fid)
            tol = 1e-10 * scale * (case.params().hlf if fid == D.KID_FLD_CALVING_HFLX else 1.0)
            assert np.abs(got - want).max() <= tol, (fid, step, np.abs(got - want).max(), tol)
        assert np.abs(cg - co).max() <= 1e-10 * scale and np.abs(hg - ho).max() <= 1e-10 * scale * case.params().hlf
    fm = o.grid_field(D.KID_FLD_FLOATING_MELT)
    assert fm.max() > 0 and np.array_equal(o.grid_field(D.KID_FLD_CALVING_HFLX)[4:-4, 4:-4], (fm * case.params().hlf)[4:-4, 4:-4])
    api.icebergs_end(b)
