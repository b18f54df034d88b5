"""Parity of the CUDA hot path (through the C ABI) with the CPU oracle on identical
seeded inputs.  Bar (north_star): cell indices, berg counts, ids, calving events
bit-exact; positions, velocities, masses within 1e-10 relative after one step."""
import numpy as np
import pytest

from common import (COMPARE_F64, Case, assert_bergs_match, by_id, grid_rel, rel_err, run_gpu, run_oracle)
from icebergs_b200 import _cdefs as D
from icebergs_b200 import api
from icebergs_b200 import synthetic as S

pytestmark = pytest.mark.gpu

NAMES = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
FLUX_FIELDS = (D.KID_FLD_FLOATING_MELT, D.KID_FLD_BERG_MELT, D.KID_FLD_BERGY_SRC, D.KID_FLD_BERGY_MELT,
               D.KID_FLD_CALVING_HFLX)


def both(case):
    return case.make_gpu(), case.make_oracle()


def compare_state(b, o, context, rtol=1e-10):
    return assert_bergs_match(b.get_bergs(NAMES), o.get_bergs(NAMES), rtol=rtol, context=context)


@pytest.mark.parametrize("old_bug", [1, 0])
def test_one_step_parity_coarse_grid(old_bug):
    case = Case(96, 48, 20000, old_bug_bilin=old_bug)
    b, o = both(case)
    compare_state(b, o, "after restart ingest")
    cg, hg = run_gpu(b, case)
    co, ho = run_oracle(o, case)
    compare_state(b, o, f"one step old_bug_bilin={old_bug}")
    for fid in FLUX_FIELDS:
        assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-10, f"grid field {fid}"
    assert np.max(rel_err(cg, co)) < 1e-10 and np.max(rel_err(hg, ho)) < 1e-10
    api.icebergs_end(b)


def test_one_step_parity_quarter_degree():
    # the BASELINE grid (1/4 degree, 1440x720) at a berg count the oracle steps in seconds
    case = Case(1440, 720, 200000)
    b, o = both(case)
    run_gpu(b, case)
    run_oracle(o, case)
    worst = compare_state(b, o, "1/4 degree, one step")
    assert max(worst.values()) < 1e-10
    for fid in FLUX_FIELDS:
        assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-10
    api.icebergs_end(b)


def test_multi_step_divergence_bound():
    """100 steps: the stated trajectory-divergence bound.  With FMA contraction off the two
    paths differ only in libm transcendentals (sin/cos/pow, <= 2 ulp), so cell indices still
    agree for all but bergs within rounding of an edge and positions stay within 1e-9
    relative (about 1e-7 degrees = 1 cm)."""
    case = Case(192, 96, 5000)
    b, o = both(case)
    run_gpu(b, case)
    run_oracle(o, case)
    for _ in range(9):
        b.step_resident(11, 1, 0.0)
        o.step_again(11, 1, 0.0)
    g, w = by_id(b.get_bergs(NAMES)), by_id(o.get_bergs(NAMES))
    assert np.array_equal(g["id"], w["id"])
    mism = int(np.sum((g["ine"] != w["ine"]) | (g["jne"] != w["jne"])))
    assert mism == 0, f"{mism} cell-index mismatches after 100 steps"
    dpos = max(np.max(np.abs(g["lon"] - w["lon"])), np.max(np.abs(g["lat"] - w["lat"])))
    assert dpos < 1e-7, f"position divergence {dpos} degrees after 100 steps"
    assert np.max(rel_err(g["mass"], w["mass"])) < 1e-9
    co, cg = o.counters(), b.counters()
    assert cg["nbergs_melted"] == co["nbergs_melted"] and cg["n_bounced"] == co["n_bounced"]
    api.icebergs_end(b)


def test_coast_bounce_and_cyclic_wrap():
    """Fast zonal current: bergs cross the periodic seam (send_bergs_to_other_pes to self,
    F:2997) and run into the analytic continents (bounce, I:7941-7998)."""
    case = Case(96, 48, 8000, dt=3 * 86400.0)
    f = case.forcing
    fast = dict(uo=np.full_like(f["uo"], 1.2), vo=np.full_like(f["vo"], 0.15), tauxa=np.full_like(f["tauxa"], 15.0))
    b, o = both(case)
    for step in range(4):
        run_gpu(b, case, **fast)
        run_oracle(o, case, **fast)
        compare_state(b, o, f"bounce/wrap step {step}", rtol=1e-8)
    co, cg = o.counters(), b.counters()
    assert co["n_bounced"] > 50 and cg["n_bounced"] == co["n_bounced"]
    assert co["n_received"] > 0
    api.icebergs_end(b)


def test_melt_to_death_counts():
    """Warm water, long steps: bergs melt away; deletion (Mnew<=0, I:3271) must hit the same
    bergs on both paths and the budgets must agree."""
    case = Case(96, 48, 6000, dt=20 * 86400.0)
    f = case.forcing
    warm = dict(sst=np.full_like(f["sst"], 12.0))
    b, o = both(case)
    for step in range(6):
        run_gpu(b, case, **warm)
        run_oracle(o, case, **warm)
        assert b.count_bergs() == o.count_bergs()
        compare_state(b, o, f"melt step {step}", rtol=1e-7)   # 6 steps of 20 days each: rounding-level differences grow
    co, cg = o.counters(), b.counters()
    assert co["nbergs_melted"] > 500 and cg["nbergs_melted"] == co["nbergs_melted"]
    api.icebergs_end(b)


def test_calving_events_bit_exact():
    """accumulate_calving + calve_icebergs (I:6153, I:6225): new bergs, their ids
    (counter*2^32 + ij, F:4165-4177), classes and start days must be bit-exact."""
    case = Case(96, 48, 500, capacity=400000)
    f = case.forcing
    rng = np.random.default_rng(7)
    calving = np.zeros_like(f["calving"])
    lat = case.init["ice_lat"]
    coast = (np.abs(lat) > 65) & (np.abs(lat) < 79) & (case.grid.wet(0) > 0)
    calving[coast] = rng.uniform(1e-4, 2e-2, size=int(coast.sum()))   # kg/m2/s: a few bergs per cell and step, several classes
    hflx = calving * -3.0e4
    b, o = both(case)
    for step in range(5):
        cg, hg = run_gpu(b, case, time=(1, 10.0 + step), calving=calving, calving_hflx=hflx)
        co, ho = run_oracle(o, case, time=(1, 10.0 + step), calving=calving, calving_hflx=hflx)
        assert b.count_bergs() == o.count_bergs(), f"count after calving step {step}"
        compare_state(b, o, f"calving step {step}", rtol=1e-9)
        assert np.max(rel_err(cg, co)) < 1e-10 and np.max(rel_err(hg, ho)) < 1e-9
    co_, cg_ = o.counters(), b.counters()
    assert co_["nbergs_calved"] > 100 and cg_["nbergs_calved"] == co_["nbergs_calved"]
    sg, hg_, ig = b.get_calving_state()
    so, ho_, io = o.get_calving_state()
    h = case.halo
    inner = (slice(None), slice(h, -h), slice(h, -h))
    assert np.array_equal(ig[inner[1:]], io[inner[1:]])
    assert np.max(rel_err(sg[inner], so[inner])) < 1e-12
    api.icebergs_end(b)


def test_sort_does_not_change_results():
    case = Case(96, 48, 10000)
    b1, b2 = case.make_gpu(), case.make_gpu()
    run_gpu(b1, case); run_gpu(b2, case)
    for _ in range(5):
        b1.step_resident(3, 1, 0.0)
        b2.step_resident(3, 1, 0.0)
        b2.sort()
    g1, g2 = by_id(b1.get_bergs(NAMES)), by_id(b2.get_bergs(NAMES))
    for k in NAMES:
        assert np.array_equal(g1[k], g2[k]), k
    api.icebergs_end(b1); api.icebergs_end(b2)


def test_full_size_properties():
    """BASELINE size (10M bergs, 1/4 degree): size-independent properties -- berg count and id
    set conserved, every berg inside its cell range, melt budget closes against the grid."""
    n = 10_000_000
    case = Case(1440, 720, n, capacity=n + 1024)
    b = case.make_gpu()
    names = ["id", "mass", "mass_of_bits", "mass_scaling", "ine", "jne", "xi", "yj"]
    s0 = by_id(b.get_bergs(names))
    run_gpu(b, case)
    s1 = by_id(b.get_bergs(names))
    assert len(s1["id"]) == n and np.array_equal(s0["id"], s1["id"])
    assert s1["ine"].min() >= 1 and s1["ine"].max() <= 1440 and s1["jne"].min() >= 1 and s1["jne"].max() <= 720
    assert s1["xi"].min() >= 0 and s1["xi"].max() < 1 and s1["yj"].min() >= 0 and s1["yj"].max() < 1
    lost = np.sum(((s0["mass"] + s0["mass_of_bits"]) - (s1["mass"] + s1["mass_of_bits"])) * s0["mass_scaling"])
    fm, area = b.grid_field(D.KID_FLD_FLOATING_MELT), b.grid_field(D.KID_FLD_AREA)
    assert abs(np.sum(fm * area) * case.dt - lost) / lost < 1e-9
    api.icebergs_end(b)


@pytest.mark.parametrize("over", [
    dict(melt_icebergs_as_ice_shelf=1),                                            # 3-equation, constant gamma (defaults)
    dict(melt_icebergs_as_ice_shelf=1, const_gamma=0),                             # 3-equation, turbulent exchange
    dict(melt_icebergs_as_ice_shelf=1, use_three_equation_model=0),                # 2-equation
    dict(use_mixed_melting=1, const_gamma=0, use_mixed_layer_salinity_for_thermo=1, melt_cutoff=10.0,
         apply_thickness_cutoff_to_bergs_melt=1),
])
def test_ice_shelf_style_basal_melt(over):
    """find_basal_melt (I:3492-3826): the two-/three-equation melt of melt_icebergs_as_ice_shelf and
    use_mixed_melting replaces / blends the basal melt rate."""
    case = Case(96, 48, 6000, dt=86400.0, **over)
    b, o = both(case)
    warm = dict(sst=np.full_like(case.forcing["sst"], 1.5))
    for step in range(3):
        run_gpu(b, case, **warm)
        run_oracle(o, case, **warm)
        compare_state(b, o, f"ice-shelf melt {over} step {step}", rtol=1e-9)
    for fid in FLUX_FIELDS:
        assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-9
    m0 = case.bergs["mass"].sum()
    assert b.get_bergs(["mass"])["mass"].sum() < 0.999 * m0       # the basal melt did act
    api.icebergs_end(b)


@pytest.mark.parametrize("over", [dict(), dict(use_new_predictive_corrective=1, old_bug_bilin=0), dict(speed_limit=0.05)])
def test_runge_kutta_stepping(over):
    """Runge_not_Verlet=.true. (the namelist default, F:733): Runge_Kutta_stepping I:7331-7679."""
    case = Case(96, 48, 12000, runge_not_verlet=1, **over)
    b, o = both(case)
    for step in range(3):
        run_gpu(b, case)
        run_oracle(o, case)
        compare_state(b, o, f"RK4 {over} step {step}", rtol=1e-9)
    for fid in FLUX_FIELDS:
        assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-9
    cg, co = b.counters(), o.counters()
    assert cg["nspeeding_tickets"] == co["nspeeding_tickets"]
    api.icebergs_end(b)


def test_runge_kutta_bounce_and_wrap():
    case = Case(96, 48, 8000, dt=2 * 86400.0, runge_not_verlet=1)
    f = case.forcing
    fast = dict(uo=np.full_like(f["uo"], 1.2), vo=np.full_like(f["vo"], 0.15), tauxa=np.full_like(f["tauxa"], 15.0))
    b, o = both(case)
    for step in range(4):
        run_gpu(b, case, **fast)
        run_oracle(o, case, **fast)
        compare_state(b, o, f"RK4 bounce/wrap step {step}", rtol=1e-8)
    co, cg = o.counters(), b.counters()
    assert co["n_bounced"] > 20 and cg["n_bounced"] == co["n_bounced"] and co["n_received"] > 0
    api.icebergs_end(b)


class PolarGrid(S.Grid):
    """Ocean all the way to the north pole: exercises the tangent-plane stepping above 89N
    (I:7767-7816, I:8066-8099) and the polar branch of pos_within_cell (F:6359-6405)."""

    def wet(self, ring=1):
        _, latc = self.centre_lonlat(ring)
        return ((latc > 60.0) & (latc < 90.0)).astype(np.float64)


@pytest.mark.parametrize("rk", [0, 1])
def test_polar_tangent_plane(rk):
    # 0.47-degree rows: the band 89N..89.5N lies below the (degenerate) row of cells that touch the pole
    gni, gnj = 48, 384
    grid = PolarGrid(gni, gnj)
    case = Case(gni, gnj, 0, grid=grid, runge_not_verlet=rk, old_bug_bilin=0, capacity=16384)
    rng = np.random.default_rng(3)
    n = 4000
    lat = rng.uniform(88.3, 89.45, n)
    lon = rng.uniform(0.5, 359.5, n)
    base, _ = S.Grid(96, 48).seed_bergs(n)
    cols = {k: v for k, v in base.items() if k not in ("ine", "jne", "id")}
    cols["lon"], cols["lat"] = lon, lat
    cols["start_lon"], cols["start_lat"] = lon.copy(), lat.copy()
    cols["uvel"] = rng.uniform(-0.3, 0.3, n)
    cols["vvel"] = rng.uniform(0.0, 0.4, n)           # heading for the pole
    case.bergs, case.counter = cols, np.zeros((gnj, gni), dtype=np.int32)
    b, o = case.make_gpu(), case.make_oracle()
    f = case.forcing
    north = dict(vo=np.full_like(f["vo"], 0.3))
    above = 0
    for step in range(6):
        run_gpu(b, case, **north)
        run_oracle(o, case, **north)
        compare_state(b, o, f"polar rk={rk} step {step}", rtol=1e-8)
        above = max(above, int(np.sum(o.get_bergs(["lat"])["lat"] > 89.0)))
    assert above > 50, f"only {above} bergs above 89N: the tangent plane was not exercised"
    api.icebergs_end(b)


def test_running_mean_of_calving_matches_oracle_and_the_closed_form():
    """tau_calving > 0: get_running_mean_calving I:5999-6038.  The mean starts from the first field (I:6010-6017), relaxes
    with alpha = tau / (tau + dt) (tau in the units the source writes, I:6020), keeps calving after the input has stopped,
    and travels through calving.res.nc (fmsio:568-569)."""
    # I:6020 DIVIDES the namelist value by the seconds of a year; restated as written, so tau_calving = 3600 * 31 536 000
    # gives tau = 3600 and alpha = 0.5 at dt = 3600 s
    case = Case(96, 48, 300, capacity=400000, tau_calving=3600.0 * 365. * 86400.)
    f = case.forcing
    rng = np.random.default_rng(11)
    lat = case.init["ice_lat"]
    coast = (np.abs(lat) > 65) & (np.abs(lat) < 79) & (case.grid.wet(0) > 0)
    calving = np.zeros_like(f["calving"])
    calving[coast] = rng.uniform(1e-4, 2e-2, size=int(coast.sum()))
    hflx = calving * -3.0e4
    b, o = both(case)
    p = case.params()
    tau = p.tau_calving / (365. * 24 * 60 * 60)
    alpha = tau / (tau + case.dt)
    assert 0.05 < alpha < 0.95, alpha
    h = case.halo
    inner = (slice(h, -h), slice(h, -h))
    mean = None
    for step in range(6):
        cin = calving if step < 3 else 0.0 * calving          # the input stops: the mean decays, bergs keep calving
        hin = hflx if step < 3 else 0.0 * hflx
        run_gpu(b, case, time=(1, 10.0 + step), calving=cin, calving_hflx=hin)
        run_oracle(o, case, time=(1, 10.0 + step), calving=cin, calving_hflx=hin)
        assert b.count_bergs() == o.count_bergs(), f"count after step {step}"
        compare_state(b, o, f"tau_calving step {step}", rtol=1e-9)
        rg, ro = b.grid_field(D.KID_FLD_RMEAN_CALVING)[inner], o.grid_field(D.KID_FLD_RMEAN_CALVING)[inner]
        assert np.array_equal(rg, ro)
        assert np.array_equal(b.grid_field(D.KID_FLD_RMEAN_CALVING_HFLX)[inner], o.grid_field(D.KID_FLD_RMEAN_CALVING_HFLX)[inner])
        wet = case.grid.wet(0)
        mean = cin * wet if mean is None else (1.0 - alpha) * (cin * wet) + alpha * mean
        assert np.max(np.abs(rg - mean)) <= 1e-12 * np.max(np.abs(mean))
    assert b.counters()["nbergs_calved"] == o.counters()["nbergs_calved"] > 100
    n3 = b.counters()["nbergs_calved"]
    # restart: a second pair picks the means up and goes on identically
    rm, rmh = b.get_calving_rmean()
    si, sh, ic = b.get_calving_state()
    b2, o2 = both(case)
    for x in (b2, o2):
        x.set_calving_state(stored_ice=si, stored_heat=sh, iceberg_counter_grd=ic)
        x.set_calving_rmean(rm, rmh)
    run_gpu(b2, case, time=(1, 16.0), calving=0.0 * calving, calving_hflx=0.0 * hflx)
    run_gpu(b, case, time=(1, 16.0), calving=0.0 * calving, calving_hflx=0.0 * hflx)
    run_oracle(o2, case, time=(1, 16.0), calving=0.0 * calving, calving_hflx=0.0 * hflx)
    assert np.array_equal(b2.grid_field(D.KID_FLD_RMEAN_CALVING)[inner], b.grid_field(D.KID_FLD_RMEAN_CALVING)[inner])
    assert np.array_equal(b2.grid_field(D.KID_FLD_RMEAN_CALVING)[inner], o2.grid_field(D.KID_FLD_RMEAN_CALVING)[inner])
    assert b.counters()["nbergs_calved"] >= n3
    api.icebergs_end(b); api.icebergs_end(b2)


def test_time_average_weight_leaves_the_spread_fields_empty_like_the_reference():
    """time_average_weight=T: the weight is spread inside the stepping stages (I:7264, I:7395...) and
    calculate_mass_on_ocean I:4984 zeroes it before sum_up_spread_fields reads it"""
    case = Case(96, 48, 3000, add_weight_to_ocean=1, time_average_weight=1, pass_fields_to_ocean_model=0)
    b, o = both(case)
    for step in range(2):
        run_gpu(b, case); run_oracle(o, case)
        compare_state(b, o, f"time_average_weight step {step}", rtol=1e-10)
        for fid in (D.KID_FLD_SPREAD_MASS, D.KID_FLD_SPREAD_AREA):
            assert np.array_equal(b.grid_field(fid), o.grid_field(fid)) and not b.grid_field(fid).any()
    api.icebergs_end(b)
