"""Namelist options of the hot path restated in the CPU oracle, checked against things that do not depend on the oracle's
own arithmetic: closed forms and conservation.  (CPU only; the CUDA library is compared with the oracle on the same options
in tests/test_parity_gpu.py, tests/test_spreading_gpu.py, tests/test_mts_gpu.py.)"""
import numpy as np
import pytest

import kid_oracle_py as O
from icebergs_b200 import _cdefs as D
from icebergs_b200 import synthetic as S

GNI, GNJ, H = 96, 48, 4


def make(n=0, dt=3600.0, **over):
    g = S.Grid(GNI, GNJ)
    kw = dict(runge_not_verlet=0, tau_is_velocity=1, Rearth=S.REARTH, add_weight_to_ocean=0, bergy_bit_erosion_fraction=0.1)
    kw.update(over)
    p = O.default_params(**kw)
    d = O.SingleDomain(GNI, GNJ, halo=H)
    o = O.Oracle(GNI, GNJ, dt, (1, 0.0), params=p, domain=d, **g.init_args())
    if n:
        b, counter = g.seed_bergs(n)
        c = np.zeros((d.njd, d.nid), dtype=np.int32)
        c[H:H + GNJ, H:H + GNI] = counter
        o.set_calving_state(iceberg_counter_grd=c)
        o.set_bergs(**b)
    return g, o


def run(o, g, t=(1, 0.0), **replace):
    f = dict(g.forcing())
    f.update(replace)
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    o.run(t, c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
    return c, h


def test_running_mean_of_calving_follows_the_closed_form():
    """get_running_mean_calving I:5999-6038: rmean starts from the first field (I:6010-6017), then
    rmean <- beta*calving + alpha*rmean with alpha = tau/(tau+dt), tau = tau_calving/(365*86400) as written at I:6020"""
    tau_nml = 7200.0 * 365.0 * 86400.0                       # tau = 7200 -> alpha = 2/3 at dt = 3600 s
    g, o = make(tau_calving=tau_nml)
    alpha = 7200.0 / (7200.0 + 3600.0)
    lat = g.init_args()["ice_lat"]
    coast = (np.abs(lat) > 65) & (np.abs(lat) < 79) & (g.wet(0) > 0)
    calving = np.zeros_like(lat)
    calving[coast] = np.random.default_rng(3).uniform(1e-4, 2e-2, int(coast.sum()))
    mean, mean_h = None, None
    for step in range(5):
        cin = calving if step < 3 else 0.0 * calving
        run(o, g, t=(1, 5.0 + step), calving=cin, calving_hflx=-3.0e4 * cin)
        want = cin * g.wet(0)
        mean = want if mean is None else (1.0 - alpha) * want + alpha * mean
        mean_h = -3.0e4 * want if mean_h is None else (1.0 - alpha) * (-3.0e4 * want) + alpha * mean_h
        got = o.grid_field(D.KID_FLD_RMEAN_CALVING)[H:-H, H:-H]
        assert np.abs(got - mean).max() <= 1e-14 * np.abs(mean).max()
        got_h = o.grid_field(D.KID_FLD_RMEAN_CALVING_HFLX)[H:-H, H:-H]
        assert np.abs(got_h - mean_h).max() <= 1e-14 * np.abs(mean_h).max()
    assert o.counters()["nbergs_calved"] > 0
    n3 = o.counters()["nbergs_calved"]
    run(o, g, t=(1, 10.0), calving=0.0 * calving, calving_hflx=0.0 * calving)
    assert o.counters()["nbergs_calved"] > n3             # the mean keeps calving after the input has stopped
    o.close()


@pytest.mark.parametrize("without_decay", [0, 1])
def test_melt_from_the_spread_mass_equals_the_melt_of_the_bergs(without_decay):
    """find_melt_using_spread_mass (I:5490-5500, I:3436-3448): the grid-integrated melt found from the spread mass before and
    after thermodynamics equals what the bergs lost, summed berg by berg in the default path -- two routes through the
    restated spreading (spread_mass_across_ocean_cells, sum_up_spread_fields) and thermodynamics that must agree"""
    tot = {}
    for fm in (0, 1):
        g, o = make(n=3000, add_weight_to_ocean=1, find_melt_using_spread_mass=fm, iceberg_melt_without_decay=without_decay)
        for _ in range(2):
            run(o, g)
        area = g.init_args()["ice_area"]
        tot[fm] = float((o.grid_field(D.KID_FLD_FLOATING_MELT)[H:-H, H:-H] * area).sum())
        if fm:
            fmelt = o.grid_field(D.KID_FLD_FLOATING_MELT)
            assert np.array_equal(o.grid_field(D.KID_FLD_CALVING_HFLX)[H:-H, H:-H], (fmelt * o.params.hlf)[H:-H, H:-H])
        o.close()
    assert tot[0] > 0 and abs(tot[1] / tot[0] - 1.0) < 1e-11, tot


def test_runge_kutta_with_mts_switches_to_verlet_and_with_footloose_is_fatal():
    """F:1303-1306 (warning, Runge_not_Verlet set to false under mts) and F:1485-1488 (FATAL with MTS, DEM or footloose)"""
    g = S.CartesianGrid()
    dom = O.SingleDomain(g.gni, g.gnj, halo=3)
    p = S.collision_params(O.default_params, mts=1, mts_sub_steps=60, explicit_inner_mts=1, contact_distance=1.75e3,
                           contact_spring_coef=1.0e-7, runge_not_verlet=1)
    o = O.Oracle(g.gni, g.gnj, 60.0, (1, 0.0), params=p, domain=dom, **g.init_args())       # accepted
    o.close()
    with pytest.raises(O.OracleFatal, match="Runge_not_Verlet must be false"):
        O.Oracle(g.gni, g.gnj, 10.0, (1, 0.0), params=S.footloose_params(O.default_params, runge_not_verlet=1),
                 domain=O.SingleDomain(g.gni, g.gnj, halo=3), **g.init_args())


def test_time_average_weight_leaves_the_spread_fields_empty():
    """I:7264 / I:7395...: the stages spread the weight, calculate_mass_on_ocean I:4984 zeroes it before it is read"""
    g, o = make(n=2000, add_weight_to_ocean=1, time_average_weight=1)
    run(o, g)
    assert not o.grid_field(D.KID_FLD_SPREAD_MASS).any()
    o.close()
    g, o = make(n=2000, add_weight_to_ocean=1, time_average_weight=0)
    run(o, g)
    assert o.grid_field(D.KID_FLD_SPREAD_MASS).max() > 0
    o.close()
