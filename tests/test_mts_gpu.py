"""The multiple-time-step scheme: evolve_icebergs_mts I:6576, accel_mts I:1278,
accel_explicit_inner_mts I:1710 -- the CUDA path against the CPU oracle on the two 8-element conglomerates of
tests/collision_tests with the switches of input_MTS_KID.nml (mts, 60 sub-steps, explicit inner steps,
force_convergence, contact_distance 1.75 km, contact_spring_coef 1e-7)."""
import numpy as np
import pytest

from common import COMPARE_F64, by_id
from icebergs_b200 import api
from icebergs_b200 import synthetic as S
from test_interactions_gpu import Pair, bond_set

pytestmark = pytest.mark.gpu

MTS_KID = dict(mts=1, mts_sub_steps=60, explicit_inner_mts=1, force_convergence=1, convergence_tolerance=1e-8,
               contact_distance=1.75e3, contact_spring_coef=1.0e-7)


def mts_pair(**over):
    kw = dict(MTS_KID)
    kw.update(over)
    return Pair(S.collision_bergs(), lambda: S.collision_params(api.default_params, **kw))


def test_mts_kid_first_steps():
    p = mts_pair()
    assert p.b.count_bergs() == 16 == p.o.count_bergs()              # README:16-22 '#= 16' for MTS_KID too
    assert bond_set(p.b.get_bonds()) == bond_set(p.o.get_bonds())
    p.step(1)
    p.check("one step", rtol=1e-10)
    g, o = by_id(p.b.get_bergs(["id", "axn_fast", "ayn_fast", "bxn_fast", "byn_fast"])), by_id(p.o.get_bergs(["id", "axn_fast", "ayn_fast", "bxn_fast", "byn_fast"]))
    for k in ("axn_fast", "ayn_fast", "bxn_fast", "byn_fast"):
        scale = max(np.abs(o[k]).max(), 1e-13)
        assert np.abs(g[k] - o[k]).max() <= 1e-9 * scale + 1e-14, k      # at rest the bond accelerations are ~1e-16 noise
    p.step(49)
    p.check("50 steps", rtol=1e-9)
    names = ["id", "axn_fast", "ayn_fast", "uo", "vo", "ssh_x", "od"]      # the environment cache of I:4673 too
    g, o = by_id(p.b.get_bergs(names)), by_id(p.o.get_bergs(names))
    assert np.abs(o["axn_fast"]).max() > 0 and np.abs(o["vo"]).max() > 0
    for k in names[1:]:
        scale = max(np.abs(o[k]).max(), 1e-13)
        assert np.abs(g[k] - o[k]).max() <= 1e-8 * scale + 1e-14, k
    p.end()


def test_mts_kid_through_contact():
    """The conglomerates meet after ~550 steps: the long-step collision force with its convergence passes, the
    bonded sub-steps (60 per step) and the contact search inside a conglomerate all act; the berg count stays 16."""
    p = mts_pair()
    gaps, passes = [], 0
    for k in range(14):
        p.step(50)
        p.check(f"{50 * (k + 1)} steps", rtol=1e-6)
        g = by_id(p.b.get_bergs(["id", "lat"]))
        half = g["lat"] < 10.0e3
        gaps.append(float(g["lat"][~half].min() - g["lat"][half].max()))
    assert min(gaps) < 1100.0 and gaps[-1] > min(gaps), f"no bounce: {gaps}"
    assert p.b.count_bergs() == 16 == p.o.count_bergs()
    p.end()


@pytest.mark.parametrize("fc", [0, 1])
def test_mts_implicit_inner_steps(fc):
    """explicit_inner_mts=.false.: the sub-steps solve accel_mts with only_interactive_forces (I:6922), with
    force_convergence they repeat until the velocity norm settles (I:6936-6968)."""
    p = mts_pair(explicit_inner_mts=0, force_convergence=fc, mts_sub_steps=20)
    p.step(1)
    p.check("one step", rtol=1e-10)
    p.step(29)
    p.check("30 steps", rtol=1e-8)
    p.end()


def test_mts_refuses_what_is_not_built():
    with pytest.raises(api.KidFatal):
        mts_pair(dem=1, iceberg_bonds_on=0, manually_initialize_bonds=0, max_bonds=0)


def test_mts_with_runge_kutta_switches_to_verlet_like_the_reference():
    """F:1303-1306: 'Multiple time stepping does not work with Runge Kutta stepping, switching to Verlet.' (a warning):
    the namelist default Runge_not_Verlet=T with mts=T runs, and runs the Verlet scheme"""
    a = mts_pair(runge_not_verlet=1)
    a.step(3)
    a.check("mts with Runge_not_Verlet=T, 3 steps", rtol=1e-8)
    b = mts_pair(runge_not_verlet=0)
    b.step(3)
    ga, gb = a.b.get_bergs(["id", "lon", "lat", "uvel"]), b.b.get_bergs(["id", "lon", "lat", "uvel"])
    oa, ob = np.argsort(ga["id"]), np.argsort(gb["id"])
    for k in ("lon", "lat", "uvel"):
        assert np.array_equal(ga[k][oa], gb[k][ob]), k
    a.end(); b.end()


# ---------------------------------------------------------------------------------------------- DEM (a15, a17)
IKID = dict(dem=1, poisson=0.3, dem_damping_coef=1.0, dem_spring_coef=4471.94)
# absolute floors: before the collision the tangential displacements (~1e-14 m on 780 m bonds), relative rotations and
# stresses are rounding noise around zero
BOND_F64 = dict(length=0.0, tangd1=1e-9, tangd2=1e-9, nstress=1e-8, sstress=1e-8, rel_rotation=1e-10)


def check_bonds(p, context, rtol):
    g, o = p.b.get_bonds(), p.o.get_bonds()
    kg = np.lexsort((g["other_id"], g["first_id"])); ko = np.lexsort((o["other_id"], o["first_id"]))
    assert np.array_equal(g["first_id"][kg], o["first_id"][ko]) and np.array_equal(g["other_id"][kg], o["other_id"][ko]), context
    assert np.array_equal(g["broken"][kg], o["broken"][ko]), f"{context}: broken flags"
    if len(o["first_id"]) == 0:
        return
    for k, floor in BOND_F64.items():
        a, b = g[k][kg], o[k][ko]
        scale = max(np.abs(b).max(), 1e-300)
        assert np.abs(a - b).max() <= rtol * scale + floor, f"{context}: bond {k} differs by {np.abs(a - b).max():.2e} (scale {scale:.2e})"


def test_ikid_dem_first_steps():
    """input_iKID.nml: the bonded elements interact through calculate_force_dem (normal + shear springs, torque from
    relative rotation, damping; Wang 2020), the pair forces are evaluated once and stored on both half-bonds."""
    p = mts_pair(**IKID)
    assert p.b.count_bergs() == 16 == p.o.count_bergs()              # README:16-22 '#= 16' for iKID
    p.step(1)
    p.check("one step", rtol=1e-10)
    check_bonds(p, "one step", 1e-9)
    names = ["id", "ang_vel", "ang_accel", "rot"]
    g, o = by_id(p.b.get_bergs(names)), by_id(p.o.get_bergs(names))
    for k in names[1:]:
        scale = max(np.abs(o[k]).max(), 1e-300)
        assert np.abs(g[k] - o[k]).max() <= 1e-8 * scale + 1e-18, k
    p.step(49)
    p.check("50 steps", rtol=1e-9)
    check_bonds(p, "50 steps", 1e-7)
    p.end()


def test_ikid_dem_through_contact():
    p = mts_pair(**IKID)
    gaps = []
    for k in range(14):
        p.step(50)
        p.check(f"{50 * (k + 1)} steps", rtol=1e-6)
        g = by_id(p.b.get_bergs(["id", "lat"]))
        half = g["lat"] < 10.0e3
        gaps.append(float(g["lat"][~half].min() - g["lat"][half].max()))
    assert min(gaps) < 1100.0 and gaps[-1] > min(gaps), f"no bounce: {gaps}"
    assert p.b.count_bergs() == 16 == p.o.count_bergs()
    check_bonds(p, "700 steps", 1e-5)
    p.end()


def test_dem_bonds_break_under_stress():
    """fracture_criterion='stress' with thresholds the collision exceeds: bonds break on the sub-steps
    (calculate_force_dem I:1137-1198) or after the long step (break_bonds_dem F:4713), conglomerates split."""
    for on_sub in (1, 0):
        p = mts_pair(**IKID, fracture_criterion_stress=1, frac_thres_n=60.0, frac_thres_t=12.0, break_bonds_on_sub_steps=on_sub)
        n0 = len(p.o.get_bonds()["first_id"])
        for k in range(14):
            p.step(50)
            p.check(f"break_on_sub={on_sub}, {50 * (k + 1)} steps", rtol=1e-6)
            check_bonds(p, f"break_on_sub={on_sub}, {50 * (k + 1)} steps", 1e-5)
        o = p.o.get_bonds()
        assert len(o["first_id"]) < n0 or o["broken"].any(), "no bond broke: thresholds too high for this case"
        p.end()


# ------------------------------------------------------------------------------ the reference's DEM beam tests
class BeamPair(Pair):
    def __init__(self, bergs, dt, **over):
        g = S.CartesianGrid(20, 20, 15000.0)
        super().__init__(bergs, lambda: S.beam_params(api.default_params, **over), grid=g, dt=dt,
                         forcing=g.forcing(ibuo=0.0, ibvo=0.0, collision_test=False))


CBEAM = dict(dem_beam_test=2, orig_dem_moment_of_inertia=1, dem_damping_coef=0.7, rho_bergs=900.0, mts_sub_steps=2000)


def test_cantilever_beam_matches_oracle():
    """dem_cbeam_test: static (clamped) elements, three rows, large deflection; 3 steps x 2000 sub-steps."""
    p = BeamPair(S.cantilever_bergs(), 100.0, **CBEAM)
    assert p.b.count_bergs() == 90 == p.o.count_bergs()
    assert bond_set(p.b.get_bonds()) == bond_set(p.o.get_bonds()) and len(p.o.get_bonds()["first_id"]) == 294
    for k in range(3):
        p.step(1)
        # accelerations are differences of ~1e10 N spring forces in a stiff, still ringing lattice (|v| ~ 10 m/s):
        # rounding differences grow over the 2000 sub-steps of a step
        if k == 0:
            p.check("step 1", rtol=1e-8)
        else:
            from common import assert_bergs_match
            from test_interactions_gpu import NAMES
            assert_bergs_match(p.b.get_bergs(NAMES), p.o.get_bergs(NAMES), rtol=1e-6, names=("lon", "lat", "uvel", "vvel", "xi", "yj"),
                               context=f"step {k + 1}")
    names = ["id", "ang_vel", "rot"]
    g, o = by_id(p.b.get_bergs(names)), by_id(p.o.get_bergs(names))
    for k in names[1:]:
        assert np.abs(g[k] - o[k]).max() <= 1e-5 * np.abs(o[k]).max(), k
    check_bonds(p, "3 steps", 1e-5)
    p.end()


def test_simply_supported_beam_known_answer_on_the_gpu():
    """dem_ssbeam_test through the CUDA path alone: 8 steps of 1e5 sub-steps (one kernel per step, the sub-step loop
    runs inside a single CTA), mid-span deflection against P l^3 / (48 E I) as in tests/test_oracle_golden.py."""
    b0 = S.beam_bergs()
    g = S.CartesianGrid(20, 20, 15000.0)
    dom = api.Domain.single(20, 20, halo=3, cyclic_x=True)
    h = api.icebergs_init(20, 20, 1.0, (1, 0.0), params=S.beam_params(api.default_params), domain=dom, capacity=1024, **g.init_args())
    h.set_bergs(**b0)
    h.set_bonds()
    f = g.forcing(ibuo=0.0, ibvo=0.0, collision_test=False)
    for k in range(8):
        c, hf = f["calving"].copy(), f["calving_hflx"].copy()
        api.icebergs_run(h, (1, k / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hf,
                         f["cn"], f["hi"], sss=f["sss"])
    b = h.get_bergs(["lon", "lat", "vvel", "start_lon"])
    order = np.argsort(b["start_lon"])
    d = b["lat"][order] - b0["lat"]
    l = 14.0
    w_mid = -1.5e5 * l ** 3 / (48.0 * 1.0e9 * (0.5 ** 3 / 12.0))
    assert abs(d[14] / w_mid - 1.0) < 0.03, (d[14], w_mid)
    assert np.abs(d[[0, 28]]).max() < 1e-6                                  # the supports do not move vertically
    api.icebergs_end(h)


def test_dem_ground_frac_switches():
    """The remaining switches of tests/dem_ground_frac_test/input.nml on the collision scenario:
    use_broken_bonds_for_substep_contact (broken bonds stay as contact pairs, I:1751, F:2692),
    constant_interaction_LW with the mean element size (F:4640), short_step_mts_grounding with partial grounding
    (I:6873-6905), hexagonal elements, fracture on the sub-steps."""
    over = dict(IKID, use_broken_bonds_for_substep_contact=1, constant_interaction_LW=1, short_step_mts_grounding=1,
                break_bonds_on_sub_steps=1, fracture_criterion_stress=1, frac_thres_n=25.0, frac_thres_t=5.0,
                cdrag_grounding=100.0, h_to_init_grounding=1000.0, mts_sub_steps=200, convergence_tolerance=1e-2)
    p = mts_pair(**over)
    n0 = len(p.o.get_bonds()["first_id"])
    for k in range(18):
        p.step(50)
        p.check(f"{50 * (k + 1)} steps", rtol=1e-6)
        check_bonds(p, f"{50 * (k + 1)} steps", 1e-5)
    o = p.o.get_bonds()
    assert o["broken"].any() or len(o["first_id"]) < n0, "no bond broke"
    assert p.b.count_bergs() == 16 == p.o.count_bergs()
    p.end()


@pytest.mark.parametrize("dem", [0, 1])
def test_reference_time_steps_of_the_collision_namelists(dem):
    """input_MTS_KID.nml / input_iKID.nml as shipped: ibdt = 3600 s with 60 sub-steps of 60 s, 20 h (the conglomerates
    meet after ~9 h; beyond ~24 h they reach the cyclic seam, which the one-rank MTS path does not cross)."""
    over = dict(MTS_KID)
    if dem:
        over.update(IKID)
    p = Pair(S.collision_bergs(), lambda: S.collision_params(api.default_params, **over), dt=3600.0)
    for k in range(4):
        p.step(5)
        p.check(f"{5 * (k + 1)} h", rtol=1e-7)
    assert p.b.count_bergs() == 16 == p.o.count_bergs()
    if dem:
        check_bonds(p, "20 h", 1e-6)
    p.end()


# ------------------------------------------------------------------ BASELINE configs[2]: a68_test scaled
def test_bonded_tabular_berg_a68_physics():
    """A square-packed bonded tabular berg (12 x 16 elements of radius 1.5 km) with the namelist of
    tests/a68_test/long_run.nml -- dem, 20 sub-steps, skip_first_outer_mts_step, constant_interaction_LW,
    contact_distance 4 km, stress fracture on the sub-steps, grounding drag on the short steps -- in a sheared current
    that carries one corner onto a shoal, where it grounds and pivots: positions, velocities, rotation and bond state
    against the CPU oracle."""
    from test_interactions_gpu import Pair
    g = S.TabularGrid()
    params = lambda: S.a68_params(api.default_params)
    p = Pair(S.tabular_berg(), params, grid=g, dt=1800.0, capacity=4096, forcing=g.forcing())
    n = 12 * 16
    assert p.b.count_bergs() == n == p.o.count_bergs()
    gb, ob = p.b.get_bonds(), p.o.get_bonds()
    assert bond_set(gb) == bond_set(ob) and len(gb["first_id"]) == 2 * (11 * 16 + 12 * 15)
    names = ["id", "ang_vel", "rot"]
    for k in range(6):
        p.step(16)
        p.check(f"{16 * (k + 1)} steps", rtol=1e-5)
        check_bonds(p, f"{16 * (k + 1)} steps", 1e-6)
        a, b = by_id(p.b.get_bergs(names)), by_id(p.o.get_bergs(names))
        for q in names[1:]:
            scale = max(np.abs(b[q]).max(), 1e-300)
            assert np.abs(a[q] - b[q]).max() <= 1e-6 * scale + 1e-16, q
    d = by_id(p.b.get_bergs(["id", "lon", "uvel", "rot"]))
    assert d["lon"].max() > 105.0e3 and np.abs(d["rot"]).max() > 0.05, "the berg should have run onto the shoal and started to pivot"
    p.end()
