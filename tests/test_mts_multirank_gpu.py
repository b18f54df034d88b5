"""transfer_mts_bergs (F:2136-2216): the multiple-time-step scheme on several ranks.  Every rank sub-steps a complete
copy of each conglomerate that reaches its tile plus the bergs within contact distance of it, with no communication
inside the sub-step loop; owners migrate between ranks with their MTS / DEM state and bond history.  The union of the
ranks' owned bergs must equal the SINGLE-rank CPU oracle: the reference's own MTS_KID and iKID collision tests run on
4 PEs (tests/collision_tests/README:16-22)."""
import numpy as np
import pytest

from common import COMPARE_F64
from icebergs_b200 import api
from icebergs_b200 import synthetic as S
from test_mts_gpu import IKID, MTS_KID
from test_multirank_gpu import _run_cartesian_ranks

pytestmark = pytest.mark.gpu

NAMES = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
# The accelerations of a conglomerate at rest in the flow are differences of large bond forces (noise around zero before
# the contact) and, with force_convergence, small differences of converged velocities: the velocities are compared.
F64_MTS = tuple(k for k in COMPARE_F64 if k not in ("axn", "ayn", "bxn", "byn"))


def _params(dem, **over):
    kw = dict(MTS_KID)
    if dem:
        kw.update(IKID)
    kw.update(over)
    return lambda: S.collision_params(api.default_params, **kw)


@pytest.mark.parametrize("dem", [0, 1])
@pytest.mark.parametrize("nranks,ibuo,ibvo", [(2, 0.05, 0.5), (4, 0.5, 0.45)])
def test_conglomerate_migrates_across_ranks(nranks, ibuo, ibvo, dem):
    """One bonded 8-element conglomerate (input_MTS_KID.nml / input_iKID.nml physics) drifting across the rank edges
    (2 ranks: north across y = 10 km; 4 ranks: north-east across the corner of the four tiles): its elements migrate
    one by one with their MTS state (environment cache, fast accelerations), with dem also angular velocity, rotation
    and the bond history; while it straddles the edge every rank involved steps a complete copy of it.  No contact, so
    no chaos: tight tolerance for all 540 steps."""
    south = {k: v[:8].copy() for k, v in S.collision_bergs().items()}
    f64 = F64_MTS + (("ang_vel", "rot") if dem else ())
    names = list(f64) + ["ine", "jne", "start_year", "id"]
    moved, o = _run_cartesian_ranks(nranks, _params(dem), south, lambda g: g.forcing(ibuo=ibuo, ibvo=ibvo, collision_test=False),
                                    540, 20, names, f64, bonds=True, rtol=1e-7)
    assert moved >= 8 and o.count_bergs() == 8
    b = o.get_bergs(["lat", "lon"])
    assert b["lat"].min() > 10.0e3, "the conglomerate should have crossed the edge completely"
    if nranks == 4:
        assert b["lon"].min() > 10.0e3


@pytest.mark.parametrize("dem", [0, 1])
@pytest.mark.parametrize("nranks", [2, 4])
def test_collision_test_across_ranks(nranks, dem):
    """input_MTS_KID.nml / input_iKID.nml on 2 ranks (tiles split at y = 10 km: the two conglomerates meet ON the rank
    edge, each rank holds its own conglomerate and a full copy of the other) and on 4 ranks (2 x 2, as the reference runs
    them; the conglomerates also cross x = 10 km during the collision).  Before the contact (~430 steps) the ranks
    reproduce the single-rank oracle to 1e-8.  The collision amplifies rounding-level differences (summation order of
    the contact forces, per-PE convergence norms I:6726) about tenfold per 100 steps, as it does between the CUDA
    path and the oracle on one rank: afterwards positions are compared to 5e-5."""
    late = lambda step: step >= 420
    moved, o = _run_cartesian_ranks(nranks, _params(dem), S.collision_bergs(), lambda g: g.forcing(), 700, 70, NAMES, F64_MTS, bonds=True,
                                    rtol_of=lambda s: (5e-5, ("lon", "lat", "mass", "thickness")) if late(s) else (1e-8, F64_MTS))
    assert o.count_bergs() == 16
    g = o.get_bergs(["lat"])
    half = g["lat"] < 10.0e3
    assert half.sum() == 8, "the conglomerates should have bounced back to their own sides"


# ---------------------------------------------------------------------------------------------- through the cyclic seam
def _seam_case(dem, nranks_x):
    """A conglomerate that drifts east through the cyclic seam of the 20 km domain (Lx = 20 km): on the GPU one rank
    that is its own E/W neighbour, or two ranks side by side.  The single-rank oracle makes no copies through the seam
    (oracle/kid_oracle_mts.inc), so it runs the same physics on a 40 km domain whose forcing repeats the 20 km one: its
    conglomerate simply keeps going east, and positions agree modulo 20 km."""
    import kid_oracle_py as O
    from icebergs_b200 import _cdefs as D
    from icebergs_b200 import parallel
    from test_multirank_gpu import _custom_domains
    from common import assert_bergs_match
    dt, L = 60.0, 20.0e3
    over = dict(MTS_KID)
    if dem:
        over.update(IKID)
    params = lambda Lx: S.collision_params(api.default_params, Lx=Lx, **over)
    bergs = {k: v[:8].copy() for k, v in S.collision_bergs().items()}
    bergs["lon"] = bergs["lon"] + 11.5e3; bergs["start_lon"] = bergs["lon"].copy()        # x = 15 .. 17.4 km
    g20, g40 = S.CartesianGrid(20, 20), S.CartesianGrid(40, 20)
    f20 = g20.forcing(ibuo=0.6, ibvo=0.02, collision_test=False)
    f40 = g40.forcing(ibuo=0.6, ibvo=0.02, collision_test=False)
    o = O.Oracle(40, 20, dt, (1, 0.0), params=params(2 * L), domain=api.Domain.single(40, 20, halo=3, cyclic_x=True), **g40.init_args())
    o.set_bergs(**bergs); o.set_bonds()
    ref0 = o.get_bergs(["id", "ine", "jne", "lon", "lat"])
    n = 8
    cols = dict(bergs, id=np.zeros(n, dtype=np.int64), ine=np.zeros(n, dtype=np.int32), jne=np.zeros(n, dtype=np.int32))
    for k in range(n):
        q = int(np.argmin((ref0["lon"] - cols["lon"][k]) ** 2 + (ref0["lat"] - cols["lat"][k]) ** 2))
        cols["id"][k], cols["ine"][k], cols["jne"][k] = ref0["id"][q], ref0["ine"][q], ref0["jne"][q]
    grp = parallel.LocalGroup(nranks_x)
    if nranks_x == 1:
        doms = [api.Domain.single(20, 20, halo=3, cyclic_x=True)]
    else:
        doms = _custom_domains(20, 20, [10, 10], [20], 3, grp)
    grids = [S.CartesianGrid(20, 20, 1.0e3, d.isc, d.iec, d.jsc, d.jec) for d in doms]
    hs = [None] * nranks_x

    def init(r):
        d = doms[r]
        hs[r] = api.icebergs_init(20, 20, dt, (1, 0.0), params=params(L), domain=d, capacity=4096, **grids[r].init_args())
        mine = (cols["ine"] >= d.isc) & (cols["ine"] <= d.iec)
        hs[r].set_bergs(**{k: np.ascontiguousarray(v[mine]) for k, v in cols.items()})
        hs[r].set_bonds()
    grp.run(init)
    f64 = ("lat", "uvel", "vvel", "mass", "thickness", "width", "length") + (("ang_vel", "rot") if dem else ())
    names = list(f64) + ["lon", "ine", "jne", "id", "start_year"]
    wrapped = 0
    for step in range(420):
        def one(r):
            f = grids[r].forcing(ibuo=0.6, ibvo=0.02, collision_test=False)
            c, h = f["calving"].copy(), f["calving_hflx"].copy()
            api.icebergs_run(hs[r], (1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h,
                             f["cn"], f["hi"], sss=f["sss"])
        grp.run(one)
        c, h = f40["calving"].copy(), f40["calving_hflx"].copy()
        o.run((1, 0.0), c, f40["uo"], f40["vo"], f40["ui"], f40["vi"], f40["tauxa"], f40["tauya"], f40["ssh"], f40["sst"], h, f40["cn"],
              f40["hi"], sss=f40["sss"])
        if step % 20 == 19:
            got = [b.get_bergs(names) for b in hs]
            got = {k: np.concatenate([p[k] for p in got]) for k in names}
            want = o.get_bergs(names)
            assert len(got["id"]) == 8 == len(want["id"]), f"step {step}: berg count"
            ga, wa = np.argsort(got["id"]), np.argsort(want["id"])
            dl = np.mod(got["lon"][ga] - want["lon"][wa] + 0.5 * L, L) - 0.5 * L
            assert np.abs(dl).max() < 1e-7 * L, f"step {step}: lon differs by {np.abs(dl).max():.3e} m (mod Lx)"
            assert np.array_equal(np.mod(got["ine"][ga] - 1, 20), np.mod(want["ine"][wa] - 1, 20)), f"step {step}: ine (mod 20)"
            wrapped = int((want["lon"] > L).sum())
            want = dict(want, ine=got["ine"][ga][np.argsort(wa)], lon=want["lon"])
            assert_bergs_match(got, want, rtol=1e-7, names=f64, context=f"seam, {nranks_x} rank(s) along x, step {step}", acc_floor=1e-13)
    for b in hs:
        api.icebergs_end(b)
    grp.close()
    o.close()
    return wrapped


@pytest.mark.parametrize("dem", [0, 1])
@pytest.mark.parametrize("nranks_x", [1, 2])
def test_conglomerate_through_the_cyclic_seam(nranks_x, dem):
    assert _seam_case(dem, nranks_x) == 8, "all eight elements should have passed x = 20 km"
