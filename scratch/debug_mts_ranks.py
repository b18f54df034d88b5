import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import COMPARE_F64, assert_bergs_match
from icebergs_b200 import api, parallel, synthetic as S
from test_mts_gpu import MTS_KID, IKID
def _run_cartesian_ranks(nranks, params, bergs, forcing_of, nsteps, check_every, names, f64, bonds, dt=60.0, rtol=1e-7,
                         yearday_of=lambda k, dt: 0.0):
    import kid_oracle_py as O
    g0 = S.CartesianGrid()
    dom0 = api.Domain.single(20, 20, halo=3, cyclic_x=True)
    o = O.Oracle(20, 20, dt, (1, 0.0), params=params(), domain=dom0, **g0.init_args())
    o.set_bergs(**bergs)
    if bonds:
        o.set_bonds()
    ref0 = o.get_bergs(["id", "ine", "jne", "lon", "lat"])
    n = len(bergs["lon"])
    cols = dict(bergs, id=np.zeros(n, dtype=np.int64), ine=np.zeros(n, dtype=np.int32), jne=np.zeros(n, dtype=np.int32))
    order = {(float(x), float(y)): k for k, (x, y) in enumerate(zip(ref0["lon"], ref0["lat"]))}
    for k in range(n):
        q = order[(float(cols["lon"][k]), float(cols["lat"][k]))]
        cols["id"][k], cols["ine"][k], cols["jne"][k] = ref0["id"][q], ref0["ine"][q], ref0["jne"][q]
    grp = parallel.LocalGroup(nranks)
    doms = [grp.domain(20, 20, r, halo=3) for r in range(nranks)]
    grids = [S.CartesianGrid(20, 20, 1.0e3, d.isc, d.iec, d.jsc, d.jec) for d in doms]
    parts = parallel.split_by_owner(cols, doms)
    hs = [None] * nranks
    _, _, ic0 = o.get_calving_state()

    def init(r):
        d = doms[r]
        hs[r] = api.icebergs_init(20, 20, dt, (1, 0.0), params=params(), domain=d, capacity=8192, **grids[r].init_args())
        cnt = np.zeros((d.njd, d.nid), dtype=np.int32)
        cnt[3:3 + d.njc, 3:3 + d.nic] = ic0[3 + d.jsc - 1:3 + d.jec, 3 + d.isc - 1:3 + d.iec]
        hs[r].set_calving_state(iceberg_counter_grd=cnt)      # the per-cell id counters follow the file-order ids
        hs[r].set_bergs(**parts[r])
        if bonds:
            hs[r].set_bonds()
    grp.run(init)
    f0 = forcing_of(g0)
    moved = 0
    for step in range(nsteps):
        t = (1, yearday_of(step, dt))

        def one(r):
            f = forcing_of(grids[r])
            c, h = f["calving"].copy(), f["calving_hflx"].copy()
            api.icebergs_run(hs[r], t, c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h,
                             f["cn"], f["hi"], sss=f["sss"])
        grp.run(one)
        c, h = f0["calving"].copy(), f0["calving_hflx"].copy()
        o.run(t, c, f0["uo"], f0["vo"], f0["ui"], f0["vi"], f0["tauxa"], f0["tauya"], f0["ssh"], f0["sst"], h, f0["cn"], f0["hi"],
              sss=f0["sss"])
        moved += sum(b.counters()["n_sent"] for b in hs)
        if step % check_every == check_every - 1:
            got = [b.get_bergs(names) for b in hs]
            got = {k: np.concatenate([p[k] for p in got]) for k in names}
            w = assert_bergs_match(got, o.get_bergs(names), rtol=1.0, names=f64, context=f"{nranks} ranks, step {step}", acc_floor=1e-13)
            print(nranks, "ranks step", step, "moved", moved, {k: "%.1e" % v for k, v in w.items() if v > 1e-9}, flush=True)
    for b in hs:
        api.icebergs_end(b)
    grp.close()
    return moved, o



NAMES = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
for nr in (2, 4):
    for dem in (0, 1):
        kw = dict(MTS_KID, convergence_tolerance=float(sys.argv[1]) if len(sys.argv) > 1 else 1e-8)
        if dem: kw.update(IKID)
        print("=== nranks", nr, "dem", dem, flush=True)
        try:
            _run_cartesian_ranks(nr, lambda: S.collision_params(api.default_params, **kw), S.collision_bergs(), lambda g: g.forcing(), 700, 50, NAMES, COMPARE_F64, bonds=True)
        except Exception as e:
            print("FAILED", str(e)[:300])
