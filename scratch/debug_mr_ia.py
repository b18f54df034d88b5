import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import kid_oracle_py as O
from common import COMPARE_F64, by_id, rel_err
from icebergs_b200 import api, parallel, synthetic as S
nranks = 2
rng = np.random.default_rng(5); n = 600
base = S.collision_bergs()
cols = {k: np.resize(v, n).copy() for k, v in base.items()}
def away(m):
    x = rng.uniform(0.0, 1.0, m)
    return np.where(x < 0.5, 1000.0 + x * 2.0 * 8000.0, 11000.0 + (x - 0.5) * 2.0 * 8000.0)
cols["lon"], cols["lat"] = away(n), away(n)
cols["start_lon"], cols["start_lat"] = cols["lon"].copy(), cols["lat"].copy()
cols["start_day"] = rng.uniform(0.0, 300.0, n)
params = lambda: S.collision_params(api.default_params, iceberg_bonds_on=0, manually_initialize_bonds=0, max_bonds=0)
g0 = S.CartesianGrid(); dom0 = api.Domain.single(20, 20, halo=3, cyclic_x=True)
o = O.Oracle(20, 20, 60.0, (1, 0.0), params=params(), domain=dom0, **g0.init_args()); o.set_bergs(**cols)
ref0 = o.get_bergs(["id", "ine", "jne", "lon"])
cols = dict(cols, id=np.zeros(n, dtype=np.int64), ine=np.zeros(n, dtype=np.int32), jne=np.zeros(n, dtype=np.int32))
order = {float(x): k for k, x in enumerate(ref0["lon"])}
for k in range(n):
    q = order[float(cols["lon"][k])]
    cols["id"][k], cols["ine"][k], cols["jne"][k] = ref0["id"][q], ref0["ine"][q], ref0["jne"][q]
grp = parallel.LocalGroup(nranks)
doms = [grp.domain(20, 20, r, halo=3) for r in range(nranks)]
grids = [S.CartesianGrid(20, 20, 1.0e3, d.isc, d.iec, d.jsc, d.jec) for d in doms]
parts = parallel.split_by_owner(cols, doms)
hs = [None] * nranks
def init(r):
    hs[r] = api.icebergs_init(20, 20, 60.0, (1, 0.0), params=params(), domain=doms[r], capacity=8192, **grids[r].init_args())
    hs[r].set_bergs(**parts[r])
grp.run(init)
f0 = g0.forcing(ibuo=0.6, ibvo=0.1, collision_test=False)
def one(r):
    f = grids[r].forcing(ibuo=0.6, ibvo=0.1, collision_test=False)
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    api.icebergs_run(hs[r], (1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
names = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
prev = None
for step in range(45):
    grp.run(one)
    c, h = f0["calving"].copy(), f0["calving_hflx"].copy()
    o.run((1, 0.0), c, f0["uo"], f0["vo"], f0["ui"], f0["vi"], f0["tauxa"], f0["tauya"], f0["ssh"], f0["sst"], h, f0["cn"], f0["hi"], sss=f0["sss"])
    got = [b.get_bergs(names) for b in hs]
    owner = np.concatenate([np.full(len(p["id"]), r) for r, p in enumerate(got)])
    got = {k: np.concatenate([p[k] for p in got]) for k in names}
    oo = np.argsort(got["id"], kind="stable"); owner = owner[oo]
    G, W = by_id(got), by_id(o.get_bergs(names))
    if len(G["id"]) != len(W["id"]):
        print("step", step, "count", len(G["id"]), len(W["id"])); break
    e = rel_err(G["lon"], W["lon"]); ev = np.abs(G["uvel"] - W["uvel"])
    sent = [b.counters()["n_sent"] for b in hs]
    print("step", step, "sent", sent, "max lon err %.2e" % e.max(), "max uvel diff %.2e" % ev.max())
    if ev.max() > 1e-8:
        for q in np.argsort(-ev)[:6]:
            print("   id", G["id"][q], "rank", owner[q], "cell", G["ine"][q], G["jne"][q], "ora", W["ine"][q], W["jne"][q], "lon %.4f %.4f lat %.3f" % (G["lon"][q], W["lon"][q], G["lat"][q]), "uvel %.6e %.6e" % (G["uvel"][q], W["uvel"][q]),
                  "prev owner", None if prev is None else prev.get(int(G["id"][q])))
        break
    prev = {int(i): int(r) for i, r in zip(G["id"], owner)}
