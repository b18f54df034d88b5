import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
from common import Case, run_oracle, by_id
from test_multirank_gpu import Ranks, NAMES
from icebergs_b200 import parallel, api
nranks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
case = Case(96, 48, 12000, dt=86400.0)
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
grp = parallel.LocalGroup(nranks)
ranks = Ranks(case, nranks, lambda r: grp.domain(case.gni, case.gnj, r, halo=case.halo), grp.run)
o = case.make_oracle()
over = dict(uo=1.2, vo=0.15, tauxa=15.0)
fast = {k: np.full_like(case.forcing[k], v) for k, v in over.items()}
for _ in range(nsteps - 1):
    ranks.step(over); run_oracle(o, case, **fast)
b0 = by_id(o.get_bergs(NAMES))
g0 = by_id(ranks.bergs())
ranks.step(over); run_oracle(o, case, **fast)
g, w = by_id(ranks.bergs()), by_id(o.get_bergs(NAMES))
print("counts", len(g["id"]), len(w["id"]))
if len(g["id"]) == len(w["id"]):
    bad = np.nonzero((g["ine"] != w["ine"]) | (g["jne"] != w["jne"]))[0]
    print("mismatches", len(bad))
    for k in bad[:12]:
        i0 = np.nonzero(b0["id"] == g["id"][k])[0][0]
        print("gpu start", g0["ine"][i0], g0["jne"][i0], "%.5f %.5f" % (g0["lon"][i0], g0["lat"][i0]), "uv %.4f %.4f" % (g0["uvel"][i0], g0["vvel"][i0]), "ora uv %.4f %.4f" % (b0["uvel"][i0], b0["vvel"][i0]))
        print("id", g["id"][k], "start", b0["ine"][i0], b0["jne"][i0], "%.4f %.4f" % (b0["lon"][i0], b0["lat"][i0]),
              "| gpu", g["ine"][k], g["jne"][k], "%.5f %.5f xi %.4f %.4f" % (g["lon"][k], g["lat"][k], g["xi"][k], g["yj"][k]),
              "| ora", w["ine"][k], w["jne"][k], "%.5f %.5f xi %.4f %.4f" % (w["lon"][k], w["lat"][k], w["xi"][k], w["yj"][k]))
else:
    missing = np.setdiff1d(w["id"], g["id"]); extra = np.setdiff1d(g["id"], w["id"])
    print("missing", missing[:10], "extra", extra[:10])
print([ (c["n_sent"], c["n_received"], c["n_bounced"]) for c in ranks.counters()], o.counters()["n_bounced"])
for r,b in enumerate(ranks.h):
    d = ranks.doms[r]
    m = b.grid_field(34)  # msk
    print("rank", r, "msk halo col sums W", m[:, :4].sum(), "E", m[:, -4:].sum(), "S", m[:4].sum(), "N", m[-4:].sum())
