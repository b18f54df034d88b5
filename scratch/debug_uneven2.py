import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import Case, by_id, run_oracle
from icebergs_b200 import api, parallel
from icebergs_b200 import _cdefs as D
from test_multirank_gpu import Ranks, NAMES
case = Case(96, 48, 12000, dt=43200.0, old_bug_bilin=0)
nr = 2
grp = parallel.LocalGroup(nr)
ranks = Ranks(case, nr, lambda r: grp.domain(case.gni, case.gnj, r, halo=case.halo), grp.run)
print("domains", [(d.isc, d.iec, d.jsc, d.jec) for d in ranks.doms])
o = case.make_oracle()
over = dict(uo=1.2, vo=0.15, tauxa=15.0)
fast = {k: np.full_like(case.forcing[k], v) for k, v in over.items()}
ids = [12884903560, 17179870760, 4294969064]
hl = case.halo
mo = o.grid_field(D.KID_FLD_MSK)
print("oracle msk rows j=18,20 cols 40..52:", mo[hl + 17, hl + 39:hl + 52], mo[hl + 19, hl + 39:hl + 52])
for r in range(nr):
    d = ranks.doms[r]
    m = ranks.h[r].grid_field(D.KID_FLD_MSK)
    if d.jsc <= 18 <= d.jec:
        i0 = max(40, d.isc - hl); i1 = min(52, d.iec + hl)
        print("rank", r, "msk row 18 cols", i0, "..", i1, m[hl + 18 - d.jsc, hl + i0 - d.isc: hl + i1 - d.isc + 1])
for step in range(1, 28):
    ranks.step(over); run_oracle(o, case, **fast)
    g, w = by_id(ranks.bergs()), by_id(o.get_bergs(NAMES))
    if len(g["id"]) != len(w["id"]): print("count differs", len(g["id"]), len(w["id"])); break
    for bid in ids:
        k = np.searchsorted(g["id"], bid)
        if k < len(g["id"]) and g["id"][k] == bid:
            if step >= 10 and (g["ine"][k] != w["ine"][k] or abs(g["lon"][k] - w["lon"][k]) > 1e-9 or step % 4 == 0):
                print("step", step, "id", bid, "gpu", g["ine"][k], g["jne"][k], "xi %.5f lon %.5f u %.4f" % (g["xi"][k], g["lon"][k], g["uvel"][k]), "| oracle", w["ine"][k], w["jne"][k], "xi %.5f lon %.5f u %.4f" % (w["xi"][k], w["lon"][k], w["uvel"][k]), flush=True)
