import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import Case, by_id, run_oracle
from icebergs_b200 import api, parallel
from test_multirank_gpu import Ranks, NAMES
case = Case(96, 48, 12000, dt=43200.0, old_bug_bilin=0)
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 2
grp = parallel.LocalGroup(nr)
ranks = Ranks(case, nr, lambda r: grp.domain(case.gni, case.gnj, r, halo=case.halo), grp.run)
o = case.make_oracle()
over = dict(uo=1.2, vo=0.15, tauxa=15.0)
fast = {k: np.full_like(case.forcing[k], v) for k, v in over.items()}
ranks.step(over); run_oracle(o, case, **fast)
iv = int(sys.argv[2]) if len(sys.argv) > 2 else 5
grp.run(lambda r: ranks.h[r].set_sort_phase(iv, 0))
steps = 1
for n in (9, 3, 1, 12, 4, 2, 16):
    if n == 1:
        ranks.step(over); run_oracle(o, case, **fast)
    else:
        ranks.resident(n); o.step_again(n, 1, 0.0)
    steps += n
    g, w = by_id(ranks.bergs()), by_id(o.get_bergs(NAMES))
    if len(g["id"]) != len(w["id"]):
        print(steps, "COUNT", len(g["id"]), len(w["id"])); break
    bad = np.nonzero((g["ine"] != w["ine"]) | (g["jne"] != w["jne"]))[0]
    dl = np.abs(g["lon"] - w["lon"]); dl = np.minimum(dl, 360 - dl)
    print("steps", steps, "cell mismatches", len(bad), "max dlon %.3e dlat %.3e" % (dl.max(), np.abs(g["lat"] - w["lat"]).max()), "sorts", [b.sorts_done() for b in ranks.h], flush=True)
    for k in bad[:6]:
        print("   id", g["id"][k], "gpu cell", g["ine"][k], g["jne"][k], "xi %.6f yj %.6f lon %.6f lat %.6f" % (g["xi"][k], g["yj"][k], g["lon"][k], g["lat"][k]),
              "| oracle", w["ine"][k], w["jne"][k], "xi %.6f yj %.6f lon %.6f lat %.6f" % (w["xi"][k], w["yj"][k], w["lon"][k], w["lat"][k]))
