import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kid_oracle_py as O
from common import by_id, rel_err
from icebergs_b200 import api, synthetic as S
rk = int(sys.argv[1]) if len(sys.argv) > 1 else 1
style = int(sys.argv[2]) if len(sys.argv) > 2 else 0
params = lambda: S.footloose_params(api.default_params, runge_not_verlet=rk, fl_style_fl_bits=style)
g = S.CartesianGrid(); dt = 10.0
dom = lambda: api.Domain.single(20, 20, halo=3, cyclic_x=True)
b = api.icebergs_init(20, 20, dt, (1, 0.0), params=params(), domain=dom(), capacity=4096, **g.init_args())
o = O.Oracle(20, 20, dt, (1, 0.0), params=params(), domain=dom(), **g.init_args())
bergs = S.footloose_bergs()
b.set_bergs(**bergs); o.set_bergs(**bergs)
f = S.footloose_forcing(g)
c, h = f["calving"].copy(), f["calving_hflx"].copy()
args = ((1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"])
api.icebergs_run(b, *args); o.run(*args)
names = ["id", "lon", "lat", "uvel", "vvel", "mass", "axn", "bxn", "fl_k", "xi", "ine"]
done = 1
for k in range(12):
    n = 250
    t = done * dt / 86400.0
    b.step_resident(n, 1, t); o.step_again(n, 1, t); done += n
    ga, oa = by_id(b.get_bergs(names)), by_id(o.get_bergs(names))
    if len(ga["id"]) != len(oa["id"]): print("count", len(ga["id"]), len(oa["id"])); break
    print("steps", done, "n", len(ga["id"]), {q: "%.1e" % rel_err(ga[q], oa[q]).max() for q in names[1:-1]}, "ids", [int(x) & 0xfff for x in ga["id"]], flush=True)
