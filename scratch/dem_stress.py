import sys
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import kid_oracle_py as O
from icebergs_b200 import api, synthetic as S
O.build()
kw = dict(mts=1, mts_sub_steps=60, explicit_inner_mts=1, force_convergence=1, convergence_tolerance=1e-8,
          contact_distance=1.75e3, contact_spring_coef=1.0e-7, dem=1, poisson=0.3, dem_damping_coef=1.0, dem_spring_coef=4471.94)
kw.update(eval("dict(%s)" % (sys.argv[1] if len(sys.argv) > 1 else "")))
g = S.CartesianGrid()
dom = api.Domain.single(g.gni, g.gnj, halo=3, cyclic_x=True)
o = O.Oracle(g.gni, g.gnj, 60.0, (1, 0.0), params=S.collision_params(api.default_params, **kw), domain=dom, **g.init_args())
o.set_bergs(**S.collision_bergs()); o.set_bonds()
f = g.forcing()
for k in range(900):
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, k * 60.0 / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
    if k % 50 == 49:
        b = o.get_bonds()
        print(k + 1, "bonds", len(b["first_id"]), "broken", int(b["broken"].sum()), "nstress max %.3e min %.3e" % (b["nstress"].max(), b["nstress"].min()), "sstress max %.3e" % b["sstress"].max())
