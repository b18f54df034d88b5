import sys, time
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import kid_oracle_py as O
from icebergs_b200 import api, synthetic as S
O.build()
def run(label, nsteps=1300, **over):
    kw = dict(mts=1, mts_sub_steps=60, explicit_inner_mts=1, force_convergence=1, convergence_tolerance=1e-8,
              contact_distance=1.75e3, contact_spring_coef=1.0e-7)
    kw.update(over)
    params = lambda: S.collision_params(api.default_params, **kw)
    g = S.CartesianGrid()
    dom = api.Domain.single(g.gni, g.gnj, halo=3, cyclic_x=True)
    o = O.Oracle(g.gni, g.gnj, 60.0, (1, 0.0), params=params(), domain=dom, **g.init_args())
    o.set_bergs(**S.collision_bergs())
    o.set_bonds()
    f = g.forcing()
    t0 = time.time()
    for k in range(nsteps):
        c, h = f["calving"].copy(), f["calving_hflx"].copy()
        o.run((1, k * 60.0 / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
        if k % 100 == 99 or k == 0:
            b = o.get_bergs(["id", "lon", "lat", "uvel", "vvel", "mass"])
            order = np.argsort(b["id"])
            lat = b["lat"][order]; v = b["vvel"][order]; m = b["mass"][order]
            print(label, k + 1, "n", o.count_bergs(), "gap", float(abs(lat[8:].mean() - lat[:8].mean())), "finite", bool(np.isfinite(lat).all()), "py", float((m * v).sum()), "vmax", float(np.abs(v).max()))
    print(label, "wall", time.time() - t0)
    o.close()
run("MTS_KID")
run("iKID", dem=1, poisson=0.3, dem_damping_coef=1.0, dem_spring_coef=4471.94)
run("STS-conglom", mts=0, force_convergence=0, explicit_inner_mts=0)
