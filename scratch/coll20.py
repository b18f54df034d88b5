import sys, time
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import kid_oracle_py as O
from icebergs_b200 import api, synthetic as S
def run(label, dt, nsteps, **over):
    g = S.CartesianGrid()
    dom = api.Domain.single(g.gni, g.gnj, halo=3, cyclic_x=True)
    o = O.Oracle(g.gni, g.gnj, dt, (1, 0.0), params=S.collision_params(api.default_params, **over), domain=dom, **g.init_args())
    o.set_bergs(**S.collision_bergs()); o.set_bonds()
    f = g.forcing()
    for k in range(nsteps):
        c, h = f["calving"].copy(), f["calving_hflx"].copy()
        o.run((1, k * dt / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
    b = o.get_bergs(["lon", "lat", "uvel", "vvel", "start_lat"])
    lo = b["start_lat"] < 10.0e3
    print(label, "n", o.count_bergs(), "sep %.1f" % (b["lat"][~lo].mean() - b["lat"][lo].mean()), "x %.1f" % b["lon"].mean(), "vmax %.3f" % np.abs(np.r_[b["uvel"], b["vvel"]]).max())
    o.close()
MTS = dict(mts=1, mts_sub_steps=60, explicit_inner_mts=1, force_convergence=1, convergence_tolerance=1e-8, contact_distance=1.75e3, contact_spring_coef=1.0e-7)
for hrs in (8, 12, 16, 20):
    run("KID      %2dh" % hrs, 60.0, 60 * hrs)
    run("KID-cong %2dh" % hrs, 60.0, 60 * hrs, contact_distance=1.75e3, contact_spring_coef=1.0e-7)
    run("MTS_KID  %2dh" % hrs, 3600.0, hrs, **MTS)
    run("iKID     %2dh" % hrs, 3600.0, hrs, **MTS, dem=1, poisson=0.3, dem_damping_coef=1.0, dem_spring_coef=4471.94)
