import sys, time
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import kid_oracle_py as O
from icebergs_b200 import api, synthetic as S
O.build()
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = S.CartesianGrid(20, 20, 15000.0)
dom = api.Domain.single(20, 20, halo=3, cyclic_x=True)
E = float(sys.argv[2]) if len(sys.argv) > 2 else 1e9
par = S.beam_params(api.default_params, dem_beam_test=2, orig_dem_moment_of_inertia=int(sys.argv[3]) if len(sys.argv) > 3 else 1, dem_damping_coef=0.7, rho_bergs=900.0, mts_sub_steps=2000, dem_spring_coef=E)
o = O.Oracle(20, 20, 100.0, (1, 0.0), params=par, domain=dom, **g.init_args())
b0 = S.cantilever_bergs()
o.set_bergs(**b0); o.set_bonds()
print("bonds", len(o.get_bonds()["first_id"]))
f = g.forcing(ibuo=0.0, ibvo=0.0, collision_test=False)
t0 = time.time()
l = 29 * 5000.0; hh = 3. * 5000.0; AI = hh ** 3 / 12.; P = -1.5e10
wtip = P * l ** 2 * (3 * l - l) / (6. * E * AI)
for k in range(nsteps):
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, k * 100. / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
    if k % 25 == 24:
        b = o.get_bergs(["id", "lon", "lat", "vvel", "start_lon", "start_lat"]); order = np.lexsort((b["start_lon"], b["start_lat"]))
        lat = b["lat"][order]
        tip = np.mean(lat[[29, 59, 89]] - b0["lat"][[29, 59, 89]])
        print(k + 1, "tip deflection %.3f (theory %.3f)  max|v| %.3e" % (tip, wtip, np.abs(b["vvel"]).max()), "wall %.1f" % (time.time() - t0), flush=True)
