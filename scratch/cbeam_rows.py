import sys, time
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import kid_oracle_py as O
from icebergs_b200 import api, synthetic as S
O.build()
rows = int(sys.argv[1]); E = float(sys.argv[2]); nsteps = int(sys.argv[3])
g = S.CartesianGrid(20, 20, 15000.0)
dom = api.Domain.single(20, 20, halo=3, cyclic_x=True)
par = S.beam_params(api.default_params, dem_beam_test=2, orig_dem_moment_of_inertia=1, dem_damping_coef=0.7, rho_bergs=900.0, mts_sub_steps=2000, dem_spring_coef=E)
o = O.Oracle(20, 20, 100.0, (1, 0.0), params=par, domain=dom, **g.init_args())
b0 = S.cantilever_bergs(rows=rows)
o.set_bergs(**b0); o.set_bonds()
f = g.forcing(ibuo=0.0, ibvo=0.0, collision_test=False)
l = 29 * 5000.0; hh = rows * 5000.0; AI = hh ** 3 / 12.; P = -1.5e10 / 3. * rows
wtip = P * l ** 3 / (3. * E * AI)
for k in range(nsteps):
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, k * 100. / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
b = o.get_bergs(["id", "lon", "lat", "vvel", "rot", "start_lon", "start_lat"]); order = np.lexsort((b["start_lon"], b["start_lat"]))
lat = b["lat"][order]
tips = [29 + 30 * r for r in range(rows)]
print("rows", rows, "tip deflection %.3f (theory %.3f) ratio %.4f max|v| %.2e" % (np.mean(lat[tips] - b0["lat"][tips]), wtip, np.mean(lat[tips] - b0["lat"][tips]) / wtip, np.abs(b["vvel"]).max()), "tip rot", b["rot"][order][tips])
if rows == 1:
    th = b["rot"][order]; y = lat - b0["lat"]; x = b["lon"][order] - b0["lon"]
    l0 = 5000.0; EI = E * AI
    for i in (0, 1, 2, 3, 10, 20, 28, 29):
        xi = i * l0
        th_t = P * (l * xi - xi * xi / 2) / EI; y_t = P * xi * xi * (3 * l - xi) / (6 * EI)
        print(i, "rot %.6e (th %.6e)  y %.4f (th %.4f)  dx %.4f" % (th[i], th_t, y[i], y_t, x[i]))
    dy = np.diff(y); avg = 0.5 * (th[1:] + th[:-1]) * l0
    print("dy - l0*avg(theta):", (dy - avg)[:5], (dy - avg)[-3:])
