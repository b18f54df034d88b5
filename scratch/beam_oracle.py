import sys, time
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import kid_oracle_py as O
from icebergs_b200 import api, synthetic as S
O.build()
nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = S.CartesianGrid(20, 20, 15000.0)
dom = api.Domain.single(20, 20, halo=3, cyclic_x=True)
o = O.Oracle(20, 20, 1.0, (1, 0.0), params=S.beam_params(api.default_params, mts_sub_steps=nsub), domain=dom, **g.init_args())
b0 = S.beam_bergs()
o.set_bergs(**b0); o.set_bonds()
print("bonds", len(o.get_bonds()["first_id"]))
f = g.forcing(ibuo=0.0, ibvo=0.0, collision_test=False)
t0 = time.time()
xa = b0["lon"] - b0["lon"][0]
l = xa.max(); P = -1.5e5; YM = 1e9; AI = 1.0 * 0.5 ** 3 / 12.0
w1 = -P * xa * (4. * xa * xa - 3. * l * l) / (48. * YM * AI)
w2 = P * (xa - l) * (l * l - 8. * l * xa + 4 * xa * xa) / (48. * YM * AI)
w = np.where(xa > 0.5 * l, w2, w1)
for k in range(nsteps):
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, k / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
    b = o.get_bergs(["id", "lon", "lat", "vvel"]); order = np.argsort(b["lon"])
    d = b["lat"][order] - b0["lat"][0]
    print(k + 1, "mid deflection %.5f (theory %.5f)  max|v| %.3e  rms err/max %.4f" % (d[14], w[14], np.abs(b["vvel"]).max(), np.sqrt(np.mean((d - w) ** 2)) / np.abs(w).max()), "wall %.1f" % (time.time() - t0), flush=True)
