import sys, time
sys.path.insert(0, '.')
import numpy as np
from icebergs_b200 import api, synthetic as S
nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
b0 = S.beam_bergs()
g = S.CartesianGrid(20, 20, 15000.0)
dom = api.Domain.single(20, 20, halo=3, cyclic_x=True)
h = api.icebergs_init(20, 20, 1.0, (1, 0.0), params=S.beam_params(api.default_params, mts_sub_steps=nsub), domain=dom, capacity=1024, **g.init_args())
h.set_bergs(**b0); h.set_bonds()
f = g.forcing(ibuo=0.0, ibvo=0.0, collision_test=False)
for k in range(nsteps):
    c, hf = f["calving"].copy(), f["calving_hflx"].copy()
    t0 = time.perf_counter()
    api.icebergs_run(h, (1, k / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hf, f["cn"], f["hi"], sss=f["sss"])
    print("step", k, "wall %.4f s -> %.2f us per sub-step" % (time.perf_counter() - t0, (time.perf_counter() - t0) / nsub * 1e6), flush=True)
api.icebergs_end(h)
