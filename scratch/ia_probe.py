"""probe: throughput of unbonded interacting bergs (interactive_icebergs_on, Verlet) on the 1/4-degree grid"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icebergs_b200 import api, synthetic as S
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = S.Grid(1440, 720)
p = S.workload_params(api.default_params, halo=4, old_bug_bilin=0, interactive_icebergs_on=1, use_new_predictive_corrective=1)
dom = api.Domain.single(1440, 720, halo=4, cyclic_x=True)
b = api.icebergs_init(1440, 720, 3600.0, (1, 0.0), params=p, domain=dom, capacity=int(1.3 * n) + 4096, **g.init_args())
cols, counter = g.seed_bergs(n)
cnt = np.zeros((dom.njd, dom.nid), dtype=np.int32); cnt[4:4 + 720, 4:4 + 1440] = counter
b.set_calving_state(iceberg_counter_grd=cnt)
b.set_bergs(**cols)
f = g.forcing()
def step():
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    api.icebergs_run(b, (1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
step(); step()
b.synchronize() if hasattr(b, "synchronize") else None
t0 = time.perf_counter()
b.step_resident(steps, 1, 0.0)
api.lib().kid_synchronize(b.handle)
dt = (time.perf_counter() - t0) / steps
print(f"interactive n={n}: {1e3*dt:.3f} ms/step, {n/dt:.3e} berg-steps/s, timing {b.last_timing()}, counters err={b.counters()['error_flags']} nbergs={b.counters()['nbergs']}")
api.icebergs_end(b)
