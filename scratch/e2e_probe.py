import time, sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from icebergs_b200 import api, synthetic as S
sys.argv = ['x']
import bench
GNI, GNJ = 1440, 720
p = S.workload_params(api.default_params)
dom = api.Domain.single(GNI, GNJ, halo=p.halo, cyclic_x=True)
grid = S.Grid(GNI, GNJ)
b = api.icebergs_init(GNI, GNJ, 3600.0, (1, 0.0), params=p, domain=dom, capacity=12_600_000, **grid.init_args())
cols, counter = grid.seed_bergs(10_000_000)
b.set_bergs(**cols); del cols
f = grid.forcing(); fp = {}; keep = []
for k, v in f.items():
    fp[k], t = bench.pinned(v); keep.append(t)
calving, hflx = fp["calving"], fp["calving_hflx"]
def run_once(zero=True):
    if zero:
        calving[...] = 0.0; hflx[...] = 0.0
    api.icebergs_run(b, (1, 0.0), calving, fp["uo"], fp["vo"], fp["ui"], fp["vi"], fp["tauxa"], fp["tauya"], fp["ssh"], fp["sst"], hflx, fp["cn"], fp["hi"], sss=fp["sss"])
for _ in range(3): run_once()
torch.cuda.synchronize()
import ctypes as C
tz = tr = 0.0
for _ in range(20):
    a = time.perf_counter()
    calving.fill(0.0); hflx.fill(0.0)
    c = time.perf_counter()
    api.icebergs_run(b, (1, 0.0), calving, fp["uo"], fp["vo"], fp["ui"], fp["vi"], fp["tauxa"], fp["tauya"], fp["ssh"], fp["sst"], hflx, fp["cn"], fp["hi"], sss=fp["sss"])
    d = time.perf_counter()
    tz += c - a; tr += d - c
print("per call: zeroing %.3f ms, icebergs_run %.3f ms" % (tz / 20 * 1e3, tr / 20 * 1e3), b.last_timing())
# raw H2D
dev = [torch.empty(v.shape, dtype=torch.float64, device="cuda") for v in keep]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    for d, h in zip(dev, keep): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
nb = sum(h.numel() * 8 for h in keep)
dt = (time.perf_counter() - t0) / 20
print("raw H2D %.1f MB in %.3f ms = %.1f GB/s" % (nb / 1e6, dt * 1e3, nb / dt / 1e9))
big = torch.empty(nb // 8, dtype=torch.float64, pin_memory=True); dbig = torch.empty(nb // 8, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("one big H2D %.3f ms = %.1f GB/s" % (dt * 1e3, nb / dt / 1e9))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): big[:2_000_000].copy_(dbig[:2_000_000], non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("D2H 16 MB %.3f ms = %.1f GB/s" % (dt * 1e3, 16e6 / dt / 1e9))
