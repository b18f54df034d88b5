"""debug: one-rank MTS conglomerate through the cyclic seam (dem from argv)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from icebergs_b200 import api, synthetic as S
from test_mts_gpu import MTS_KID, IKID
dem = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dt, L = 60.0, 20.0e3
over = dict(MTS_KID)
if dem: over.update(IKID)
bergs = {k: v[:8].copy() for k, v in S.collision_bergs().items()}
bergs["lon"] = bergs["lon"] + 11.5e3; bergs["start_lon"] = bergs["lon"].copy()
g20 = S.CartesianGrid(20, 20)
f = g20.forcing(ibuo=0.6, ibvo=0.02, collision_test=False)
h = api.icebergs_init(20, 20, dt, (1, 0.0), params=S.collision_params(api.default_params, Lx=L, **over), domain=api.Domain.single(20, 20, halo=3, cyclic_x=True), capacity=4096, **g20.init_args())
h.set_bergs(**bergs); h.set_bonds()
names = ["id", "lon", "lat", "ine", "jne", "halo_berg", "uvel", "conglom_id", "n_bonds"]
def dump(tag):
    g = h.get_bergs(names, include_halo=True)
    print(tag, "slots", len(g["id"]))
    order = np.lexsort((g["lon"], g["id"]))
    for q in order:
        print("  id", int(g["id"][q]) & 0xffff, "halo", int(g["halo_berg"][q]), "lon %.1f lat %.1f" % (g["lon"][q], g["lat"][q]), "cell", g["ine"][q], g["jne"][q], "u %.3e" % g["uvel"][q], "conglom", g["conglom_id"][q], "nb", g["n_bonds"][q])
    b = h.get_bonds()
    print("  bonds", len(b["first_id"]))
for step in range(420):
    try:
        c, hf = f["calving"].copy(), f["calving_hflx"].copy()
        api.icebergs_run(h, (1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hf, f["cn"], f["hi"], sss=f["sss"])
    except Exception as e:
        print("step", step, "FAILED", str(e)[:150])
        try: dump("after failure")
        except Exception as e2: print("dump failed", e2)
        break
    g = h.get_bergs(["lon", "ine"])
    if step % 20 == 0 or g["lon"].max() > 19.0e3:
        if step % 5 == 0: dump(f"step {step}")
