"""debug: single-rank DEM collision case with the MTS image machinery"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import by_id
from icebergs_b200 import api, synthetic as S
from test_mts_gpu import mts_pair, IKID
p = mts_pair(**IKID)
names = ["id", "lon", "lat", "ine", "jne", "halo_berg", "uvel", "vvel", "rot"]
def show(tag):
    g = p.b.get_bergs(names, include_halo=True)
    print(tag, "slots", len(g["id"]), "halo codes", np.unique(g["halo_berg"], return_counts=True))
    gb, ob = p.b.get_bonds(), p.o.get_bonds()
    print("  bonds gpu", len(gb["first_id"]), "oracle", len(ob["first_id"]), "broken", gb["broken"].sum(), ob["broken"].sum())
show("after set_bonds")
for k in range(3):
    p.step(1)
    a, b = by_id(p.b.get_bergs(names)), by_id(p.o.get_bergs(names))
    for q in ("lon", "lat", "uvel", "vvel", "rot"):
        print("  step", k + 1, q, np.abs(a[q] - b[q]).max())
    print("  ine", a["ine"], b["ine"])
    show(f"after step {k+1}")
g = p.b.get_bergs(names + ["conglom_id", "n_bonds"], include_halo=True)
for k in range(len(g["id"])):
    print(int(g["id"][k]) & 0xffff, "halo", g["halo_berg"][k], "lon %.1f lat %.1f" % (g["lon"][k], g["lat"][k]), "cell", g["ine"][k], g["jne"][k], "conglom", g["conglom_id"][k], "nb", g["n_bonds"][k])
