import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from icebergs_b200 import api, synthetic as S
from test_mts_gpu import BeamPair, CBEAM
p = BeamPair(S.cantilever_bergs(), 100.0, **dict(CBEAM, remove_unused_bergs=0))
names = ["id", "lon", "lat", "ine", "jne", "halo_berg", "conglom_id", "n_bonds", "static_berg"]
def show(tag):
    g = p.b.get_bergs(names, include_halo=True)
    print(tag, "slots", len(g["id"]), "halo codes", np.unique(g["halo_berg"], return_counts=True), "congloms", np.unique(g["conglom_id"], return_counts=True))
    print("  n_bonds", np.unique(g["n_bonds"], return_counts=True), "cells i", np.unique(g["ine"]), "j", np.unique(g["jne"]))
    return g
g = show("after set_bonds")
h = g["halo_berg"] > 0
import collections
print("image lon range by conglom:")
for c in np.unique(g["conglom_id"][h]):
    m = h & (g["conglom_id"] == c)
    print("  conglom", c, "n", m.sum(), "lon", g["lon"][m].min(), g["lon"][m].max())
bd = p.b.get_bonds()
print("owned bonds", len(bd["first_id"]))
try:
    p.step(1)
except Exception as e:
    print("STEP FAILED", e)
g = show("after step")
bd = p.b.get_bonds()
print("bonds", len(bd["first_id"]))
