import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
from common import by_id, rel_err
from test_interactions_gpu import Pair, NAMES
from icebergs_b200 import api, synthetic as S
rng = np.random.default_rng(11); n = 800
base = S.collision_bergs()
cols = {k: np.resize(v, n).copy() for k, v in base.items()}
cols["lon"] = rng.uniform(50.0, 19950.0, n); cols["lat"] = rng.uniform(1050.0, 18950.0, n)
cols["start_lon"], cols["start_lat"] = cols["lon"].copy(), cols["lat"].copy()
cols["start_day"] = rng.uniform(0.0, 300.0, n)
params = lambda: S.collision_params(api.default_params, iceberg_bonds_on=0, manually_initialize_bonds=0, max_bonds=0)
g = S.CartesianGrid()
p = Pair(cols, params, grid=g, bonds=False, forcing=g.forcing(ibuo=0.6, ibvo=0.0, collision_test=False), capacity=16384)
for k in range(12):
    p.step(1)
    gb, ob = p.b.get_bergs(NAMES), p.o.get_bergs(NAMES)
    print("step", k + 1, "counts", len(gb["id"]), len(ob["id"]), end=" ")
    if len(gb["id"]) != len(ob["id"]):
        print("missing", np.setdiff1d(ob["id"], gb["id"])[:5], "extra", np.setdiff1d(gb["id"], ob["id"])[:5]); break
    G, W = by_id(gb), by_id(ob)
    worst = {k2: float(rel_err(G[k2], W[k2]).max()) for k2 in ("lon", "lat", "uvel", "vvel", "axn", "bxn", "xi", "lon_old")}
    bad = np.nonzero((G["ine"] != W["ine"]) | (G["jne"] != W["jne"]))[0]
    print("cell mismatches", len(bad), {a: "%.1e" % b for a, b in worst.items()})
    for q in bad[:5]:
        print("   id", G["id"][q], "gpu", G["ine"][q], G["jne"][q], "%.6f %.6f xi %.5f" % (G["lon"][q], G["lat"][q], G["xi"][q]),
              "ora", W["ine"][q], W["jne"][q], "%.6f %.6f xi %.5f" % (W["lon"][q], W["lat"][q], W["xi"][q]))
    big = np.argsort(-rel_err(G["lon"], W["lon"]))[:3]
    for q in big:
        print("   lon worst id", G["id"][q], "%.9f vs %.9f" % (G["lon"][q], W["lon"][q]), "cell", G["ine"][q], G["jne"][q], "uvel %.6e %.6e" % (G["uvel"][q], W["uvel"][q]))
