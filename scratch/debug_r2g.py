"""debug (round 2, pass g): (A) ang_vel of the migrating DEM conglomerate on 2 ranks vs the oracle, (B) the 3600 s DEM collision
case on one rank step by step"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import by_id, COMPARE_F64
from icebergs_b200 import api, parallel, synthetic as S
from test_mts_gpu import MTS_KID, IKID
from test_interactions_gpu import Pair
import test_multirank_gpu as T
import common

which = sys.argv[1] if len(sys.argv) > 1 else "AB"

if "B" in which:
    print("=== B: dem, dt=3600, one rank", flush=True)
    over = dict(MTS_KID); over.update(IKID)
    p = Pair(S.collision_bergs(), lambda: S.collision_params(api.default_params, **over), dt=3600.0)
    names = ["id", "lon", "lat", "ine", "jne", "halo_berg", "uvel", "vvel", "rot", "conglom_id", "n_bonds"]
    for k in range(20):
        try:
            p.step(1)
        except Exception as e:
            print("STEP", k + 1, "FAILED", str(e)[:200])
            try:
                g = p.b.get_bergs(names, include_halo=True)
                for q in range(len(g["id"])):
                    print(int(g["id"][q]) & 0xffff, "halo", g["halo_berg"][q], "lon %.1f lat %.1f" % (g["lon"][q], g["lat"][q]), "cell", g["ine"][q], g["jne"][q],
                          "u %.3e v %.3e" % (g["uvel"][q], g["vvel"][q]), "conglom", g["conglom_id"][q], "nb", g["n_bonds"][q])
            except Exception as e2:
                print("  get_bergs failed", str(e2)[:200])
            o = by_id(p.o.get_bergs(["id", "lon", "lat", "ine", "jne"]))
            print("  oracle lon", np.round(o["lon"], 1), "lat", np.round(o["lat"], 1))
            break
        a, b = by_id(p.b.get_bergs(names)), by_id(p.o.get_bergs(["id", "lon", "lat", "uvel", "vvel", "rot", "ine", "jne"]))
        g = p.b.get_bergs(names, include_halo=True)
        print("step", k + 1, "dlon %.2e dlat %.2e du %.2e drot %.2e" % tuple(np.abs(a[q] - b[q]).max() for q in ("lon", "lat", "uvel", "rot")),
              "slots", len(g["id"]), "codes", dict(zip(*[x.tolist() for x in np.unique(g["halo_berg"], return_counts=True)])),
              "lon range %.0f %.0f lat range %.0f %.0f" % (a["lon"].min(), a["lon"].max(), a["lat"].min(), a["lat"].max()), flush=True)

if "A" in which:
    print("=== A: migrating DEM conglomerate, 2 ranks", flush=True)
    real = common.assert_bergs_match
    def loud(got, want, rtol=1e-10, names=COMPARE_F64, context="", acc_floor=0.0):
        g, w = by_id(got), by_id(want)
        for k in ("ang_vel", "rot", "uvel", "lat"):
            if k in g:
                print(context, k, "max|want| %.3e max|diff| %.3e" % (np.abs(w[k]).max(), np.abs(g[k] - w[k]).max()), flush=True)
        return {}
    T.assert_bergs_match = loud
    kw = dict(MTS_KID); kw.update(IKID)
    south = {k: v[:8].copy() for k, v in S.collision_bergs().items()}
    f64 = tuple(COMPARE_F64) + ("ang_vel", "rot")
    names = list(f64) + ["ine", "jne", "start_year", "id"]
    moved, o = T._run_cartesian_ranks(2, lambda: S.collision_params(api.default_params, **kw), south,
                                      lambda g: g.forcing(ibuo=0.05, ibvo=0.5, collision_test=False), 650, 25, names, f64, bonds=True, rtol=1e-7)
    b = o.get_bergs(["lat", "lon"])
    print("moved", moved, "lat min", b["lat"].min())
