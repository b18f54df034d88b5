"""debug: the multi-rank bench sequence on ONE GPU (2 in-process ranks): plain calls, announced calls, long resident runs"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from icebergs_b200 import api, parallel, synthetic as S
GNI, GNJ, DT = 1440, 720, 3600.0
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_per = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_500_000
mode = sys.argv[3] if len(sys.argv) > 3 else "all"
grp = parallel.LocalGroup(nr)
doms = [grp.domain(GNI, GNJ, r, halo=4) for r in range(nr)]
grids = [S.Grid(GNI, GNJ, d.isc, d.iec, d.jsc, d.jec) for d in doms]
hs = [None] * nr
fs = [None] * nr
def init(r):
    p = S.workload_params(api.default_params)
    hs[r] = api.icebergs_init(GNI, GNJ, DT, (1, 0.0), params=p, domain=doms[r], capacity=int(n_per * 1.25) + 4096, **grids[r].init_args())
    cols, _ = grids[r].seed_bergs(n_per, stream=r)
    hs[r].set_bergs(**cols)
    f = grids[r].forcing()
    fs[r] = [{k: np.ascontiguousarray(v).copy() for k, v in f.items()} for _ in range(2)]
grp.run(init)
def args_of(r, k, c, h):
    f = fs[r][k % 2]
    return (c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"]), f["sss"]
def stage(name, fn):
    t0 = time.perf_counter()
    try:
        grp.run(fn)
        print(name, "ok %.2fs" % (time.perf_counter() - t0), "bergs", [b.count_bergs() for b in hs], "sent", [b.counters()["n_sent"] for b in hs], flush=True)
    except Exception as e:
        print(name, "FAILED", str(e)[:200], flush=True)
        raise
def plain(r, n=1):
    for k in range(n):
        c, h = np.zeros_like(fs[r][0]["calving"]), np.zeros_like(fs[r][0]["calving_hflx"])
        a, sss = args_of(r, k, c, h)
        api.icebergs_run(hs[r], (1, 0.0), *a, sss=sss)
def announced(r, n):
    pairs = [(np.zeros_like(fs[r][0]["calving"]), np.zeros_like(fs[r][0]["calving_hflx"])) for _ in range(3)]
    a, sss = args_of(r, 0, *pairs[0])
    api.icebergs_prefetch(hs[r], *a, sss=sss)
    for k in range(n):
        if k + 1 < n:
            for q in pairs[(k + 1) % 3]: q.fill(0.0)
            a1, s1 = args_of(r, k + 1, *pairs[(k + 1) % 3])
            api.icebergs_prefetch(hs[r], *a1, sss=s1)
        a, sss = args_of(r, k, *pairs[k % 3])
        api.icebergs_run(hs[r], (1, 0.0), *a, sss=sss)
stage("first run", lambda r: plain(r))
stage("resident 25", lambda r: hs[r].step_resident(25, 1, 0.0))
if mode in ("all", "plain"):
    stage("plain x23", lambda r: plain(r, 23))
if mode in ("all", "pre"):
    stage("announced x23", lambda r: announced(r, 23))
pat = sys.argv[4] if len(sys.argv) > 4 else "16x5"
n1, n2 = [int(x) for x in pat.split("x")]
done = 25
for rep in range(n2):
    stage(f"resident {n1} (steps {done + 1}..{done + n1})", lambda r: hs[r].step_resident(n1, 1, 0.0))
    done += n1
