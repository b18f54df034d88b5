import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import Case, by_id
from icebergs_b200 import api
import test_prefetch_gpu as T
case = Case(96, 48, 3000)
sets = T._forcing_sets(case, 2)
a, b = case.make_gpu(), case.make_gpu()
c0, h0 = sets[0]["calving"].copy(), sets[0]["calving_hflx"].copy()
T._run(a, sets[0], c0, h0)
c1, h1 = sets[0]["calving"].copy(), sets[0]["calving_hflx"].copy()
T._run(b, sets[0], c1, h1)
print("two plain handles: calving equal", np.array_equal(c0, c1), "max diff", np.abs(c0 - c1).max(), "hflx equal", np.array_equal(h0, h1), np.abs(h0 - h1).max(), "nan", np.isnan(c0).sum(), np.isnan(h0).sum())
print("calving in max", sets[0]["calving"].max(), "out max", c0.max(), "hflx in", np.abs(sets[0]["calving_hflx"]).max(), "out", np.abs(h0).max())
print("counts", a.count_bergs(), b.count_bergs())
