import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'oracle'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
from common import *
case = Case(96, 48, 20000)
b, o = case.make_gpu(), case.make_oracle()
run_gpu(b, case); run_oracle(o, case)
names = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
g, w = by_id(b.get_bergs(names)), by_id(o.get_bergs(names))
for k in COMPARE_F64:
    e = rel_err(g[k], w[k])
    print(k, 'max rel err', e.max(), 'n>1e-10', int((e > 1e-10).sum()))
e = rel_err(g['mass_of_bits'], w['mass_of_bits'])
bad = np.nonzero(e > 1e-6)[0][:8]
for k in bad:
    print('id', g['id'][k], 'gpu bits', g['mass_of_bits'][k], 'cpu bits', w['mass_of_bits'][k], 'mass', g['mass'][k], w['mass'][k], 'T', g['thickness'][k], w['thickness'][k], 'W', g['width'][k], 'L', g['length'][k], 'lat', g['lat'][k], 'ine,jne', g['ine'][k], g['jne'][k])
