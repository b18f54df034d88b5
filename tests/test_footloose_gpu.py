"""Footloose calving (footloose_calving I:2503, calve_fl_icebergs I:6405,
adjust_fl_berg_interactivity I:2765, the FL-bits branch of thermodynamics I:3031-3064) on the
reference's own footloose test (tests/footloose_tests): CUDA path against the CPU oracle."""
import numpy as np
import pytest

import kid_oracle_py as O
from common import COMPARE_F64, assert_bergs_match
from icebergs_b200 import _cdefs as D
from icebergs_b200 import api
from icebergs_b200 import synthetic as S

pytestmark = pytest.mark.gpu

F64 = tuple(COMPARE_F64) + ("mass_of_fl_bits", "mass_of_fl_bergy_bits", "fl_k")
NAMES = list(F64) + ["ine", "jne", "start_year", "id"]


def run_pair(params, bergs, nsteps, chunk, dt=10.0, check_every=1):
    g = S.CartesianGrid()
    dom = lambda: api.Domain.single(20, 20, halo=3, cyclic_x=True)
    b = api.icebergs_init(20, 20, dt, (1, 0.0), params=params(), domain=dom(), capacity=4096, **g.init_args())
    o = O.Oracle(20, 20, dt, (1, 0.0), params=params(), domain=dom(), **g.init_args())
    b.set_bergs(**bergs); o.set_bergs(**bergs)
    f = S.footloose_forcing(g)
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    args = ((1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"])
    api.icebergs_run(b, *args)
    o.run(*args)
    done, k = 1, 0
    counts = []
    while done < nsteps:
        n = min(chunk, nsteps - done)
        t = done * dt / 86400.0
        b.step_resident(n, 1, t)
        o.step_again(n, 1, t)
        done += n
        k += 1
        assert b.count_bergs() == o.count_bergs(), f"count after {done} steps"
        counts.append(o.count_bergs())
        if k % check_every == 0:
            assert_bergs_match(b.get_bergs(NAMES), o.get_bergs(NAMES), rtol=1e-7, names=F64, context=f"{done} steps")
    for fid in (D.KID_FLD_FL_BITS_SRC, D.KID_FLD_FL_BITS_MELT, D.KID_FLD_FLOATING_MELT, D.KID_FLD_BERGY_SRC):
        gf, of = b.grid_field(fid), o.grid_field(fid)
        assert np.max(np.abs(gf - of)) <= 1e-7 * max(np.max(np.abs(of)), 1e-300), f"grid field {fid}"
    cg, co = b.counters(), o.counters()
    assert cg["nbergs_calved_fl"] == co["nbergs_calved_fl"]
    api.icebergs_end(b)
    o.close()
    return counts, co


def test_footloose_bits_then_bergs_from_bits():
    """fl_style='fl_bits' (the shipped test): the foot accumulates (fl_k), k child-areas go to the FL
    bits, and once the bits exceed new_berg_from_fl_bits_mass_thres a berg is made from them."""
    params = lambda: S.footloose_params(api.default_params)
    counts, co = run_pair(params, S.footloose_bergs(), 24001, 2000)
    assert counts[-1] >= 4 and co["nbergs_calved_fl"] >= 2, counts


def test_footloose_new_bergs_style():
    """fl_style='new_bergs': every calving event makes a child berg (ids from the parent's cell counter)."""
    params = lambda: S.footloose_params(api.default_params, fl_style_fl_bits=0)
    counts, co = run_pair(params, S.footloose_bergs(), 12001, 1000)
    assert co["nbergs_calved_fl"] >= 4, co


def test_footloose_children_with_interactions_on():
    """new_bergs style with interactive_icebergs_on: children start non-interacting (fl_k = -1,
    I:6423-6428) and switch to -2 once out of contact range (adjust_fl_berg_interactivity I:2765)."""
    params = lambda: S.footloose_params(api.default_params, fl_style_fl_bits=0, interactive_icebergs_on=1)
    counts, co = run_pair(params, S.footloose_bergs(), 8001, 1000)
    assert co["nbergs_calved_fl"] >= 2, co
