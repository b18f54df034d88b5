"""The multiple-time-step scheme: evolve_icebergs_mts I:6576, accel_mts I:1278,
accel_explicit_inner_mts I:1710 -- the CUDA path against the CPU oracle on the two 8-element conglomerates of
tests/collision_tests with the switches of input_MTS_KID.nml (mts, 60 sub-steps, explicit inner steps,
force_convergence, contact_distance 1.75 km, contact_spring_coef 1e-7)."""
import numpy as np
import pytest

from common import COMPARE_F64, by_id
from icebergs_b200 import api
from icebergs_b200 import synthetic as S
from test_interactions_gpu import Pair, bond_set

pytestmark = pytest.mark.gpu

MTS_KID = dict(mts=1, mts_sub_steps=60, explicit_inner_mts=1, force_convergence=1, convergence_tolerance=1e-8,
               contact_distance=1.75e3, contact_spring_coef=1.0e-7)


def mts_pair(**over):
    kw = dict(MTS_KID)
    kw.update(over)
    return Pair(S.collision_bergs(), lambda: S.collision_params(api.default_params, **kw))


def test_mts_kid_first_steps():
    p = mts_pair()
    assert p.b.count_bergs() == 16 == p.o.count_bergs()              # README:16-22 '#= 16' for MTS_KID too
    assert bond_set(p.b.get_bonds()) == bond_set(p.o.get_bonds())
    p.step(1)
    p.check("one step", rtol=1e-10)
    g, o = by_id(p.b.get_bergs(["id", "axn_fast", "ayn_fast", "bxn_fast", "byn_fast"])), by_id(p.o.get_bergs(["id", "axn_fast", "ayn_fast", "bxn_fast", "byn_fast"]))
    for k in ("axn_fast", "ayn_fast"):
        scale = max(np.abs(o[k]).max(), 1e-13)
        assert np.abs(g[k] - o[k]).max() <= 1e-9 * scale + 1e-16, k
    p.step(49)
    p.check("50 steps", rtol=1e-9)
    p.end()


def test_mts_kid_through_contact():
    """The conglomerates meet after ~550 steps: the long-step collision force with its convergence passes, the
    bonded sub-steps (60 per step) and the contact search inside a conglomerate all act; the berg count stays 16."""
    p = mts_pair()
    gaps, passes = [], 0
    for k in range(14):
        p.step(50)
        p.check(f"{50 * (k + 1)} steps", rtol=1e-6)
        g = by_id(p.b.get_bergs(["id", "lat"]))
        half = g["lat"] < 10.0e3
        gaps.append(float(g["lat"][~half].min() - g["lat"][half].max()))
    assert min(gaps) < 1100.0 and gaps[-1] > min(gaps), f"no bounce: {gaps}"
    assert p.b.count_bergs() == 16 == p.o.count_bergs()
    p.end()


@pytest.mark.parametrize("fc", [0, 1])
def test_mts_implicit_inner_steps(fc):
    """explicit_inner_mts=.false.: the sub-steps solve accel_mts with only_interactive_forces (I:6922), with
    force_convergence they repeat until the velocity norm settles (I:6936-6968)."""
    p = mts_pair(explicit_inner_mts=0, force_convergence=fc, mts_sub_steps=20)
    p.step(1)
    p.check("one step", rtol=1e-10)
    p.step(29)
    p.check("30 steps", rtol=1e-8)
    p.end()


def test_mts_refuses_what_is_not_built():
    with pytest.raises(api.KidFatal):
        mts_pair(dem=1)
