"""Runge_Kutta_stepping (I:7331-7679, the namelist default F:733) with interactions and with footloose calving
(SURVEY 8 f4): every accel call of the four stages evaluates interactive_force with the berg's own *_old position and
the stage velocity (I:2152-2161, I:2216-2227); positions are written in the first loop of evolve_icebergs, *_old in the
second (I:7178-7197).  The CUDA path against the CPU oracle on the reference's collision test."""
import numpy as np
import pytest

from common import by_id
from icebergs_b200 import api
from icebergs_b200 import synthetic as S
from test_interactions_gpu import Pair, bond_set

pytestmark = pytest.mark.gpu


def test_rk4_collision_test_first_steps():
    params = lambda: S.collision_params(api.default_params, runge_not_verlet=1)
    p = Pair(S.collision_bergs(), params)
    assert p.b.count_bergs() == 16 == p.o.count_bergs()
    assert bond_set(p.b.get_bonds()) == bond_set(p.o.get_bonds())
    p.step(1)
    p.check("RK4, one step", rtol=1e-10)
    p.step(49)
    p.check("RK4, 50 steps", rtol=1e-9)
    p.end()


def test_rk4_collision_test_through_contact():
    """the conglomerates meet after ~500 steps: contact springs and damping inside the RK stages"""
    params = lambda: S.collision_params(api.default_params, runge_not_verlet=1)
    p = Pair(S.collision_bergs(), params)
    gaps = []
    for k in range(12):
        p.step(50)
        # (after the contact at ~500 steps rounding differences grow about tenfold per 100 steps, as with Verlet stepping)
        # (byn = ay - ayn/2 is a small difference of large terms: 1e-13 m/s2 absolute is already 1e-6 relative late in the approach)
        p.check(f"RK4, {50 * (k + 1)} steps", rtol=1e-6 if k < 6 else (1e-5 if k < 9 else 1e-3))
        g = by_id(p.b.get_bergs(["id", "lat"]))
        half = g["lat"] < 10.0e3
        gaps.append(float(g["lat"][~half].min() - g["lat"][half].max()))
    assert min(gaps) < 1500.0, f"the conglomerates never met: {gaps}"
    assert p.b.count_bergs() == 16 == p.o.count_bergs()
    p.end()


def test_rk4_unbonded_interacting_bergs_old_predictive_corrective():
    """no bonds (contact forces only), the original predictor-corrector (use_new_predictive_corrective=F: the drag and
    the damping use the mean of the old and the new velocity, I:2196-2203)"""
    params = lambda: S.collision_params(api.default_params, runge_not_verlet=1, iceberg_bonds_on=0, manually_initialize_bonds=0,
                                        max_bonds=0, use_new_predictive_corrective=0)
    p = Pair(S.collision_bergs(), params, bonds=False)
    for k in range(8):
        p.step(50)
        p.check(f"RK4 unbonded, {50 * (k + 1)} steps", rtol=1e-6)
    p.end()


def test_rk4_with_footloose_is_refused():
    """RK4 + footloose is the reference's own FATAL ('Runge_not_Verlet must be false to use MTS, DEM, or footloose!',
    F:1485-1488): refused at init with the same text"""
    with pytest.raises(api.KidFatal, match="footloose"):
        g = S.CartesianGrid()
        api.icebergs_init(20, 20, 10.0, (1, 0.0), params=S.footloose_params(api.default_params, runge_not_verlet=1),
                          domain=api.Domain.single(20, 20, halo=3, cyclic_x=True), capacity=1024, **g.init_args())
