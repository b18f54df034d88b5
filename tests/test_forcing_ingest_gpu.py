"""Forcing ingest on the device (I:5203-5383, invert_tau_for_du I:8272-8296): every staggering the
reference accepts -- B-grid copies, C-grid velocities averaged to the corners, C-grid and A-grid wind stress --
with the stress inverted to a wind speed (tau_is_velocity=.false.), SST handed over in Kelvin, NaNs and land
values scrubbed.  The ingested fields are compared with the CPU oracle's field by field, then the bergs after
steps taken with them."""
import numpy as np
import pytest

from common import COMPARE_F64, Case, assert_bergs_match, grid_rel
from icebergs_b200 import _cdefs as D
from icebergs_b200 import api

pytestmark = pytest.mark.gpu

NAMES = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
INGESTED = (D.KID_FLD_UO, D.KID_FLD_VO, D.KID_FLD_UI, D.KID_FLD_VI, D.KID_FLD_UA, D.KID_FLD_VA, D.KID_FLD_SSH,
            D.KID_FLD_SST, D.KID_FLD_SSS, D.KID_FLD_CN, D.KID_FLD_HI)
FLUXES = (D.KID_FLD_FLOATING_MELT, D.KID_FLD_BERG_MELT, D.KID_FLD_BERGY_SRC, D.KID_FLD_BERGY_MELT)


def _forcing(case, kelvin, with_nans):
    f = {k: v.copy() for k, v in case.forcing.items()}
    # a wind STRESS (N/m2) instead of a wind speed: what the coupler passes when tau_is_velocity=.false.
    f["tauxa"] = 0.12 * np.cos(np.linspace(0.0, 5.0, f["tauxa"].size)).reshape(f["tauxa"].shape) + 0.03
    f["tauya"] = 0.08 * np.sin(np.linspace(0.0, 9.0, f["tauya"].size)).reshape(f["tauya"].shape)
    f["tauxa"][3, 5] = 0.0; f["tauya"][3, 5] = 0.0          # |tau| == 0: invert_tau_for_du's guarded branch
    if kelvin:
        f["sst"] = f["sst"] + 273.15
    if with_nans:                                            # I:5372-5381
        f["uo"][7, 9] = np.nan; f["cn"][11, 4] = np.nan; f["sst"][2, 2] = np.nan; f["tauxa"][20, 20] = np.nan
    return f


def _run(obj, run, f, stagger, stress_stagger):
    calving, hflx = f["calving"].copy(), f["calving_hflx"].copy()
    run(obj, (1, 0.0), calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hflx,
        f["cn"], f["hi"], stagger=stagger, stress_stagger=stress_stagger, sss=f["sss"])
    return calving, hflx


@pytest.mark.parametrize("stagger,stress_stagger,kelvin,with_nans", [
    (api.BGRID_NE, api.BGRID_NE, False, False),
    (api.CGRID_NE, api.CGRID_NE, True, False),
    (api.CGRID_NE, api.AGRID, True, True),
    (api.BGRID_NE, api.AGRID, False, True),
    (api.BGRID_NE, api.CGRID_NE, True, False),
])
def test_staggered_forcing_matches_oracle(stagger, stress_stagger, kelvin, with_nans):
    case = Case(96, 48, 5000, tau_is_velocity=0, old_bug_bilin=0)
    f = _forcing(case, kelvin, with_nans)
    b, o = case.make_gpu(), case.make_oracle()
    for step in range(3):
        cg, hg = _run(b, api.icebergs_run, f, stagger, stress_stagger)
        co, ho = _run(o, lambda obj, *a, **k: obj.run(*a, **k), f, stagger, stress_stagger)
        if step == 0:
            for fid in INGESTED:
                got, want = b.grid_field(fid), o.grid_field(fid)
                assert not np.isnan(got).any(), f"field {fid}: NaN survived the scrub"
                assert grid_rel(got, want) < 1e-14, f"ingested field {fid} (stagger {stagger}, stress {stress_stagger})"
            if kelvin:
                assert b.grid_field(D.KID_FLD_SST).max() < 50.0, "SST was not converted from Kelvin"
        ctx = f"stagger={stagger} stress_stagger={stress_stagger} kelvin={kelvin} step {step}"
        assert_bergs_match(b.get_bergs(NAMES), o.get_bergs(NAMES), context=ctx, rtol=1e-10 if step == 0 else 1e-9)
        for fid in FLUXES:
            assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-9, f"flux field {fid}, {ctx}"
        assert np.allclose(cg, co, rtol=1e-9, atol=1e-18) and np.allclose(hg, ho, rtol=1e-9, atol=1e-18)
    api.icebergs_end(b)
    o.close()


def test_unrecognised_stagger_is_fatal():
    """error_mesg('KID, iceberg_run', 'Unrecognized value of stagger!', FATAL), I:5261 / I:5314"""
    case = Case(48, 24, 100)
    b = case.make_gpu()
    f = case.forcing
    with pytest.raises(api.KidFatal, match="Unrecognized value of stagger"):
        _run(b, api.icebergs_run, f, 7, api.BGRID_NE)
    with pytest.raises(api.KidFatal, match="Unrecognized value of stress_stagger"):
        _run(b, api.icebergs_run, f, api.BGRID_NE, 9)
    api.icebergs_end(b)
