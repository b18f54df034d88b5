"""kid_prefetch_forcing: the next call's inputs announced early travel on a copy stream into a second staging set.
The berg state must be the one of the plain call sequence, bit for bit (the returned flux fields up to the order of
their atomic sums); a prefetch the next call does not match is dropped."""
import numpy as np
import pytest

from common import Case, by_id
from icebergs_b200 import api

pytestmark = pytest.mark.gpu

NAMES = ["id", "lon", "lat", "uvel", "vvel", "mass", "thickness", "ine", "jne", "mass_of_bits", "heat_density"]


def _forcing_sets(case, n):
    """n different forcing sets (scaled currents / winds, a moving warm patch) and calving that varies in time"""
    rng = np.random.default_rng(11)
    sets = []
    for k in range(n):
        f = {q: np.ascontiguousarray(v, dtype=np.float64).copy() for q, v in case.forcing.items()}
        f["uo"] *= 1.0 + 0.1 * k; f["vo"] *= 1.0 - 0.05 * k; f["tauxa"] *= 1.0 + 0.2 * k
        f["sst"] = f["sst"] + 0.3 * k
        f["calving"] = f["calving"] + (rng.random(f["calving"].shape) < 0.002) * 2.0e-6
        sets.append(f)
    return sets


def _run(b, f, c, h):
    api.icebergs_run(b, (1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"],
                     sss=f["sss"])


def _announce(b, f, c, h):
    api.icebergs_prefetch(b, c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])


def test_prefetched_calls_equal_plain_calls():
    case = Case(96, 48, 6000)
    sets = _forcing_sets(case, 4)
    plain, pre = case.make_gpu(), case.make_gpu()
    outs = {"plain": [], "pre": []}
    inout = [(f["calving"].copy(), f["calving_hflx"].copy()) for f in sets]
    for k, f in enumerate(sets):
        c, h = inout[k][0].copy(), inout[k][1].copy()
        _run(plain, f, c, h)
        outs["plain"].append((c, h))
    pairs = [(c.copy(), h.copy()) for c, h in inout]
    _announce(pre, sets[0], *pairs[0])
    for k, f in enumerate(sets):
        if k + 1 < len(sets):
            _announce(pre, sets[k + 1], *pairs[k + 1])         # while step k is still to run
        _run(pre, f, *pairs[k])
        outs["pre"].append(pairs[k])
    for k in range(len(sets)):
        # (the returned fields hold sums of per-berg melt fluxes added with atomics: equal up to the summation order)
        assert np.allclose(outs["plain"][k][0], outs["pre"][k][0], rtol=1e-12, atol=0.0), f"calving returned by call {k}"
        assert np.allclose(outs["plain"][k][1], outs["pre"][k][1], rtol=1e-12, atol=0.0), f"calving_hflx returned by call {k}"
    a, b = by_id(plain.get_bergs(NAMES)), by_id(pre.get_bergs(NAMES))
    assert len(a["id"]) == len(b["id"])
    for q in NAMES:
        assert np.array_equal(a[q], b[q]), q
    api.icebergs_end(plain); api.icebergs_end(pre)


def test_unmatched_prefetch_is_dropped():
    case = Case(96, 48, 3000)
    sets = _forcing_sets(case, 2)
    plain, pre = case.make_gpu(), case.make_gpu()
    c0, h0 = sets[0]["calving"].copy(), sets[0]["calving_hflx"].copy()
    _run(plain, sets[0], c0, h0)
    c1, h1 = sets[1]["calving"].copy(), sets[1]["calving_hflx"].copy()
    _announce(pre, sets[1], c1, h1)                          # announced: set 1 ...
    c2, h2 = sets[0]["calving"].copy(), sets[0]["calving_hflx"].copy()
    _run(pre, sets[0], c2, h2)                               # ... but the call brings set 0
    assert np.allclose(c0, c2, rtol=1e-12, atol=0.0) and np.allclose(h0, h2, rtol=1e-12, atol=0.0)
    a, b = by_id(plain.get_bergs(NAMES)), by_id(pre.get_bergs(NAMES))
    for q in NAMES:
        assert np.array_equal(a[q], b[q]), q
    api.icebergs_end(plain); api.icebergs_end(pre)
