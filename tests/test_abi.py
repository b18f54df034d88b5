"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/kid_b200.h declares, refuses to compute without a device, and carries the
reference's namelist defaults."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from icebergs_b200 import _cdefs as D
from icebergs_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "kid_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kid_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kid_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_struct_mirror_sizes_are_consistent():
    # _cdefs parses the header itself; spot-check layout facts the ABI relies on
    assert C.sizeof(D.KidBergColumns) == 8 * len(D.KidBergColumns._fields_)
    assert D.KID_NCLASSES == 10
    assert D.KidParams.initial_mass_s.size == 80


def test_reference_namelist_defaults():
    p = api.default_params()
    # F:686-822
    assert p.halo == 4 and p.rho_bergs == 850.0 and p.Lx == 360.0 and p.Rearth == 6360000.0
    assert p.runge_not_verlet == 1 and p.old_bug_bilin == 1 and p.use_roundoff_fix == 1
    assert p.spring_coef == 1e-8 and p.radial_damping_coef == 1e-4 and p.tangental_damping_coef == 2e-5
    assert p.LoW_ratio == 1.5 and p.bergy_bit_erosion_fraction == 0.0 and p.use_operator_splitting == 1
    assert p.allow_bergs_to_roll == 1 and p.use_updated_rolling_scheme == 0 and p.h_to_init_grounding == 100.0
    assert list(p.initial_mass_s)[:3] == [8.8e7, 4.1e8, 3.3e9] and list(p.mass_scaling_n)[:3] == [200, 50, 25]
    assert abs(sum(p.distribution_s) - 0.99) < 1e-12


def test_domain_layout_matches_mpp_define_layout():
    # 8 ranks on 1440x720 -> layout (4,2) (idiv = nint(sqrt(npes*ni/nj)))
    seen = set()
    for r in range(8):
        d = api.Domain.decomposed(1440, 720, r, 8)
        assert (d.layout_x, d.layout_y) == (4, 2)
        assert d.nic == 360 and d.njc == 360
        seen.add((d.isc, d.jsc))
        px, py = r % 4, r // 4
        assert d.pe_E == (px + 1) % 4 + 4 * py and d.pe_W == (px - 1) % 4 + 4 * py
        assert d.pe_N == (px + 4 * (py + 1) if py == 0 else -1)
        assert d.pe_S == (px + 4 * (py - 1) if py == 1 else -1)
    assert len(seen) == 8
    d = api.Domain.single(20, 20, halo=3, cyclic_x=True)
    assert (d.isd, d.ied, d.jsd, d.jed) == (-2, 23, -2, 23) and d.pe_E == 0 and d.pe_N == -1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from icebergs_b200 import synthetic as S
    g = S.Grid(24, 12)
    with pytest.raises(api.KidFatal) as e:
        api.icebergs_init(24, 12, 3600.0, (1, 0.0), params=S.workload_params(api.default_params), **g.init_args())
    assert e.value.code == D.KID_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "icebergs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "kid_oracle" not in txt.replace("oracle/kid_oracle.c halo_update", ""), f"{f} references the oracle"
