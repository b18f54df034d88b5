"""icebergs_b200 -- B200-native per-timestep hot path of the KID iceberg model.

Public surface = the reference's own: ``icebergs_init``, ``icebergs_run``,
``icebergs_end`` (src/icebergs.F90:65-66).  All compute happens in the CUDA library
``lib/libkid_b200.so`` behind the C ABI of ``include/kid_b200.h``.
"""
from .api import (AGRID, BGRID_NE, CGRID_NE, Domain, Icebergs, KidFatal, default_params, icebergs_end,
                  icebergs_init, icebergs_run, set_params)

__all__ = ["icebergs_init", "icebergs_run", "icebergs_end", "Icebergs", "Domain", "KidFatal", "default_params",
           "set_params", "BGRID_NE", "CGRID_NE", "AGRID"]
