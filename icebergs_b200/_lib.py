"""ctypes binding of libkid_b200.so (the C ABI declared in include/kid_b200.h).

The library is built in-tree by ``icebergs_b200.build.build()`` (nvcc, sm_100a).
There is no fallback: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _cdefs as D

_HERE = os.path.dirname(os.path.abspath(__file__))
# KID_B200_LIB: another build of the same library (kernel-tuning experiments: scratch/ variants); default in-tree
LIB_PATH = os.environ.get("KID_B200_LIB") or os.path.join(_HERE, "lib", "libkid_b200.so")

# every symbol include/kid_b200.h declares
EXPORTS = [
    "kid_default_params", "kid_single_domain", "kid_define_domain", "kid_init", "kid_set_bergs",
    "kid_get_bergs", "kid_count_bergs", "kid_set_bonds", "kid_get_bonds", "kid_set_calving_state",
    "kid_get_calving_state", "kid_run", "kid_prefetch_forcing", "kid_set_forcing", "kid_step_resident", "kid_last_timing",
    "kid_kernel_launches", "kid_get_counters", "kid_get_grid_field", "kid_stock", "kid_incr_mass",
    "kid_sort_bergs", "kid_synchronize", "kid_end", "kid_last_error", "kid_version",
    "kid_nccl_unique_id", "kid_nccl_init", "kid_nccl_destroy", "kid_pack_width",
    "kid_local_comm_create", "kid_local_comm_destroy", "kid_owner_rank", "kid_set_sort_phase", "kid_sorts_done", "kid_last_slow_count",
    "kid_unit_hexagon_into_quadrants", "kid_unit_point_in_triangle", "kid_record_posn", "kid_trajectory_count", "kid_get_trajectory",
    "kid_set_calving_rmean", "kid_get_calving_rmean",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). icebergs_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.kid_default_params.argtypes = [C.POINTER(D.KidParams)]
    lib.kid_default_params.restype = None
    lib.kid_single_domain.argtypes = [C.POINTER(D.KidDomain)] + [C.c_int32] * 6
    lib.kid_single_domain.restype = None
    lib.kid_define_domain.argtypes = [C.POINTER(D.KidDomain)] + [C.c_int32] * 8
    lib.kid_define_domain.restype = C.c_int32
    lib.kid_init.argtypes = [C.POINTER(_vp), C.POINTER(D.KidParams), C.POINTER(D.KidDomain), C.c_int32,
                             C.c_double, C.c_int64] + [_vp] * 9 + [C.c_int32]
    lib.kid_init.restype = C.c_int32
    lib.kid_set_bergs.argtypes = [_vp, C.c_int64, C.POINTER(D.KidBergColumns)]
    lib.kid_get_bergs.argtypes = [_vp, C.POINTER(C.c_int64), C.POINTER(D.KidBergColumns), C.c_int32]
    lib.kid_count_bergs.argtypes = [_vp, C.POINTER(C.c_int64)]
    lib.kid_set_bonds.argtypes = [_vp, C.c_int64, C.POINTER(D.KidBondColumns)]
    lib.kid_get_bonds.argtypes = [_vp, C.POINTER(C.c_int64), C.POINTER(D.KidBondColumns)]
    lib.kid_set_calving_state.argtypes = [_vp, _vp, _vp, _vp]
    lib.kid_get_calving_state.argtypes = [_vp, _vp, _vp, _vp]
    lib.kid_set_calving_rmean.argtypes = [_vp, _vp, _vp]
    lib.kid_get_calving_rmean.argtypes = [_vp, _vp, _vp]
    lib.kid_run.argtypes = [_vp, C.c_int32, C.c_double] + [_vp] * 12 + [C.c_int32, C.c_int32] + [_vp] * 4
    lib.kid_set_forcing.argtypes = [_vp] + [_vp] * 12 + [C.c_int32, C.c_int32, _vp]
    lib.kid_prefetch_forcing.argtypes = [_vp] + [_vp] * 13
    lib.kid_record_posn.argtypes = [_vp]
    lib.kid_trajectory_count.argtypes = [_vp, C.POINTER(C.c_int64)]
    lib.kid_get_trajectory.argtypes = [_vp, C.POINTER(C.c_int64), C.POINTER(D.KidTrajColumns), C.c_int32]
    lib.kid_step_resident.argtypes = [_vp, C.c_int32, C.c_int32, C.c_double]
    lib.kid_last_timing.argtypes = [_vp, _dp]
    lib.kid_kernel_launches.argtypes = [_vp]
    lib.kid_kernel_launches.restype = C.c_int64
    lib.kid_get_counters.argtypes = [_vp, C.POINTER(D.KidCounters)]
    lib.kid_get_grid_field.argtypes = [_vp, C.c_int32, _vp]
    lib.kid_stock.argtypes = [_vp, C.c_int32, _dp]
    lib.kid_incr_mass.argtypes = [_vp, _vp]
    lib.kid_sort_bergs.argtypes = [_vp]
    lib.kid_set_sort_phase.argtypes = [_vp, C.c_int32, C.c_int32]
    lib.kid_sorts_done.argtypes = [_vp]
    lib.kid_sorts_done.restype = C.c_int64
    lib.kid_last_slow_count.argtypes = [_vp]
    lib.kid_last_slow_count.restype = C.c_int64
    lib.kid_unit_hexagon_into_quadrants.argtypes = [C.c_int32] + [C.c_double] * 4 + [_dp]
    lib.kid_unit_point_in_triangle.argtypes = [C.c_int32, _dp, _ip, _dp]
    lib.kid_synchronize.argtypes = [_vp]
    lib.kid_end.argtypes = [C.POINTER(_vp)]
    lib.kid_last_error.argtypes = [_vp]
    lib.kid_last_error.restype = C.c_char_p
    lib.kid_version.restype = C.c_char_p
    lib.kid_nccl_unique_id.argtypes = [C.c_char_p, C.c_int32]
    lib.kid_nccl_init.argtypes = [C.POINTER(_vp), C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.kid_nccl_destroy.argtypes = [_vp]
    lib.kid_pack_width.argtypes = []
    lib.kid_local_comm_create.argtypes = [C.POINTER(_vp), C.c_int32]
    lib.kid_local_comm_destroy.argtypes = [_vp]
    lib.kid_owner_rank.argtypes = [C.POINTER(D.KidDomain), C.c_int32, C.c_int32]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:  # default: set explicit int32 status
            fn.restype = C.c_int32
    return lib
