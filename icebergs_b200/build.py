"""In-tree build of the CUDA library (nvcc, sm_100a only).  `python -m icebergs_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "kid_b200.cu")
DEPS = [os.path.join(_HERE, "csrc", f) for f in
        ("kid_b200.cu", "kid_kernels.cuh", "kid_physics.cuh", "kid_geom.cuh", "kid_device.cuh", "kid_comm.cuh",
         "kid_interact.cuh", "kid_spread.cuh", "kid_mts.cuh", "kid_sort.cuh", "kid_step_tma.cuh", "kid_traj.cuh")]
DEPS.append(os.path.join(os.path.dirname(_HERE), "include", "kid_b200.h"))
OUT = os.path.join(_HERE, "lib", "libkid_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # no implicit FMA contraction: expressions whose exact cancellation the reference relies on keep
              # the plain IEEE sequence; interpolation and momentum sums use explicit fma() (kid_physics.cuh)
              "-fmad=false",
              "-shared", "-Xcompiler", "-fPIC", "-ldl"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("KID_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libkid_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
