#!/usr/bin/env python
"""Writes the fixtures of tests/golden/.

1. known_answers.json -- the portable known answers the REFERENCE's own unit tests hold for this
   path, transcribed with the file:line they come from (the reference is Fortran + FMS and cannot
   be executed in this image, so its numbers are transcribed, not generated).
2. oracle_step_48x24.npz -- a small seeded case stepped by the CPU oracle (regression fixture for
   the oracle itself: NOT a reference pin; it freezes the restated arithmetic so that a later edit
   of oracle/kid_oracle.c that changes results is noticed).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

S3 = 3.0 ** 0.5


def known_answers():
    H = 1.0
    S = 2.0 * H / S3
    A = (3.0 * S3 / 2.0) * S * S          # I:280
    return {
        "_source": "NOAA-GFDL/icebergs unit tests, transcribed (I: = src/icebergs.F90, F: = src/icebergs_framework.F90)",
        "id_round_trip": {"cite": "F:7319-7325", "i": 1440 * 1080, "counter": 2 ** 30 + 2 ** 4 + 1,
                          "id": (2 ** 30 + 2 ** 4 + 1) * 2 ** 32 + 1440 * 1080},
        "bilin_corners": {"cite": "F:7313-7316",
                          "cases": [{"xi": 0.0, "yj": 1.0, "corner": "NW(i-1,j)"}, {"xi": 1.0, "yj": 1.0, "corner": "NE(i,j)"},
                                    {"xi": 1.0, "yj": 0.0, "corner": "SE(i,j-1)"}, {"xi": 0.0, "yj": 0.0, "corner": "SW(i-1,j-1)"}]},
        "yearday": {"cite": "F:4431-4441", "cases": [[1, 1, 0, 0, 0, 0.0], [3, 2, 12, 0, 0, 63.5], [12, 31, 0, 0, 0, 371.0]]},
        "point_in_triangle": {"cite": "I:234-242", "A": [-2.695732526092343e-12, 0.204344508198090],
                              "B": [-2.695750202346321e-12, -8.433062639672301e-02],
                              "C": [0.249999999997304, 6.000694090068343e-02], "q": [0.0, 0.0], "inside": True},
        "hexagon": {"cite": "I:261-348", "H": H, "theta": 0.0, "tol": 1e-10, "area": A,
                    "cases": [
                        {"name": "origin", "x0": 0.0, "y0": 0.0, "Q": [A / 4, A / 4, A / 4, A / 4]},
                        {"name": "2a x>0", "x0": S, "y0": 0.0, "Q": [A / 2, 0.0, 0.0, A / 2]},
                        {"name": "2b x<0", "x0": -S, "y0": 0.0, "Q": [0.0, A / 2, A / 2, 0.0]},
                        {"name": "2c y>0", "x0": 0.0, "y0": H, "Q": [A / 2, A / 2, 0.0, 0.0]},
                        {"name": "2d y<0", "x0": 0.0, "y0": -H, "Q": [0.0, 0.0, A / 2, A / 2]},
                        {"name": "3a corners x>0", "x0": S / 2, "y0": 0.0, "Q": [2.5 * A / 6, 0.5 * A / 6, 0.5 * A / 6, 2.5 * A / 6]},
                        {"name": "3b corners x<0", "x0": -S / 2, "y0": 0.0, "Q": [0.5 * A / 6, 2.5 * A / 6, 2.5 * A / 6, 0.5 * A / 6]},
                    ]},
        "restart_counts": {"cite": "tests/collision_tests/README:16-22, tests/footloose_tests/input.nml:1, tests/dem_ground_frac_test/input.nml:7-10",
                           "KID": 16, "MTS_KID": 16, "iKID": 16, "footloose": 12, "ground_frac": 69},
        "calving_tables": {"cite": "F:787-796",
                           "initial_mass_s": [8.8e7, 4.1e8, 3.3e9, 1.8e10, 3.8e10, 7.5e10, 1.2e11, 2.2e11, 3.9e11, 7.4e11],
                           "distribution_s": [0.24, 0.12, 0.15, 0.18, 0.12, 0.07, 0.03, 0.03, 0.03, 0.02],
                           "mass_scaling_s": [2000, 200, 50, 20, 10, 5, 2, 1, 1, 1],
                           "initial_thickness_s": [40., 67., 133., 175., 250., 250., 250., 250., 250., 250.]},
    }


def oracle_regression():
    from common import COMPARE_F64, Case, by_id, run_oracle
    case = Case(48, 24, 400)
    o = case.make_oracle()
    names = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
    run_oracle(o, case)
    o.step_again(3, 1, 0.0)
    b = by_id(o.get_bergs(names))
    return {k: b[k] for k in names}


def main():
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known_answers(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "oracle_step_48x24.npz"), **oracle_regression())
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
