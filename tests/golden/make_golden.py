#!/usr/bin/env python
"""Writes the fixtures of tests/golden/.

1. known_answers.json -- the portable known answers the REFERENCE's own unit tests hold for this
   path, transcribed with the file:line they come from (the reference is Fortran + FMS and cannot
   be executed in this image, so its numbers are transcribed, not generated).
2. oracle_step_48x24.npz -- a small seeded case stepped by the CPU oracle (regression fixture for
   the oracle itself: NOT a reference pin; it freezes the restated arithmetic so that a later edit
   of oracle/kid_oracle.c that changes results is noticed).
3. oracle_mts_20h.npz -- the same for the MTS / DEM restatement: tests/collision_tests with the
   reference's MTS_KID and iKID namelists (3600 s steps, 60 sub-steps) after 20 h.
4. oracle_fold_90x24.npz -- the same for the tripolar fold (FOLD_NORTH_EDGE): 400 bergs on the analytic bipolar
   cap after 8 steps of 6 h (a third of them cross the fold), the halo rows beyond the fold of a corner, a centre
   and a B-grid vector field, and the spread mass with its weights turned by 180 degrees.

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
        "dem_beams": {"cite": "tests/dem_ssbeam_test/animate_trajectories.py:143-157, tests/dem_cbeam_test/animate_trajectories.py:149-157",
                      "simply_supported": {"P": -1.5e5, "E": 1.0e9, "l": 14.0, "I": 0.5 ** 3 / 12.0, "w_mid": -1.5e5 * 14.0 ** 3 / (48.0 * 1.0e9 * 0.5 ** 3 / 12.0)},
                      "cantilever": {"P": -1.5e10, "E": 1.0e9, "l": 29 * 5000.0, "I": (3 * 5000.0) ** 3 / 12.0,
                                     "w_tip_linear": -1.5e10 * (29 * 5000.0) ** 3 / (3.0 * 1.0e9 * (3 * 5000.0) ** 3 / 12.0)}},
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


MTS_NML = dict(mts=1, mts_sub_steps=60, explicit_inner_mts=1, force_convergence=1, convergence_tolerance=1e-8,
               contact_distance=1.75e3, contact_spring_coef=1.0e-7)
IKID_NML = dict(MTS_NML, dem=1, poisson=0.3, dem_damping_coef=1.0, dem_spring_coef=4471.94)
MTS_NAMES = ["id", "lon", "lat", "uvel", "vvel", "axn_fast", "ayn_fast", "ang_vel", "rot", "ine", "jne"]


def mts_run(over, nsteps=20, dt=3600.0):
    import kid_oracle_py as O
    from icebergs_b200 import api
    from icebergs_b200 import synthetic as S
    g = S.CartesianGrid()
    dom = api.Domain.single(g.gni, g.gnj, halo=3, cyclic_x=True)
    o = O.Oracle(g.gni, g.gnj, dt, (1, 0.0), params=S.collision_params(api.default_params, **over), domain=dom, **g.init_args())
    o.set_bergs(**S.collision_bergs())
    o.set_bonds()
    f = g.forcing()
    for k in range(nsteps):
        c, h = f["calving"].copy(), f["calving_hflx"].copy()
        o.run((1, k * dt / 86400.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h,
              f["cn"], f["hi"], sss=f["sss"])
    b = o.get_bergs(MTS_NAMES)
    o.close()
    order = np.argsort(b["id"], kind="stable")
    return {k: v[order] for k, v in b.items()}


def mts_regression():
    out = {}
    for tag, over in (("mts_kid", MTS_NML), ("ikid", IKID_NML)):
        for k, v in mts_run(over).items():
            out[f"{tag}.{k}"] = v
    return out


def fold_run():
    import kid_oracle_py as O
    from icebergs_b200 import _cdefs as D
    from icebergs_b200 import synthetic as S
    gni, gnj, halo = 90, 24, 4
    g = S.BipolarCapGrid(gni, gnj)
    p = O.default_params(runge_not_verlet=0, bergy_bit_erosion_fraction=0.1, tau_is_velocity=1, old_bug_bilin=0, Rearth=S.REARTH,
                         add_weight_to_ocean=1, use_old_spreading=0, hexagonal_icebergs=1, grid_is_regular=0, halo=halo)
    d = O.SingleDomain(gni, gnj, halo=halo, cyclic_x=True)
    d.c.fold_north = 1
    d.c.pe_N = 0
    o = O.Oracle(gni, gnj, 21600.0, (1, 0.0), params=p, domain=d, **g.init_args())
    bergs, counter = g.seed_bergs(400)
    c = np.zeros((d.njd, d.nid), dtype=np.int32)
    c[halo:halo + gnj, halo:halo + gni] = counter
    o.set_calving_state(iceberg_counter_grd=c)
    o.set_bergs(**bergs)
    f = g.forcing()
    for _ in range(8):
        cv, hf = f["calving"].copy(), f["calving_hflx"].copy()
        o.run((1, 0.0), cv, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hf, f["cn"], f["hi"], sss=f["sss"])
    names = ["id", "lon", "lat", "uvel", "vvel", "xi", "yj", "mass", "mass_of_bits", "ine", "jne"]
    b = o.get_bergs(names)
    order = np.argsort(b["id"], kind="stable")
    out = {f"berg.{k}": v[order] for k, v in b.items()}
    for name, fid in (("lat", D.KID_FLD_LAT), ("area", D.KID_FLD_AREA), ("uo", D.KID_FLD_UO), ("vo", D.KID_FLD_VO)):
        out[f"halo.{name}"] = o.grid_field(fid)[halo + gnj:, :]              # the rows beyond the fold
    out["spread_mass"] = o.grid_field(D.KID_FLD_SPREAD_MASS)[halo:halo + gnj, halo:halo + gni]
    o.close()
    return out


def main():
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known_answers(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "oracle_step_48x24.npz"), **oracle_regression())
    np.savez_compressed(os.path.join(HERE, "oracle_mts_20h.npz"), **mts_regression())
    np.savez_compressed(os.path.join(HERE, "oracle_fold_90x24.npz"), **fold_run())
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
