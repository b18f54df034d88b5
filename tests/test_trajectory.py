"""Trajectory sampling, SURVEY 8(f2): record_posn F:5328-5498 and the iceberg_trajectories.nc schema of write_trajectory
(fmsio:1575-2047).  The CUDA path against the CPU oracle: which bergs are sampled (area thresholds, bonded bergs,
save_all_traj_year) is bit-exact, the sampled state -- including the environment thermodynamics left on the berg
(this%uo ... this%hi, I:2890) -- within 1e-10."""
import numpy as np
import pytest

from common import Case, run_gpu, run_oracle
from icebergs_b200 import api
from icebergs_b200 import restart_io as R
from icebergs_b200 import synthetic as S

F64 = ("lon", "lat", "day", "mass", "start_mass", "thickness", "mass_of_bits", "uvel", "vvel", "mass_scaling", "uvel_prev", "vvel_prev",
       "heat_density", "width", "length", "uo", "vo", "ui", "vi", "ua", "va", "ssh_x", "ssh_y", "sst", "sss", "cn", "hi", "axn", "ayn",
       "bxn", "byn", "halo_berg", "static_berg", "od")


def _keyed(t):
    order = np.lexsort((t["day"], t["year"], t["id"]))
    return {k: v[order] for k, v in t.items()}


def _compare(got, want, names=F64, rtol=1e-10, context=""):
    assert len(got["id"]) == len(want["id"]), f"{context}: {len(got['id'])} samples on the GPU, {len(want['id'])} in the oracle"
    g, w = _keyed(got), _keyed(want)
    assert np.array_equal(g["id"], w["id"]) and np.array_equal(g["year"], w["year"]), f"{context}: which bergs were sampled"
    for k in names:
        scale = np.maximum(np.maximum(np.abs(g[k]), np.abs(w[k])), 1e-300)
        err = np.where(g[k] == w[k], 0.0, np.abs(g[k] - w[k]) / scale)
        floor = 1e-13 * max(np.abs(w[k]).max(), 1e-300) if len(w[k]) else 0.0
        bad = (err > rtol) & (np.abs(g[k] - w[k]) > floor)
        assert not bad.any(), f"{context}: {k} rel err {err[bad].max():.3e}"


@pytest.mark.gpu
def test_record_posn_free_drift_matches_oracle():
    """Free-drifting bergs with an area threshold that selects a part of the population; samples after steps 1, 3 and 4
    (three records for a berg that stays above the threshold), then save_all_traj_year in the past: everything."""
    case = Case(96, 48, 3000, traj_area_thres=0.05)             # km^2: the smaller mass classes fall below it
    b, o = case.make_gpu(), case.make_oracle()
    for k in range(4):
        t = (1, 0.5 * k)
        run_gpu(b, case, t); run_oracle(o, case, t)
        if k != 1:
            b.record_posn(); o.record_posn()
    assert 0 < len(b.get_trajectory(clear=False)["id"]) < 3 * b.count_bergs()
    _compare(b.get_trajectory(), o.get_trajectory(), context="area threshold")
    assert len(b.get_trajectory()["id"]) == 0                   # cleared
    api.icebergs_end(b); o.close()
    case = Case(96, 48, 2000, traj_area_thres=1.0e9, save_all_traj_year=0.0)
    b, o = case.make_gpu(), case.make_oracle()
    run_gpu(b, case, (1, 0.0)); run_oracle(o, case, (1, 0.0))
    b.record_posn(); o.record_posn()
    got, want = b.get_trajectory(), o.get_trajectory()
    assert len(got["id"]) == b.count_bergs()
    _compare(got, want, context="save_all_traj_year")
    api.icebergs_end(b); o.close()


@pytest.mark.gpu
def test_record_posn_bonded_dem_conglomerates():
    """iKID collision bergs (mts, dem, bonds): bonded bergs are always sampled; the record carries the MTS environment
    cache, the fast accelerations, n_bonds and the rotation state; written and read back in the reference's schema."""
    from test_interactions_gpu import Pair
    from test_mts_gpu import IKID, MTS_KID
    over = dict(MTS_KID); over.update(IKID); over.update(traj_area_thres=1.0e9)
    p = Pair(S.collision_bergs(), lambda: S.collision_params(api.default_params, **over))
    for k in range(3):
        p.step(2)
        p.b.record_posn(); p.o.record_posn()
    got, want = p.b.get_trajectory(), p.o.get_trajectory()
    assert len(got["id"]) == 3 * 16
    names = tuple(k for k in F64 if k not in ("axn", "ayn", "bxn", "byn")) + ("axn_fast", "ayn_fast", "bxn_fast", "byn_fast")
    _compare(got, want, names=[k for k in names if not k.endswith("_fast")], rtol=1e-8, context="iKID")
    g, w = _keyed(got), _keyed(want)
    assert np.array_equal(g["n_bonds"], w["n_bonds"]) and g["n_bonds"].max() >= 3
    p.end()


def test_trajectory_file_schema(tmp_path):
    """write_trajectory's variable lists (fmsio:1925-1990) for the short, fl and long forms, and a round trip."""
    n = 7
    rng = np.random.default_rng(5)
    names = {f[0] for f in __import__("icebergs_b200._cdefs", fromlist=["x"]).KidTrajColumns._fields_}
    traj = {k: rng.random(n) for k in names}
    traj["id"] = (np.arange(n, dtype=np.int64) % 3 + 1) * 2 ** 32 + 17
    traj["year"] = np.full(n, 3, dtype=np.int32); traj["n_bonds"] = np.arange(n, dtype=np.int32)
    short = [v[0] for v in R.trajectory_variables(save_short_traj=True, save_fl_traj=False)]
    assert short == ["lon", "lat", "year", "day", "id_cnt", "id_ij"]
    fl = [v[0] for v in R.trajectory_variables(save_short_traj=True, save_fl_traj=True, footloose=True)]
    assert fl == short + ["mass", "start_mass", "thickness", "mass_of_bits", "uvel", "vvel", "mass_scaling", "mass_of_fl_bits",
                          "mass_of_fl_bergy_bits", "fl_k"]
    long_ = [v[0] for v in R.trajectory_variables(save_short_traj=False, save_fl_traj=True, mts=True, iceberg_bonds_on=True, dem=True)]
    assert long_[-8:] == ["axn_fast", "ayn_fast", "bxn_fast", "byn_fast", "n_bonds", "ang_vel", "ang_accel", "rot"] and "od" in long_
    p = str(tmp_path / "iceberg_trajectories.nc")
    R.write_trajectory(p, traj, save_short_traj=False, save_fl_traj=True, mts=True, iceberg_bonds_on=True, dem=True)
    back = R.read_trajectory(p)
    order = np.lexsort((traj["day"], traj["year"], traj["id"]))
    assert np.array_equal(back["id"], traj["id"][order]) and np.allclose(back["uo"], traj["uo"][order])
    assert np.array_equal(back["n_bonds"], traj["n_bonds"][order])
