"""The reference's restart files as the on-disk interface: NetCDF-3 classic files with
the variable names of src/icebergs_fmsio.F90:261-337 / 473-493 / 564-566."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from icebergs_b200 import restart_io as R
from icebergs_b200 import synthetic as S


def test_berg_restart_round_trip(tmp_path):
    cols, _ = S.Grid(96, 48).seed_bergs(500)
    cols["mass_of_fl_bits"] = np.linspace(0.0, 1.0e9, 500)
    cols["axn_fast"] = np.linspace(-1e-6, 1e-6, 500)          # mts / dem runs carry these too
    cols["rot"] = np.linspace(0.0, 0.3, 500)
    p = str(tmp_path / "icebergs.res.nc")
    R.write_restart_bergs(p, cols)
    f = netcdf_file(p, "r", mmap=False)
    assert f.dimensions["i"] is None                                   # unlimited, like the reference's files
    for name in ("lon", "lat", "uvel", "vvel", "mass", "axn", "ayn", "bxn", "byn", "ine", "jne", "thickness", "width", "length",
                 "start_lon", "start_lat", "start_year", "id_cnt", "id_ij", "start_day", "start_mass", "mass_scaling",
                 "mass_of_bits", "heat_density", "static_berg"):
        assert name in f.variables, name                               # fmsio:261-337
    f.close()
    back = R.read_restart_bergs(p)
    for k, v in cols.items():
        assert np.array_equal(back[k], v), k
    assert np.array_equal(back["id"], cols["id"])                      # id = id_cnt*2^32 + id_ij, F:7276-7296


def test_legacy_makeberg_file_has_no_ids(tmp_path):
    """The makeberg scripts of the reference tests write 'iceberg_num' instead of id_cnt/id_ij: no id column
    comes back, so kid_set_bergs generates the ids in file order (fmsio:841-845)."""
    p = str(tmp_path / "legacy.res.nc")
    b = S.footloose_bergs()
    f = netcdf_file(p, "w", version=1)
    f.createDimension("i", None)
    for name in ("lon", "lat", "uvel", "vvel", "mass", "thickness", "width", "length", "start_lon", "start_lat", "start_day",
                 "start_mass", "mass_scaling"):
        v = f.createVariable(name, "d", ("i",)); v[:2] = b[name]
    for name in ("ine", "jne", "start_year", "iceberg_num"):
        v = f.createVariable(name, "i", ("i",)); v[:2] = np.array([1, 2], dtype=np.int32)
    f.close()
    cols = R.read_restart_bergs(p, ignore_ij_restart=True)
    assert "id" not in cols and "ine" not in cols and np.array_equal(cols["lon"], b["lon"])
    assert np.array_equal(cols["axn"], np.zeros(2)) and np.array_equal(cols["static_berg"], np.zeros(2))


def test_bond_and_calving_round_trip(tmp_path):
    ids = np.array([5 * 2 ** 32 + 77, 9 * 2 ** 32 + 12], dtype=np.int64)
    bonds = dict(first_id=ids, other_id=ids[::-1].copy(), first_ine=np.array([3, 4], np.int32), first_jne=np.array([5, 6], np.int32),
                 other_ine=np.array([4, 3], np.int32), other_jne=np.array([6, 5], np.int32))
    bonds.update(tangd1=np.array([1e-3, -1e-3]), tangd2=np.array([2e-3, -2e-3]), nstress=np.array([5.0, 5.0]), sstress=np.array([1.0, 1.0]),
                 rel_rotation=np.array([1e-4, -1e-4]), broken=np.array([0, 1], np.int32))       # dem bond history, fmsio:484-493
    p = str(tmp_path / "bonds_iceberg.res.nc")
    R.write_restart_bonds(p, bonds)
    back = R.read_restart_bonds(p)
    for k, v in bonds.items():
        assert np.array_equal(back[k], v), k
    si, sh, ic = np.random.rand(10, 6, 8), np.random.rand(6, 8), np.arange(48, dtype=np.int32).reshape(6, 8)
    q = str(tmp_path / "calving.res.nc")
    R.write_restart_calving(q, si, sh, ic)
    a, b_, c = R.read_restart_calving(q)
    assert np.array_equal(a, si) and np.array_equal(b_, sh) and np.array_equal(c, ic)


@pytest.mark.gpu
def test_restart_through_files_continues(tmp_path):
    """Stop, write icebergs.res.nc + calving.res.nc, start a new handle from the files: the continued run
    equals the uninterrupted one.  Not bit for bit, in the reference either: the file holds lon/lat and the
    reader recomputes xi,yj from them (fmsio:871-880), an ulp-level perturbation; ids/cells are exact."""
    from common import Case, by_id, run_gpu
    from icebergs_b200 import api
    case = Case(96, 48, 5000)
    a = case.make_gpu()
    for _ in range(3):
        run_gpu(a, case)
    names = [v[0] for v in R.BERG_VARS if not v[0].startswith("id_")] + ["id"]
    R.write_restart_bergs(str(tmp_path / "icebergs.res.nc"), a.get_bergs(names))
    si, sh, ic = a.get_calving_state()
    R.write_restart_calving(str(tmp_path / "calving.res.nc"), si, sh, ic)
    for _ in range(3):
        run_gpu(a, case)
    b = api.icebergs_init(case.gni, case.gnj, case.dt, (1, 0.0), params=case.params(), domain=case.domain(), capacity=case.capacity,
                          **case.init)
    b.set_calving_state(*R.read_restart_calving(str(tmp_path / "calving.res.nc")))
    b.set_bergs(**R.read_restart_bergs(str(tmp_path / "icebergs.res.nc")))
    for _ in range(3):
        run_gpu(b, case)
    ga, gb = by_id(a.get_bergs(names)), by_id(b.get_bergs(names))
    for k in names:
        if ga[k].dtype.kind == "i":
            assert np.array_equal(ga[k], gb[k]), k
        else:
            assert np.allclose(ga[k], gb[k], rtol=1e-10, atol=1e-13), k
    api.icebergs_end(a); api.icebergs_end(b)
