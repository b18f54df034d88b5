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
    assert R.read_restart_calving_rmean(q) == (None, None)                 # a file without the running means (tau_calving = 0)
    rm, rmh = np.random.rand(6, 8), -np.random.rand(6, 8)
    R.write_restart_calving(q, si, sh, ic, rmean_calving=rm, rmean_calving_hflx=rmh)     # fmsio:568-569
    a, b_ = R.read_restart_calving_rmean(q)
    assert np.array_equal(a, rm) and np.array_equal(b_, rmh)
    assert np.array_equal(R.read_restart_calving(q)[1], sh)


def test_calving_restart_in_the_fms_layout(tmp_path):
    """calving.res.nc as FMS writes it (fmsio:564-566): compute-domain points only, global index space, a Time record
    axis.  A file shaped like that is cut to each rank's compute domain and padded to the data domain the library holds;
    what a rank writes back has the same layout (the halo-padded, Time-less file was not the reference's)."""
    from scipy.io import netcdf_file
    from icebergs_b200 import api
    gni, gnj, nk = 12, 8, 10
    rng = np.random.default_rng(3)
    gsi, gsh = rng.random((nk, gnj, gni)), rng.random((gnj, gni))
    gic = np.arange(gnj * gni, dtype=np.int32).reshape(gnj, gni)
    q = str(tmp_path / "calving.res.nc")
    f = netcdf_file(q, "w", version=1)               # the fixture: written here field by field, not by the module under test
    f.createDimension("Time", None)
    f.createDimension("xaxis_1", gni); f.createDimension("yaxis_1", gnj); f.createDimension("zaxis_1", nk)
    v = f.createVariable("Time", "d", ("Time",)); v[0] = 1.0
    v = f.createVariable("stored_ice", "d", ("Time", "zaxis_1", "yaxis_1", "xaxis_1")); v[0] = gsi
    v = f.createVariable("stored_heat", "d", ("Time", "yaxis_1", "xaxis_1")); v[0] = gsh
    v = f.createVariable("iceberg_counter_grd", "i", ("Time", "yaxis_1", "xaxis_1")); v[0] = gic
    f.close()
    merged_si, merged_sh, merged_ic = np.zeros_like(gsi), np.zeros_like(gsh), np.zeros_like(gic)
    for rank in range(2):
        d = api.Domain.decomposed(gni, gnj, rank, 2, halo=2)
        si, sh, ic = R.read_restart_calving(q, domain=d)
        h = d.halo
        assert si.shape == (nk, d.njd, d.nid) and sh.shape == (d.njd, d.nid) and ic.shape == (d.njd, d.nid)
        js, is_ = slice(d.jsc - 1, d.jec), slice(d.isc - 1, d.iec)
        assert np.array_equal(si[:, h:h + d.njc, h:h + d.nic], gsi[:, js, is_])
        assert np.array_equal(ic[h:h + d.njc, h:h + d.nic], gic[js, is_])
        assert si[:, :h].sum() == 0 and si[:, :, :h].sum() == 0            # halos: zero until the first halo update
        out = str(tmp_path / f"calving.res.nc.{rank:04d}")
        R.write_restart_calving(out, si, sh, ic, domain=d, global_file=True)
        g = netcdf_file(out, "r", mmap=False)
        assert g.variables["stored_ice"].dimensions == ("Time", "zaxis_1", "yaxis_1", "xaxis_1")
        assert g.variables["stored_ice"].shape[1:] == (nk, gnj, gni)
        merged_si += np.array(g.variables["stored_ice"][0]); merged_sh += np.array(g.variables["stored_heat"][0])
        merged_ic += np.array(g.variables["iceberg_counter_grd"][0])
        g.close()
        # a per-tile file (compute domain only) is read back onto the same rank
        tile = str(tmp_path / f"tile{rank}.nc")
        R.write_restart_calving(tile, si, sh, ic, domain=d)
        si2, sh2, ic2 = R.read_restart_calving(tile, domain=d)
        assert np.array_equal(si2, si) and np.array_equal(sh2, sh) and np.array_equal(ic2, ic)
    assert np.array_equal(merged_si, gsi) and np.array_equal(merged_sh, gsh) and np.array_equal(merged_ic, gic)


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
    R.write_restart_calving(str(tmp_path / "calving.res.nc"), si, sh, ic, domain=a.domain, global_file=True)
    for _ in range(3):
        run_gpu(a, case)
    b = api.icebergs_init(case.gni, case.gnj, case.dt, (1, 0.0), params=case.params(), domain=case.domain(), capacity=case.capacity,
                          **case.init)
    b.set_calving_state(*R.read_restart_calving(str(tmp_path / "calving.res.nc"), domain=b.domain))
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
