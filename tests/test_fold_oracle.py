"""Tripolar fold (FOLD_NORTH_EDGE F:649, F:933, F:3138-3196, I:6110-6123) in the CPU oracle, checked against
things that do not depend on the oracle: the analytic continuation of a bipolar cap (icebergs_b200.synthetic.BipolarCapGrid:
cell (i, gnj+k) IS cell (gni+1-i, gnj+1-k) turned by 180 degrees) pins the halo mapping of every position the reference
updates (centre, corner, C-grid faces, B-grid vectors), and bergs that drift across the fold line must keep a continuous
track in the stereographic plane."""
import numpy as np

import kid_oracle_py as O
from icebergs_b200 import _cdefs as D
from icebergs_b200 import synthetic as S

GNI, GNJ, HALO = 90, 24, 4


def fold_domain(halo=HALO):
    d = O.SingleDomain(GNI, GNJ, halo=halo, cyclic_x=True)
    d.c.fold_north = 1
    d.c.pe_N = 0
    return d


def make(n=0, dt=21600.0, **over):
    g = S.BipolarCapGrid(GNI, GNJ)
    kw = dict(runge_not_verlet=0, bergy_bit_erosion_fraction=0.1, tau_is_velocity=1, old_bug_bilin=0, Rearth=S.REARTH,
              add_weight_to_ocean=0, grid_is_regular=0, halo=HALO)
    kw.update(over)
    p = O.default_params(**kw)
    d = fold_domain(kw["halo"])
    o = O.Oracle(GNI, GNJ, dt, (1, 0.0), params=p, domain=d, **g.init_args())
    bergs = None
    if n:
        bergs, counter = g.seed_bergs(n)
        c = np.zeros((d.njd, d.nid), dtype=np.int32)
        c[HALO:HALO + GNJ, HALO:HALO + GNI] = counter
        o.set_calving_state(iceberg_counter_grd=c)
        o.set_bergs(**bergs)
    return g, o, d, bergs


def run(o, g, **replace):
    f = dict(g.forcing())
    f.update(replace)
    calving, hflx = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, 0.0), calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hflx,
          f["cn"], f["hi"], sss=f["sss"])


def dd_index(d):
    i = np.arange(d.isd, d.ied + 1)
    j = np.arange(d.jsd, d.jed + 1)
    return np.meshgrid(i, j)


def test_fold_halos_of_the_grid_match_the_analytic_continuation():
    g, o, d, _ = make()
    i, j = dd_index(d)
    north = j > GNJ                                                    # the rows beyond the fold
    lon, lat = g.lonlat(i, j)
    got_lon, got_lat = o.grid_field(D.KID_FLD_LON), o.grid_field(D.KID_FLD_LAT)
    assert np.abs(got_lat - lat)[north].max() < 1e-9                   # corners (position=CORNER)
    assert np.abs(np.mod(got_lon - lon + 180.0, 360.0) - 180.0)[north].max() < 1e-9
    dx = g._dist(g._xyz(i - 1, j), g._xyz(i, j))                       # north faces
    dy = g._dist(g._xyz(i, j - 1), g._xyz(i, j))                       # east faces
    assert np.abs(o.grid_field(D.KID_FLD_DX) / dx - 1.0)[north].max() < 1e-9
    assert np.abs(o.grid_field(D.KID_FLD_DY) / dy - 1.0)[north].max() < 1e-9
    d1 = g._xyz(i, j) - g._xyz(i - 1, j - 1)
    d2 = g._xyz(i - 1, j) - g._xyz(i, j - 1)
    area = 0.5 * S.REARTH ** 2 * np.sqrt((np.cross(d1, d2, axis=0) ** 2).sum(axis=0))
    assert np.abs(o.grid_field(D.KID_FLD_AREA) / area - 1.0)[north].max() < 1e-9      # centres
    o.close()


def test_fold_halos_of_vectors_and_scalars_match_the_analytic_continuation():
    g, o, d, _ = make()
    run(o, g)
    i, j = dd_index(d)
    wide = S.BipolarCapGrid(GNI, GNJ, isc=d.isd + 1, iec=d.ied - 1, jsc=d.jsd + 1, jec=d.jed - 1)   # ring 1 = the data domain
    want = wide.forcing()
    wet = (g.__class__(GNI, GNJ, isc=d.isd + 1, iec=d.ied - 1, jsc=d.jsd + 1, jec=d.jed - 1).wet(1) > 0.5)
    north = (j > GNJ) & wet
    for fid, key in ((D.KID_FLD_UO, "uo"), (D.KID_FLD_VO, "vo"), (D.KID_FLD_UI, "ui"), (D.KID_FLD_SSH, "ssh"),
                     (D.KID_FLD_CN, "cn"), (D.KID_FLD_HI, "hi")):
        got = o.grid_field(fid)
        # B-grid components flip sign with the grid directions, scalars do not: both are what the continuation gives
        assert np.abs(got - want[key])[north].max() < 1e-9, key
    assert north.sum() > 200
    o.close()


def _plane(lon, lat):
    r = 2.0 * np.tan(0.5 * np.radians(90.0 - lat))
    return r * np.cos(np.radians(lon)), r * np.sin(np.radians(lon))


def test_bergs_cross_the_fold_on_a_continuous_track():
    g, o, d, bergs = make(n=600)
    names = ["lon", "lat", "ine", "jne", "id", "xi", "yj", "mass"]
    prev = o.get_bergs(names)
    order = np.argsort(prev["id"])
    prev = {k: v[order] for k, v in prev.items()}
    crossed = 0
    for step in range(14):
        run(o, g)
        cur = o.get_bergs(names)
        assert len(cur["id"]) == 600, f"step {step}: bergs lost at the fold"
        order = np.argsort(cur["id"])
        cur = {k: v[order] for k, v in cur.items()}
        assert (cur["jne"] <= GNJ).all() and (cur["ine"] >= 1).all() and (cur["ine"] <= GNI).all()
        assert (cur["xi"] > -1e-9).all() and (cur["xi"] < 1 + 1e-9).all() and (cur["yj"] > -1e-9).all() and (cur["yj"] < 1 + 1e-9).all()
        X0, Y0 = _plane(prev["lon"], prev["lat"])
        X1, Y1 = _plane(cur["lon"], cur["lat"])
        step_len = np.hypot(X1 - X0, Y1 - Y0) * S.REARTH
        # a berg that crossed sits in the mirror column on the other side of the fold line Y = 0
        hop = (np.sign(Y0) != np.sign(Y1)) & (prev["jne"] == GNJ) & (cur["jne"] == GNJ)
        crossed += int(hop.sum())
        # ... and did not jump on the way: < 2.5 m/s x dt (0.9 m/s stream, 7 m/s wind).  (Bergs that meet the land around
        # the geographic pole are put back by the coast bounce of adjust_index_and_ground: not part of this check.)
        open_water = hop & (cur["lat"] < 85.0)
        assert not open_water.any() or step_len[open_water].max() < 2.5 * 21600.0, step_len[open_water].max()
        assert (np.abs(cur["ine"][hop] - (GNI + 1 - prev["ine"][hop])) <= 1).all()
        # the cell the berg is filed under contains it
        for q in np.nonzero(hop)[0][:5]:
            assert O.lib().oracle_is_point_in_cell(o._h, cur["lon"][q], cur["lat"][q], int(cur["ine"][q]), int(cur["jne"][q]))
        prev = cur
    assert crossed > 40, crossed
    assert o.counters()["error_flags"] == 0
    o.close()


def test_mass_spread_across_the_fold_is_not_lost():
    """sum_up_spread_fields I:6077-6150: the weights a berg in row gnj puts on the cells beyond the fold come back, turned by
    180 degrees (I:6110-6123), to the cells on the other side -- without the fold they are lost in the halo"""
    tot = {}
    for fold in (1, 0):
        g = S.BipolarCapGrid(GNI, GNJ)
        bergs, _ = g.seed_bergs(600)
        p = O.default_params(runge_not_verlet=0, tau_is_velocity=1, old_bug_bilin=0, Rearth=S.REARTH, add_weight_to_ocean=1,
                             grid_is_regular=0, halo=HALO, use_old_spreading=1)
        d = fold_domain()
        d.c.fold_north = fold
        if not fold:
            d.c.pe_N = -1
        o = O.Oracle(GNI, GNJ, 21600.0, (1, 0.0), params=p, domain=d, **g.init_args())
        top = bergs["jne"] == GNJ
        o.set_bergs(**{k: v[top] for k, v in bergs.items()})
        f = g.forcing()
        run(o, g, **{k: 0.0 * f[k] for k in ("uo", "vo", "ui", "vi", "tauxa", "tauya", "ssh")})      # the bergs stay put
        sm = o.grid_field(D.KID_FLD_SPREAD_MASS)[HALO:HALO + GNJ, HALO:HALO + GNI]
        area = g.init_args()["ice_area"]
        tot[fold] = float((sm * area).sum())
        b = o.get_bergs(["mass", "mass_scaling", "mass_of_bits"])
        tot["bergs", fold] = float(((b["mass"] + b["mass_of_bits"]) * b["mass_scaling"]).sum())
        o.close()
    assert tot[1] > 1.02 * tot[0]                       # the top row's northward share is a good part of the total
    assert abs(tot[1] / tot["bergs", 1] - 1.0) < 1e-12  # nothing is lost
