"""Tripolar fold (FOLD_NORTH_EDGE F:649, F:933, F:3138-3196, I:6110-6123) through the CUDA library: halo rows
beyond the folded edge for every position the reference updates, bergs handed across the fold (on one rank and between
ranks), and the 180-degree turn of the mass-spreading weights -- against the CPU oracle, whose fold is pinned by the
analytic continuation of the bipolar cap (tests/test_fold_oracle.py)."""
import ctypes as C

import numpy as np
import pytest

from common import COMPARE_F64, Case, assert_bergs_match, grid_rel, run_gpu, run_oracle
from icebergs_b200 import _cdefs as D
from icebergs_b200 import api, parallel
from icebergs_b200 import synthetic as S
from test_multirank_gpu import Ranks

pytestmark = pytest.mark.gpu

GNI, GNJ = 90, 24
NAMES = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
STATIC = dict(lon=D.KID_FLD_LON, lat=D.KID_FLD_LAT, lonc=D.KID_FLD_LONC, latc=D.KID_FLD_LATC, dx=D.KID_FLD_DX, dy=D.KID_FLD_DY,
              area=D.KID_FLD_AREA, msk=D.KID_FLD_MSK, cos=D.KID_FLD_COS, sin=D.KID_FLD_SIN, depth=D.KID_FLD_OCEAN_DEPTH)
FORCING = dict(uo=D.KID_FLD_UO, vo=D.KID_FLD_VO, ui=D.KID_FLD_UI, vi=D.KID_FLD_VI, ua=D.KID_FLD_UA, va=D.KID_FLD_VA,
               ssh=D.KID_FLD_SSH, sst=D.KID_FLD_SST, cn=D.KID_FLD_CN, hi=D.KID_FLD_HI)


class FoldCase(Case):
    def __init__(self, n, dt=21600.0, **over):
        kw = dict(grid_is_regular=0, old_bug_bilin=0)
        kw.update(over)
        super().__init__(GNI, GNJ, n, dt=dt, grid=S.BipolarCapGrid(GNI, GNJ), **kw)

    def domain(self):
        d = api.Domain.single(self.gni, self.gnj, halo=self.halo, cyclic_x=True)
        d.c.fold_north = 1
        d.c.pe_N = 0                    # mpp_get_neighbor_pe across the fold: this PE itself
        return d


def layout_domain(grp, lx, ly, rank, halo):
    """mpp_define_domains for an lx x ly layout of the folded grid: even split, cyclic in x, and -- as FMS hands it out --
    the PE on the other side of the fold as the northern neighbour of the top row"""
    d = D.KidDomain()
    px, py = rank % lx, rank // lx
    xs = [k * (GNI // lx) + min(k, GNI % lx) + 1 for k in range(lx + 1)]
    ys = [k * (GNJ // ly) + min(k, GNJ % ly) + 1 for k in range(ly + 1)]
    d.gni, d.gnj = GNI, GNJ
    d.isc, d.iec, d.jsc, d.jec = xs[px], xs[px + 1] - 1, ys[py], ys[py + 1] - 1
    d.isd, d.ied, d.jsd, d.jed = d.isc - halo, d.iec + halo, d.jsc - halo, d.jec + halo
    d.cyclic_x, d.cyclic_y, d.fold_north = 1, 0, 1
    d.rank, d.nranks, d.layout_x, d.layout_y = rank, lx * ly, lx, ly
    d.pe_E, d.pe_W = (px + 1) % lx + lx * py, (px - 1) % lx + lx * py
    d.pe_S = px + lx * (py - 1) if py > 0 else -1
    d.pe_N = px + lx * (py + 1) if py < ly - 1 else (lx - 1 - px) + lx * py
    d.device = grp.devices[rank]
    d.nccl_comm = grp._g.value
    d.comm_kind = D.KID_COMM_LOCAL
    return api.Domain(d)


def test_static_halos_beyond_the_fold_match_oracle():
    case = FoldCase(0)
    b, o = case.make_gpu(), case.make_oracle()
    for name, fid in STATIC.items():
        got, want = b.grid_field(fid), o.grid_field(fid)
        assert np.array_equal(got, want), f"{name}: max diff {np.abs(got - want).max()}"
    api.icebergs_end(b)
    o.close()


@pytest.mark.parametrize("stagger,stress", [(D.KID_BGRID_NE, D.KID_BGRID_NE), (D.KID_CGRID_NE, D.KID_CGRID_NE),
                                            (D.KID_BGRID_NE, D.KID_AGRID)])
def test_forcing_halos_beyond_the_fold_match_oracle(stagger, stress):
    """B-grid vectors (sign flip, fold row, pole points), C-grid and A-grid stress pairs, centred scalars"""
    case = FoldCase(0, tau_is_velocity=0)
    b, o = case.make_gpu(), case.make_oracle()
    calving, hflx, f = case.run_args()
    api.icebergs_run(b, (1, 0.0), calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hflx,
                     f["cn"], f["hi"], stagger=stagger, stress_stagger=stress, sss=f["sss"])
    calving, hflx, f = case.run_args()
    o.run((1, 0.0), calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hflx, f["cn"],
          f["hi"], stagger=stagger, stress_stagger=stress, sss=f["sss"])
    for name, fid in FORCING.items():
        got, want = b.grid_field(fid), o.grid_field(fid)
        assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1.0), f"{name}: max diff {np.abs(got - want).max()}"
    api.icebergs_end(b)
    o.close()


@pytest.mark.parametrize("rk", [0, 1])
def test_bergs_cross_the_fold_like_the_oracle(rk):
    case = FoldCase(3000, runge_not_verlet=rk)
    b, o = case.make_gpu(), case.make_oracle()
    crossed = 0
    for step in range(12):
        before = b.get_bergs(["ine", "jne", "id", "lat"])
        run_gpu(b, case); run_oracle(o, case)
        got, want = b.get_bergs(NAMES), o.get_bergs(NAMES)
        assert_bergs_match(got, want, rtol=1e-10 if step == 0 else 1e-8, context=f"fold rk={rk} step {step}")
        ob, og = np.argsort(before["id"]), np.argsort(got["id"])
        hop = (before["jne"][ob] == GNJ) & (got["jne"][og] == GNJ) & (np.abs(before["ine"][ob] + got["ine"][og] - (GNI + 1)) <= 1) \
            & (np.abs(before["ine"][ob] - got["ine"][og]) > 2)
        crossed += int(hop.sum())
        for fid in (D.KID_FLD_FLOATING_MELT, D.KID_FLD_BERG_MELT, D.KID_FLD_BERGY_SRC):
            assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-9
    assert crossed > 150, crossed
    cg, co = b.counters(), o.counters()
    assert cg["error_flags"] == 0 and cg["nbergs"] == co["nbergs"] == 3000
    api.icebergs_end(b)
    o.close()


@pytest.mark.parametrize("old_bug", [0, 1])
def test_mass_spread_across_the_fold_matches_oracle_and_is_conserved(old_bug):
    """old_bug_rotated_weights (F:38) skips the 180-degree turn of the weights beyond the fold (I:6110): both ways"""
    case = FoldCase(3000, add_weight_to_ocean=1, use_old_spreading=0, hexagonal_icebergs=1, pass_fields_to_ocean_model=1,
                    old_bug_rotated_weights=old_bug)
    b, o = case.make_gpu(), case.make_oracle()
    for step in range(3):
        run_gpu(b, case); run_oracle(o, case)
        for fid in (D.KID_FLD_SPREAD_MASS, D.KID_FLD_SPREAD_AREA, D.KID_FLD_SPREAD_UVEL, D.KID_FLD_SPREAD_VVEL, D.KID_FLD_USTAR_ICEBERG):
            assert grid_rel(b.grid_field(fid), o.grid_field(fid)) < 1e-10, (fid, step)
    hl = case.halo
    sm = b.grid_field(D.KID_FLD_SPREAD_MASS)[hl:hl + GNJ, hl:hl + GNI]
    got = b.get_bergs(["mass", "mass_of_bits", "mass_scaling"])
    total = ((got["mass"] + got["mass_of_bits"]) * got["mass_scaling"]).sum()
    if not old_bug:
        assert abs((sm * case.init["ice_area"]).sum() / total - 1.0) < 1e-10   # nothing is lost in the halo beyond the fold
    api.icebergs_end(b)
    o.close()


@pytest.mark.parametrize("lx,ly", [(2, 1), (1, 2), (2, 2), (3, 2)])
def test_ranks_on_the_folded_grid_match_single_rank_oracle(lx, ly):
    """bergs that cross the fold change rank (the owner of the mirrored column), halo rows beyond the fold come from the
    other ranks of the top row"""
    case = FoldCase(3000, add_weight_to_ocean=1, use_old_spreading=0, hexagonal_icebergs=1)
    grp = parallel.LocalGroup(lx * ly)
    ranks = Ranks(case, lx * ly, lambda r: layout_domain(grp, lx, ly, r, case.halo), grp.run)
    o = case.make_oracle()
    hl = case.halo
    for name in ("lat", "dx", "dy", "area", "msk", "cos", "sin"):       # the rows beyond the fold on the ranks of the top row
        want = o.grid_field(STATIC[name])
        for r, h in enumerate(ranks.h):
            d = ranks.doms[r]
            if d.jec < GNJ:
                continue
            got = h.grid_field(STATIC[name])[hl + d.njc:, hl:hl + d.nic]
            sub = want[hl + GNJ:, hl + d.isc - 1:hl + d.iec]
            assert np.array_equal(got, sub), f"{name} rank {r}: {np.abs(got - sub).max()}"
    sent = 0
    for step in range(10):
        ranks.step()
        run_oracle(o, case)
        assert_bergs_match(ranks.bergs(), o.get_bergs(NAMES), rtol=1e-8, context=f"fold {lx}x{ly} ranks, step {step}")
        assert ranks.owners_ok()
        sent += sum(c["n_sent"] for c in ranks.counters())
        for fid in (D.KID_FLD_FLOATING_MELT, D.KID_FLD_BERG_MELT, D.KID_FLD_SPREAD_MASS):
            want = o.grid_field(fid)[hl:hl + GNJ, hl:hl + GNI]
            assert grid_rel(ranks.field(fid), want) < 1e-9, f"grid field {fid} step {step}"
    assert all(c["error_flags"] == 0 for c in ranks.counters())
    if lx > 1:
        assert sent > 100, sent
    ranks.end()
    grp.close()


def test_nccl_ranks_on_the_folded_grid_match_single_rank_oracle():
    """The same check over NCCL (one rank per GPU, threads of this process): the strip the top-row ranks share and the bergs
    that cross the fold travel by ncclSend / ncclRecv.  Needs >= 2 GPUs (gpurun --gpus 2)."""
    import threading
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    lx, ly = (2, 2) if ngpu >= 4 else (2, 1)
    nranks = lx * ly
    case = FoldCase(3000, add_weight_to_ocean=1, use_old_spreading=0, hexagonal_icebergs=1)
    uid = parallel.nccl_unique_id()
    comms = [None] * nranks

    def run_ranks(fn):
        out, err = [None] * nranks, [None] * nranks

        def work(r):
            try:
                out[r] = fn(r)
            except BaseException as e:  # noqa: BLE001
                err[r] = e
        th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
        [t.start() for t in th]
        [t.join() for t in th]
        for e in err:
            if e is not None:
                raise e
        return out

    def mk(r):
        comms[r] = parallel.nccl_comm(uid, nranks, r, r)
    run_ranks(mk)

    class G:            # what layout_domain needs of a group
        devices = list(range(nranks))

    def dom(r):
        class One:
            pass
        g = One(); g.devices = G.devices; g._g = One(); g._g.value = comms[r]
        d = layout_domain(g, lx, ly, r, case.halo)
        d.c.comm_kind = D.KID_COMM_NCCL
        return d
    ranks = Ranks(case, nranks, dom, run_ranks)
    o = case.make_oracle()
    hl = case.halo
    sent = 0
    for step in range(8):
        ranks.step()
        run_oracle(o, case)
        assert_bergs_match(ranks.bergs(), o.get_bergs(NAMES), rtol=1e-8, context=f"fold over NCCL, {lx}x{ly} ranks, step {step}")
        sent += sum(c["n_sent"] for c in ranks.counters())
        want = o.grid_field(D.KID_FLD_SPREAD_MASS)[hl:hl + GNJ, hl:hl + GNI]
        assert grid_rel(ranks.field(D.KID_FLD_SPREAD_MASS), want) < 1e-9
    assert sent > 100, sent
    ranks.end()
    for c in comms:
        api.lib().kid_nccl_destroy(c)
