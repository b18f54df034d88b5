"""Multi-rank path: the cell grid is decomposed like mpp_define_layout, bergs that
leave a tile are packed on the device and shipped to their owner (send_bergs_to_other_pes
F:2997), halos travel like mpp_update_domains.  Free-drifting bergs do not interact, so the
union of the ranks' bergs must equal the SINGLE-rank CPU oracle on the same inputs.

Ranks run as host threads of this process: on one GPU through the in-process group (device to
device copies), on >= 2 GPUs through NCCL."""
import numpy as np
import pytest

from common import COMPARE_F64, Case, assert_bergs_match, grid_rel, run_oracle
from icebergs_b200 import _cdefs as D
from icebergs_b200 import api, parallel
from icebergs_b200 import synthetic as S

pytestmark = pytest.mark.gpu

NAMES = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
FLUX_FIELDS = (D.KID_FLD_FLOATING_MELT, D.KID_FLD_BERG_MELT, D.KID_FLD_BERGY_SRC, D.KID_FLD_BERGY_MELT)


class Ranks:
    """nranks library handles over one Case, driven like MPI ranks."""

    def __init__(self, case, nranks, domain_of, run_ranks):
        self.case, self.nranks, self.run_ranks = case, nranks, run_ranks
        self.doms = [domain_of(r) for r in range(nranks)]
        self.grids = [type(case.grid)(case.gni, case.gnj, d.isc, d.iec, d.jsc, d.jec) for d in self.doms]
        parts = parallel.split_by_owner(case.bergs, self.doms)
        self.h = [None] * nranks

        def init(r):
            d = self.doms[r]
            b = api.icebergs_init(case.gni, case.gnj, case.dt, (1, 0.0), params=case.params(), domain=d,
                                  capacity=case.capacity, **self.grids[r].init_args())
            cnt = np.zeros((d.njd, d.nid), dtype=np.int32)
            hl = case.halo
            cnt[hl:hl + d.njc, hl:hl + d.nic] = case.counter[d.jsc - 1:d.jec, d.isc - 1:d.iec]
            b.set_calving_state(iceberg_counter_grd=cnt)
            b.set_bergs(**parts[r])
            self.h[r] = b
        run_ranks(init)

    def step(self, over=None):
        def one(r):
            f = self.grids[r].forcing()
            for k, v in (over or {}).items():
                f[k] = np.full_like(f[k], v)
            calving, hflx = f["calving"].copy(), f["calving_hflx"].copy()
            api.icebergs_run(self.h[r], (1, 0.0), calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"],
                             f["ssh"], f["sst"], hflx, f["cn"], f["hi"], sss=f["sss"])
        self.run_ranks(one)

    def resident(self, n):
        self.run_ranks(lambda r: self.h[r].step_resident(n, 1, 0.0))

    def bergs(self):
        parts = [b.get_bergs(NAMES) for b in self.h]
        return {k: np.concatenate([p[k] for p in parts]) for k in NAMES}

    def owners_ok(self):
        for r, b in enumerate(self.h):
            g = b.get_bergs(["ine", "jne"])
            d = self.doms[r]
            if len(g["ine"]) and not ((g["ine"] >= d.isc) & (g["ine"] <= d.iec) & (g["jne"] >= d.jsc) & (g["jne"] <= d.jec)).all():
                return False
        return True

    def field(self, fid):
        """Stitches the compute domains of a grid field into the global (gnj, gni) array."""
        out = np.zeros((self.case.gnj, self.case.gni))
        hl = self.case.halo
        for r, b in enumerate(self.h):
            d = self.doms[r]
            out[d.jsc - 1:d.jec, d.isc - 1:d.iec] = b.grid_field(fid)[hl:hl + d.njc, hl:hl + d.nic]
        return out

    def counters(self):
        return [b.counters() for b in self.h]

    def end(self):
        for b in self.h:
            api.icebergs_end(b)


def check_against_oracle(case, ranks, steps, over, rtol):
    o = case.make_oracle()
    fast = {k: np.full_like(case.forcing[k], v) for k, v in (over or {}).items()}
    moved = 0
    for step in range(steps):
        ranks.step(over)
        run_oracle(o, case, **fast)
        assert_bergs_match(ranks.bergs(), o.get_bergs(NAMES), rtol=rtol, context=f"{ranks.nranks} ranks, step {step}")
        assert ranks.owners_ok(), "a rank holds a berg outside its compute domain"
        moved += sum(c["n_sent"] for c in ranks.counters())
        hl = case.halo
        for fid in FLUX_FIELDS:
            want = o.grid_field(fid)[hl:hl + case.gnj, hl:hl + case.gni]
            assert grid_rel(ranks.field(fid), want) < 1e-9, f"grid field {fid} step {step}"
    return moved


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_in_process_ranks_match_single_rank_oracle(nranks):
    # half-day steps in a fast current: bergs cross tile edges (and the cyclic seam) every step but
    # stay within the 4-cell walk of adjust_index_and_ground, whose data-domain clamps (I:7941-7998)
    # would otherwise make an N-PE run of the reference itself differ from a 1-PE run
    case = Case(96, 48, 12000, dt=43200.0, old_bug_bilin=0)
    grp = parallel.LocalGroup(nranks)
    ranks = Ranks(case, nranks, lambda r: grp.domain(case.gni, case.gnj, r, halo=case.halo), grp.run)
    moved = check_against_oracle(case, ranks, 8, dict(uo=1.2, vo=0.15, tauxa=15.0), rtol=1e-8)
    assert moved > 100, f"only {moved} bergs migrated: the case does not exercise the exchange"
    ranks.end()
    grp.close()


def test_in_process_ranks_resident_steps_and_sort():
    """Many resident steps (periodic cell sort compacts the slots the leavers freed)."""
    case = Case(96, 48, 6000)
    grp = parallel.LocalGroup(4)
    ranks = Ranks(case, 4, lambda r: grp.domain(case.gni, case.gnj, r, halo=case.halo), grp.run)
    o = case.make_oracle()
    ranks.step(); run_oracle(o, case)
    ranks.resident(40); o.step_again(40, 1, 0.0)
    assert_bergs_match(ranks.bergs(), o.get_bergs(NAMES), rtol=1e-9, context="4 ranks, 41 steps")
    assert ranks.owners_ok()
    ranks.end()
    grp.close()


def test_overlapped_migration_in_resident_steps():
    """kid_step_resident defers the migration of a step into the next step's kernel (second stream, arrivals
    stepped separately): every berg still takes each step once, and the flux fields the caller sees -- the last
    step's -- hold exactly that step's melt."""
    case = Case(96, 48, 12000, dt=43200.0, old_bug_bilin=0)
    grp = parallel.LocalGroup(4)
    ranks = Ranks(case, 4, lambda r: grp.domain(case.gni, case.gnj, r, halo=case.halo), grp.run)
    o = case.make_oracle()
    over = dict(uo=1.2, vo=0.15, tauxa=15.0)
    fast = {k: np.full_like(case.forcing[k], v) for k, v in over.items()}
    ranks.step(over); run_oracle(o, case, **fast)
    sent0 = sum(c["n_sent"] for c in ranks.counters())
    ranks.resident(12); o.step_again(12, 1, 0.0)
    assert_bergs_match(ranks.bergs(), o.get_bergs(NAMES), rtol=1e-8, context="4 ranks, 13 steps, overlapped migration")
    assert ranks.owners_ok()
    assert sum(c["n_sent"] for c in ranks.counters()) > sent0 + 500, "the case does not exercise the exchange"
    hl = case.halo
    for fid in FLUX_FIELDS:
        want = o.grid_field(fid)[hl:hl + case.gnj, hl:hl + case.gni]
        assert grid_rel(ranks.field(fid), want) < 1e-9, f"grid field {fid} after the resident steps"
    ranks.end()
    grp.close()


def test_uneven_resident_calls_with_frequent_sorts():
    """Resident calls of different lengths (the last two steps of a call migrate synchronously, the others overlap their
    migration with the next step's kernel and append the arrivals at a tile-aligned slot), plain icebergs_run calls in
    between and a cell sort every 3 steps: the store shrinks and grows by different amounts from one sort epoch to the
    next, so slots beyond its end hold stale data of earlier epochs (the column arrays rotate through the sort's spares).
    Regression: the gap below the aligned arrivals once kept stale ALIVE flags (round 2, found by bench.py --gpus 2).
    14 steps only: later on a few bergs of this fast-current case jump many cells in a step, where an N-PE run of the
    reference differs from its own 1-PE run (the data-domain clamps of I:7941-7998)."""
    case = Case(96, 48, 12000, dt=43200.0, old_bug_bilin=0)
    grp = parallel.LocalGroup(2)
    ranks = Ranks(case, 2, lambda r: grp.domain(case.gni, case.gnj, r, halo=case.halo), grp.run)
    o = case.make_oracle()
    over = dict(uo=1.2, vo=0.15, tauxa=15.0)
    fast = {k: np.full_like(case.forcing[k], v) for k, v in over.items()}
    ranks.step(over); run_oracle(o, case, **fast)
    grp.run(lambda r: ranks.h[r].set_sort_phase(3, 0))
    steps = 1
    for n in (4, 2, 1, 3, 3):
        if n == 1:
            ranks.step(over); run_oracle(o, case, **fast)
        else:
            ranks.resident(n); o.step_again(n, 1, 0.0)
        steps += n
        assert_bergs_match(ranks.bergs(), o.get_bergs(NAMES), rtol=1e-8, context=f"2 ranks, {steps} steps, uneven resident calls")
    assert ranks.owners_ok()
    assert sum(b.sorts_done() for b in ranks.h) >= 2 * 4
    ranks.end()
    grp.close()


def test_nccl_ranks_match_single_rank_oracle():
    """Same check over NCCL: one rank per GPU (threads of this process), needs >= 2 GPUs."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    nranks = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    case = Case(96, 48, 12000, dt=43200.0, old_bug_bilin=0)
    uid = parallel.nccl_unique_id()
    import threading
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
    ranks = Ranks(case, nranks,
                  lambda r: api.Domain.decomposed(case.gni, case.gnj, r, nranks, halo=case.halo, device=r, comm=comms[r]),
                  run_ranks)
    moved = check_against_oracle(case, ranks, 8, dict(uo=1.2, vo=0.15, tauxa=15.0), rtol=1e-8)
    assert moved > 100
    ranks.end()
    for c in comms:
        api.lib().kid_nccl_destroy(c)


@pytest.mark.parametrize("nranks", [2, 4])
def test_interacting_bergs_across_ranks(nranks):
    """Contact forces across tile edges: halo copies travel between ranks (update_halo_icebergs
    F:1800), owners migrate (F:2997), and the union must still equal the single-rank oracle."""
    import kid_oracle_py as O
    rng = np.random.default_rng(5)
    n = 600
    base = S.collision_bergs()
    cols = {k: np.resize(v, n).copy() for k, v in base.items()}
    # Without bonds the reference builds no halo copies before the first step (I:150-167), so on N PEs the
    # first step misses contacts across tile edges that one PE sees: nobody starts within contact range
    # (2 x 390 m) of a tile edge or of the seam; from step 2 on the copies exist in every decomposition.
    def away(m):
        x = rng.uniform(0.0, 1.0, m)
        return np.where(x < 0.5, 1000.0 + x * 2.0 * 8000.0, 11000.0 + (x - 0.5) * 2.0 * 8000.0)
    cols["lon"], cols["lat"] = away(n), away(n)
    cols["start_lon"], cols["start_lat"] = cols["lon"].copy(), cols["lat"].copy()
    cols["start_day"] = rng.uniform(0.0, 300.0, n)
    params = lambda: S.collision_params(api.default_params, iceberg_bonds_on=0, manually_initialize_bonds=0, max_bonds=0)
    g0 = S.CartesianGrid()
    dom0 = api.Domain.single(20, 20, halo=3, cyclic_x=True)
    o = O.Oracle(20, 20, 60.0, (1, 0.0), params=params(), domain=dom0, **g0.init_args())
    o.set_bergs(**cols)
    ref0 = o.get_bergs(["id", "ine", "jne", "lon"])           # ids are generated in file order (legacy iceberg_num)
    cols = dict(cols, id=np.zeros(n, dtype=np.int64), ine=np.zeros(n, dtype=np.int32), jne=np.zeros(n, dtype=np.int32))
    order = {float(x): k for k, x in enumerate(ref0["lon"])}
    for k in range(n):
        q = order[float(cols["lon"][k])]
        cols["id"][k], cols["ine"][k], cols["jne"][k] = ref0["id"][q], ref0["ine"][q], ref0["jne"][q]
    grp = parallel.LocalGroup(nranks)
    doms = [grp.domain(20, 20, r, halo=3) for r in range(nranks)]
    grids = [S.CartesianGrid(20, 20, 1.0e3, d.isc, d.iec, d.jsc, d.jec) for d in doms]
    parts = parallel.split_by_owner(cols, doms)
    hs = [None] * nranks

    def init(r):
        hs[r] = api.icebergs_init(20, 20, 60.0, (1, 0.0), params=params(), domain=doms[r], capacity=8192, **grids[r].init_args())
        hs[r].set_bergs(**parts[r])
    grp.run(init)
    f0 = g0.forcing(ibuo=0.6, ibvo=0.1, collision_test=False)

    def one(r):
        f = grids[r].forcing(ibuo=0.6, ibvo=0.1, collision_test=False)
        c, h = f["calving"].copy(), f["calving_hflx"].copy()
        api.icebergs_run(hs[r], (1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h,
                         f["cn"], f["hi"], sss=f["sss"])
    names = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
    moved = 0
    for step in range(80):
        grp.run(one)
        c, h = f0["calving"].copy(), f0["calving_hflx"].copy()
        o.run((1, 0.0), c, f0["uo"], f0["vo"], f0["ui"], f0["vi"], f0["tauxa"], f0["tauya"], f0["ssh"], f0["sst"], h, f0["cn"],
              f0["hi"], sss=f0["sss"])
        moved += sum(b.counters()["n_sent"] for b in hs)
        if step % 10 == 9:
            got = [b.get_bergs(names) for b in hs]
            got = {k: np.concatenate([p[k] for p in got]) for k in names}
            assert_bergs_match(got, o.get_bergs(names), rtol=1e-8, context=f"{nranks} ranks, interacting, step {step}", acc_floor=1e-13)
    assert moved > 5
    for b in hs:
        api.icebergs_end(b)
    grp.close()


def _run_cartesian_ranks(nranks, params, bergs, forcing_of, nsteps, check_every, names, f64, bonds, dt=60.0, rtol=1e-7,
                         yearday_of=lambda k, dt: 0.0, rtol_of=None):
    import kid_oracle_py as O
    g0 = S.CartesianGrid()
    dom0 = api.Domain.single(20, 20, halo=3, cyclic_x=True)
    o = O.Oracle(20, 20, dt, (1, 0.0), params=params(), domain=dom0, **g0.init_args())
    o.set_bergs(**bergs)
    if bonds:
        o.set_bonds()
    ref0 = o.get_bergs(["id", "ine", "jne", "lon", "lat"])
    n = len(bergs["lon"])
    cols = dict(bergs, id=np.zeros(n, dtype=np.int64), ine=np.zeros(n, dtype=np.int32), jne=np.zeros(n, dtype=np.int32))
    order = {(float(x), float(y)): k for k, (x, y) in enumerate(zip(ref0["lon"], ref0["lat"]))}
    for k in range(n):
        q = order[(float(cols["lon"][k]), float(cols["lat"][k]))]
        cols["id"][k], cols["ine"][k], cols["jne"][k] = ref0["id"][q], ref0["ine"][q], ref0["jne"][q]
    grp = parallel.LocalGroup(nranks)
    doms = [grp.domain(20, 20, r, halo=3) for r in range(nranks)]
    grids = [S.CartesianGrid(20, 20, 1.0e3, d.isc, d.iec, d.jsc, d.jec) for d in doms]
    parts = parallel.split_by_owner(cols, doms)
    hs = [None] * nranks
    _, _, ic0 = o.get_calving_state()

    def init(r):
        d = doms[r]
        hs[r] = api.icebergs_init(20, 20, dt, (1, 0.0), params=params(), domain=d, capacity=8192, **grids[r].init_args())
        cnt = np.zeros((d.njd, d.nid), dtype=np.int32)
        cnt[3:3 + d.njc, 3:3 + d.nic] = ic0[3 + d.jsc - 1:3 + d.jec, 3 + d.isc - 1:3 + d.iec]
        hs[r].set_calving_state(iceberg_counter_grd=cnt)      # the per-cell id counters follow the file-order ids
        hs[r].set_bergs(**parts[r])
        if bonds:
            hs[r].set_bonds()
    grp.run(init)
    f0 = forcing_of(g0)
    moved = 0
    for step in range(nsteps):
        t = (1, yearday_of(step, dt))

        def one(r):
            f = forcing_of(grids[r])
            c, h = f["calving"].copy(), f["calving_hflx"].copy()
            api.icebergs_run(hs[r], t, c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h,
                             f["cn"], f["hi"], sss=f["sss"])
        grp.run(one)
        c, h = f0["calving"].copy(), f0["calving_hflx"].copy()
        o.run(t, c, f0["uo"], f0["vo"], f0["ui"], f0["vi"], f0["tauxa"], f0["tauya"], f0["ssh"], f0["sst"], h, f0["cn"], f0["hi"],
              sss=f0["sss"])
        moved += sum(b.counters()["n_sent"] for b in hs)
        if step % check_every == check_every - 1:
            got = [b.get_bergs(names) for b in hs]
            got = {k: np.concatenate([p[k] for p in got]) for k in names}
            tol, cmp_names = rtol, f64
            if rtol_of:                      # (rtol, names) may depend on the step: chaos after a collision
                tol, cmp_names = rtol_of(step)
            assert_bergs_match(got, o.get_bergs(names), rtol=tol, names=cmp_names, context=f"{nranks} ranks, step {step}", acc_floor=1e-13)
    for b in hs:
        api.icebergs_end(b)
    grp.close()
    return moved, o


def test_bonded_conglomerates_across_ranks():
    """One bonded 8-element conglomerate of tests/collision_tests drifting north across the edge between
    2 ranks (split in y): its elements migrate one by one, so bonds span the ranks for a while -- they
    travel with the migrating elements and with the halo copies (F:3336-3354)."""
    params = lambda: S.collision_params(api.default_params)
    names = list(COMPARE_F64) + ["ine", "jne", "start_year", "id"]
    south = {k: v[:8].copy() for k, v in S.collision_bergs().items()}
    moved, o = _run_cartesian_ranks(2, params, south, lambda g: g.forcing(ibuo=0.05, ibvo=0.5, collision_test=False), 300, 20,
                                    names, COMPARE_F64, bonds=True)
    assert moved >= 4 and o.count_bergs() == 8
    lat = o.get_bergs(["lat"])["lat"]
    assert lat.min() < 10000.0 < lat.max(), "the conglomerate should straddle the rank edge at the end"


def test_footloose_across_ranks():
    params = lambda: S.footloose_params(api.default_params, fl_style_fl_bits=0)
    f64 = tuple(COMPARE_F64) + ("mass_of_fl_bits", "mass_of_fl_bergy_bits", "fl_k")
    names = list(f64) + ["ine", "jne", "start_year", "id"]
    moved, o = _run_cartesian_ranks(2, params, S.footloose_bergs(), lambda g: S.footloose_forcing(g), 6000, 1000, names, f64,
                                    bonds=False, dt=10.0, yearday_of=lambda k, dt: k * dt / 86400.0)
    assert o.counters()["nbergs_calved_fl"] >= 2


def _custom_domains(gni, gnj, xsizes, ysizes, halo, grp, rank_of=None):
    """Tiles with the given extents along each axis, as mpp_define_domains would hand them out for an
    mpp_compute_extent that does not split evenly; rank_of(px, py) numbers the PEs (default px + lx*py)."""
    lx, ly = len(xsizes), len(ysizes)
    x0 = np.r_[1, 1 + np.cumsum(xsizes)]
    y0 = np.r_[1, 1 + np.cumsum(ysizes)]
    rank_of = rank_of or (lambda px, py: px + lx * py)
    doms = {}
    for py in range(ly):
        for px in range(lx):
            c = D.KidDomain()
            c.gni, c.gnj = gni, gnj
            c.isc, c.iec, c.jsc, c.jec = int(x0[px]), int(x0[px + 1] - 1), int(y0[py]), int(y0[py + 1] - 1)
            c.isd, c.ied, c.jsd, c.jed = c.isc - halo, c.iec + halo, c.jsc - halo, c.jec + halo
            c.cyclic_x, c.cyclic_y = 1, 0
            r = rank_of(px, py)
            c.rank, c.nranks = r, lx * ly
            c.layout_x, c.layout_y = lx, ly
            c.pe_E, c.pe_W = rank_of((px + 1) % lx, py), rank_of((px - 1) % lx, py)
            c.pe_N = rank_of(px, py + 1) if py + 1 < ly else -1
            c.pe_S = rank_of(px, py - 1) if py > 0 else -1
            c.device = 0
            c.nccl_comm = grp._g.value
            c.comm_kind = D.KID_COMM_LOCAL
            doms[r] = api.Domain(c)
    return [doms[r] for r in range(lx * ly)]


@pytest.mark.parametrize("numbering", ["row_major", "shuffled"])
def test_uneven_tiles_and_any_rank_numbering(numbering):
    """The library takes the decomposition from the ranks' own compute domains and neighbour PEs (what the shim
    passes from mpp_define_domains), not from a formula: a 100x50 grid on 3x2 PEs with extents 33/34/33 (a
    remainder in the middle, as a symmetric mpp_compute_extent leaves it) and, second, PEs numbered in another order."""
    gni, gnj = 100, 50
    case = Case(gni, gnj, 9000, dt=43200.0, old_bug_bilin=0)
    grp = parallel.LocalGroup(6)
    perm = [0, 1, 2, 3, 4, 5] if numbering == "row_major" else [4, 0, 3, 5, 1, 2]
    doms = _custom_domains(gni, gnj, [33, 34, 33], [24, 26], case.halo, grp, rank_of=lambda px, py: perm[px + 3 * py])

    class R(Ranks):
        def __init__(self):
            self.case, self.nranks, self.run_ranks = case, 6, grp.run
            self.doms = doms
            self.grids = [S.Grid(gni, gnj, d.isc, d.iec, d.jsc, d.jec) for d in doms]
            own = np.full((gnj + 1, gni + 1), -1)
            for r, d in enumerate(doms):
                own[d.jsc:d.jec + 1, d.isc:d.iec + 1] = r
            owner = own[case.bergs["jne"], case.bergs["ine"]]
            parts = [{k: np.ascontiguousarray(v[owner == r]) for k, v in case.bergs.items()} for r in range(6)]
            self.h = [None] * 6

            def init(r):
                d = doms[r]
                b = api.icebergs_init(gni, gnj, case.dt, (1, 0.0), params=case.params(), domain=d, capacity=case.capacity,
                                      **self.grids[r].init_args())
                cnt = np.zeros((d.njd, d.nid), dtype=np.int32)
                hl = case.halo
                cnt[hl:hl + d.njc, hl:hl + d.nic] = case.counter[d.jsc - 1:d.jec, d.isc - 1:d.iec]
                b.set_calving_state(iceberg_counter_grd=cnt)
                b.set_bergs(**parts[r])
                self.h[r] = b
            grp.run(init)

    ranks = R()
    moved = check_against_oracle(case, ranks, 6, dict(uo=1.2, vo=0.15, tauxa=15.0), rtol=1e-8)
    assert moved > 100
    ranks.end()
    grp.close()
