#!/usr/bin/env python
"""bench.py -- berg-steps/sec (dyn+thermo) of the KID hot path on N B200s.

Workload (BASELINE.json configs[1] / BASELINE.md): synthetic free drift + melt, seeded point
bergs on a 1/4-degree global grid (1440x720) with analytic currents/winds, dt=3600 s, Verlet,
bergy bits on.  N=1: 10M bergs on one GPU (plus, in the same line, the 12.5M-berg point that is
the k=1 base of the weak-scaling series).  N>1: weak scaling, the grid is decomposed like
mpp_define_layout and every rank owns 12.5M bergs of its own tile; bergs that leave a tile
migrate over NCCL, and before the timed region a small fast-current case is pushed through the
same NCCL communicator and compared with the single-rank CPU oracle ("parity_nccl").

`--workload bonded`: BASELINE.json configs[2], a square-packed bonded tabular berg (DEM bonds,
MTS sub-steps, a68_test physics) -- element-steps/s, its own algorithmic bytes.
`--workload interactive`: BASELINE.json configs[0] scaled up, 10 M unbonded bergs that collide through
interactive_force on the grid of the drift workload (bench_interactive.py).

One JSON line on rank 0 (contract in the task statement).  `--impl reference` times the CPU
oracle (the reference is Fortran+FMS and cannot be built in this image) on the host cores; that
arm never imports icebergs_b200 nor loads libkid_b200.so.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GNI, GNJ = 1440, 720
DT = 3600.0
B_BERG = 290.0      # algorithmic bytes per berg-step
B_CELL = 224.0      # algorithmic bytes per occupied-cell-step
B_NEIGHBOUR = 88.0  # per neighbour visited (interacting / bonded populations)
METRIC = "berg_steps_per_sec"
UNIT = "berg-steps/s"
N1_BERGS = 10_000_000          # configs[1]
WEAK_BERGS = 12_500_000        # per GPU, configs[4] / BASELINE.md weak series


def ncu_traffic():
    """dram bytes of one k_step launch on the N=1 workload from the committed `ncu --set full` summary."""
    for name in ("r2_kstep_fast_summary.txt", "r1i_kstep_final_summary.txt"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        rd = wr = None
        for ln in open(p):
            f = ln.split()
            if len(f) >= 2 and f[0] == "dram__bytes_read.sum":        # "name value [Gbyte]"
                rd = float(f[1])
            if len(f) >= 2 and f[0] == "dram__bytes_write.sum":
                wr = float(f[1])
        if rd is not None and wr is not None:
            return (rd + wr) * 1e9, f"ncu --set full capture of this workload, profiles/{name}"
    return None, None


def workload_text(n_per, gni=None, gnj=None):
    gni, gnj = gni or GNI, gnj or GNJ
    return (f"free drift + melt, {n_per} seeded bergs per GPU on the {'1/4-degree ' if gni == 1440 else ''}{gni}x{gnj} grid, "
            f"analytic currents/winds, dt=3600 s, Verlet, bergy bits on")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def nsamples(self):
        return len(self.lines)

    def mark(self, t0, t1):
        """a period under load (settle loop, timed region, e2e loop)"""
        self.marks.append((t0, t1))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if any(a <= t <= b + 0.06 for a, b in self.marks)]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm),
                "window": "nvidia-smi -lms 50 over the timed region, the e2e loop and 0.6 s of the same resident stepping kept up after them"}


def trace(msg):
    """KID_BENCH_TRACE=1: stage markers on stderr (which stage a hung multi-rank run is in)"""
    if os.environ.get("KID_BENCH_TRACE"):
        sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:.1f}s] {msg}\n")
        sys.stderr.flush()


_T0 = time.perf_counter()


def pinned(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    out = t.numpy()
    out[...] = a
    return out, t


# ------------------------------------------------------------------ CPU legs (oracle only; no product import)
def cpu_oracle_run(nbergs, steps, nthreads, warm=1):
    """The CPU leg: the oracle port of the reference path on the host cores (-O3 -march=native -fopenmp build)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kid_oracle_py as O
    S = O.load_by_path("synthetic", os.path.join("icebergs_b200", "synthetic.py"))
    grid = S.Grid(GNI, GNJ)
    p = S.workload_params(O.default_params)
    dom = O.SingleDomain(GNI, GNJ, halo=p.halo, cyclic_x=True)
    o = O.Oracle(GNI, GNJ, DT, (1, 0.0), params=p, domain=dom, **grid.init_args())
    bergs, _ = grid.seed_bergs(nbergs)
    o.set_bergs(**bergs)
    del bergs
    f = grid.forcing()
    calving, hflx = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, 0.0), calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hflx,
          f["cn"], f["hi"], sss=f["sss"])          # first step (also ingests the forcing)
    if warm > 0:
        o.step_again(warm, 1, 0.0, nthreads)
    t0 = time.perf_counter()
    o.step_again(steps, 1, 0.0, nthreads)
    wall = time.perf_counter() - t0
    o.close()
    return nbergs * steps / wall, wall


def cpu_build_info():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kid_oracle_py as O
    O.use_fast_build()
    return {"omp_threads": int(O.lib().oracle_omp_max_threads()),
            "build": "gcc -O3 -march=native -fopenmp (oracle/Makefile `fast`, compiled on this box)"}


def reference_arm(args, rank):
    if rank != 0:
        return
    info = cpu_build_info()
    cores = info["omp_threads"]
    nb = args.cpu_bergs or N1_BERGS
    n_per = N1_BERGS if args.gpus == 1 else WEAK_BERGS
    steps = max(args.steps, 1)
    v, wall = cpu_oracle_run(nb, steps, cores, warm=min(max(args.warmup, 1), 3))
    same = nb == n_per
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(n_per),
                   "sample": (f"each step = one pass over {'all' if same else 'a bounded sample of'} {nb} bergs on the same grid and forcing"),
                   "same_config": same,
                   "note": "reference = NOAA-GFDL/icebergs is Fortran+FMS (no Fortran compiler here): CPU oracle port, " + info["build"]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{nb} bergs x {steps} steps, {wall:.1f} s wall, {cores} OpenMP threads (omp_get_max_threads) over cell rows"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ NCCL data-path parity (N > 1)
def nccl_parity_check(rank, world, local_rank, comm):
    """The 96x48 / 12k-berg / fast-current case of tests/test_multirank_gpu.py over the SAME NCCL communicator the
    timed run uses; rank 0 compares the union of the ranks' bergs with the single-rank CPU oracle (F:2997-3247)."""
    import torch.distributed as dist
    from icebergs_b200 import _cdefs as D
    from icebergs_b200 import api, parallel
    from icebergs_b200 import synthetic as S
    gni, gnj, n, dt, steps = 96, 48, 12000, 43200.0, 8
    over = dict(uo=1.2, vo=0.15, tauxa=15.0)
    names = ["lon", "lat", "uvel", "vvel", "axn", "ayn", "bxn", "byn", "xi", "yj", "mass", "thickness", "width", "length",
             "mass_of_bits", "ine", "jne", "id"]
    g0 = S.Grid(gni, gnj)
    cols, counter = g0.seed_bergs(n)
    mk_params = lambda dflt: S.workload_params(dflt, halo=4, old_bug_bilin=0)
    dom = api.Domain.decomposed(gni, gnj, rank, world, halo=4, device=local_rank, comm=comm, comm_kind=D.KID_COMM_NCCL)
    grid = S.Grid(gni, gnj, dom.isc, dom.iec, dom.jsc, dom.jec)
    b = api.icebergs_init(gni, gnj, dt, (1, 0.0), params=mk_params(api.default_params), domain=dom, capacity=4 * n,
                          **grid.init_args())
    cnt = np.zeros((dom.njd, dom.nid), dtype=np.int32)
    cnt[4:4 + dom.njc, 4:4 + dom.nic] = counter[dom.jsc - 1:dom.jec, dom.isc - 1:dom.iec]
    b.set_calving_state(iceberg_counter_grd=cnt)
    mine = (cols["ine"] >= dom.isc) & (cols["ine"] <= dom.iec) & (cols["jne"] >= dom.jsc) & (cols["jne"] <= dom.jec)
    b.set_bergs(**{k: np.ascontiguousarray(v[mine]) for k, v in cols.items()})
    f = grid.forcing()
    for k, v in over.items():
        f[k] = np.full_like(f[k], v)
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    api.icebergs_run(b, (1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h,
                     f["cn"], f["hi"], sss=f["sss"])
    sent = b.counters()["n_sent"]
    b.step_resident(steps - 1, 1, 0.0)          # incl. the migration overlapped with the next step's kernel
    sent_total = b.counters()["n_sent"]
    got = b.get_bergs(names)
    api.icebergs_end(b)
    parts = [None] * world
    dist.all_gather_object(parts, (got, int(sent_total), int(sent)))
    if rank != 0:
        return None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kid_oracle_py as O
    got = {k: np.concatenate([p[0][k] for p in parts]) for k in names}
    d0 = O.SingleDomain(gni, gnj, halo=4, cyclic_x=True)
    o = O.Oracle(gni, gnj, dt, (1, 0.0), params=mk_params(O.default_params), domain=d0, **g0.init_args())
    c0 = np.zeros((d0.njd, d0.nid), dtype=np.int32)
    c0[4:4 + gnj, 4:4 + gni] = counter
    o.set_calving_state(iceberg_counter_grd=c0)
    o.set_bergs(**cols)
    f0 = g0.forcing()
    for k, v in over.items():
        f0[k] = np.full_like(f0[k], v)
    c, h = f0["calving"].copy(), f0["calving_hflx"].copy()
    o.run((1, 0.0), c, f0["uo"], f0["vo"], f0["ui"], f0["vi"], f0["tauxa"], f0["tauya"], f0["ssh"], f0["sst"], h, f0["cn"],
          f0["hi"], sss=f0["sss"])
    o.step_again(steps - 1, 1, 0.0, 1)
    want = o.get_bergs(names)
    o.close()
    res = {"case": f"{gni}x{gnj}, {n} bergs, dt={dt:.0f} s, {steps} steps, uo=1.2 m/s: union of {world} NCCL ranks vs the single-rank CPU oracle",
           "migrated": int(sum(p[1] for p in parts)), "count_gpu": int(len(got["id"])), "count_oracle": int(len(want["id"]))}
    ok = len(got["id"]) == len(want["id"])
    worst = 0.0
    if ok:
        ga, wa = np.argsort(got["id"], kind="stable"), np.argsort(want["id"], kind="stable")
        for k in ("id", "ine", "jne"):
            ok = ok and bool(np.array_equal(got[k][ga], want[k][wa]))
        for k in names:
            if k in ("id", "ine", "jne"):
                continue
            a, w = got[k][ga], want[k][wa]
            scale = np.maximum(np.maximum(np.abs(a), np.abs(w)), 1e-300)
            e = np.where(a == w, 0.0, np.abs(a - w) / scale)
            floor = 1e-8 if k in ("xi", "yj") else 1e-13 * max(float(np.max(np.abs(w))), 1e-300)
            e = np.where(np.abs(a - w) <= floor, 0.0, e)
            worst = max(worst, float(e.max()) if len(e) else 0.0)
        ok = ok and worst <= 1e-8
    res.update(ok=bool(ok), worst_rel=worst, rtol=1e-8, integers="id, ine, jne bit-exact" if ok else "MISMATCH")
    return res


# ------------------------------------------------------------------ the GPU arm
def run_drift(args, rank, world, local_rank, dom, n_per, clocks, do_e2e, seed_stream=None):
    """One population on this rank's tile: resident throughput, (optionally) end-to-end, per-phase times."""
    import torch
    import torch.distributed as dist
    from icebergs_b200 import api
    from icebergs_b200 import synthetic as S
    multi = world > 1
    p = S.workload_params(api.default_params)
    grid = S.Grid(GNI, GNJ, dom.isc, dom.iec, dom.jsc, dom.jec)
    bergs = api.icebergs_init(GNI, GNJ, DT, (1, 0.0), params=p, domain=dom, capacity=int(n_per * 1.25) + 4096,
                              **grid.init_args())
    # seeded in chunks: bounded host memory at 1e8 bergs
    occupied_cells = np.zeros(GNI * GNJ, dtype=bool)
    chunk, done, counter = 12_500_000, 0, None
    while done < n_per:
        m = min(chunk, n_per - done)
        cols, counter = grid.seed_bergs(m, stream=rank if seed_stream is None else seed_stream, start=done, counter0=counter)
        occupied_cells[(cols["jne"].astype(np.int64) - 1) * GNI + (cols["ine"].astype(np.int64) - 1)] = True
        bergs.set_bergs(**cols)
        del cols
        done += m
    occupied = int(occupied_cells.sum())
    trace(f"seeded {n_per} bergs")
    f = grid.forcing()
    keep, fp = [], {}
    for k, v in f.items():
        fp[k], t = pinned(v)
        keep.append(t)
    calving, hflx = fp["calving"], fp["calving_hflx"]
    # calving / calving_hflx are intent(inout): the caller hands in this step's calving (none in this workload) and
    # gets the unused calving and the heat flux back.  As a coupler would, the bench double-buffers the pair: the
    # next step's arrays are prepared (zeroed) on a helper thread while the call on the current pair is in flight.
    from concurrent.futures import ThreadPoolExecutor
    pair = [(calving, hflx)]
    a, ta = pinned(np.zeros_like(calving)); b_, tb = pinned(np.zeros_like(hflx))
    keep += [ta, tb]
    pair.append((a, b_))
    pool = ThreadPoolExecutor(max_workers=1)
    state = {"cur": 0, "fut": None}

    def prepare(k):
        pair[k][0].fill(0.0)
        pair[k][1].fill(0.0)

    prepare(0)

    def run_once():
        if state["fut"] is not None:
            state["fut"].result()
        c, h = pair[state["cur"]]
        state["cur"] ^= 1
        state["fut"] = pool.submit(prepare, state["cur"])
        api.icebergs_run(bergs, (1, 0.0), c, fp["uo"], fp["vo"], fp["ui"], fp["vi"], fp["tauxa"], fp["tauya"],
                         fp["ssh"], fp["sst"], h, fp["cn"], fp["hi"], sss=fp["sss"])

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    interval = int(os.environ.get("KID_SORT_INTERVAL", "32"))

    def place_sorts(k):
        # the periodic cell sort is part of the step: every window of k steps is charged ceil(k/interval) sorts,
        # whatever k is (a 20-step window gets one, a 50-step window two), the last one on the window's last step
        bergs.set_sort_phase(interval, interval - 1 - ((k - 1) % interval))

    # ---- HBM-resident throughput: forcing on the device, K steps
    run_once()                                  # uploads forcing, first step
    trace("first icebergs_run done")
    place_sorts(args.warmup)
    bergs.step_resident(args.warmup, 1, 0.0)    # W untimed warm-up steps (ends on a sort)
    place_sorts(args.steps)
    trace("warm-up done")
    barrier()
    l0, s0 = bergs.kernel_launches(), bergs.sorts_done()
    t0 = time.perf_counter()
    bergs.step_resident(args.steps, 1, 0.0)
    barrier()
    wall = time.perf_counter() - t0
    trace("timed resident steps done")
    if clocks is not None:
        clocks.mark(t0, t0 + wall)
    l1, s1 = bergs.kernel_launches(), bergs.sorts_done()
    tm = bergs.last_timing()
    dev_ms = tm["total"]
    nst = max(tm["_"], 1.0)
    kern_ms, comm_ms, sort_ms = tm["momentum+thermodyn"] / nst, tm["communication"] / nst, tm["sort"] / nst
    sorts = int(s1 - s0)
    n_alive = bergs.count_bergs()
    slow_frac = bergs.slow_fraction()
    t_all = torch.tensor([dev_ms, wall * 1e3, float(n_alive), kern_ms, float(occupied), comm_ms, sort_ms, float(sorts)],
                         dtype=torch.float64, device="cuda")
    per_rank = None
    if multi:
        tmax = t_all.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_all.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, kern_ms, comm_ms, sort_ms = float(tmax[0]), float(tmax[1]), float(tmax[3]), float(tmax[5]), float(tmax[6])
        n_total = float(tsum[2])
        gathered = [torch.zeros_like(t_all) for _ in range(world)]
        dist.all_gather(gathered, t_all)
        per_rank = {"kernel_ms": [round(float(g[3]), 4) for g in gathered], "migration_ms": [round(float(g[5]), 4) for g in gathered],
                    "bergs": [int(g[2]) for g in gathered], "occupied_cells": [int(g[4]) for g in gathered],
                    "sorts_in_window": [int(g[7]) for g in gathered]}
    else:
        wall_ms, n_total = wall * 1e3, float(n_alive)
    out = {"dev_ms": dev_ms, "wall_ms": wall_ms, "n_total": n_total, "kern_ms": kern_ms, "comm_ms": comm_ms, "sort_ms": sort_ms,
           "sorts": sorts, "sort_ms_per_call": (tm["sort"] / sorts) if sorts else None, "occupied": occupied,
           "launches": int(l1 - l0), "per_rank": per_rank, "interval": interval, "n_alive_rank0": int(n_alive),
           "slow_fraction": slow_frac,
           "value": n_total * args.steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / args.steps}

    # ---- end to end through icebergs_run: host buffers, H2D of the forcing + D2H of the returns
    if do_e2e:
        # (a) the plain call: copy in, step, copy out, one after the other
        for _ in range(3):
            run_once()
        place_sorts(args.steps)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run_once()
        barrier()
        sync_wall = time.perf_counter() - t0
        trace("e2e: plain calls done")
        if state["fut"] is not None:
            state["fut"].result()
        # (b) the same calls with the next step's forcing announced early (kid_prefetch_forcing): two sets of pinned
        # forcing arrays alternate (a coupler fills one while the other is in use), three inout pairs rotate (in use,
        # announced, being zeroed).  Every step still moves all 13 fields to the device and the two inout fields back.
        fq = {}
        for k, v in fp.items():
            fq[k], t = pinned(v.copy())
            keep.append(t)
        sets = [fp, fq]
        c3, t3 = pinned(np.zeros_like(calving)); h3, t4 = pinned(np.zeros_like(hflx))
        keep += [t3, t4]
        pairs3 = [pair[0], pair[1], (c3, h3)]
        for c_, h_ in pairs3:
            c_.fill(0.0); h_.fill(0.0)

        def args_of(k):
            fs, (c_, h_) = sets[k % 2], pairs3[k % 3]
            return (c_, fs["uo"], fs["vo"], fs["ui"], fs["vi"], fs["tauxa"], fs["tauya"], fs["ssh"], fs["sst"], h_, fs["cn"], fs["hi"])

        def pipelined(n, k0):
            fut = None
            a0 = args_of(k0)
            api.icebergs_prefetch(bergs, *a0, sss=sets[k0 % 2]["sss"])
            for k in range(k0, k0 + n):
                a_now = args_of(k)
                if fut is not None:
                    fut.result()                                  # the pair announced next is zeroed
                if k + 1 < k0 + n:
                    api.icebergs_prefetch(bergs, *args_of(k + 1), sss=sets[(k + 1) % 2]["sss"])
                fut = pool.submit(lambda q=pairs3[(k + 2) % 3]: (q[0].fill(0.0), q[1].fill(0.0)))
                api.icebergs_run(bergs, (1, 0.0), *a_now, sss=sets[k % 2]["sss"])
            fut.result()
            return k0 + n

        kk = pipelined(3, 0)
        place_sorts(args.steps)
        barrier()
        t0 = time.perf_counter()
        pipelined(args.steps, kk)
        barrier()
        e2e_wall = time.perf_counter() - t0
        trace("e2e: announced calls done")
        if clocks is not None:
            clocks.mark(t0 - sync_wall, t0 + e2e_wall)
        tw = torch.tensor([e2e_wall, sync_wall], dtype=torch.float64, device="cuda")
        if multi:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        h2d = sum(fp[k].nbytes for k in ("calving", "uo", "vo", "ui", "vi", "tauxa", "tauya", "ssh", "sst",
                                          "calving_hflx", "cn", "hi", "sss"))
        d2h = calving.nbytes + hflx.nbytes
        out["e2e"] = {"value": n_total * args.steps / float(tw[0]), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(tw[0]) / args.steps,
                      "unpipelined_ms_per_step": 1e3 * float(tw[1]) / args.steps,
                      "unpipelined_value": n_total * args.steps / float(tw[1]),
                      "note": "icebergs_run through the C ABI with pinned host arrays: 13 forcing fields H2D and the two inout "
                              "fields D2H every step (host-side wall clock, max over ranks).  value: the caller announces the "
                              "next step's arrays with icebergs_prefetch (kid_prefetch_forcing) before each call, so their copy "
                              "overlaps the step in flight (PCIe-bound: h2d_bytes_per_step at ~50 GB/s); unpipelined_*: the same "
                              "calls without the announcement (copy, step, copy back in sequence).  The caller rotates the "
                              "inout pairs (the next one is zeroed on a helper thread while the call runs)"}
    if clocks is not None or multi:
        # the timed region lasts tens of ms and nvidia-smi samples every 50 ms: the same load is kept up for about another
        # half second so that the clocks / throttle reasons are seen under exactly this kernel mix (untimed).  Every rank
        # takes the SAME number of steps: with N > 1 each step holds a migration exchange between the ranks.
        t_a = time.perf_counter()
        for _ in range(30):
            bergs.step_resident(16, 1, 0.0)
        if clocks is not None:
            clocks.mark(t_a, time.perf_counter())
    api.icebergs_end(bergs)
    pool.shutdown()
    return out


def main():
    global GNI, GNJ
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="drift", choices=["drift", "bonded", "interactive"])
    ap.add_argument("--bergs-per-gpu", type=int, default=0, help="0 = 10M at N=1, 12.5M per GPU at N>1")
    ap.add_argument("--cpu-bergs", type=int, default=0, help="reference arm / cpu_baseline population (0 = the workload's own)")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--gni", type=int, default=GNI, help="diagnostics only: another grid size (the metric is quoted on 1440x720)")
    ap.add_argument("--gnj", type=int, default=GNJ)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak-base", action="store_true", help="N=1: skip the extra 12.5M-berg point (k=1 of the weak series)")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the NCCL data-path parity check")
    ap.add_argument("--elements", type=int, default=0, help="--workload bonded: elements of the tabular berg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload == "bonded":
            import bench_bonded
            bench_bonded.reference_arm(args, rank)
        elif args.workload == "interactive":
            import bench_interactive
            bench_interactive.reference_arm(args, rank)
        else:
            reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    GNI, GNJ = args.gni, args.gnj
    if args.workload == "bonded":
        import bench_bonded
        bench_bonded.main(args, rank, world, local_rank)
        return
    if args.workload == "interactive":
        import bench_interactive
        bench_interactive.main(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from icebergs_b200 import api

    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_per = args.bergs_per_gpu or (N1_BERGS if world == 1 else WEAK_BERGS)
    halo = 4
    parity = None
    if multi:
        from icebergs_b200 import parallel
        dom = parallel.make_domain(GNI, GNJ, rank, world, halo=halo, device=local_rank)
        trace("domain + communicator ready")
        if not args.no_parity:
            parity = nccl_parity_check(rank, world, local_rank, dom.c.nccl_comm)
            trace(f"NCCL parity check done: {parity}")
            flag = torch.tensor([0 if (parity is None or parity["ok"]) else 1], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag[0]):
                if rank == 0:
                    print(json.dumps({"error": "NCCL data-path parity check failed", "parity_nccl": parity}), flush=True)
                dist.destroy_process_group()
                sys.exit(3)
    else:
        dom = api.Domain.single(GNI, GNJ, halo=halo, cyclic_x=True, device=local_rank)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks is not None:
        clocks.start()
    r = run_drift(args, rank, world, local_rank, dom, n_per, clocks, not args.no_e2e)
    weak_base = None
    if world == 1 and not args.no_weak_base and n_per == N1_BERGS and GNI == 1440:
        w = run_drift(args, rank, world, local_rank, dom, WEAK_BERGS, None, False)
        weak_base = {"bergs": WEAK_BERGS, "value": w["value"], "ms_per_step": w["ms_per_step"], "kernel_ms": w["kern_ms"],
                     "sorts_in_window": w["sorts"], "sort_ms_per_call": w["sort_ms_per_call"],
                     "note": "k=1 point of the weak-scaling series (BASELINE.md: 1.25e7 bergs per GPU at k=1,2,4,8), same run, same "
                             "grid; the N>1 lines of this bench run this population per GPU"}
    clk = clocks.stop() if clocks is not None else None

    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic bytes of one launch on rank 0: the bergs alive at the end of the window (a few melt on the way)
        alg_bytes = B_BERG * r["n_alive_rank0"] + B_CELL * r["occupied"]
        achieved = alg_bytes / (r["kern_ms"] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(n_per),
                       "bergs_total": int(r["n_total"]), "layout": [int(dom.layout_x), int(dom.layout_y)],
                       "l2_policy": f"inputs larger than L2 ({B_BERG * n_per / 1e9:.2f} GB of berg state per step vs 126 MB L2)",
                       "timed_region": "kid_step_resident(K): fused dyn+thermo kernel, flux-field zeroing, the periodic cell sort "
                                       "(placed so that every window holds ceil(K/sort_interval) sorts)"
                                       + (", NCCL migration" if multi else ""),
                       "sorts_in_window": r["sorts"], "sort_ms_per_call": r["sort_ms_per_call"],
                       "wall_ms_per_step": r["wall_ms"] / args.steps, "migration_ms_per_step": r["comm_ms"],
                       "sort_ms_per_step": r["sort_ms"],
                       "physics": "namelist defaults except Verlet stepping, bergy_bit_erosion_fraction=0.1, tau_is_velocity, "
                                  "add_weight_to_ocean off (the metric is dyn+thermo; mass spreading is SURVEY 8f1)",
                       "sort_interval": r["interval"], "per_rank": r["per_rank"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "traffic": traffic if (world == 1 and n_per == N1_BERGS and GNI == 1440) else None,
                         "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel": "k_step_fast (fused evolve_icebergs + thermodynamics)",
                         "alg_bytes_per_launch": alg_bytes, "kernel_ms": r["kern_ms"],
                         "kernel_note": "k_step_fast + k_step_slow (the bergs the fast kernel defers: cell walks beyond one hop, "
                                        "coast bounces, exchange) timed together with CUDA events on the launching stream",
                         "slow_list_fraction": r["slow_fraction"]},
            "gpu_launches": r["launches"], "clocks": clk,
        }
        if "e2e" in r:
            line["e2e"] = r["e2e"]
        if weak_base:
            line["weak_base"] = weak_base
        if parity is not None:
            line["parity_nccl"] = parity
        if world == 1 and not args.no_cpu:
            info = cpu_build_info()
            cores = info["omp_threads"]
            nb = args.cpu_bergs or n_per
            v, cw = cpu_oracle_run(nb, args.cpu_steps, cores)
            nb1 = min(nb, 1_000_000)
            v1, cw1 = cpu_oracle_run(nb1, 2, 1, warm=0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "single_thread": {"value": v1, "cores": 1, "sample": f"{nb1} bergs x 2 steps, {cw1:.1f} s wall"},
                                    "build": info["build"],
                                    "sample": f"{nb} bergs x {args.cpu_steps} steps of the same workload (after 2 untimed steps), "
                                              f"{cw:.1f} s wall, OpenMP over cell rows on {cores} threads (omp_get_max_threads)"}
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
