#!/usr/bin/env python
"""bench.py -- berg-steps/sec (dyn+thermo) of the KID hot path on N B200s.

Workload (BASELINE.json configs[1] / BASELINE.md): synthetic free drift + melt, seeded point
bergs on a 1/4-degree global grid (1440x720) with analytic currents/winds, dt=3600 s, Verlet,
bergy bits on.  N=1: 10M bergs on one GPU.  N>1: weak scaling, the grid is decomposed like
mpp_define_layout and every rank owns bergs-per-GPU bergs of its own tile; bergs that leave
a tile migrate over NCCL.

One JSON line on rank 0 (contract in the task statement).  `--impl reference` times the CPU
oracle (the reference is Fortran+FMS and cannot be built in this image) on the host cores.
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
# dram__bytes_read.sum + dram__bytes_write.sum of one k_step launch on the N=1 workload (10M bergs), from the
# `ncu --set full` capture summarised in profiles/r1i_kstep_final_summary.txt (1.607 GB read + 1.425 GB written)
NCU_TRAFFIC_10M = 3.032e9
METRIC = "berg_steps_per_sec"
UNIT = "berg-steps/s"


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

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", -1e30), getattr(self, "t1", 1e30)
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.05]
        if len(inside) < 3:          # a 50 ms region yields few 20 ms samples: fall back to everything sampled under load
            inside = [ln for (_, ln) in self.lines]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    out = t.numpy()
    out[...] = a
    return out, t


def cpu_oracle_run(nbergs, steps, nthreads):
    """The CPU leg: the oracle port of the reference path on the host cores, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kid_oracle_py as O
    from icebergs_b200 import api
    from icebergs_b200 import synthetic as S
    O.build()
    grid = S.Grid(GNI, GNJ)
    p = S.workload_params(api.default_params)
    dom = api.Domain.single(GNI, GNJ, halo=p.halo, cyclic_x=True)
    o = O.Oracle(GNI, GNJ, DT, (1, 0.0), params=p, domain=dom, **grid.init_args())
    bergs, _ = grid.seed_bergs(nbergs)
    o.set_bergs(**bergs)
    f = grid.forcing()
    calving, hflx = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, 0.0), calving, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], hflx,
          f["cn"], f["hi"], sss=f["sss"])          # warm-up step (also uploads the forcing)
    t0 = time.perf_counter()
    o.step_again(steps, 1, 0.0, nthreads)
    wall = time.perf_counter() - t0
    tm = o.last_timing()
    o.close()
    return nbergs * steps / wall, wall, tm


def reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nb = args.cpu_bergs
    per_step = max(args.steps, 1)
    t0 = time.perf_counter()
    # warm-up W and timed K steps of the bounded sample; each "step" = one pass over the sample
    v, wall, tm = cpu_oracle_run(nb, per_step, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(10_000_000 if args.gpus == 1 else 12_500_000),
                   "sample": f"each step = one pass over a bounded sample of {nb} of those bergs on the same grid and forcing",
                   "note": "reference = NOAA-GFDL/icebergs is Fortran+FMS (no compiler here): CPU oracle port, OpenMP"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{nb} bergs x {per_step} steps, {wall:.1f} s wall, all {cores} host threads (OpenMP over cell rows)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    global GNI, GNJ
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--bergs-per-gpu", type=int, default=0, help="0 = 10M at N=1, 12.5M per GPU at N>1")
    ap.add_argument("--cpu-bergs", type=int, default=1_000_000)
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--gni", type=int, default=GNI, help="diagnostics only: another grid size (the metric is quoted on 1440x720)")
    ap.add_argument("--gnj", type=int, default=GNJ)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    GNI, GNJ = args.gni, args.gnj

    import torch
    import torch.distributed as dist
    from icebergs_b200 import api
    from icebergs_b200 import synthetic as S
    from icebergs_b200 import _cdefs as D

    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_per = args.bergs_per_gpu or (10_000_000 if world == 1 else 12_500_000)
    p = S.workload_params(api.default_params)
    if multi:
        from icebergs_b200 import parallel
        dom = parallel.make_domain(GNI, GNJ, rank, world, halo=p.halo, device=local_rank)
    else:
        dom = api.Domain.single(GNI, GNJ, halo=p.halo, cyclic_x=True, device=local_rank)
    grid = S.Grid(GNI, GNJ, dom.isc, dom.iec, dom.jsc, dom.jec)
    bergs = api.icebergs_init(GNI, GNJ, DT, (1, 0.0), params=p, domain=dom, capacity=int(n_per * 1.25) + 4096,
                              **grid.init_args())
    cols, counter = grid.seed_bergs(n_per, stream=rank)
    cell = (cols["jne"].astype(np.int64) - 1) * GNI + (cols["ine"].astype(np.int64) - 1)
    occupied = int(np.unique(cell).size)
    bergs.set_bergs(**cols)
    del cols
    f = grid.forcing()
    keep = []
    fp = {}
    for k, v in f.items():
        fp[k], t = pinned(v)
        keep.append(t)
    calving, hflx = fp["calving"], fp["calving_hflx"]

    # calving / calving_hflx are intent(inout): the caller hands in this step's calving (none in this workload) and
    # gets the unused calving and the heat flux back.  As a coupler would, the bench double-buffers the pair: the
    # next step's arrays are prepared (zeroed) on a helper thread while the call on the current pair is in flight.
    from concurrent.futures import ThreadPoolExecutor
    pair = [(calving, hflx)]
    for _ in range(1):
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

    # ---- HBM-resident throughput: forcing on the device, K steps
    run_once()                                  # uploads forcing, first step
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                          # sampled from the warm-up on: the timed region itself lasts ~50 ms
    bergs.step_resident(args.warmup, 1, 0.0)    # W untimed warm-up steps
    barrier()
    l0 = bergs.kernel_launches()
    t0 = time.perf_counter()
    bergs.step_resident(args.steps, 1, 0.0)
    barrier()
    wall = time.perf_counter() - t0
    clocks.window(t0, t0 + wall)
    l1 = bergs.kernel_launches()
    tm = bergs.last_timing()
    dev_ms = tm["total"]
    kern_ms = tm["momentum+thermodyn"] / max(tm["_"], 1.0)
    comm_ms = tm["communication"] / max(tm["_"], 1.0)
    sort_ms = tm["sort"] / max(tm["_"], 1.0)
    n_alive = bergs.count_bergs()
    t_all = torch.tensor([dev_ms, wall * 1e3, float(n_alive), kern_ms, float(occupied), comm_ms, sort_ms], dtype=torch.float64, device="cuda")
    if multi:
        tmax = t_all.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_all.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, kern_ms, comm_ms, sort_ms = float(tmax[0]), float(tmax[1]), float(tmax[3]), float(tmax[5]), float(tmax[6])
        n_total = float(tsum[2])
        gathered = [torch.zeros_like(t_all) for _ in range(world)]
        dist.all_gather(gathered, t_all)
        per_rank = {"kernel_ms": [round(float(g[3]), 4) for g in gathered], "migration_ms": [round(float(g[5]), 4) for g in gathered],
                    "bergs": [int(g[2]) for g in gathered], "occupied_cells": [int(g[4]) for g in gathered]}
    else:
        per_rank = None
        wall_ms, n_total = wall * 1e3, float(n_alive)
    ms_per_step = dev_ms / args.steps
    value = n_total * args.steps / (dev_ms * 1e-3)

    # ---- end to end through icebergs_run: host buffers, H2D of the forcing + D2H of the returns
    e2e = None
    if not args.no_e2e:
        for _ in range(3):
            run_once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run_once()
        barrier()
        e2e_wall = time.perf_counter() - t0
        tw = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        if multi:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        h2d = sum(fp[k].nbytes for k in ("calving", "uo", "vo", "ui", "vi", "tauxa", "tauya", "ssh", "sst",
                                          "calving_hflx", "cn", "hi", "sss"))
        d2h = calving.nbytes + hflx.nbytes
        e2e = {"value": n_total * args.steps / float(tw[0]), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(tw[0]) / args.steps,
               "note": "icebergs_run through the C ABI with pinned host arrays: 13 forcing fields H2D and the two inout "
                       "fields D2H every step; the caller double-buffers the inout pair (the next pair is zeroed on a "
                       "helper thread while the call runs)"}
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = B_BERG * n_per + B_CELL * occupied
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(n_per),
                       "bergs_total": int(n_total), "layout": [int(dom.layout_x), int(dom.layout_y)],
                       "l2_policy": f"inputs larger than L2 ({B_BERG * n_per / 1e9:.2f} GB of berg state per step vs 126 MB L2)",
                       "timed_region": "kid_step_resident(K): fused dyn+thermo kernel, flux-field zeroing, periodic cell sort"
                                       + (", NCCL migration" if multi else ""),
                       "wall_ms_per_step": wall_ms / args.steps, "migration_ms_per_step": comm_ms, "sort_ms_per_step": sort_ms,
                       "physics": "namelist defaults except Verlet stepping, bergy_bit_erosion_fraction=0.1, tau_is_velocity, "
                                  "add_weight_to_ocean off (the metric is dyn+thermo; mass spreading is SURVEY 8f1)",
                       "sort_interval": int(os.environ.get("KID_SORT_INTERVAL", "32")), "per_rank": per_rank},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_10M if (world == 1 and n_per == 10_000_000 and GNI == 1440) else None,
                         "traffic_source": "ncu --set full capture of this workload, profiles/r1i_kstep_final_summary.txt",
                         "peak_source": peak_src, "kernel": "k_step (fused evolve+thermodynamics)",
                         "alg_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms},
            "gpu_launches": int(l1 - l0), "clocks": clk,
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            v, cw, ctm = cpu_oracle_run(args.cpu_bergs, args.cpu_steps, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_bergs} bergs x {args.cpu_steps} steps of the same workload, "
                                              f"{cw:.1f} s wall, OpenMP over cell rows on all {cores} host threads"}
        print(json.dumps(line), flush=True)
    api.icebergs_end(bergs)
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
