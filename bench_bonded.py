"""bench.py --workload bonded: BASELINE.json configs[2], "tests/a68_test scaled" -- a bonded tabular berg.

A square-packed rectangle of bonded elements (radius 1.5 km, DEM bonds) with the physics of tests/a68_test/long_run.nml
(MTS scheme: 60 sub-steps of dt/60 per step, dem, contact_distance 4 km, stress fracture on the sub-steps, grounding drag on
the short steps) in a sheared current, on the synthetic Cartesian grid of icebergs_b200/synthetic.py (the A68a forcing of the
reference test is FTP-only, tests/a68_test/get_data.sh:4).  One "step" = one icebergs_run call = one long MTS step
(evolve_icebergs_mts I:6576 with its 60 fast sub-steps, transfer / connect_all_bonds / set_conglom_ids, thermodynamics).

Metric: element-steps per second (unit "berg-steps/s": every element takes the long step), also quoted as element
sub-steps per second.  The population is 1e2..1e4 elements: the path is latency-bound (sub-step sweeps with grid-wide
dependencies), not HBM-bound; `roofline` reports the algorithmic bytes of SURVEY 8(d) (290 B per element + 88 B per
neighbour visited) against the measured HBM peak to say exactly that.

The CPU legs (cpu_baseline, --impl reference) run the oracle port of the same step on one host thread (the reference's
bonded path is serial per PE; its tests spread 16 bergs over 4 PEs)."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC, UNIT = "berg_steps_per_sec", "berg-steps/s"
DT = 1800.0
SUBSTEPS = 60
B_BERG, B_NEIGHBOUR = 290.0, 88.0


def shape_of(elements):
    """nx x ny elements: the 12 x 16 berg of tests/test_mts_gpu.py::test_bonded_tabular_berg_a68_physics scaled up at the
    same 3:4 aspect, kept west of the shoal and inside the 300 km domain."""
    n = elements or 432          # A68a in December 2020: ~3900 km2 = ~430 elements of (3 km)^2 (tests/a68_test/makeberg/RUN:7)
    nx = int(min(29, max(2, round((0.75 * n) ** 0.5))))
    ny = int(min(90, max(2, -(-n // nx))))
    return nx, ny


def workload_text(nx, ny):
    return (f"bonded tabular berg: {nx} x {ny} = {nx * ny} square-packed DEM-bonded elements (r = 1.5 km) with the physics of "
            f"tests/a68_test/long_run.nml (mts, {SUBSTEPS} sub-steps, dem, contact_distance 4 km, stress fracture, grounding), "
            f"60x60 Cartesian grid of 5 km cells, sheared current + shoal, dt={DT:.0f} s")


def berg_columns(S, nx, ny):
    r = 1500.0
    x0 = 96.0e3 - 2.0 * r * (nx - 1)            # east edge 3 km short of the shoal, as in the parity test
    y0 = max(15.0e3, 84.0e3 - r * (ny - 1))     # centred on y = 84 km (the parity test's berg), inside the domain
    return S.tabular_berg(nx=nx, ny=ny, r=r, x0=x0, y0=y0)


def neighbours_per_element(nx, ny):
    """bonded neighbours per element of the square packing (what the DEM pair sweep visits)"""
    return (2.0 * ((nx - 1) * ny + nx * (ny - 1))) / (nx * ny)


def oracle_run(elements, steps, warm):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kid_oracle_py as O
    O.use_fast_build()
    S = O.load_by_path("synthetic", os.path.join("icebergs_b200", "synthetic.py"))
    nx, ny = shape_of(elements)
    g = S.TabularGrid()
    p = S.a68_params(O.default_params)
    dom = O.SingleDomain(g.gni, g.gnj, halo=p.halo, cyclic_x=True)
    o = O.Oracle(g.gni, g.gnj, DT, (1, 0.0), params=p, domain=dom, **g.init_args())
    o.set_bergs(**berg_columns(S, nx, ny))
    o.set_bonds()
    f = g.forcing()

    def step():
        c, h = f["calving"].copy(), f["calving_hflx"].copy()
        o.run((1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    wall = time.perf_counter() - t0
    n = o.count_bergs()
    o.close()
    return n * steps / wall, wall, n, (nx, ny)


def reference_arm(args, rank):
    if rank != 0:
        return
    steps = max(args.steps, 1)
    v, wall, n, (nx, ny) = oracle_run(args.elements, steps, min(max(args.warmup, 1), 3))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(nx, ny), "same_config": True,
                       "note": "reference = NOAA-GFDL/icebergs is Fortran+FMS (no Fortran compiler here): CPU oracle port, "
                               "gcc -O3 -march=native (oracle/Makefile `fast`), one thread (the bonded path is serial per PE)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{n} elements x {steps} long steps ({SUBSTEPS} sub-steps each), {wall:.1f} s wall"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main(args, rank, world, local_rank):
    """replicas only: a conglomerate is sub-stepped whole on every rank that touches it (transfer_mts_bergs), so N ranks
    run N independent bergs here; N = 1 is the figure."""
    import torch
    from icebergs_b200 import api
    from icebergs_b200 import synthetic as S
    sys.path.insert(0, ROOT)
    from bench import ClockSampler, measured_peak, pinned
    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nx, ny = shape_of(args.elements)
    g = S.TabularGrid()
    p = S.a68_params(api.default_params)
    dom = api.Domain.single(g.gni, g.gnj, halo=p.halo, cyclic_x=True, device=local_rank)
    b = api.icebergs_init(g.gni, g.gnj, DT, (1, 0.0), params=p, domain=dom, capacity=max(8192, 8 * nx * ny), **g.init_args())
    b.set_bergs(**berg_columns(S, nx, ny))
    b.set_bonds()
    f = g.forcing()
    keep, fp = [], {}
    for k, v in f.items():
        fp[k], t = pinned(v)
        keep.append(t)

    def step():
        fp["calving"].fill(0.0); fp["calving_hflx"].fill(0.0)
        api.icebergs_run(b, (1, 0.0), fp["calving"], fp["uo"], fp["vo"], fp["ui"], fp["vi"], fp["tauxa"], fp["tauya"], fp["ssh"],
                         fp["sst"], fp["calving_hflx"], fp["cn"], fp["hi"], sss=fp["sss"])

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks is not None:
        clocks.start()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = b.kernel_launches()
    t0 = time.perf_counter()
    dyn_ms = 0.0
    for _ in range(args.steps):
        step()
        dyn_ms += b.last_timing()["momentum+thermodyn"]
    barrier()
    wall = time.perf_counter() - t0
    if clocks is not None:
        clocks.mark(t0, t0 + wall)
    l1 = b.kernel_launches()
    n = b.count_bergs()
    bonds = len(b.get_bonds()["first_id"])
    tw = torch.tensor([wall, dyn_ms], dtype=torch.float64, device="cuda")
    ntot = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
    if multi:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        dist.all_reduce(ntot, op=dist.ReduceOp.SUM)
    if clocks is not None:
        t_a = time.perf_counter()
        while time.perf_counter() - t_a < 0.6:
            step()
        clocks.mark(t_a, time.perf_counter())
    clk = clocks.stop() if clocks is not None else None
    api.icebergs_end(b)
    if rank == 0:
        peak, peak_src = measured_peak()
        wall_s, dyn = float(tw[0]), float(tw[1]) / args.steps
        value = float(ntot[0]) * args.steps / wall_s
        nb = neighbours_per_element(nx, ny)
        # per sub-step every element is read and written once and visits its bonded neighbours
        alg_bytes = (B_BERG + B_NEIGHBOUR * nb) * n * SUBSTEPS
        achieved = alg_bytes / (dyn * 1e-3) / 1e9
        h2d = sum(fp[k].nbytes for k in ("calving", "uo", "vo", "ui", "vi", "tauxa", "tauya", "ssh", "sst", "calving_hflx", "cn", "hi", "sss"))
        d2h = fp["calving"].nbytes + fp["calving_hflx"].nbytes
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(nx, ny), "elements": int(n), "half_bonds": int(bonds), "substeps_per_step": SUBSTEPS,
                           "element_substeps_per_sec": value * SUBSTEPS,
                           "us_per_substep": 1e3 * dyn / SUBSTEPS,
                           "l2_policy": "the element state (a few MB at most) lives in L2 / shared memory by design: this path is "
                                        "latency-bound, not HBM-bound",
                           "timed_region": "icebergs_run through the C ABI (host arrays in, inout fields back): forcing ingest, "
                                           "evolve_icebergs_mts (parts 1-3), transfer / sort / connect_all_bonds / set_conglom_ids, "
                                           "thermodynamics; host wall clock, max over ranks",
                           "replicas": "N>1: one independent berg per GPU (the path does not shard: a conglomerate is "
                                       "sub-stepped whole on every rank it touches)"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                             "peak_source": peak_src, "kernel": "the MTS sub-step loop (k_mts_substeps_one_cta, or one kernel per sweep)",
                             "alg_bytes_per_step": alg_bytes, "kernel_ms": dyn,
                             "note": "algorithmic bytes = (290 + 88 x bonded neighbours) per element and sub-step; the loop is bound by "
                                     "the latency of its grid-wide dependencies, the fraction only says how far from a bandwidth "
                                     "problem this population is"},
                "gpu_launches": int(l1 - l0), "clocks": clk,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": 1e3 * wall_s / args.steps,
                        "note": "the timed region IS the end-to-end call (this workload has no device-resident variant)"}}
        if world == 1 and not args.no_cpu:
            v, cw, cn, _ = oracle_run(args.elements, max(2, min(args.steps, 10)), 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "build": "gcc -O3 -march=native (oracle/Makefile `fast`)",
                                    "sample": f"{cn} elements x {max(2, min(args.steps, 10))} long steps of the same workload, {cw:.1f} s wall, "
                                              "one thread (the bonded path is serial per PE)"}
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()
