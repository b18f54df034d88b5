"""bench.py --workload interactive: BASELINE.json configs[0] scaled up -- unbonded bergs that collide through
interactive_force (I:480-607, calculate_force I:611-804) on the 1/4-degree grid of the drift workload.

Physics: the drift workload's namelist plus interactive_icebergs_on=T under Verlet stepping with the predictor-corrector
evaluation of the contact forces (use_new_predictive_corrective, forced T under Verlet I:2012): every step each berg walks
the bergs of its 3x3 cells twice (I:2153, I:2217), the store is re-sorted by cell and the halo copies are rebuilt
(update_halo_icebergs F:1800).  One "step" = one such step for every berg; metric berg-steps/s.

N > 1: the tiles of the drift workload (mpp_define_layout), bergs and halo copies travel over NCCL every step.

The CPU legs (cpu_baseline, --impl reference) run the oracle port (gcc -O3 -march=native -fopenmp, all host threads in the
first sweep of evolve_icebergs) on a bounded sample: fewer bergs on the same grid, stated in `sample`."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC, UNIT = "berg_steps_per_sec", "berg-steps/s"
GNI, GNJ, DT = 1440, 720, 3600.0
B_BERG, B_NEIGHBOUR = 290.0, 88.0


def params_of(default_params, S):
    return S.workload_params(default_params, halo=4, old_bug_bilin=0, interactive_icebergs_on=1, use_new_predictive_corrective=1)


def workload_text(n):
    return (f"interacting free bergs: {n} seeded bergs per GPU on the 1/4-degree 1440x720 grid, contact forces through "
            f"interactive_force (3x3 cells, predictor + corrector), analytic currents/winds, dt={DT:.0f} s, Verlet, bergy bits on")


def candidates_per_berg(ine, jne):
    """mean number of bergs in the 3x3 cells around a berg (what one evaluation of interactive_force visits)"""
    cnt = np.zeros((GNJ + 2, GNI + 2))
    np.add.at(cnt, (jne, ine), 1.0)
    cnt[:, 0] += cnt[:, GNI]; cnt[:, GNI + 1] += cnt[:, 1]           # cyclic in x
    s = sum(np.roll(np.roll(cnt, dj, 0), di, 1) for dj in (-1, 0, 1) for di in (-1, 0, 1))
    return float((cnt[1:-1, 1:-1] * s[1:-1, 1:-1]).sum() / max(cnt[1:-1, 1:-1].sum(), 1.0))


def oracle_run(n, steps, warm):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kid_oracle_py as O
    O.use_fast_build()
    S = O.load_by_path("synthetic", os.path.join("icebergs_b200", "synthetic.py"))
    g = S.Grid(GNI, GNJ)
    p = params_of(O.default_params, S)
    dom = O.SingleDomain(GNI, GNJ, halo=4, cyclic_x=True)
    o = O.Oracle(GNI, GNJ, DT, (1, 0.0), params=p, domain=dom, **g.init_args())
    cols, counter = g.seed_bergs(n)
    c0 = np.zeros((dom.njd, dom.nid), dtype=np.int32)
    c0[4:4 + GNJ, 4:4 + GNI] = counter
    o.set_calving_state(iceberg_counter_grd=c0)
    o.set_bergs(**cols)
    f = g.forcing()
    c, h = f["calving"].copy(), f["calving_hflx"].copy()
    o.run((1, 0.0), c, f["uo"], f["vo"], f["ui"], f["vi"], f["tauxa"], f["tauya"], f["ssh"], f["sst"], h, f["cn"], f["hi"], sss=f["sss"])
    threads = max(1, int(O.lib().oracle_omp_max_threads()))
    if warm > 0:
        o.step_again(warm, 1, 0.0, threads)
    t0 = time.perf_counter()
    o.step_again(steps, 1, 0.0, threads)
    wall = time.perf_counter() - t0
    nb = o.count_bergs()
    o.close()
    return nb * steps / wall, wall, nb, threads


def reference_arm(args, rank):
    if rank != 0:
        return
    n = args.cpu_bergs or 1_000_000
    steps = max(min(args.steps, 3), 1)
    v, wall, nb, threads = oracle_run(n, steps, 1)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.bergs_per_gpu or 10_000_000), "same_config": False,
                       "sample": f"each step = one pass over a bounded sample of {nb} bergs on the same grid and forcing "
                                 "(fewer candidates per 3x3 cells than the GPU arm's population: the per-berg CPU cost is flattered)",
                       "note": "reference = NOAA-GFDL/icebergs is Fortran+FMS (no Fortran compiler here): CPU oracle port, "
                               "gcc -O3 -march=native -fopenmp (oracle/Makefile `fast`)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{nb} bergs x {steps} steps, {wall:.1f} s wall; the first sweep of evolve_icebergs and the "
                                       "thermodynamics run on all threads, the cell lists and halo copies on one"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main(args, rank, world, local_rank):
    import torch
    from icebergs_b200 import api, parallel
    from icebergs_b200 import synthetic as S
    sys.path.insert(0, ROOT)
    from bench import ClockSampler, measured_peak, pinned
    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_per = args.bergs_per_gpu or 10_000_000
    dom = parallel.make_domain(GNI, GNJ, rank, world, halo=4, device=local_rank)
    grid = S.Grid(GNI, GNJ, dom.isc, dom.iec, dom.jsc, dom.jec)
    p = params_of(api.default_params, S)
    b = api.icebergs_init(GNI, GNJ, DT, (1, 0.0), params=p, domain=dom, capacity=int(1.4 * n_per) + 65536, **grid.init_args())
    cols, _ = grid.seed_bergs(n_per, stream=rank)
    cand = candidates_per_berg(cols["ine"], cols["jne"]) if not multi else None
    b.set_bergs(**cols)
    del cols
    f = grid.forcing()
    keep, fp = [], {}
    for k, v in f.items():
        fp[k], t = pinned(v)
        keep.append(t)

    def run_call():
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
    run_call()
    b.step_resident(args.warmup, 1, 0.0)
    barrier()
    l0 = b.kernel_launches()
    t0 = time.perf_counter()
    b.step_resident(args.steps, 1, 0.0)
    api.lib().kid_synchronize(b.handle)
    barrier()
    wall = time.perf_counter() - t0
    tm = b.last_timing()
    if clocks is not None:
        clocks.mark(t0, t0 + wall)
    l1 = b.kernel_launches()
    n_now = b.counters()["nbergs"]
    # end to end: icebergs_run with pinned host arrays, every call copies its 13 fields in and the two inout fields back
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        run_call()
    barrier()
    e2e_wall = time.perf_counter() - t1
    if clocks is not None:
        clocks.mark(t1, t1 + e2e_wall)
    err = b.counters()["error_flags"]
    tw = torch.tensor([wall, e2e_wall, tm["momentum+thermodyn"], tm["sort"]], dtype=torch.float64, device="cuda")
    ntot = torch.tensor([float(n_now)], dtype=torch.float64, device="cuda")
    if multi:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        dist.all_reduce(ntot, op=dist.ReduceOp.SUM)
    clk = clocks.stop() if clocks is not None else None
    api.icebergs_end(b)
    if rank == 0:
        peak, peak_src = measured_peak()
        wall_s, e2e_s = float(tw[0]), float(tw[1])
        dyn_ms = float(tw[2]) / args.steps
        value = float(ntot[0]) * args.steps / wall_s
        h2d = sum(fp[k].nbytes for k in ("calving", "uo", "vo", "ui", "vi", "tauxa", "tauya", "ssh", "sst", "calving_hflx", "cn", "hi", "sss"))
        d2h = fp["calving"].nbytes + fp["calving_hflx"].nbytes
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(n_per), "bergs_total": int(ntot[0]), "device_error_flags": int(err),
                           "l2_policy": "inputs larger than L2 (the berg state alone is 2.9 GB at 10 M bergs vs 126 MB L2)",
                           "timed_region": "kid_step_resident(K): interaction records, first sweep with interactive_force twice per "
                                           "berg, second sweep + thermodynamics, halo copies, cell sort and bond bookkeeping of every "
                                           "step; host wall clock around the call, max over ranks",
                           "momentum_thermo_ms_per_step": dyn_ms, "sort_ms_per_step": float(tw[3]) / args.steps},
                "gpu_launches": int(l1 - l0), "clocks": clk,
                "e2e": {"value": float(ntot[0]) * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / e2e_steps,
                        "note": "icebergs_run through the C ABI with pinned host arrays (13 fields in, the two inout fields back), "
                                "plain calls, host wall clock, max over ranks"}}
        if cand is not None:
            alg = (B_BERG + B_NEIGHBOUR * cand) * float(ntot[0])
            line["roofline"] = {"bound": "hbm", "achieved": alg / (dyn_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": alg / (dyn_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                                "kernel": "k_ia_prepare + k_ia_velocity + k_step<SPLIT> (timed together with CUDA events on the launching stream)",
                                "alg_bytes_per_launch": alg, "kernel_ms": dyn_ms, "candidates_per_berg": cand,
                                "note": "algorithmic bytes = 290 B per berg + 88 B per berg of its 3x3 cells (SURVEY 8d), counted for ONE "
                                        "walk of the cells; the candidates come out of L1/L2 (bergs of a cell share them), so this figure "
                                        "can exceed what HBM delivers: it says how many candidate records per second the walk gets through"}
        if world == 1 and not args.no_cpu:
            cn = args.cpu_bergs or 1_000_000
            v, cw, cnb, threads = oracle_run(cn, 2, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "build": "gcc -O3 -march=native -fopenmp (oracle/Makefile `fast`)",
                                    "sample": f"{cnb} bergs x 2 steps of the same workload on the same grid (fewer candidates per berg "
                                              f"than the GPU arm's {n_per}), {cw:.1f} s wall"}
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()
