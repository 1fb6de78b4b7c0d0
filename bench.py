#!/usr/bin/env python
"""bench.py — headline benchmark of the two hot paths (contract: see the task statement / DESIGN.md §measurement).

  python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

One JSON line on rank 0. A "step" is one pass of the hot path over one batch of synthetic input:
  sssp_*  : one mz_sssp_batch over `trees_per_step` (root link, entry time) trees per GPU   → edge relaxations/s
  sim_*   : one mz_sim_run over the whole horizon of the synthetic network               → vehicle-link traversals/s
The default workload is the north star's target size: `sim_1m` (1.02 M links, 10.2 M trips, ONE network — partitioned over
the ranks for N > 1, strong scaling) as the line's headline metric, with BASELINE.json's second metric measured in the same
run on config 4 (`sssp_cfg4`, origins sharded over the ranks) in the line's "sssp" sub-object.
`config` holds only the workload's static description, so both arms (ours / --impl reference) print the same object;
what a run did (links, events, visits …) goes to "work".
`value`  = device-resident throughput (inputs and outputs in HBM), timed with CUDA events on the engine's stream.
`e2e`    = the same metric through the host-buffer C-ABI call (H2D of the step's inputs + D2H of its results inside).
`cpu_baseline` = the reference's own CPU code (oracle/_ref, kind "reference") or the C port (oracle/, kind "port")
                 timed on this box's host cores on a bounded sample of the same workload (rank 0, N=1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# ------------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (recipe: B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# DRAM traffic of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum per launch) from the committed
# `ncu --set full` captures (profiles/r01_ncu_*_full_summary.txt); only valid for the default sizes of these workloads
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu captures
NCU_TRAFFIC = {"sim_grid100": (0.2792e9, "profiles/r02_ncu_k_sim_async_cfg2_full_summary.txt"),
               "sim_1m": (849.5e9, "profiles/r02_ncu_k_sim_async_sim_1m_sections.txt (dram__bytes.sum.per_second x duration)"),
               "sssp_grid100": (54.31e9, "profiles/r01_ncu_k_fast_run_full_summary.txt"),
               "sssp_cfg4": (1167.3e9, "profiles/r02_ncu_k_fast_run_cfg4_full_summary.txt")}


def ncu_traffic(name, default_size):
    if default_size and name in NCU_TRAFFIC:
        return {"traffic": NCU_TRAFFIC[name][0], "traffic_source": NCU_TRAFFIC[name][1]}
    return {"traffic": None}


# ------------------------------------------------------------------------------------------------ SSSP workloads
SSSP_WORKLOADS = {
    # config 4 of BASELINE.json: 1M-node / 2.5M-link thinned grid, 16-period table
    # the table spans the longest path from the latest entry time (16 x 7200 s), as a histtimes file covering the
    # analysis horizon does; trees that still run past its end are redone by the exact kernel and reported
    "sssp_cfg4": dict(nx=1000, ny=1000, keep=0.625, len_lo=50.0, len_hi=500.0, P=16, periodlength=7200.0,
                      n_entries=4, trees_per_step=1024, n_dests=16, cpu_trees=6,
                      desc="all-origin time-dependent SSSP, 1000x1000 grid thinned to ~2.5M links, P=16, "
                           "4 entry times per origin root link"),
    # the route-generation part of config 2 (100x100 grid, ~39.6k links)
    "sssp_grid100": dict(nx=100, ny=100, keep=1.0, len_lo=300.0, len_hi=700.0, P=4, periodlength=3600.0,
                         n_entries=2, trees_per_step=8192, n_dests=16, cpu_trees=256,
                         desc="route generation on the 100x100 grid of config 2 (39.6k links), P=4, "
                              "2 entry times per origin root link"),
}


def sssp_inputs(w, rank, world, step, n_trees):
    """Deterministic (root, entry, dests) for one step of one rank: origins sharded across ranks."""
    rng = np.random.default_rng(1234 + 7919 * step + 104729 * rank)
    L, N = w["n_links"], w["n_nodes"]
    # shortest_paths_all runs one tree per (origin root link, entry time): every sampled root gets all entry times
    E = w["n_entries"]
    roots = np.repeat(rng.integers(0, L, (n_trees + E - 1) // E), E)[:n_trees].astype(np.int32)
    entries = (np.tile(np.arange(E), (n_trees + E - 1) // E)[:n_trees] * w["periodlength"]).astype(np.float64)
    dests = rng.integers(0, N, n_trees * w["n_dests"]).astype(np.int32)
    dest_off = (np.arange(n_trees + 1, dtype=np.int64) * w["n_dests"])
    return roots, entries, dest_off, dests


def sssp_config(name, n_trees, mode="auto"):
    """Static description of an SSSP workload — identical in both arms."""
    w = SSSP_WORKLOADS[name]
    return {"workload": name, "desc": w["desc"], "grid": f"{w['nx']}x{w['ny']}", "periods": w["P"],
            "entry_times_per_root": w["n_entries"], "trees_per_step_per_gpu": n_trees, "kernel": mode,
            "l2": "per-step tree state (16 B x links x trees) exceeds the 126 MB L2; no flush needed",
            "sharding": "origins (trees) sharded across ranks, graph+table replicated, no collective"}


def build_sssp_workload(name):
    from mezzo_b200 import synth
    w = dict(SSSP_WORKLOADS[name])
    gl = synth.grid_links(w["nx"], w["ny"], seed=42, len_lo=w["len_lo"], len_hi=w["len_hi"], keep_prob=w["keep"])
    w.update(n_nodes=gl["n_nodes"], n_links=gl["n_links"], up=gl["up"], dn=gl["dn"])
    w["T"] = synth.link_time_table(gl["length"], w["P"], seed=42, fifo=True)
    return w


def sssp_cpu_reference(w, n_trees, kind_pref="reference"):
    """Times the reference's own labelCorrecting (oracle/_ref) — or the C port if _ref is absent — on n_trees trees."""
    import orc
    spec = orc.GraphSpec(w["n_nodes"], w["n_links"], w["up"], w["dn"])
    roots, entries, _, _ = sssp_inputs(w, 0, 1, 0, n_trees)
    if kind_pref == "reference" and orc.ref_lib() is not None:
        rg = orc.RefGraph(spec)
        rg.set_linktimes(w["P"], w["periodlength"], w["T"])
        kind = "reference"
        og = orc.OracleGraph(spec)  # only to count canonical relaxations of the same trees (untimed)
        og.set_linktimes(w["P"], w["periodlength"], w["T"])
        t0 = time.perf_counter()
        lib, h = rg.lib, rg.h
        for r, e in zip(roots, entries):
            lib.ref_label_correcting(h, int(r), float(e), None, None, None)
        dt = time.perf_counter() - t0
        _, _, _, _, canon = og.batch(roots, entries, want_pred=False)
    else:
        og = orc.OracleGraph(spec)
        og.set_linktimes(w["P"], w["periodlength"], w["T"])
        kind = "port"
        t0 = time.perf_counter()
        _, _, _, _, canon = og.batch(roots, entries, want_pred=False)
        dt = time.perf_counter() - t0
    return canon / dt, kind, dt, canon


def run_sssp(args, name):
    import torch
    import mezzo_b200
    from mezzo_b200 import _lib

    rank, world, local = dist_env()
    w = build_sssp_workload(name)
    n_trees = args.trees_per_step or w["trees_per_step"]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    g = mezzo_b200.Graph.from_arrays(w["n_nodes"], w["n_links"], w["up"], w["dn"], device=local)
    g.set_linktimes(w["T"], w["P"], w["periodlength"])
    mode = mezzo_b200.MZ_SSSP_EXACT if args.sssp_mode == "exact" else mezzo_b200.MZ_SSSP_AUTO
    L, N = w["n_links"], w["n_nodes"]

    # ---- device-resident arm: roots/entries and all outputs live in HBM; timed with CUDA events on the engine stream
    stream = torch.cuda.Stream(device=dev)
    g.set_stream(stream.cuda_stream)
    g.set_option(_lib.MZ_OPT_CHUNK_TREES, n_trees)  # one chunk per step so outputs are written in place
    out = (torch.empty((n_trees, L), dtype=torch.int32, device=dev),
           torch.empty((n_trees, N), dtype=torch.int32, device=dev),
           torch.empty((n_trees, N), dtype=torch.float64, device=dev))
    steps_in = []
    for s in range(args.warmup + args.steps):
        r, e, _, _ = sssp_inputs(w, rank, world, s, n_trees)
        steps_in.append((r, e))  # host copies are validated/uploaded by the call; 12 B per tree, negligible

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    canon_total, eval_total, reached_total, main_ms, kern_ms, launches = 0, 0, 0, 0.0, 0.0, 0
    redo_total, fast_steps = 0, 0
    for s in range(args.warmup):
        g.sssp_batch(*steps_in[s], mode=mode, out=out)
    # The metric's unit — canonical relaxations — is counted by a kernel of its own (k_count_canonical) that
    # neither the reference nor a mezzo_lib caller runs: count the timed steps' inputs in an untimed pass, then time the
    # same inputs with the counter switched off (MZ_OPT_COUNT_CANONICAL = 0). Results of the two passes are identical.
    for s in range(args.warmup, args.warmup + args.steps):
        g.sssp_batch(*steps_in[s], mode=mode, out=out)
        st = g.last_stats()
        canon_total += st["canonical_relax"]
        reached_total += st["reached_links"]
    g.set_option(_lib.MZ_OPT_COUNT_CANONICAL, 0)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for s in range(args.warmup, args.warmup + args.steps):
            g.sssp_batch(*steps_in[s], mode=mode, out=out)
            st = g.last_stats()
            eval_total += st["evaluated_relax"]
            main_ms += st["main_kernel_ms"]
            kern_ms += st["kernel_ms"]
            launches += st["n_launches"]
            redo_total += st["n_exact_redo"]
            fast_steps += st["fast_steps"]
        ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    dev_ms = ev0.elapsed_time(ev1)
    main_launches = args.steps

    # ---- e2e arm: host buffers through the reference-facing call (trees + route extraction for the OD destinations)
    g.set_stream(None)
    del out  # the device-resident arm's output tensors (22 GB on config 4): the e2e call sizes its chunks from free memory
    torch.cuda.empty_cache()
    g.set_option(_lib.MZ_OPT_CHUNK_TREES, 0)
    g.set_option(_lib.MZ_OPT_COUNT_CANONICAL, 1)
    e2e_in = [sssp_inputs(w, rank, world, 1000 + s, n_trees) for s in range(1 + args.steps)]
    g.sssp_routes(e2e_in[0][0], e2e_in[0][1], e2e_in[0][2], e2e_in[0][3], mode=mode)
    barrier()
    t0 = time.perf_counter()
    e2e_canon, h2d, d2h, e2e_chunks, e2e_kernel_ms = 0, 0, 0, 0, 0.0
    for s in range(1, 1 + args.steps):
        r, e, doff, d = e2e_in[s]
        poff, links = g.sssp_routes(r, e, doff, d, mode=mode)
        st = g.last_stats()
        e2e_canon += st["canonical_relax"]
        e2e_chunks += st["n_chunks"]
        e2e_kernel_ms += st["kernel_ms"]
        h2d += r.nbytes + e.nbytes + d.nbytes + d.size * 4 + (d.size + 1) * 8  # inputs + pair->tree map + offsets
        d2h += links.nbytes + d.size * 4
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks: time = max, work = sum
    vals = torch.tensor([dev_ms, e2e_s, main_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([canon_total, eval_total, e2e_canon, launches], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(vals, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(sums, op=torch.distributed.ReduceOp.SUM)
    dev_ms_max, e2e_s_max, main_ms_max = vals.tolist()
    canon_all, eval_all, e2e_canon_all, launches_all = sums.tolist()
    g.close()
    if rank != 0:
        return None

    peaks, peak_src = measured_peaks()
    # algorithmic bytes (DESIGN.md / SURVEY §8d): 21 B per canonical relaxation + 12 B per reached link (label+pred
    # update) + 16 B/link-slot init + 4 B/(link+node slot) pred output, per tree — for the dominant (tree) kernel only
    # the relaxation terms apply: 21 B/relaxation + 12 B per reached link; init/output belong to the other kernels.
    alg_bytes_main = 21.0 * canon_total + 12.0 * reached_total
    achieved = alg_bytes_main / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
    line = {
        "metric": "SSSP edge relaxations/s", "value": canon_all / (dev_ms_max * 1e-3), "unit": "relaxations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": sssp_config(name, n_trees, args.sssp_mode),
        "work": {"links": L, "nodes": N, "canonical_relaxations_per_step": canon_all / args.steps,
                 "exact_redo_trees_per_step": redo_total / args.steps, "near_far_steps_per_step": fast_steps / args.steps},
        "e2e": {"value": e2e_canon_all / e2e_s_max, "unit": "relaxations/s", "ms_per_step": 1e3 * e2e_s_max / args.steps,
                "device_ms_per_step": e2e_kernel_ms / args.steps, "chunks_per_step": e2e_chunks / args.steps,
                "h2d_bytes_per_step": h2d // args.steps,
                "d2h_bytes_per_step": d2h // args.steps,
                "call": f"mz_sssp_routes (trees + {w['n_dests']} destination paths per tree), host buffers"},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "peak_source": peak_src,
                     **ncu_traffic(name, args.sssp_mode == "auto" and n_trees == w["trees_per_step"]),
                     "kernel": "k_sssp_exact" if args.sssp_mode == "exact" else "k_fast_run",
                     "kernel_ms_per_launch": main_ms / max(1, main_launches),
                     "algorithmic_bytes_per_launch": alg_bytes_main / max(1, main_launches),
                     "evaluated_over_canonical": eval_total / max(1, canon_total),
                     "all_kernels_ms_per_step": kern_ms / args.steps},
    }
    if world == 1 and not args.no_cpu_baseline:
        v, kind, dt, canon = sssp_cpu_reference(w, args.cpu_trees or w["cpu_trees"])
        line["cpu_baseline"] = {"value": v, "unit": "relaxations/s", "cores": 1, "kind": kind,
                                "sample": f"{args.cpu_trees or w['cpu_trees']} trees of the same workload, {dt:.1f} s, "
                                          f"single thread (mezzo_lib is single-threaded); host has {os.cpu_count()} cpus"}
    return line


def run_sssp_reference(args, name):
    """--impl reference: the reference's own CPU labelCorrecting on this box's host cores, same config/metric."""
    rank, world, _ = dist_env()
    if rank != 0:
        return None
    w = build_sssp_workload(name)
    n = args.cpu_trees or max(1, w["cpu_trees"] // 2)
    for _ in range(args.warmup and 1):
        sssp_cpu_reference(w, 1)
    tot_c, tot_t, kind = 0, 0.0, "reference"
    for _ in range(args.steps):
        v, kind, dt, canon = sssp_cpu_reference(w, n)
        tot_c += canon
        tot_t += dt
    val = tot_c / tot_t
    return {"impl": "reference", "metric": "SSSP edge relaxations/s", "value": val, "unit": "relaxations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": sssp_config(name, args.trees_per_step or w["trees_per_step"]),
            "cpu_baseline": {"value": val, "unit": "relaxations/s", "cores": 1, "kind": kind,
                             "sample": f"{n} trees per step, single thread; host has {os.cpu_count()} cpus"},
            "e2e": {"value": val, "unit": "relaxations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}



# ------------------------------------------------------------------------------------------------ link-update workloads
SIM_WORKLOADS = {
    # config 2 of BASELINE.json: 100x100 grid, 200k vehicles, 1 h horizon, single-class demand (SURVEY.md §8d recipe:
    # link length U(300,700) m, O/D connectors >= 200 m at every 5th junction, one N(1.8,0.3) server per turning)
    "sim_grid100": dict(gen=dict(nx=100, ny=100, seed=42, od_every=5, n_od=4000, total_trips=200_000, horizon=3600.0,
                                 connector_len=250),
                        cpu_horizon=600.0, cpu_sample_desc="the same network and demand rate, first 600 s of the horizon",
                        desc="100x100 grid, 200k vehicles, 1 h horizon, one N(1.8,0.3) server per turning"),
    "sim_grid40": dict(gen=dict(nx=40, ny=40, seed=42, od_every=5, n_od=600, total_trips=60_000, horizon=3600.0,
                                connector_len=250),
                       cpu_horizon=3600.0, cpu_sample_desc="the whole workload",
                       desc="40x40 grid, 60k vehicles, 1 h horizon (the survey's probe size)"),
    # the north star's target size: ~1 M links, 10 M trips in 1 h. OD pairs are local (destination within 4 steps of the
    # OD lattice, mean route 24 links) and routes are staircases, because an all-sources Dijkstra does not fit the host prep.
    # With --gpus N this ONE network is partitioned over the N GPUs (strong scaling).
    "sim_grid250": dict(gen=dict(nx=250, ny=250, seed=42, od_every=5, n_od=50_000, total_trips=2_500_000, horizon=3600.0,
                                 connector_len=250, od_radius=4),
                        cpu_horizon=300.0, scaling="strong",
                        cpu_gen=dict(nx=100, ny=100, n_od=8000, total_trips=400_000),
                        cpu_sample_desc="spatial + temporal sample: a 100x100-junction window of the grid (same generator, "
                                        "link lengths, OD density and demand rate per OD pair), first 300 s of the horizon",
                        desc="250x250 grid (255 k links), 2.5 M vehicles, 1 h horizon, local OD pairs (sim_1m at quarter size)"),
    "sim_1m": dict(gen=dict(nx=500, ny=500, seed=42, od_every=5, n_od=200_000, total_trips=10_000_000, horizon=3600.0,
                            connector_len=250, od_radius=4),
                   # The reference cannot load the whole network in bounded time: Network::add_od_routes is quadratic in the OD
                   # pairs (200 k pairs x 200 k routes; > 18 min of init on one core, measured). Its arm therefore runs a
                   # spatial sample — the same generator on a 100x100 window with the same OD density and demand rate.
                   cpu_horizon=300.0, scaling="strong",
                   cpu_gen=dict(nx=100, ny=100, n_od=8000, total_trips=400_000),
                   cpu_sample_desc="spatial + temporal sample: a 100x100-junction window of the grid (1/25 of the network; same "
                                   "generator, link lengths, OD density and demand rate per OD pair), first 300 s of the horizon",
                   desc="500x500 grid (1.02 M links, 3.08 M turnings), 10.2 M vehicles, 1 h horizon, local OD pairs"),
    # the same network with NON-LOCAL demand: destinations up to 25 OD-lattice steps away (125 junctions = 25 % of the grid,
    # mean route 119 links), 1.4 M trips — about the same number of traversals per horizon as sim_1m, but most trips cross
    # several partition cuts when the network is split over GPUs
    "sim_1m_far": dict(gen=dict(nx=500, ny=500, seed=42, od_every=5, n_od=200_000, total_trips=1_200_000, horizon=3600.0,
                                connector_len=250, od_radius=25),
                       cpu_horizon=300.0, scaling="strong",
                       cpu_gen=dict(nx=100, ny=100, n_od=8000, total_trips=48_000),
                       cpu_sample_desc="spatial + temporal sample: a 100x100-junction window of the grid (same generator, OD "
                                       "density and demand rate per OD pair; destinations up to 25 lattice steps away), first "
                                       "300 s of the horizon",
                       desc="500x500 grid (1.02 M links), 1.4 M vehicles with destinations up to 125 junctions away (mean "
                            "route 119 links), 1 h horizon"),
}


def sim_config(name, world):
    """The workload's static description — identical in both arms (it does not depend on anything a run measures)."""
    w = SIM_WORKLOADS[name]
    g = w["gen"]
    strong = w.get("scaling", "weak") == "strong"
    return {"workload": name, "desc": w["desc"], "grid": f"{g['nx']}x{g['ny']}" + ("" if strong or world == 1 else f" per GPU, x{world}"),
            "od_pairs": g["n_od"], "trips_per_horizon": g["total_trips"], "horizon_s": g["horizon"],
            "l2": "state is re-initialised (memset + upload) between timed steps; the working set (queues + vehicles + "
                  "exit log) exceeds the 126 MB L2",
            "sharding": "single GPU" if world == 1 else
                        "one network, nodes partitioned over the ranks (graph-growing bisection); cut links are replicated and "
                        "written through NVLink peer memory, pairwise key synchronisation, no collective"}


def build_sim_workload(name, horizon=None, world=1, sample=False):
    from mezzo_b200 import flatten, synth
    w = dict(SIM_WORKLOADS[name])
    gen = dict(w["gen"])
    if sample:  # the bounded CPU sample of the workload (see SIM_WORKLOADS)
        gen.update(w.get("cpu_gen", {}))
        horizon = w["cpu_horizon"]
    if world > 1 and w.get("scaling", "weak") == "weak":
        # weak scaling: every GPU gets a grid region of the named size — the network is gx x gy times larger,
        # with proportionally more OD pairs and trips (2 -> 2x1, 4 -> 2x2, 8 -> 4x2 regions)
        gx = {1: 1, 2: 2, 4: 2, 8: 4}.get(world, world)
        gy = world // gx
        gen["nx"] *= gx
        gen["ny"] *= gy
        gen["n_od"] *= world
        gen["total_trips"] *= world
        w["desc"] += f" — scaled {gx}x{gy} for {world} GPUs (weak scaling: one named-size region per GPU)"
    if horizon is not None and horizon < gen["horizon"]:
        # same network and demand RATE, shorter horizon: the bounded CPU sample
        gen["total_trips"] = int(round(gen["total_trips"] * horizon / gen["horizon"]))
        gen["n_periods"] = max(1, int(round(4 * horizon / gen["horizon"])))
        gen["horizon"] = horizon
    w["spec"] = synth.sim_grid(**gen)
    w["flat"] = flatten.FlatNet(w["spec"])
    return w


_SIM_SAMPLE = {}


def sim_cpu_reference(name, also_gpu=False):
    """The UNMODIFIED reference (oracle/_ref/mezzo_ref_run, Network::step timed inside the harness) on a bounded sample
    of the workload, built once per process; falls back to the C port (oracle/sim_oracle.c) when the compiled reference is
    absent. Traversals are counted from the reference's own run: link exits + vehicles still on links."""
    import refrun
    w = SIM_WORKLOADS[name]
    if name not in _SIM_SAMPLE:
        _SIM_SAMPLE[name] = build_sim_workload(name, sample=True)
    ws = _SIM_SAMPLE[name]
    if refrun.have_ref():
        r = refrun.run_reference(ws["spec"])
        secs = float(r["info"]["step_seconds"])
        trav = int(r["info"]["n_exits"]) + int(r["info"]["n_on_links"])
        kind = "reference"
    else:
        import orc
        o = orc.oracle_sim(ws["flat"])
        secs, trav, kind = o["seconds"], o["stats"]["n_traversals"], "port"
    st = ws["flat"].struct
    sample = (f"{w['cpu_sample_desc']}: {int(st.n_links)} links, {int(st.n_trips)} trips, {trav} traversals in {secs:.1f} s of "
              f"Network::step, single thread (mezzo_lib is single-threaded); host has {os.cpu_count()} cpus")
    out = {"value": trav / secs, "unit": "traversals/s", "cores": 1, "kind": kind, "sample": sample}
    if also_gpu:  # the GPU engine on exactly that sample (a small, latency-bound network): the like-for-like ratio
        import mezzo_b200
        sim = mezzo_b200.Simulation(ws["flat"])
        sim.step()
        sim.reset()
        sim.step()
        g = sim.stats()
        sim.close()
        out["gpu_on_same_sample"] = {"value": g["n_traversals"] / (g["kernel_ms"] * 1e-3), "unit": "traversals/s",
                                     "traversals": int(g["n_traversals"]), "traversals_match": int(g["n_traversals"]) == trav}
    return out, secs, trav


def run_sim(args, name):
    import torch
    import mezzo_b200

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        return run_sim_partitioned(args, name, rank, world, local)
    w = build_sim_workload(name)
    flat = w["flat"]
    st_ = flat.struct
    sim = mezzo_b200.Simulation(flat, device=local)
    if os.environ.get("MZ_SIM_KERNEL"):  # tuning aid: 1 = warp per node, 2 = thread per node (default: by network size)
        sim.set_option(4, int(os.environ["MZ_SIM_KERNEL"]))
    stream = torch.cuda.Stream(device=dev)
    sim.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sim.reset()
        sim.step()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    dev_ms, trav, launches, supersteps, events = 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        sim.reset()  # Network::reset — state re-initialisation is not part of Network::step
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record(stream)
            sim.step()
            ev1.record(stream)
        torch.cuda.synchronize()
        dev_ms += ev0.elapsed_time(ev1)
        s = sim.stats()
        trav += s["n_traversals"]
        launches += s["n_launches"]
        supersteps += s["n_supersteps"]
        events += s["n_events"]
    barrier()
    clocks = sampler.stop() if sampler else None
    sim.set_stream(None)
    sim.close()

    # e2e: the host-buffer path a mezzo_lib caller takes — upload the flattened network and trip table, run the
    # horizon, read back arrival times and link-time averages (what Network::writeall needs)
    barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    e2e_trav = 0
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        sim2 = mezzo_b200.Simulation(flat, device=local)
        if os.environ.get("MZ_SIM_KERNEL"):
            sim2.set_option(4, int(os.environ["MZ_SIM_KERNEL"]))
        sim2.step()
        arr = sim2.arrivals()
        lt = sim2.linktimes()
        e2e_trav += sim2.stats()["n_traversals"]
        sim2.close()
        if os.environ.get("MZ_TIMING"):
            print(f"[bench] e2e iteration {1e3 * (time.perf_counter() - t1):.1f} ms", file=sys.stderr)
    e2e_s = time.perf_counter() - t0
    h2d = sum(getattr(flat, f).nbytes for f, _ in type(st_)._fields_ if isinstance(getattr(flat, f, None), np.ndarray))
    d2h = arr.nbytes + lt.nbytes
    # the same with a WARM engine — what a caller running several replications does (Network::reset + step): the network
    # and the trip table stay on the device, every iteration resets the state, runs the horizon and reads the results back
    sim3 = mezzo_b200.Simulation(flat, device=local)
    if os.environ.get("MZ_SIM_KERNEL"):
        sim3.set_option(4, int(os.environ["MZ_SIM_KERNEL"]))
    sim3.step()
    t0 = time.perf_counter()
    warm_trav = 0
    for _ in range(e2e_steps):
        sim3.reset()
        sim3.step()
        sim3.arrivals()
        sim3.linktimes()
        warm_trav += sim3.stats()["n_traversals"]
    warm_s = time.perf_counter() - t0
    sim3.close()

    vals = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    sums = torch.tensor([trav, e2e_trav, launches], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(vals, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(sums, op=torch.distributed.ReduceOp.SUM)
    if rank != 0:
        return None
    dev_ms_max, e2e_s_max = vals.tolist()
    trav_all, e2e_trav_all, launches_all = sums.tolist()
    peaks, peak_src = measured_peaks()
    # algorithmic bytes (SURVEY.md §8d restated for the event-driven engine, which has no per-window sweep):
    # per executed event 128 B (action record 32 B read + 32 B write, the in- and out-link records 2 x 32 B), per traversal
    # another 152 B (queue entry 32 B out + 32 B in, travel-time accumulators 32 B read + 32 B write, exit-log record 16 B,
    # route successor 8 B). One kernel (k_sim_async) runs the whole horizon.
    # (SURVEY.md §8d's literal per-window formula, 96 B per traversal + (64 L + 32 A) per window, gives about half of this
    # for the same run; the per-event restatement is the one used for `achieved`)
    alg_bytes = 128.0 * events + 152.0 * trav
    main_launches = args.steps
    ms_per_launch = dev_ms / max(1, main_launches)
    achieved = alg_bytes / max(1, main_launches) / (ms_per_launch * 1e-3) / 1e9
    line = {
        "metric": "vehicle-link traversals/s", "value": trav_all / (dev_ms_max * 1e-3), "unit": "traversals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True, "scaling": w.get("scaling", "weak"), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": sim_config(name, world),
        "work": {"links": int(st_.n_links), "nodes": int(st_.n_nodes), "turnings": int(st_.n_turnings),
                 "trips": int(st_.n_trips), "traversals_per_step": trav // max(1, args.steps),
                 "events_per_step": events // max(1, args.steps), "node_visits_per_step": supersteps // max(1, args.steps)},
        "e2e": {"value": e2e_trav_all / e2e_s_max, "unit": "traversals/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "call": "mz_sim_create + mz_sim_run + mz_sim_get_arrivals + mz_sim_get_linktimes, host buffers",
                "warm_engine": {"value": warm_trav / warm_s, "unit": "traversals/s", "h2d_bytes_per_step": 0,
                                "d2h_bytes_per_step": int(d2h),
                                "call": "mz_sim_reset + mz_sim_run + mz_sim_get_arrivals + mz_sim_get_linktimes on an engine "
                                        "created once (replications of one scenario); not the headline e2e: no input upload"}},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "peak_source": peak_src, **ncu_traffic(name, True),
                     # mz_sim_run picks the thread-per-node kernel from 600 k nodes per GPU (sim.cu: kThreadPerNodeFrom)
                     "kernel": {"1": "k_sim_async", "2": "k_sim_async_g"}.get(
                         os.environ.get("MZ_SIM_KERNEL", ""), "k_sim_async_g" if int(st_.n_nodes) >= 600000 else "k_sim_async"),
                     "kernel_ms_per_launch": ms_per_launch,
                     "algorithmic_bytes_per_launch": alg_bytes / max(1, main_launches),
                     "note": "latency-bound discrete-event simulation: the run time is the longest chain of dependent "
                             "events (each ~6 dependent 32-byte loads), not a stream of bytes"},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _, _ = sim_cpu_reference(name, also_gpu=True)
    return line


def partitioned_parity_check(rank, world, local):
    """N > 1, untimed, before the timed region: two of the reference's own golden networks (tests/golden/sim_golden.npz —
    exit logs written by the compiled, unmodified reference; deterministic servers, so the bar is bit-exactness) run
    PARTITIONED over all N ranks; rank 0 compares the merged exit log with the golden vectors and with the single-GPU
    engine, word for word. GPUTEST leases one GPU, so this is where the multi-GPU parity shows up in a driver record."""
    import mezzo_b200
    import simcases
    from mezzo_b200 import dist as mzdist, flatten
    from mezzo_b200.sim import CudaPartBackend
    golden = np.load(os.path.join(ROOT, "tests", "golden", "sim_golden.npz"))
    verdict = "ok"
    for case in ("det_ties_small", "det_two_lane_congested"):
        f = flatten.FlatNet(simcases.make(case))
        ps = mzdist.PartitionedSimulation(f, rank, world, CudaPartBackend(device=local))
        ps.step()
        ex, arr = ps.exit_log(), ps.arrivals()
        ps.close()
        if rank == 0:
            lidx = np.array([f.link_index[int(l)] for l in golden[f"{case}_lid"]], np.int32)
            same = (len(ex) == len(lidx) and np.array_equal(ex["vid"], golden[f"{case}_vid"]) and np.array_equal(ex["link"], lidx)
                    and np.array_equal(ex["entry"], golden[f"{case}_entry"]) and np.array_equal(ex["exit"], golden[f"{case}_exit"]))
            one = mezzo_b200.Simulation(f, device=local)
            one.step()
            e1 = one.exit_log()
            same = same and all(np.array_equal(e1[k], ex[k]) for k in ("vid", "link", "entry", "exit")) and \
                np.array_equal(one.arrivals(), arr)
            one.close()
            if not same:
                verdict = f"MISMATCH on {case}"
    return f"{verdict} ({world} ranks, golden det_ties_small + det_two_lane_congested: exit logs == reference vectors == 1 GPU)"


def run_sim_partitioned(args, name, rank, world, local):
    """N > 1: ONE network partitioned over the ranks (graph-growing bisection), every rank runs the persistent node-visit
    kernel on its nodes, neighbouring nodes on different GPUs synchronise pairwise through NVLink peer memory."""
    import torch
    from mezzo_b200 import dist as mzdist, partition
    from mezzo_b200.sim import CudaPartBackend
    dev = torch.device("cuda", local)
    w = build_sim_workload(name, world=world)
    flat = w["flat"]
    st_ = flat.struct
    parity = partitioned_parity_check(rank, world, local)  # untimed: N ranks vs the reference's golden vectors and vs one GPU
    ps = mzdist.PartitionedSimulation(flat, rank, world, CudaPartBackend(device=local))
    cut = int(partition.cut_links(ps.node_rank, flat.link_in_node, flat.link_out_node).sum())
    for _ in range(args.warmup):
        ps.reset()
        ps.step()
    sampler = ClockSampler(local) if rank == 0 else None
    torch.distributed.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    dev_ms, trav, events, visits = 0.0, 0, 0, 0
    for _ in range(args.steps):
        ps.reset()
        ps.step()
        s = ps.stats()  # kernel_ms = max over ranks of the device time of the rank's kernel (CUDA events on its stream)
        dev_ms += s["kernel_ms"]
        trav += s["n_traversals"]
        events += s["n_events"]
        visits += s["n_supersteps"]
    torch.distributed.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    torch.distributed.barrier()
    t0 = time.perf_counter()
    warm_trav = 0
    warm_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(warm_steps):
        ps.reset()
        ps.step()
        ps.arrivals()
        ps.linktimes()
        warm_trav += ps.stats()["n_traversals"]
    torch.distributed.barrier()
    warm_s = time.perf_counter() - t0
    ps.close()
    # e2e: create (upload) + run + merged read-out through the public multi-rank API, host buffers
    torch.distributed.barrier()
    t0 = time.perf_counter()
    e2e_trav = 0
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(e2e_steps):
        ps2 = mzdist.PartitionedSimulation(flat, rank, world, CudaPartBackend(device=local), node_rank=ps.node_rank)
        ps2.step()
        arr = ps2.arrivals()
        lt = ps2.linktimes()
        e2e_trav += ps2.stats()["n_traversals"]
        ps2.close()
    torch.distributed.barrier()
    e2e_s = time.perf_counter() - t0
    vals = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(vals, op=torch.distributed.ReduceOp.MAX)
    if rank != 0:
        return None
    e2e_s_max = vals.tolist()[0]
    h2d = sum(getattr(flat, f).nbytes for f, _ in type(st_)._fields_ if isinstance(getattr(flat, f, None), np.ndarray))
    peaks, peak_src = measured_peaks()
    alg_bytes = 128.0 * events + 152.0 * trav
    ms_per_launch = dev_ms / max(1, args.steps)
    achieved = alg_bytes / max(1, args.steps) / (ms_per_launch * 1e-3) / 1e9 / world  # per GPU
    return {
        "metric": "vehicle-link traversals/s", "value": trav / (dev_ms * 1e-3), "unit": "traversals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": w.get("scaling", "weak"), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": sim_config(name, world),
        "work": {"links": int(st_.n_links), "nodes": int(st_.n_nodes), "turnings": int(st_.n_turnings),
                 "trips": int(st_.n_trips), "traversals_per_step": trav // max(1, args.steps),
                 "events_per_step": events // max(1, args.steps), "node_visits_per_step": visits // max(1, args.steps),
                 "cut_links": cut, "parity_check": parity},
        "e2e": {"value": e2e_trav / e2e_s_max, "unit": "traversals/s", "h2d_bytes_per_step": int(h2d) * world,
                "d2h_bytes_per_step": int(arr.nbytes + lt.nbytes) * world, "steps": e2e_steps,
                "call": "PartitionedSimulation: mz_sim_create_part + attach peers + mz_sim_run + merged arrivals/link "
                        "times, host buffers (every rank uploads the replicated static network)",
                "warm_engine": {"value": warm_trav / warm_s, "unit": "traversals/s", "h2d_bytes_per_step": 0,
                                "d2h_bytes_per_step": int(arr.nbytes + lt.nbytes) * world,
                                "call": "reset + step + merged arrivals / link times on engines created once (replications of "
                                        "one scenario); not the headline e2e: no input upload"}},
        "gpu_launches": 2 * args.steps * world,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                     "kernel": "k_sim_async_g" if int(st_.n_nodes) // world >= 600000 else "k_sim_async",
                     "kernel_ms_per_launch": ms_per_launch,
                     "algorithmic_bytes_per_launch": alg_bytes / max(1, args.steps) / world,
                     "note": "per GPU; latency-bound discrete-event simulation (see DESIGN.md)"},
    }


def run_sim_reference(args, name):
    rank, world, _ = dist_env()
    if rank != 0:
        return None
    tot_t, tot_s, cb = 0, 0.0, None
    for _ in range(max(1, args.steps)):
        cb, secs, trav = sim_cpu_reference(name)
        tot_t += trav
        tot_s += secs
    w = SIM_WORKLOADS[name]
    val = tot_t / tot_s
    cb = dict(cb, value=val)
    return {"impl": "reference", "metric": "vehicle-link traversals/s", "value": val, "unit": "traversals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, args.steps),
            "higher_is_better": True, "scaling": w.get("scaling", "weak"), "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": sim_config(name, world),
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "traversals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MZ_BENCH_WORKLOAD", "default"))
    ap.add_argument("--sssp-mode", default="auto", choices=["exact", "auto"])
    ap.add_argument("--trees-per-step", type=int, default=0)
    ap.add_argument("--cpu-trees", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5, help="iterations of the host-buffer (e2e) arm, at most --steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank, world, local = dist_env()
    name = args.workload
    # default = the north star's target: sim_1m as the headline metric + BASELINE.json's second metric (config 4) in "sssp"
    combined = name == "default"
    if combined:
        name = "sim_1m"

    def sub(line):  # the second metric as a sub-object of the same JSON line
        keys = ("metric", "value", "unit", "ms_per_step", "scaling", "dtype", "config", "work", "e2e", "gpu_launches", "roofline",
                "cpu_baseline")
        return {k: line[k] for k in keys if k in line}

    if args.impl == "reference":
        line = run_sssp_reference(args, name) if name.startswith("sssp") else run_sim_reference(args, name)
        if combined and rank == 0:
            line["sssp"] = sub(run_sssp_reference(args, "sssp_cfg4"))
        if rank == 0:
            print(json.dumps(line), flush=True)
        return

    import torch
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        line = run_sssp(args, name) if name.startswith("sssp") else run_sim(args, name)
        if combined:
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            l2 = run_sssp(args, "sssp_cfg4")
            if rank == 0:
                line["sssp"] = sub(l2)
                line["gpu_launches"] = int(line.get("gpu_launches", 0)) + int(l2.get("gpu_launches", 0))
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
