#!/usr/bin/env python
"""Benchmark of the B200 BO engine — prints ONE JSON line (see the contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (N > 1: launched by torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      CPU arm: the oracle port on the host cores

Workload (BASELINE.json configs[4], the configuration the metric is quoted on; it fits one GPU): 100 000 independent BO
trials on synthetic Wildcat Wells instances, sharded over the N GPUs (strong scaling: the same trials at every N),
10 random initial lattice points each, fixed GP hyper-parameters (the authors' fitted values), Matérn-5/2, EI over the
100 x 100 candidate lattice, 100 iterations, tol disabled so every trial runs all iterations, one final all-gather of
the gap curves.  One step = the whole job once.
metric = BO iterations per second (trials x iterations / time), whole job over all GPUs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

AUTHOR_RAW = (0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869425163269042)
FAMILIES = [(s, r) for s in (0.2, 0.4, 0.6, 0.8) for r in (0.2, 0.4, 0.6, 0.8)]


def softplus(v):
    return v if v > 20.0 else float(np.log1p(np.exp(v)))


THETA = (AUTHOR_RAW[0], AUTHOR_RAW[1], softplus(AUTHOR_RAW[2]), softplus(AUTHOR_RAW[3]))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trials", type=int, default=100000,
                    help="trials of the WHOLE job, sharded over the GPUs (BASELINE configs[4]: strong scaling)")
    ap.add_argument("--trials-per-gpu", type=int, default=0,
                    help="> 0: that many trials on every GPU instead (weak scaling)")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--n0", type=int, default=10)
    ap.add_argument("--surfaces", type=int, default=256)
    ap.add_argument("--cpu-trials-per-core", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dpp", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configs[2]/configs[3] figures")
    ap.add_argument("--posterior", default="auto", choices=["auto", "persistent", "full", "incremental"],
                    help="auto / persistent: the whole loop of a trial in one CTA, O(n) posterior update per candidate, kernel "
                         "table in shared memory (csrc/bo_persist.cu); full: O(n^2) fused recomputation per iteration (the "
                         "north-star kernel K-B); incremental: O(n) update with the (mu, var) state in HBM")
    ap.add_argument("--no-kb-extra", action="store_true", help="skip the K-B (posterior='full') slice beside the headline")
    return ap.parse_args()


def init_sets(T, n0, seed):
    """T random n0-subsets of the 10 000 lattice nodes (distinct inside a set), as flat indices."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 10000, size=(T, n0))
    while True:
        s = np.sort(idx, axis=1)
        bad = np.nonzero((np.diff(s, axis=1) == 0).any(axis=1))[0]
        if len(bad) == 0:
            return idx
        idx[bad] = rng.integers(0, 10000, size=(len(bad), n0))


# ------------------------------------------------------------------------------------------ CPU arm

def _cpu_worker(job):
    seed, sm, ra, n0, iters, trial_seed = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import gp as ogp, wildcat as owc, dpp as odpp
    _, surf = owc.wildcatwells(seed, sm, ra)
    grid = odpp.construct_pool()
    obj = surf[:100, :100].reshape(-1)
    idx = init_sets(1, n0, trial_seed)[0]
    X0 = grid[idx]
    y0 = obj[idx].astype(np.float64)
    t0 = time.perf_counter()
    out = ogp.bo_loop(grid, obj, X0, y0, iters, theta=THETA, kernel_id=3, tol=-1.0)
    return out["nit"], time.perf_counter() - t0


def cpu_sample(cores, trials_per_core, n0, iters):
    """Oracle port of the same BO iteration on the host cores: one trial per process at a time (how the
    reference fans trials out over ipyparallel engines, Scripts/Experiment2.py:110-141)."""
    import multiprocessing as mp
    jobs = [(i % 7, FAMILIES[i % 16][0], FAMILIES[i % 16][1], n0, iters, 1000 + i) for i in range(cores * trials_per_core)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(0, 0.2, 0.2, n0, 2, 1)] * cores)       # import + warm-up, untimed
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    its = sum(r[0] for r in res)
    return its / wall, its, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup):
        cpu_sample(cores, 1, args.n0, max(2, args.iters // 10))
    t_all = 0.0
    for _ in range(max(1, args.steps)):
        v, its, wall = cpu_sample(cores, args.cpu_trials_per_core, args.n0, args.iters)
        vals.append(v)
        t_all += wall
    value = float(np.mean(vals))
    sample = f"{cores * args.cpu_trials_per_core} trials x {args.iters} iterations per step, one trial per core (numpy fp64 oracle port, 1 BLAS thread per process)"
    line = {
        "impl": "reference", "metric": "bo_iterations_per_s", "value": value, "unit": "BO iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong" if total_trials(args, max(1, args.gpus))[1] else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)), "sample_trials_per_step": cores * args.cpu_trials_per_core,
        "cpu_baseline": {"value": value, "unit": "BO iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "BO iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated oracle (oracle/gp.py), not the botorch path: botorch/gpytorch are not installable here",
    }
    print(json.dumps(line), flush=True)


def total_trials(args, world):
    """(trials of the whole job, strong?) — default: BASELINE configs[4] as written, 100 000 trials sharded over the GPUs."""
    if args.trials_per_gpu > 0:
        return args.trials_per_gpu * world, False
    return args.trials, True


def workload_config(args, world):
    total, strong = total_trials(args, world)
    return {"workload": "BASELINE configs[4]: fixed-theta BO trials on 2-D Wildcat Wells, EI over the 100x100 lattice, "
                        "sharded over the GPUs, one final all-gather of the gap curves",
            "trials_total": total, "iterations": args.iters,
            "n_init": args.n0, "grid": "100x100", "kernel": "matern52", "acquisition": "EI", "tol": "disabled",
            "hyperparameters": "fixed (authors' modes, smoothness 0.2 / ruggedness 0.2)", "surfaces": args.surfaces,
            "l2_policy": "inputs larger than L2: every step streams a fresh batch's observations / curves / picks through HBM "
                         "(0.4 GB per 100 000 trials) and rewrites every trial's per-candidate state",
            "parallelism": f"trial-sharded x{world}" + (" (strong: the same trials at every N)" if strong else " (weak)")}


# ------------------------------------------------------------------------------------------ GPU arm

class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def kb_full_roofline(prof, T, G, ms_total, peak, iters, extra=None):
    """Roofline object of the fused O(n^2) grid posterior + EI + arg-max kernel (K-B) from the engine's per-launch events."""
    from b200bo import ops
    kb = [(n, e0.elapsed_time(e1)) for name, n, e0, e1 in prof if name == "grid_acq"]
    ka = [(n, e0.elapsed_time(e1)) for name, n, e0, e1 in prof if name == "gp_chol"]
    flops = sum(T * G * (n * n + 17 * n + 30) for n, _ in kb)
    kb_ms = sum(ms for _, ms in kb)
    ka_ms = sum(ms for _, ms in ka)
    achieved = flops / (kb_ms * 1e-3) / 1e12 if kb_ms > 0 else 0.0
    kb_avg_ms = kb_ms / max(1, len(kb))
    bytes_mean = float(np.mean([T * (8 * ops.wfrag_doubles(n) + 24 * n + 56) for n, _ in kb])) if kb else None
    out = {"bound": "tensor", "bound_detail": "FP64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA), which shares the SM's FP64 "
                                              "pipe with DFMA (profiles/fp64_peaks_r1.json); not HBM (see hbm_view)",
           "kernel": "grid_acq_lattice8_kernel<NI,THREADS> (K-B, byte-coordinate lattice tile)", "achieved": achieved,
           "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
           "hbm_view": {"algorithmic_bytes_per_launch_mean": bytes_mean,
                        "achieved_gbs": bytes_mean / (kb_avg_ms * 1e-3) / 1e9 if kb else None},
           "flops_model": "F_B = G (n^2 + 17 n + 30) per trial-iteration, exp/sqrt/erfc = 1 flop (SURVEY.md 8d)",
           "kernel_share_of_step": kb_ms / ms_total if ms_total else None,
           "gp_chol_share_of_step": ka_ms / ms_total if ms_total else None, "launches_timed": len(kb), "trials": T,
           "frac_by_n": {str(n): round(T * G * (n * n + 17 * n + 30) / (ms * 1e-3) / 1e12 / peak, 3)
                         for n, ms in kb[:iters] if n in (10, 16, 24, 32, 48, 64, 80, 96, 109)}}
    out.update(extra or {})
    return out


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as fh:
            return json.load(fh)
    except Exception:
        return None


def run_b200(args):
    import torch
    from b200bo import dist, engine, ops
    import BO
    import objectives

    rank, world, local = dist.init()
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    iters, n0, S = args.iters, args.n0, args.surfaces
    total, strong = total_trials(args, world)
    lo, hi = dist.shard_range(total, rank, world)
    T = hi - lo
    lat = ops.Lattice(100, 100)
    grid = lat.points(dev)
    G = lat.G

    # ---- problem instances: S synthetic Wildcat Wells surfaces (K-E), objective table at the candidates
    gen = objectives.objectives()
    gen.device = str(dev)
    seeds = [s // 16 for s in range(S)]
    fam = [FAMILIES[s % 16] for s in range(S)]
    make_surfaces = lambda: gen.wildcatwells_batch(seeds, [f[0] for f in fam], [f[1] for f in fam], device=dev)
    surf = make_surfaces()
    obj_dev = surf[:, :100, :100].reshape(S, G).contiguous()
    # the whole job's initial sets (a function of the global trial index: every N runs the same 100 000 trials)
    idx_all = init_sets(total, n0, 12345)
    sid_all = (np.arange(total) % S).astype(np.int32)
    grid_h = grid.cpu().numpy()
    obj_h = obj_dev.cpu().numpy()
    X0_all = torch.from_numpy(grid_h[idx_all]).pin_memory()
    y0_all = torch.from_numpy(obj_h[sid_all[:, None], idx_all].astype(np.float64)).pin_memory()
    sid_allt = torch.from_numpy(sid_all).pin_memory()
    X0_d, y0_d, sid_d = X0_all[lo:hi].to(dev), y0_all[lo:hi].to(dev), sid_allt[lo:hi].to(dev)
    theta = torch.tensor(THETA, dtype=torch.float64, device=dev)

    bo = engine.BatchedBO(grid, obj_dev, sid_d, X0_d, y0_d, iters, theta=theta, kernel_id=ops.MATERN52, acq_id=ops.EI,
                          tol=-1.0, lattice=lat, check_every=0, posterior=args.posterior)
    mode = "persistent" if bo.persistent else ("incremental" if bo.incremental else "full")
    yall_host = torch.empty((total, iters + 2), dtype=torch.float64).pin_memory()
    nit_host = torch.empty(total, dtype=torch.int32).pin_memory()

    def step_resident():
        bo.reset(X0_d, y0_d)
        res = bo.run()
        curves = dist.gather_curves(res.yall, dist.shard_counts(total, world))    # the path's only exchange (N > 1)
        return res, curves

    def step_e2e():
        # what a user of the reference-facing module calls, host buffers in, host curves out: surfaces (K-E), the batched
        # BO entry (engine construction, kernel table, H2D of this rank's initial sets, the loop, the gather), D2H
        table = make_surfaces()[:, :100, :100].reshape(S, G).contiguous()
        res = BO.BOloop_batched(table, sid_allt, X0_all, y0_all, iters, theta=np.asarray(THETA), grid=lat, tol=-1.0,
                                device=str(dev), posterior=args.posterior, devices="ranks")
        yall_host.copy_(res.yall, non_blocking=True)
        nit_host.copy_(res.nit, non_blocking=True)
        torch.cuda.synchronize(dev)
        return res, res.yall

    def timed(fn, k):
        dist.barrier(dev)
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(k):
            last = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        dist.barrier(dev)
        return dist.max_over_ranks(e0.elapsed_time(e1), dev), last

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize(dev)
    # roofline denominators measured now (GPU warm): FP64 FMA issue rate (never below the committed micro-benchmark of
    # this pool's B200, profiles/fp64_peaks_r1.json: DFMA and DMMA share the pipe) and shared-memory read bandwidth
    peak_probe = ops.fp64_fma_peak_tflops(dev)
    pf = load_json("profiles", "fp64_peaks_r1.json") or {}
    peak_file = max(pf.get("dfma_tflops", 0.0), pf.get("dmma_tflops_acc4", 0.0)) or None
    fp64_peak = max(peak_probe, peak_file or 0.0)
    smem_peak = ops.smem_lds_peak_gbs(dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    bo.profile = []
    ms_total, (res, curves) = timed(step_resident, args.steps)
    prof = bo.profile
    bo.profile = None
    clocks = sampler.stop() if rank == 0 else None
    launches = bo.launches * args.steps
    nit_sum = int(res.nit.sum().item())
    status_max = int(res.status.max().item())
    picked_headline = res.picked[:1480].clone()
    for _ in range(2):                                       # allocator / lazy-module warm-up of the host-facing entry
        step_e2e()
    ms_e2e, (res_e2e, _) = timed(step_e2e, args.steps)
    e2e_same = bool(torch.equal(res_e2e.yall[lo:hi], res.yall))

    hbm_peak_gbs = (load_json("MEASURED_PEAKS.json") or {}).get("hbm_gbs")
    if mode == "persistent":
        # dominant kernel = bo_fixed_loop_kernel (K-F). Its hot loop is ONE 8-byte kernel-table look-up out of shared
        # memory per (candidate, observation) pair of every appended row: G * R (R + 1) / 2 look-ups per trial,
        # R = n0 + nit - 1 rows folded. Roofline = shared-memory read bandwidth (128 B / clk / SM), measured above.
        kf_ms = sum(e0.elapsed_time(e1) for name, n, e0, e1 in prof if name == "bo_fixed_loop")
        R = (res.nit.to(torch.int64) + (n0 - 1))
        lookups = int((R * (R + 1) // 2).sum().item()) * G
        nl = max(1, sum(1 for name, *_ in prof if name == "bo_fixed_loop"))
        bytes_per_launch = 8.0 * lookups
        ach = bytes_per_launch * nl / (kf_ms * 1e-3) / 1e9 if kf_ms > 0 else 0.0
        tr = load_json("profiles", "r2_kf_traffic.json") or {}
        # HBM view: per trial the kernel reads 24 n0 + 36 B of inputs, the objective value of each pick, and writes
        # 24 nit + 8 (max_iter + 2) + 4 max_iter + 16 B of results; the per-candidate state stays in L2
        hbm_bytes = T * (24 * n0 + 36 + 4 * iters + 24 * iters + 8 * (iters + 2) + 4 * iters + 16)
        roofline = {"bound": "smem", "bound_detail": "shared-memory read bandwidth (LDS.64, 128 B/clk/SM): the kernel table "
                    "lives in shared memory and the loop is one 8-byte look-up + one DFMA per (candidate, observation) pair; "
                    "neither HBM (hbm_view) nor the FP64 pipe (fp64_view) is near its limit",
                    "kernel": "bo_fixed_loop_kernel<100,10> (K-F, csrc/bo_persist.cu)", "achieved": ach, "peak": smem_peak,
                    "unit": "GB/s", "frac": ach / smem_peak if smem_peak else None,
                    "peak_source": "b200bo_smem_lds_probe measured in this run (conflict-free LDS.64, one 1024-thread CTA per "
                                   "SM); nominal 128 B/clk/SM x 148 SMs x SM clock",
                    "bytes_model": "8 B x G x R (R + 1) / 2 per trial, R = n0 + nit - 1 appended rows (SURVEY.md 8d counts "
                                   "the O(n^2) recomputation; this kernel updates the same posterior in O(n) per candidate)",
                    "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": kf_ms / nl,
                    "traffic": tr.get("dram_bytes_per_trial", None) and tr["dram_bytes_per_trial"] * T,
                    "traffic_source": tr.get("source"),
                    "hbm_view": {"algorithmic_bytes_per_launch": hbm_bytes,
                                 "achieved_gbs": hbm_bytes * nl / (kf_ms * 1e-3) / 1e9 if kf_ms else None,
                                 "peak_gbs_of_measured": hbm_peak_gbs},
                    "fp64_view": {"dfma_per_lookup": 1, "achieved_tflops": 2.0 * lookups * nl / (kf_ms * 1e-3) / 1e12 if kf_ms else None,
                                  "peak_tflops": fp64_peak},
                    "kernel_share_of_step": kf_ms / ms_total if ms_total else None, "launches_timed": nl}
    elif mode == "incremental":
        # dominant kernel = grid_inc_kernel (+ the row append inside the same timed bracket): HBM-side model
        # 32 B per candidate and launch (read + write of the two fp64 state words), first launch writes only
        ki = [(n, e0.elapsed_time(e1)) for name, n, e0, e1 in prof if name == "grid_inc"]
        ki_ms = sum(ms for _, ms in ki)
        hbm_bytes = sum(T * G * (16 if n == 1 else 32) for n, _ in ki)
        hbm_peak = hbm_peak_gbs or 6650.0
        ach = hbm_bytes / (ki_ms * 1e-3) / 1e9 if ki_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": "grid_inc_kernel (incremental K-B) + gp_append_kernel", "achieved": ach,
                    "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                    "bytes_model": "32 B per candidate-iteration (two fp64 state words read + written)",
                    "note": "issue-bound above n ~ 20 (n table look-ups + n FMA per candidate), HBM-bound below",
                    "kernel_share_of_step": ki_ms / ms_total if ms_total else None, "launches_timed": len(ki)}
    else:
        tr = load_json("profiles", "r1_grid_acq_traffic.json") or {}
        roofline = kb_full_roofline(prof, T, G, ms_total, fp64_peak, iters,
                                    {"traffic": tr.get("dram_bytes_per_launch"), "traffic_trials": tr.get("trials"),
                                     "peak_probe_this_run": peak_probe, "peak_microbench_file": peak_file})
    total_its = int(dist.sum_over_ranks(nit_sum, dev))
    per_step_ms = ms_total / args.steps
    value = total_its / (per_step_ms * 1e-3)
    e2e_val = total_its / (ms_e2e / args.steps * 1e-3)
    h2d = (X0_all[lo:hi].numel() * 8 + y0_all[lo:hi].numel() * 8 + sid_allt[lo:hi].numel() * 4) * 1
    d2h = yall_host.numel() * 8 + nit_host.numel() * 4

    line = {"metric": "bo_iterations_per_s", "value": value, "unit": "BO iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world), "roofline": roofline,
            "e2e": {"value": e2e_val, "unit": "BO iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps, "bytes_are": "per rank",
                    "entry": "objectives.wildcatwells_batch -> BO.BOloop_batched(host arrays, devices='ranks') -> curves on the host",
                    "same_curves_as_resident_run": e2e_same},
            "engine": {"posterior": mode, "trials_this_rank": T},
            "gpu_launches": int(launches), "clocks": clocks, "status_bits_seen": status_max,
            "gathered_curves_shape": list(curves.shape)}

    if rank == 0 and world == 1 and mode == "persistent" and not args.no_kb_extra:
        # the north-star kernel itself — the fused O(n^2) grid posterior + EI + arg-max (K-B) — on a slice of the same
        # workload, against the FP64 pipe; and its picks beside the headline path's
        Tk = min(T, 1480)
        bo2 = engine.BatchedBO(grid, obj_dev, sid_d[:Tk], X0_d[:Tk], y0_d[:Tk], iters, theta=theta, kernel_id=ops.MATERN52,
                               acq_id=ops.EI, tol=-1.0, lattice=lat, check_every=0, posterior="full")

        def step_full():
            bo2.reset(X0_d[:Tk], y0_d[:Tk])
            return bo2.run()

        step_full()
        bo2.profile = []
        ms_full, r2 = timed(step_full, 1)
        kb = kb_full_roofline(bo2.profile, Tk, G, ms_full, fp64_peak, iters)
        kb["value"] = int(r2.nit.sum().item()) / (ms_full * 1e-3)
        kb["unit_value"] = "BO iterations/s"
        kb["trials_with_identical_picks_vs_headline"] = float((r2.picked == picked_headline[:Tk]).all(dim=1).float().mean().item())
        line["kb_full"] = kb
        del bo2
    if rank == 0 and world == 1 and not args.no_dpp:
        line["dpp"] = bench_dpp(dev)
    if rank == 0 and world == 1 and not args.no_configs:
        line["other_configs"] = bench_other_configs(dev, surf, S)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, its, wall = cpu_sample(cores, args.cpu_trials_per_core, n0, iters)
        line["cpu_baseline"] = {"value": v, "unit": "BO iterations/s", "cores": cores, "kind": "port",
                                "sample": f"{cores * args.cpu_trials_per_core} trials x {iters} iterations of the same workload "
                                          f"({wall:.1f} s), one trial per core, numpy fp64 oracle port (oracle/gp.py)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


def bench_other_configs(dev, surf, S):
    """Secondary figures (not the headline): BASELINE configs[2] (per-iteration hyper-parameter fit, 1 000 trials) and
    configs[3] (200 x 200 lattice, fixed theta; a 1 184-trial slice of the 10 000), plus the bootstrap of a
    (1000 x 102) curve matrix (results.py:787-794)."""
    import torch
    from b200bo import engine, fit, ops
    out = {}
    rng = np.random.default_rng(99)

    def run(name, lat, T, iters, theta, tol, reps=1, groups=1, posterior="full"):
        G = lat.G
        pts = lat.points(dev)
        if lat.h == 1.0:
            obj = surf[:, :100, :100].reshape(S, G).contiguous()
        else:
            sidx = torch.arange(S, dtype=torch.int32, device=dev).repeat_interleave(G)
            obj = ops.wildcat_eval(surf.reshape(S, 101, 101), pts.repeat(S, 1), sidx).reshape(S, G).float().contiguous()
        idx = torch.as_tensor(np.stack([rng.choice(G, 10, replace=False) for _ in range(T)]), device=dev)
        sid = torch.as_tensor((np.arange(T) % S).astype(np.int32), device=dev)
        X0 = pts[idx]
        y0 = obj[sid.long()[:, None], idx].double()
        # groups > 1: the batch cut into contiguous groups advanced concurrently on separate CUDA streams
        # (engine.run_concurrent; same results trial by trial)
        cuts = [T * g // groups for g in range(groups + 1)]
        bos = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            fitter = None if theta is not None else fit.LBFGSFitter(hi - lo, dev)
            bos.append(engine.BatchedBO(pts, obj, sid[lo:hi], X0[lo:hi], y0[lo:hi], iters, theta=theta, tol=tol,
                                        fitter=fitter, lattice=lat, posterior=posterior))
        streams = [torch.cuda.Stream(dev) for _ in bos]
        go = (lambda: [bos[0].run()]) if groups == 1 else (lambda: engine.run_concurrent(bos, streams))
        go()                                                       # warm-up
        torch.cuda.synchronize(dev)
        account = theta is None and groups == 1                    # K-C roofline: per-launch events + evaluation counts
        if account:
            bos[0].profile = []
            bos[0].fitter.account = True
            bos[0].fitter.flops.zero_(); bos[0].fitter.evals.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for bo, lo, hi in zip(bos, cuts[:-1], cuts[1:]):
                bo.reset(X0[lo:hi], y0[lo:hi])
            res = engine.concat_results(go())
        e1.record()
        torch.cuda.synchronize(dev)
        sec = e0.elapsed_time(e1) * 1e-3 / reps
        nit = int(res.nit.sum().item())
        out[name] = {"trials": T, "grid": f"{lat.Gx}x{lat.Gy}", "iterations_max": iters, "tol": tol, "seconds": sec,
                     "bo_iterations": nit, "bo_iterations_per_s": nit / sec, "status_bits": int(res.status.max().item()),
                     "concurrent_groups": groups}
        if theta is None:
            out[name]["fit_evaluations_mean_last"] = float(torch.cat([b.fitter.iters[:, 1] for b in bos]).float().mean().item())
        if account:
            fit_ms = sum(a.elapsed_time(b) for nm, _, a, b in bos[0].profile if nm == "fit")
            flops, evals = float(bos[0].fitter.flops.item()), float(bos[0].fitter.evals.item())
            peak = 37.055
            out[name]["roofline_kc"] = {"bound": "tensor", "bound_detail": "FP64 pipe (DFMA; profiles/fp64_peaks_r1.json)",
                                        "kernel": "gp_fit_sweep_kernel<NB,OW> (K-C, fused L-BFGS fit)",
                                        "achieved": flops / (fit_ms * 1e-3) / 1e12 if fit_ms else None, "peak": peak,
                                        "unit": "TFLOP/s", "frac": flops / (fit_ms * 1e-3) / 1e12 / peak if fit_ms else None,
                                        "flops_model": "n^3 + 6 n^2 per marginal-likelihood evaluation (SURVEY.md 8d) x evaluations run",
                                        "evaluations": evals, "fit_ms_total": fit_ms, "fit_share_of_run": fit_ms * 1e-3 / (sec * reps)}
            bos[0].profile = None
            bos[0].fitter.account = False

    theta = torch.tensor(THETA, dtype=torch.float64, device=dev)
    run("configs[2]_fit_every_iteration_tol_disabled", ops.Lattice(100, 100), 1000, 100, None, -1.0)
    run("configs[2]_fit_every_iteration_tol_0.1", ops.Lattice(100, 100), 1000, 100, None, 0.1)
    run("configs[2]_fit_every_iteration_tol_disabled_3_streams", ops.Lattice(100, 100), 1000, 100, None, -1.0, groups=3)
    run("configs[2]_fit_every_iteration_tol_0.1_3_streams", ops.Lattice(100, 100), 1000, 100, None, 0.1, groups=3)
    run("configs[3]_grid200_fixed_theta_slice", ops.Lattice(200, 200, 0.0, 0.0, 0.5), 1184, 100, theta, -1.0)
    run("configs[3]_grid200_fixed_theta_slice_incremental", ops.Lattice(200, 200, 0.0, 0.0, 0.5), 1184, 100, theta, -1.0,
        posterior="incremental")
    # configs[0]: ONE BO run through the reference-facing API (optimizer.optimize -> BOloop_fixparam), 10 initial
    # points, fixed theta, EI over the 100 x 100 lattice, 50 iterations — latency of a single trial, wall clock
    import objectives as objmod
    from Optimizers import optimizer
    synth = objmod.objectives()
    synth.device = str(dev)
    synth.bounds = [(0, 100), (0, 100)]
    synth.seed = 0
    synth.args = {"N": 1, "Smoothness": 0.2, "rug_freq": 1, "rug_amp": 0.2}
    t0 = time.perf_counter()
    f = synth.generate_cont()
    t_surface = time.perf_counter() - t0
    opt = optimizer()
    opt.max_iter = 50
    opt.opt = "BO-fixparam"
    opt.options["tol"] = -1.0
    opt.options["device"] = str(dev)
    opt.bounds = synth.bounds
    names = ("likelihood.noise_covar.raw_noise", "mean_module.constant", "covar_module.raw_outputscale",
             "covar_module.base_kernel.raw_lengthscale")
    opt.paramset = {k: [v] for k, v in zip(names, AUTHOR_RAW)}
    pick = np.random.default_rng(0).choice(10000, 10, replace=False)
    opt.train = np.stack([pick % 100, pick // 100], axis=1).astype(np.float64)
    opt.optimize(f)                                                # warm-up
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        r1 = opt.optimize(f)
    torch.cuda.synchronize(dev)
    t_run = (time.perf_counter() - t0) / 3
    out["configs[0]_single_run_through_optimizer_api"] = {"iterations": int(r1.nit), "seconds_per_run": t_run,
                                                           "bo_iterations_per_s": int(r1.nit) / t_run,
                                                           "wildcat_surface_and_interpolant_seconds": t_surface,
                                                           "note": "wall clock incl. Python, objective table, result packing; "
                                                                   "one trial cannot fill the GPU - latency figure"}
    # the same single trial at engine level: eager launches vs the whole run replayed as one CUDA graph
    lat1 = ops.Lattice(100, 100)
    pts1 = lat1.points(dev)
    obj1 = surf[:1, :100, :100].reshape(1, -1).contiguous()
    X1 = pts1[torch.as_tensor(pick, device=dev)].unsqueeze(0)
    y1 = obj1[0][torch.as_tensor(pick, device=dev)].double().unsqueeze(0)
    bo1 = engine.BatchedBO(pts1, obj1, torch.zeros(1, dtype=torch.int32, device=dev), X1, y1, 50, theta=theta, tol=-1.0,
                           lattice=lat1, check_every=0)
    lat_ms = {}
    for mode in ("eager", "graph"):
        runner = bo1.run if mode == "eager" else bo1.run_graphed
        for _ in range(2):
            bo1.reset(X1, y1); runner()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(10):
            bo1.reset(X1, y1); runner()
        torch.cuda.synchronize(dev)
        lat_ms[mode] = (time.perf_counter() - t0) / 10 * 1e3
    out["configs[0]_single_trial_engine_latency_ms"] = {"eager": lat_ms["eager"], "cuda_graph_replay": lat_ms["graph"],
                                                        "iterations": 50, "note": "BatchedBO.run vs run_graphed, wall clock per run"}
    # bootstrap of a (1000 trials x 102) gap-curve matrix, 1000 replicates; the oracle restatement timed on a slice
    import random
    import utils
    from oracle import bootstrap as ob
    data = np.maximum.accumulate(rng.uniform(0, 100, size=(1000, 102)), axis=1)[:, ::-1].copy()
    idx = torch.as_tensor(rng.integers(0, 1000, size=(1000, 1000)).astype(np.int32), device=dev)
    dd = torch.as_tensor(data, device=dev)
    for _ in range(2):
        ops.bootstrap_ci(dd, idx)
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.bootstrap_ci(dd, idx)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 5
    t0 = time.perf_counter()
    random.seed(0)
    utils.bootstrap(data, trials=1000, device=str(dev))
    t_api = time.perf_counter() - t0
    t0 = time.perf_counter()
    random.seed(0)
    ob.bootstrap(data[:, :4], trials=50)
    t_cpu = (time.perf_counter() - t0) * (102 / 4) * (1000 / 50)
    out["bootstrap_1000x102_1000_replicates"] = {"kernels_ms": ms, "gathered_GB_per_s": 1000 * 1000 * 102 * 8 / (ms * 1e-3) / 1e9,
                                                 "api_seconds_incl_host_index_draws": t_api,
                                                 "cpu_port_seconds_extrapolated": t_cpu,
                                                 "cpu_sample": "oracle restatement of utils.bootstrap on 4 of 102 columns x 50 of "
                                                               "1000 replicates, scaled linearly, 1 core"}
    return out


def bench_dpp(dev):
    """Secondary metric of BASELINE.json: ranked-DPP sets scored per second (k = 10, gamma = 1e-5)."""
    import torch
    from b200bo import ops
    lat = ops.Lattice(100, 100)
    pool = lat.points(dev)
    idx = torch.as_tensor(init_sets(100_000, 10, 7).astype(np.int32), device=dev)
    big = idx.repeat(100, 1).contiguous()                                 # 1e7 sets for a throughput figure
    out = {}
    for name, t in (("n1e5", idx), ("n1e7", big)):
        for _ in range(3):
            ops.dpp_logdet_score(pool, t, 1e-5)
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            ops.dpp_logdet_score(pool, t, 1e-5)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"sets": int(t.shape[0]), "ms": ms, "sets_per_s": t.shape[0] / (ms * 1e-3)}
    # lattice-table variant (what data_gen.generate() uses on the constructX pool): bit-identical scores
    table = ops.dpp_table(lat, 1e-5, dev)
    for name, t in (("lattice_n1e5", idx), ("lattice_n1e7", big)):
        for _ in range(3):
            ops.dpp_logdet_score_lattice(lat, table, t)
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.dpp_logdet_score_lattice(lat, table, t)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 5
        out[name] = {"sets": int(t.shape[0]), "ms": ms, "sets_per_s": t.shape[0] / (ms * 1e-3)}
    out["lattice_identical_to_generic"] = bool(torch.equal(ops.dpp_logdet_score_lattice(lat, table, idx),
                                                           ops.dpp_logdet_score(pool, idx, 1e-5)))
    # end to end at config-2 size: score + mean/std/z-score + device radix sort + gather of the sorted sets (csrc/dpp_rank.cu)
    def ev(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def rank_dev():
        ld = ops.dpp_logdet_score_lattice(lat, table, idx)
        stats, z = ops.dpp_zscore(ld)
        return ops.dpp_rank(z, idx, pool)

    def rank_torch():
        ld = ops.dpp_logdet_score_lattice(lat, table, idx)
        stats, z = ops.dpp_zscore(ld)
        zs, order = torch.sort(z, stable=True)
        return zs, order, idx[order], pool[idx.long()[order]]

    ms = ev(rank_dev)
    ms_t = ev(rank_torch)
    a, b = rank_dev(), rank_torch()
    out["rank_n1e5"] = {"sets": 100000, "ms": ms, "sets_per_s": 1e5 / (ms * 1e-3),
                        "what": "score (lattice table) + z-score + hand-written stable radix sort + gather of the sorted sets",
                        "ms_with_torch_sort_and_indexing": ms_t,
                        "same_order_as_torch_stable_sort": bool(torch.equal(a[1].long(), b[1]))}
    z1 = ops.dpp_zscore(ops.dpp_logdet_score_lattice(lat, table, idx))[1]
    zbig = ops.dpp_zscore(ops.dpp_logdet_score_lattice(lat, table, big))[1]
    out["sort_only"] = {"n1e5_ms": ev(lambda: ops.dpp_rank(z1)), "n1e5_torch_ms": ev(lambda: torch.sort(z1, stable=True)),
                        "n1e7_ms": ev(lambda: ops.dpp_rank(zbig), reps=3, warm=1),
                        "n1e7_torch_ms": ev(lambda: torch.sort(zbig, stable=True), reps=3, warm=1),
                        "what": "b200bo_dpp_rank on the z-scores alone (8 x 8-bit stable LSD passes) beside torch.sort(stable=True)"}
    # end to end with HOST buffers: pinned index sets in, sorted z-scores + permutation out (config 2 through the C ABI)
    idx_h = idx.cpu().pin_memory()
    zs_h = torch.empty(100000, dtype=torch.float64).pin_memory()
    ord_h = torch.empty(100000, dtype=torch.int32).pin_memory()

    def e2e_host():
        d = idx_h.to(dev, non_blocking=True)
        ld = ops.dpp_logdet_score_lattice(lat, table, d)
        stats, z = ops.dpp_zscore(ld)
        zs, order, _, _ = ops.dpp_rank(z)
        zs_h.copy_(zs, non_blocking=True); ord_h.copy_(order, non_blocking=True)
        torch.cuda.synchronize(dev)

    ms = ev(e2e_host)
    out["e2e_host_n1e5"] = {"sets": 100000, "ms": ms, "sets_per_s": 1e5 / (ms * 1e-3), "h2d_bytes": 4000000, "d2h_bytes": 1200000,
                            "what": "pinned (N,10) int32 index sets -> device -> score + z-score + rank -> sorted scores and "
                                    "permutation back on the host"}
    # roofline of the scorer by SURVEY.md 8(d): 623 flops per set (k = 10; exp / log / sqrt / div = 1 flop) against the
    # FP64 pipe; the slot view counts what the pipe really issues (exp 22, sqrt 15, div 14, log 51 DFMA-slots:
    # profiles/fp64_peaks_r1.json) — with ONE log per set (product of the pivots): 2586 slots per set for the generic
    # kernel and 411 for the lattice-table variant
    peak = 37.055
    out["roofline"] = {"bound": "fp64 pipe", "flops_per_set": 623, "peak_tflops": peak,
                       "generic": {"achieved_tflops": out["n1e7"]["sets_per_s"] * 623 / 1e12,
                                   "frac": out["n1e7"]["sets_per_s"] * 623 / 1e12 / peak,
                                   "pipe_slot_frac": out["n1e7"]["sets_per_s"] * 2586 * 2 / 1e12 / peak},
                       "lattice": {"achieved_tflops": out["lattice_n1e7"]["sets_per_s"] * 623 / 1e12,
                                   "frac": out["lattice_n1e7"]["sets_per_s"] * 623 / 1e12 / peak,
                                   "pipe_slot_frac": out["lattice_n1e7"]["sets_per_s"] * 411 * 2 / 1e12 / peak},
                       "hbm_bytes_per_set": 48, "hbm_gbs_lattice": out["lattice_n1e7"]["sets_per_s"] * 48 / 1e9}
    # the reference's generate() path (un-rank + score + rank) at its own size, N = 10 000 (3.22 s in SURVEY.md §6)
    import data_gen
    d = data_gen.diverse_data_generator()
    d.device = str(dev)
    d.options["bounds"] = [(0, 100), (0, 100)]
    d.options["N_samples"] = 10000
    d.gamma = 1e-5
    d.options["seed"] = 0
    d.training_size = 10
    d.options["time-debug"] = True
    d.generate()
    d.generate()
    tt = d.tracker["time-tracking"]
    out["generate_n1e4"] = {"seconds": tt["Total"], "host_rng_and_unrank_share": tt["sampling"], "scoring_share": tt["scoring"],
                            "note": "per-element seeds from numpy's legacy generator on the host (as data_gen.py:866-867); the "
                                    "seeded Mersenne-Twister randint per element (csrc/mt_rank.cu, bit-exact vs CPython), "
                                    "un-ranking, scoring, statistics and the sort run on the GPU; reference: 3.22 s (SURVEY.md 6)"}
    # gammacalc (data_gen.py:758-842; 15.6 s in the reference): 16 gammas x 200 sets, one multi-gamma launch
    d2 = data_gen.diverse_data_generator()
    d2.device = str(dev)
    d2.options["bounds"] = [(0, 100), (0, 100)]
    d2.options["seed"] = 0
    d2.training_size = 10
    d2.gammacalc(grange=(-6, 1))                                    # warm-up on a short range
    d2.gamma_correlation_dict = {}
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    pair = d2.gammacalc()
    torch.cuda.synchronize(dev)
    out["gammacalc_16x16_n200"] = {"seconds": time.perf_counter() - t0, "pair": [int(v) for v in pair] if pair is not None else None,
                                   "reference_seconds": 15.6, "note": "wall clock incl. the 136 host-side rank regressions"}
    # CPU: the oracle's twoD_score restatement (numpy slogdet), one core, bounded sample
    from oracle import dpp as odpp
    pool_h = pool.cpu().numpy()
    sets = idx[:20000].cpu().numpy()
    t0 = time.perf_counter()
    for s in sets:
        odpp.score_lu(pool_h[s], 1e-5)
    dt = time.perf_counter() - t0
    out["cpu_port"] = {"sets_per_s": len(sets) / dt, "cores": 1, "sample": f"{len(sets)} sets, numpy restatement of twoD_score"}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
