"""K-C probe: fused-fit latency / throughput by observation count, evaluation-count distribution, and a config-3 run.

    python tools/fit_probe.py            (on a GPU box)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
sys.path.insert(0, ROOT)
from b200bo import engine, fit, ops  # noqa: E402
import objectives  # noqa: E402

dev = torch.device("cuda:0")
FAM = [(s, r) for s in (0.2, 0.4, 0.6, 0.8) for r in (0.2, 0.4, 0.6, 0.8)]
gen = objectives.objectives()
S = 64
surf = gen.wildcatwells_batch([s // 16 for s in range(S)], [FAM[s % 16][0] for s in range(S)], [FAM[s % 16][1] for s in range(S)], device=dev)
obj = surf[:, :100, :100].reshape(S, -1).contiguous()
lat = ops.Lattice(100, 100)
pts = lat.points(dev)
rng = np.random.default_rng(0)
out = {"fit": []}


def fit_case(T, n):
    idx = torch.as_tensor(np.stack([rng.choice(10000, n, replace=False) for _ in range(T)]), device=dev)
    sid = torch.arange(T, device=dev) % S
    X = pts[idx].contiguous()
    y = obj[sid[:, None], idx].double().contiguous()
    nn = torch.full((T,), n, dtype=torch.int32, device=dev)
    f = fit.LBFGSFitter(T, dev)
    theta = torch.zeros((T, 4), dtype=torch.float64, device=dev)
    for _ in range(2):
        f.fit_into(X, y, nn, 3, theta, n_hint=n)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f.fit_into(X, y, nn, 3, theta, n_hint=n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ev = f.iters[:, 1].float()
    tot = float(ev.sum())
    # n^3/2 FMA (sweep) + n^2/2 entries x 2 exp (22 slots) x ~14 FMA per evaluation
    fma = tot * (n ** 3 / 2 + n * n / 2 * (2 * 22 + 14))
    rec = {"T": T, "n": n, "ms": ms, "evals_mean": float(ev.mean()), "evals_p99": float(ev.quantile(0.99)), "evals_max": float(ev.max()),
           "evals_per_s": tot / (ms * 1e-3), "us_per_eval_critical_path": ms * 1e3 / float(ev.max()),
           "fp64_tflops_model": 2 * fma / (ms * 1e-3) / 1e12}
    out["fit"].append(rec)
    print(rec, flush=True)


for T in (1000, 8000):
    for n in (10, 20, 35, 60, 63, 64, 90, 110):
        fit_case(T, n)

# config-3 shape: 1000 trials, fit before every acquisition, 100 iterations; tol honoured / disabled
for tol, name in ((0.1, "tol0.1"), (-1.0, "tol_off")):
    T = 1000
    idx = np.stack([rng.choice(10000, 10, replace=False) for _ in range(T)])
    sid = torch.as_tensor((np.arange(T) % S).astype(np.int32), device=dev)
    X0 = pts[torch.as_tensor(idx, device=dev)]
    y0 = obj[sid.long()[:, None], torch.as_tensor(idx, device=dev)].double()
    fitter = fit.LBFGSFitter(T, dev)
    bo = engine.BatchedBO(pts, obj, sid, X0, y0, 100, theta=None, tol=tol, fitter=fitter, lattice=lat)
    bo.run()
    bo.reset(X0, y0)
    torch.cuda.synchronize(); t0 = time.time(); res = bo.run(); torch.cuda.synchronize(); dt = time.time() - t0
    nit = int(res.nit.sum())
    rec = {"T": T, "tol": tol, "seconds": dt, "bo_it_per_s": nit / dt, "mean_nit": nit / T}
    out["config3_" + name] = rec
    print(name, rec, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fit_probe.json"), "w"), indent=1)
