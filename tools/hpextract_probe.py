"""Experiment3-2-1-shaped workload: 200 independent fits on 1 010 ... 1 209 points (Matern-3/2), one CTA each.

    python tools/hpextract_probe.py [n_fits] [n_first]
"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
sys.path.insert(0, ROOT)
from b200bo import fit, ops  # noqa: E402
import objectives  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_first = int(sys.argv[2]) if len(sys.argv) > 2 else 1010
dev = torch.device("cuda:0")
gen = objectives.objectives()
surf = gen.wildcatwells_batch([0], [0.4], [0.4], device=dev)[0, :100, :100].reshape(-1).double()
rng = np.random.default_rng(0)
order = rng.permutation(10000)                     # init set + optgammarand points: a fixed random order, growing prefix
n = np.arange(n_first, n_first + F).astype(np.int32)
n_max = int(n.max())
pts = ops.Lattice(100, 100).points(dev)
idx = torch.as_tensor(order[:n_max].copy(), device=dev)
X = pts[idx].unsqueeze(0).repeat(F, 1, 1).contiguous()
y = surf[idx].unsqueeze(0).repeat(F, 1).contiguous()
nn = torch.as_tensor(n, device=dev)
for _ in range(1):
    r = fit.fit_large(X[:8], y[:8], nn[:8], 2)
torch.cuda.synchronize()
t0 = time.perf_counter()
r = fit.fit_large(X, y, nn, 2)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ev = r["iters"][:, 1].double()
fma = float((ev * (torch.as_tensor(((n + 63) // 64 * 64).astype(np.float64), device=dev) ** 3) / 2).sum())
rec = {"fits": F, "n_first": n_first, "n_last": int(n.max()), "seconds": dt, "evals_mean": float(ev.mean()), "evals_max": float(ev.max()),
       "fp64_tflops": 2 * fma / dt / 1e12, "seconds_per_evaluation_per_cta": dt / float(ev.max()),
       "theta_median": r["theta"].median(dim=0).values.tolist()}
print(json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "hpextract_probe.json"), "w"), indent=1)
