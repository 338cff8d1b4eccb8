"""Fused register-sweep fit (gp_fit_sweep_kernel) against the blocked DMMA fit (gp_fit_large_kernel) at mid n (development aid)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200")); sys.path.insert(0, ROOT)
from b200bo import fit, ops
import objectives
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
lat = ops.Lattice(100, 100); pts = lat.points(dev)
gen = objectives.objectives()
surf = gen.wildcatwells_batch([0, 1, 2, 3], [0.2, 0.4, 0.6, 0.8], [0.2, 0.4, 0.6, 0.8], device=dev)
obj = surf[:, :100, :100].reshape(4, -1).double()
T = int(os.environ.get("T", 296))
for n in [int(v) for v in os.environ.get("NS", "48,64,80,96,110,127").split(",")]:
    idx = torch.as_tensor(np.stack([rng.choice(10000, n, replace=False) for _ in range(T)]), device=dev)
    X = pts[idx].contiguous()
    y = obj[torch.arange(T, device=dev) % 4][torch.arange(T, device=dev)[:, None], idx].contiguous()
    nn = torch.full((T,), n, dtype=torch.int32, device=dev)
    ft = fit.LBFGSFitter(T, dev); ft.lpt = False
    th = torch.zeros((T, 4), dtype=torch.float64, device=dev)
    row = {"n": n, "T": T, "mma": os.environ.get("B200BO_KC_MMA", "1")}
    cands = [("sweep", lambda: ft.fit_into(X, y, nn, 3, th, n_hint=n))]
    if os.environ.get("LARGE"):
        cands.append(("large", lambda: fit.fit_large(X, y, nn, 3)))
    for name, fn in cands:
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        row[name + "_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        if name == "sweep":
            row["evals_mean"] = float(ft.iters[:, 1].float().mean()); row["evals_max"] = int(ft.iters[:, 1].max())
            m1 = ft.mll.clone()
        else:
            row["large_evals_mean"] = float(out["iters"][:, 1].float().mean())
            row["max_rel_mll_diff"] = float(((out["mll"] - m1).abs() / m1.abs()).max())
    print(json.dumps(row), flush=True)
