"""Latency of ONE marginal-likelihood evaluation per trial (register-tiled sweep vs MMA path), T trials per launch (development aid)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200")); sys.path.insert(0, ROOT)
from b200bo import ops
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
pts = ops.Lattice(100, 100).points(dev)
T = int(os.environ.get("T", 296))
for n in [int(v) for v in os.environ.get("NS", "72,96,110,120").split(",")]:
    idx = torch.as_tensor(np.stack([rng.choice(10000, n, replace=False) for _ in range(T)]), device=dev)
    X = pts[idx].contiguous()
    y = (40 + 25 * torch.sin(X[..., 0] / 9.0) * torch.cos(X[..., 1] / 13.0)).contiguous()
    nn = torch.full((T,), n, dtype=torch.int32, device=dev)
    raw = torch.tensor([[0.05, 40.0, 150.0, 9.0]], dtype=torch.float64, device=dev).repeat(T, 1).contiguous()
    row = {"n": n, "T": T}
    for name, fn in (("sweep", ops.mll_value_grad), ("mma", ops.mll_value_grad_mma)):
        for _ in range(3):
            fn(X, y, nn, raw, 3, n_hint=n)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn(X, y, nn, raw, 3, n_hint=n)
        e1.record(); torch.cuda.synchronize()
        row[name + "_us_per_launch"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    print(json.dumps(row), flush=True)
