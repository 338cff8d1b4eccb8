"""Tiny invocations of every kernel added in round 1, sized for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
from b200bo import engine, fit, ops  # noqa: E402
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
lat = ops.Lattice(100, 100)
pts = lat.points(dev)


def data(T, n):
    idx = torch.as_tensor(np.stack([rng.choice(10000, n, replace=False) for _ in range(T)]), device=dev)
    X = pts[idx].contiguous()
    y = (40 + 25 * torch.sin(X[..., 0] / 9.0) * torch.cos(X[..., 1] / 13.0)).contiguous()
    return X, y, torch.full((T,), n, dtype=torch.int32, device=dev)


for n in (7, 20, 63, 70, 100):                       # 8x8 and 16x16 owner grids, chunked staging (NB = 6, 7)
    X, y, nn = data(3, n)
    raw = torch.tensor([[0.3, 30.0, 3.0, 2.0]], dtype=torch.float64, device=dev).repeat(3, 1).contiguous()
    ops.mll_value_grad(X, y, nn, raw, 3, n_hint=n)
    f = fit.LBFGSFitter(3, dev, max_iter=6)
    th = torch.zeros((3, 4), dtype=torch.float64, device=dev)
    f.fit_into(X, y, nn, 3, th, n_hint=n)
X, y, nn = data(2, 150)
raw = torch.tensor([[0.3, 30.0, 3.0, 2.0]], dtype=torch.float64, device=dev).repeat(2, 1).contiguous()
fit.mll_value_grad_large(X, y, nn, raw, 2)
fit.fit_large(X, y, nn, 2, max_iter=2)
d = torch.as_tensor(rng.uniform(0, 100, size=(40, 7)), device=dev)
ix = torch.as_tensor(rng.integers(0, 40, size=(33, 40)).astype(np.int32), device=dev)
ops.bootstrap_ci(d, ix)
sets = torch.as_tensor(np.stack([rng.choice(10000, 10, replace=False) for _ in range(300)]).astype(np.int32), device=dev)
ops.dpp_logdet_score(pts, sets, 1e-5)
ops.dpp_logdet_score_lattice(lat, ops.dpp_table(lat, 1e-5, dev), sets)
lo, hi = ops.seeded_randint_u128(np.arange(-50, 50), 2 ** 100 + 7, dev)
ops.unrank_u128(10000, 10, *ops.seeded_randint_u128(np.arange(64), __import__("math").comb(10000, 10), dev))
# K-A / K-B (TMA-staged L^-1 fragments, 128- and 256-thread CTAs), generic K-B, incremental K-B, bookkeeping
obj = (40 + 25 * torch.sin(pts[:, 0] / 9.0) * torch.cos(pts[:, 1] / 13.0)).float().reshape(1, -1).contiguous()
theta = torch.tensor([0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869467], dtype=torch.float64, device=dev)
for n0, iters, kw in ((10, 3, {}), (70, 3, {}), (10, 3, {"lattice": None}), (10, 3, {"posterior": "incremental", "check_every": 0}),
                      (10, 2, {"append": False})):
    T = 3
    idx = torch.as_tensor(np.stack([rng.choice(10000, n0, replace=False) for _ in range(T)]), device=dev)
    sid = torch.zeros(T, dtype=torch.int32, device=dev)
    engine.BatchedBO(pts, obj, sid, pts[idx], obj[0][idx].double(), iters, theta=theta, tol=-1.0, **kw).run()
torch.cuda.synchronize()
print("sanitizer targets done")
