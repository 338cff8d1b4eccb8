"""Small single-kernel workloads for ncu captures: python tools/ncu_targets.py {dpp|inc|large|boot|wildcat|kf|rank}"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
sys.path.insert(0, ROOT)
from b200bo import engine, fit, ops  # noqa: E402
import objectives  # noqa: E402
what = sys.argv[1]
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
lat = ops.Lattice(100, 100)
pts = lat.points(dev)
if what == "dpp":
    idx = torch.as_tensor(np.stack([rng.choice(10000, 10, replace=False) for _ in range(100000)]).astype(np.int32), device=dev).repeat(100, 1).contiguous()
    for _ in range(2):
        ops.dpp_logdet_score(pts, idx, 1e-5)
elif what == "boot":
    data = torch.as_tensor(rng.uniform(0, 100, size=(1000, 102)), device=dev)
    idx = torch.as_tensor(rng.integers(0, 1000, size=(1000, 1000)).astype(np.int32), device=dev)
    for _ in range(2):
        ops.bootstrap_ci(data, idx)
elif what == "large":
    T, n = 148, 1210
    order = rng.permutation(10000)[:n]
    X = pts[torch.as_tensor(order, device=dev)].unsqueeze(0).repeat(T, 1, 1).contiguous()
    y = (40 + 25 * torch.sin(X[..., 0] / 9.0) * torch.cos(X[..., 1] / 13.0)).contiguous()
    nn = torch.full((T,), n, dtype=torch.int32, device=dev)
    raw = torch.tensor([[0.05, 40.0, 150.0, 9.0]], dtype=torch.float64, device=dev).repeat(T, 1).contiguous()
    for _ in range(2):
        fit.mll_value_grad_large(X, y, nn, raw, 2)
elif what == "wildcat":
    gen = objectives.objectives()
    for _ in range(2):
        gen.wildcatwells_batch(list(range(1024)), [0.4] * 1024, [0.4] * 1024, device=dev)
elif what == "inc":
    gen = objectives.objectives()
    S, T = 64, 2368
    surf = gen.wildcatwells_batch(list(range(S)), [0.4] * S, [0.4] * S, device=dev)
    obj = surf[:, :100, :100].reshape(S, -1).contiguous()
    idx = torch.as_tensor(np.stack([rng.choice(10000, 10, replace=False) for _ in range(T)]), device=dev)
    sid = (torch.arange(T, device=dev) % S).int()
    theta = torch.tensor([0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869467], dtype=torch.float64, device=dev)
    bo = engine.BatchedBO(pts, obj, sid, pts[idx], obj[sid.long()[:, None], idx].double(), 60, theta=theta, tol=-1.0, lattice=lat,
                          check_every=0, posterior="incremental")
    bo.run()
elif what == "kf":
    # the persistent fixed-theta loop on the bench shape: T trials (env T, default 12 500) x 100 iterations, ONE launch
    gen = objectives.objectives()
    S, T = 256, int(os.environ.get("T", 12500))
    fam = [(s, r) for s in (0.2, 0.4, 0.6, 0.8) for r in (0.2, 0.4, 0.6, 0.8)]
    surf = gen.wildcatwells_batch([s // 16 for s in range(S)], [fam[s % 16][0] for s in range(S)], [fam[s % 16][1] for s in range(S)], device=dev)
    obj = surf[:, :100, :100].reshape(S, -1).contiguous()
    idx = torch.as_tensor(np.stack([rng.choice(10000, 10, replace=False) for _ in range(T)]), device=dev)
    sid = (torch.arange(T, device=dev) % S).int()
    theta = torch.tensor([0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869467], dtype=torch.float64, device=dev)
    bo = engine.BatchedBO(pts, obj, sid, pts[idx], obj[sid.long()[:, None], idx].double(), 100, theta=theta, tol=-1.0, lattice=lat,
                          check_every=0, posterior="persistent")
    bo.run()
elif what == "fit":
    # the fused hyper-parameter fit (K-C) on T trials of N observations (env T, N): one launch
    T, n = int(os.environ.get("T", 296)), int(os.environ.get("N", 110))
    gen = objectives.objectives()
    surf = gen.wildcatwells_batch([0, 1, 2, 3], [0.2, 0.4, 0.6, 0.8], [0.2, 0.4, 0.6, 0.8], device=dev)
    obj = surf[:, :100, :100].reshape(4, -1).double()
    idx = torch.as_tensor(np.stack([rng.choice(10000, n, replace=False) for _ in range(T)]), device=dev)
    X = pts[idx].contiguous()
    y = obj[torch.arange(T, device=dev) % 4][torch.arange(T, device=dev)[:, None], idx].contiguous()
    nn = torch.full((T,), n, dtype=torch.int32, device=dev)
    ft = fit.LBFGSFitter(T, dev)
    th = torch.zeros((T, 4), dtype=torch.float64, device=dev)
    for _ in range(2):
        ft.fit_into(X, y, nn, 3, th, n_hint=n)
    torch.cuda.synchronize()
    print("evals mean / max", float(ft.iters[:, 1].float().mean()), int(ft.iters[:, 1].max()))
elif what == "rank":
    z = torch.as_tensor(rng.normal(size=10_000_000), device=dev)
    for _ in range(2):
        ops.dpp_rank(z)
torch.cuda.synchronize()
print("done", what)
