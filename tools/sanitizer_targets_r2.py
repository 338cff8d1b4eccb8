"""Tiny invocations of the kernels added in round 2, sized for compute-sanitizer (memcheck / racecheck / synccheck):
the persistent loop (both variants, several table slots, early stops, the fallback path), the radix sort + gather,
count-less, window pick and the multi-gamma scorer."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
from b200bo import engine, ops  # noqa: E402
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
lat = ops.Lattice(100, 100)
pts = lat.points(dev)
obj = (40 + 25 * torch.sin(pts[:, 0] / 9.0) * torch.cos(pts[:, 1] / 13.0)).float().reshape(1, -1).contiguous()
base = np.array([0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869467])
T, n0, iters = 5, 6, 4
idx = torch.as_tensor(np.stack([rng.choice(10000, n0, replace=False) for _ in range(T)]), device=dev)
sid = torch.zeros(T, dtype=torch.int32, device=dev)
theta = torch.as_tensor(np.stack([base, base * [2, 1, 0.7, 1.5], base, base * [2, 1, 0.7, 1.5], base * [0.5, 1, 1, 2]]), device=dev)
for tol in (-1.0, 60.0):
    r = engine.BatchedBO(pts, obj, sid, pts[idx], obj[0][idx].double(), iters, theta=theta, tol=tol, posterior="persistent").run()
    torch.cuda.synchronize()
    print("persistent", os.environ.get("B200BO_KF_DUAL", "1"), "tol", tol, "nit", r.nit.tolist(), "status", r.status.tolist())
# an odd lattice through the generic template instance
lat2 = ops.Lattice(37, 23, 1.0, 2.0, 0.5)
p2 = lat2.points(dev)
o2 = (p2[:, 0] * 1.5 - p2[:, 1]).float().reshape(1, -1).contiguous()
i2 = torch.as_tensor(np.stack([rng.choice(lat2.G, 5, replace=False) for _ in range(3)]), device=dev)
r = engine.BatchedBO(p2, o2, torch.zeros(3, dtype=torch.int32, device=dev), p2[i2], o2[0][i2].double(), 3,
                     theta=torch.as_tensor(base, device=dev), tol=-1.0, posterior="persistent", acq_id=ops.UCB, acq_param=2.0).run()
torch.cuda.synchronize()
print("odd lattice nit", r.nit.tolist())
z = torch.as_tensor(rng.normal(size=5000).round(1), device=dev)
pos = torch.as_tensor(rng.integers(0, 10000, size=(5000, 10)).astype(np.int32), device=dev)
zs, order, ps, dist = ops.dpp_rank(z, pos, pts)
print("count_less", ops.dpp_count_less(zs, 0.05))
mask = torch.zeros((5000 + 31) // 32, dtype=torch.int32, device=dev)
for r_ in (0, 3, 100, 249):
    ops.dpp_window_pick(mask, 5000, 4750, 5000, r_ % 240, ps, pts)
ops.dpp_logdet_score_multigamma(pts, pos[:300].contiguous(), [1e-6, 1e-5, 1e-3])
torch.cuda.synchronize()
print("sanitizer targets (round 2) done")
