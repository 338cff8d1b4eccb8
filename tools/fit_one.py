"""One fused-fit launch for profiling: python tools/fit_one.py T n [reps]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
from b200bo import fit, ops  # noqa: E402
T, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
pts = ops.Lattice(100, 100).points(dev)
idx = torch.as_tensor(np.stack([rng.choice(10000, n, replace=False) for _ in range(T)]), device=dev)
X = pts[idx].contiguous()
y = (40 + 25 * torch.sin(X[..., 0] / 9.0) * torch.cos(X[..., 1] / 13.0) + 3 * torch.randn(T, n, device=dev, dtype=torch.float64)).contiguous()
nn = torch.full((T,), n, dtype=torch.int32, device=dev)
f = fit.LBFGSFitter(T, dev)
theta = torch.zeros((T, 4), dtype=torch.float64, device=dev)
for _ in range(reps):
    f.fit_into(X, y, nn, 3, theta, n_hint=n)
torch.cuda.synchronize()
print("evals mean", f.iters[:, 1].float().mean().item(), "max", f.iters[:, 1].max().item())
