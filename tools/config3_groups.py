"""configs[2] (1 000 trials, fit before every acquisition) as one batch and as 2 / 3 / 4 concurrent groups."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200")); sys.path.insert(0, ROOT)
from b200bo import ops
import BO, objectives
dev = torch.device("cuda:0")
FAM = [(s, r) for s in (0.2, 0.4, 0.6, 0.8) for r in (0.2, 0.4, 0.6, 0.8)]
gen = objectives.objectives(); S = 64
surf = gen.wildcatwells_batch([s // 16 for s in range(S)], [FAM[s % 16][0] for s in range(S)], [FAM[s % 16][1] for s in range(S)], device=dev)
lat = ops.Lattice(100, 100); pts = lat.points(dev); obj = surf[:, :100, :100].reshape(S, -1).contiguous()
rng = np.random.default_rng(99)
T = int(os.environ.get("T", 1000))
idx = torch.as_tensor(np.stack([rng.choice(10000, 10, replace=False) for _ in range(T)]), device=dev)
sid = torch.as_tensor((np.arange(T) % S).astype(np.int32), device=dev)
X0 = pts[idx]; y0 = obj[sid.long()[:, None], idx].double()
for tol in (0.1, -1.0):
    ref = None
    for groups in [int(g) for g in os.environ.get("GROUPS", "1,2,3,4,6").split(",")]:
        BO.BOloop_batched(obj, sid, X0, y0, 100, theta=None, grid=lat, tol=tol, groups=groups); torch.cuda.synchronize()
        t0 = time.time(); res = BO.BOloop_batched(obj, sid, X0, y0, 100, theta=None, grid=lat, tol=tol, groups=groups)
        torch.cuda.synchronize(); dt = time.time() - t0
        its = int(res.nit.sum())
        same = True if ref is None else bool(torch.equal(ref.picked, res.picked) and torch.equal(ref.yall, res.yall))
        if ref is None: ref = res
        print("tol %5.1f groups %d: %.3f s, %d BO it, %.0f it/s, identical to one batch: %s, mean final gap %.4f" % (tol, groups, dt, its, its / dt, same, float((100.0 - res.yall[:, -1]).mean())), flush=True)

