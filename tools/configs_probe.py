"""Exercise BASELINE configs 3 (per-iteration fit) and 4 (200x200 lattice, fixed theta) at reduced size."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
sys.path.insert(0, ROOT)
from b200bo import ops, engine, fit
import objectives

dev = torch.device("cuda:0")
FAM = [(s, r) for s in (0.2, 0.4, 0.6, 0.8) for r in (0.2, 0.4, 0.6, 0.8)]
gen = objectives.objectives()
S = 64
surf = gen.wildcatwells_batch([s // 16 for s in range(S)], [FAM[s % 16][0] for s in range(S)], [FAM[s % 16][1] for s in range(S)], device=dev)
rng = np.random.default_rng(0)
theta = torch.tensor([0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869467], dtype=torch.float64, device=dev)

def run(name, lat, T, iters, theta, tol):
    G = lat.G
    pts = lat.points(dev)
    obj = ops.wildcat_eval(surf.reshape(S, 101, 101), pts.repeat(S, 1), torch.arange(S, dtype=torch.int32, device=dev).repeat_interleave(G)).reshape(S, G).float().contiguous()
    idx = np.stack([rng.choice(G, 10, replace=False) for _ in range(T)])
    # initial points must be integer-lattice nodes for the DPP pool; on the half-step lattice any node is fine
    sid = torch.as_tensor((np.arange(T) % S).astype(np.int32), device=dev)
    X0 = pts[torch.as_tensor(idx, device=dev)]
    y0 = obj[sid.long()[:, None], torch.as_tensor(idx, device=dev)].double()
    fitter = None if theta is not None else fit.LBFGSFitter(T, dev)
    bo = engine.BatchedBO(pts, obj, sid, X0, y0, iters, theta=theta, tol=tol, fitter=fitter, lattice=lat)
    bo.profile = []
    torch.cuda.synchronize(); t0 = time.time(); res = bo.run(); torch.cuda.synchronize(); dt = time.time() - t0
    nit = res.nit.sum().item()
    kb = sum(e0.elapsed_time(e1) for nm, n, e0, e1 in bo.profile if nm == "grid_acq")
    ka = sum(e0.elapsed_time(e1) for nm, n, e0, e1 in bo.profile if nm == "gp_chol")
    gap = (100 - res.yall[:, -1]).mean().item()
    print(f"{name}: T={T} iters={iters} G={G}: {dt:.2f} s, {nit / dt:.0f} it/s, K-B {kb:.0f} ms, K-A {ka:.0f} ms, mean nit {nit / T:.1f}, "
          f"mean final gap {gap:.3f}, status bits {int(res.status.max())}, launches {bo.launches}", flush=True)
    if fitter is not None:
        print("   fit rounds last", fitter.last_rounds, "mean iters", fitter.iters[:, 0].float().mean().item())
    return res

run("config4-like fixed theta 200x200", ops.Lattice(200, 200, 0.0, 0.0, 0.5), int(os.environ.get("T4", 1184)), 100, theta, -1.0)
run("config3-like per-iteration fit 100x100", ops.Lattice(100, 100), int(os.environ.get("T3", 1000)), int(os.environ.get("I3", 100)), None, 0.1)
