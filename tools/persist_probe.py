"""Timing probe of the persistent fixed-theta loop (csrc/bo_persist.cu) on the bench shape (development aid):
T trials x ITERS iterations on the 100x100 lattice, 256 Wildcat-like surfaces, tol disabled; compares picks with the
full O(n^2) kernel on a subset and prints the acquisition statistics (B200BO_KF_STATS=1)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
sys.path.insert(0, ROOT)
from b200bo import engine, ops  # noqa: E402
from oracle import gp as ogp  # noqa: E402

dev = torch.device("cuda:0")
T = int(os.environ.get("T", 2960))
ITERS = int(os.environ.get("ITERS", 100))
N0 = int(os.environ.get("N0", 10))
S = 64
import objectives  # noqa: E402

gen = objectives.objectives()
rng = np.random.default_rng(0)
surf = gen.wildcatwells_batch(list(range(S)), [0.2] * S, [0.2] * S, device=dev)
lat = ops.Lattice(100, 100)
grid = lat.points(dev)
obj = surf[:, :100, :100].reshape(S, -1).contiguous()
theta = torch.as_tensor(ogp.raw_to_theta((0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869425163269042)), device=dev)
idx = torch.as_tensor(np.stack([rng.choice(10000, N0, replace=False) for _ in range(T)]), device=dev)
sid = torch.as_tensor((np.arange(T) % S).astype(np.int32), device=dev)
X0 = grid[idx].contiguous()
y0 = obj[sid.long()[:, None], idx].double()
out = {}
for mode in os.environ.get("MODES", "persistent").split(","):
    bo = engine.BatchedBO(grid, obj, sid, X0, y0, ITERS, theta=theta, tol=-1.0, lattice=lat, check_every=0, posterior=mode)
    res = bo.run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e30
    for rep in range(int(os.environ.get("REPS", 2))):
        bo.reset(X0, y0)
        torch.cuda.synchronize()
        e0.record(); res = bo.run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    its = int(res.nit.sum())
    row = dict(mode=mode, T=T, iters=ITERS, ms=round(best, 3), bo_it_per_s=round(its / best * 1e3), status=int(res.status.max()))
    if mode == "persistent":
        lookups = T * 10000 * sum(range(1, N0 + ITERS))            # G x (m + 1) for m = 0 .. n0 + iters - 2
        row["lookups_per_s"] = lookups / best * 1e3
        row["smem_frac"] = row["lookups_per_s"] / (16 * 148 * 1.965e9)
        if os.environ.get("B200BO_KF_STATS"):
            st = bo.ws_loop[64:64 + 128].view(torch.int64).cpu().numpy()
            row["acquisitions"] = int(st[0]); row["exact_evals_per_acq"] = float(st[1]) / max(1, st[0])
            row["batches_per_acq"] = float(st[2]) / max(1, st[0])
            names = ["append", "main_loop", "state_seed_select", "seed_eval_sync", "prune", "survivors_reduce", "bookkeeping"]
            tot = float(sum(st[3:10]))
            row["phase_cycles_per_acq"] = {n: round(float(st[3 + i]) / max(1, st[0])) for i, n in enumerate(names)}
            row["phase_share"] = {n: round(float(st[3 + i]) / max(1.0, tot), 3) for i, n in enumerate(names)}
        out["picked_persistent"] = res.picked[:256].clone()
    if mode == "full":
        out["picked_full"] = res.picked[:256].clone()
    print(json.dumps(row), flush=True)
    del bo
if "picked_full" in out and "picked_persistent" in out:
    same = (out["picked_full"] == out["picked_persistent"]).all(dim=1).float().mean().item()
    print(json.dumps({"identical_pick_sequences_of_first_256": same}))
