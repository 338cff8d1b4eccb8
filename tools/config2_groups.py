"""configs[2] (1 000 trials, fit before every acquisition, tol disabled) by number of concurrent groups (development aid)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200")); sys.path.insert(0, ROOT)
import BO, objectives
from b200bo import ops
dev = "cuda:0"
gen = objectives.objectives()
S, T = 256, int(os.environ.get("T", 1000))
FAM = [(s, r) for s in (0.2, 0.4, 0.6, 0.8) for r in (0.2, 0.4, 0.6, 0.8)]
surf = gen.wildcatwells_batch([s // 16 for s in range(S)], [FAM[s % 16][0] for s in range(S)], [FAM[s % 16][1] for s in range(S)], device=dev)
obj = surf[:, :100, :100].reshape(S, -1).contiguous()
rng = np.random.default_rng(0)
lat = ops.Lattice(100, 100); pts = lat.points(dev)
idx = torch.as_tensor(np.stack([rng.choice(10000, 10, replace=False) for _ in range(T)]), device=dev)
sid = (torch.arange(T, device=dev) % S).int()
X0 = pts[idx]; y0 = obj[sid.long()[:, None], idx].double()
for g in [int(v) for v in os.environ.get("NGROUPS", "1,3").split(",")]:
    ts = []
    for rep in range(int(os.environ.get("REPS", 4))):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = BO.BOloop_batched(obj, sid, X0, y0, 100, theta=None, tol=float(os.environ.get("TOL", -1.0)), device=dev, groups=g)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    dt = min(ts[1:])
    print(json.dumps({"lpt": os.environ.get("B200BO_FIT_LPT", "1"), "mma": os.environ.get("B200BO_KC_MMA", "1"), "groups": g, "seconds_best": round(dt, 4), "seconds_all": [round(t, 3) for t in ts],
                      "bo_it_per_s": round(int(r.nit.sum()) / dt)}), flush=True)
