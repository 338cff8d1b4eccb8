"""Where does the host-facing batched entry spend its time? (development aid) Times the stages of one
BO.BOloop_batched call on the bench workload with wall clock + device synchronisation."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200")); sys.path.insert(0, ROOT)
from b200bo import engine, ops
import BO
dev = torch.device("cuda:0")
T = int(os.environ.get("T", 100000))
def tick(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(json.dumps({label: round((time.perf_counter() - t0) * 1e3, 2)}), flush=True); return r
th = torch.tensor([0.0017, 41.8, 190.4, 9.87], dtype=torch.float64, device=dev)
thT = th.expand(T, 4).contiguous().clone()
for rep in range(2):
    tick("unique_dim0", lambda: torch.unique(thT, dim=0, return_inverse=True))
    tick("zeros_X", lambda: torch.zeros((T, 110, 2), dtype=torch.float64, device=dev))
    lat = ops.Lattice(100, 100)
    X = torch.zeros((T, 110, 2), dtype=torch.float64, device=dev)
    tick("contains", lambda: bool(lat.contains(X[:, :10]).all()))
    tick("from_points", lambda: ops.Lattice.from_points(lat.points(dev)))
import objectives, cProfile, pstats
gen = objectives.objectives(); gen.device = str(dev)
S = 256
FAM = [(s, r) for s in (0.2, 0.4, 0.6, 0.8) for r in (0.2, 0.4, 0.6, 0.8)]
seeds = [s // 16 for s in range(S)]; fam = [FAM[s % 16] for s in range(S)]
mk = lambda: gen.wildcatwells_batch(seeds, [f[0] for f in fam], [f[1] for f in fam], device=dev)
surf = tick("surfaces", mk); surf = tick("surfaces", mk)
obj = surf[:, :100, :100].reshape(S, -1).contiguous()
rng = np.random.default_rng(0)
idx = rng.integers(0, 10000, size=(T, 10))
sid = (np.arange(T) % S).astype(np.int32)
grid_h = lat.points(dev).cpu().numpy(); obj_h = obj.cpu().numpy()
X0 = torch.from_numpy(grid_h[idx]).pin_memory(); y0 = torch.from_numpy(obj_h[sid[:, None], idx].astype(np.float64)).pin_memory()
sidt = torch.from_numpy(sid).pin_memory()
call = lambda: BO.BOloop_batched(obj, sidt, X0, y0, 100, theta=np.array([0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869425163269042]),
                                 grid=lat, tol=-1.0, device=str(dev), devices="ranks")
tick("BOloop_batched", call)
pr = cProfile.Profile(); pr.enable(); tick("BOloop_batched", call); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
