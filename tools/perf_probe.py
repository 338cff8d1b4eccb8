"""Quick kernel timing probe (development aid): K-A / K-B time vs n, and a stepwise BO run."""
import os, sys, time, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
sys.path.insert(0, ROOT)
from b200bo import ops, engine
from oracle import gp as ogp

dev = torch.device("cuda:0")
T = int(os.environ.get("T", 2048))
G = 10000
ax = np.arange(0, 100, 1.0)
grid = torch.as_tensor(np.stack(np.meshgrid(ax, ax), axis=-1).reshape(-1, 2), device=dev)
rng = np.random.default_rng(0)
theta = torch.as_tensor(ogp.raw_to_theta((0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869425163269042)), device=dev).expand(T, 4).contiguous()
obj = torch.as_tensor(rng.uniform(0, 100, size=(1, G)).astype(np.float32), device=dev)

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

rows = []
lat = ops.Lattice.from_points(grid)
ktab = ops.kernel_table(theta[:1].contiguous(), 3, lat)
NS = [int(v) for v in os.environ.get('NS', '10,16,24,32,40,48,64,80,96,110,128').split(',')]
for n in NS:
    n_max = n
    idx = np.stack([rng.choice(G, n, replace=False) for _ in range(T)])
    X = grid[torch.as_tensor(idx, device=dev)].contiguous()
    y = torch.as_tensor(rng.uniform(0, 100, size=(T, n)), device=dev)
    nn = torch.full((T,), n, dtype=torch.int32, device=dev)
    alpha = torch.empty((T, n_max), dtype=torch.float64, device=dev)
    wfrag = torch.empty((T, ops.wfrag_doubles(n_max)), dtype=torch.float64, device=dev)
    status = torch.zeros(T, dtype=torch.int32, device=dev)
    best = y.max(dim=1).values.contiguous()
    ws = ops.grid_acq_workspace(T, G, dev)
    ai = torch.empty(T, dtype=torch.int32, device=dev); av = torch.empty(T, dtype=torch.float64, device=dev)
    ta = timeit(lambda: ops.gp_assemble_chol_into(X, y, nn, theta, 3, alpha, wfrag, status))
    tb = timeit(lambda: ops.grid_posterior_acq_argmax(grid, X, wfrag, alpha, nn, theta, best, 3, status=status, ws=ws, arg_idx=ai, arg_val=av))
    tl = timeit(lambda: ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, X, wfrag, alpha, nn, theta, best, 3, status=status, ws=ws, arg_idx=ai, arg_val=av))
    fb = T * G * (n * n + 17 * n + 30)
    rows.append(dict(n=n, T=T, ka_ms=round(ta,3), kb_ms=round(tb,3), kb_frac=round(fb / tb / 1e9 / 36.6,3), kbl_ms=round(tl,3), kbl_tflops=round(fb / tl / 1e9,2), kbl_frac=round(fb / tl / 1e9 / 36.6,3)))
    print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "perf_probe.json"), "w"))
# stepwise BO run
if os.environ.get('NO_BO'): sys.exit(0)
Tb = int(os.environ.get("TB", 1024)); iters = int(os.environ.get("ITERS", 50))
idx = np.stack([rng.choice(G, 10, replace=False) for _ in range(Tb)])
X0 = grid[torch.as_tensor(idx, device=dev)].contiguous()
y0 = obj[0][torch.as_tensor(idx, device=dev)].double()
bo = engine.BatchedBO(grid, obj, torch.zeros(Tb, dtype=torch.int32, device=dev), X0, y0, iters, theta=theta[0], tol=-1.0)
torch.cuda.synchronize(); t0 = time.time(); res = bo.run(); torch.cuda.synchronize(); dt = time.time() - t0
print("BO run: T=%d iters=%d  %.3f s  %.0f it/s" % (Tb, iters, dt, Tb * iters / dt), "nit", res.nit[:4].tolist(), "status", int(res.status.max()))
