import sys, torch, numpy as np
sys.path.insert(0, "jmd-diversity-in-bayesian-optimization_b200")
from b200bo import ops
lat = ops.Lattice(100, 100); pool = lat.points("cuda:0"); rng = np.random.default_rng(7)
idx = torch.as_tensor(rng.integers(0, 10000, size=(100000, 10)).astype(np.int32), device="cuda:0").repeat(100, 1).contiguous()
table = ops.dpp_table(lat, 1e-5, "cuda:0")
for name, fn in (("generic", lambda: ops.dpp_logdet_score(pool, idx, 1e-5)), ("lattice", lambda: ops.dpp_logdet_score_lattice(lat, table, idx))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
    print(name, round(ms, 3), "ms", round(1e7 / ms / 1e6, 2), "G sets/s")
