"""BASELINE configs[3] and configs[4] at FULL size on one GPU, through posterior="auto" (K-F / incremental) and through
the O(n^2) kernel (posterior="full"), with the pick sequences of the two compared trial by trial.

    python tools/full_configs.py            -> gpurun_out/full_configs.json
configs[3]: 10 000 trials over the 16 Wildcat Wells families, fixed theta, 200 x 200 lattice (step 0.5), 100 iterations.
configs[4]: 100 000 trials x 100 iterations, 100 x 100 lattice, fixed theta (what 8 GPUs would share), on ONE GPU.
"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
sys.path.insert(0, ROOT)
from b200bo import engine, ops  # noqa: E402
import objectives  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
gen = objectives.objectives()
S = 256
surf = gen.wildcatwells_batch([s // 16 for s in range(S)], [bench.FAMILIES[s % 16][0] for s in range(S)],
                              [bench.FAMILIES[s % 16][1] for s in range(S)], device=dev)
theta = torch.tensor(bench.THETA, dtype=torch.float64, device=dev)
out = {}
picks = {}


def run(name, lat, T, iters, posterior="full"):
    G = lat.G
    pts = lat.points(dev)
    if lat.h == 1.0:
        obj = surf[:, :100, :100].reshape(S, G).contiguous()
    else:
        sidx = torch.arange(S, dtype=torch.int32, device=dev).repeat_interleave(G)
        obj = ops.wildcat_eval(surf.reshape(S, 101, 101), pts.repeat(S, 1), sidx).reshape(S, G).float().contiguous()
    rng = np.random.default_rng(1)
    idx = torch.as_tensor(bench.init_sets(T, 10, 5) % G if G == 10000 else np.stack([rng.choice(G, 10, replace=False) for _ in range(T)]), device=dev)
    sid = torch.as_tensor((np.arange(T) % S).astype(np.int32), device=dev)
    X0 = pts[idx]
    y0 = obj[sid.long()[:, None], idx].double()
    bo = engine.BatchedBO(pts, obj, sid, X0, y0, iters, theta=theta, tol=-1.0, lattice=lat, check_every=0, posterior=posterior)
    path = "persistent" if bo.persistent else ("incremental" if bo.incremental else "full")
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); res = bo.run(); e1.record(); torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    nit = int(res.nit.sum())
    flops = sum(T * G * (n * n + 17 * n + 30) for n in range(10, 10 + iters))
    gap = (100 - res.yall[:, -1]).mean().item()
    picks[name] = res.picked.clone()
    out[name] = {"trials": T, "grid": f"{lat.Gx}x{lat.Gy}", "iterations": iters, "posterior": path, "seconds": sec, "bo_iterations_per_s": nit / sec,
                 "F_B_tflops_per_s": flops / sec / 1e12, "mean_final_gap": gap, "status_bits": int(res.status.max()),
                 "mem_allocated_GB": torch.cuda.max_memory_allocated() / 1e9}
    print(name, out[name], flush=True)
    del bo, res
    torch.cuda.empty_cache()


for post in ("auto", "full"):
    run(f"configs[3]_full_10000_trials_200x200_{post}", ops.Lattice(200, 200, 0.0, 0.0, 0.5), 10000, 100, post)
    run(f"configs[4]_full_100000_trials_100x100_one_gpu_{post}", ops.Lattice(100, 100), 100000, 100, post)
for c in ("configs[3]_full_10000_trials_200x200", "configs[4]_full_100000_trials_100x100_one_gpu"):
    same = (picks[c + "_auto"] == picks[c + "_full"]).all(dim=1)
    out[c + "_trials_with_identical_picks_auto_vs_full"] = [int(same.sum()), int(same.numel())]
    print(c, "identical pick sequences:", int(same.sum()), "of", int(same.numel()), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "full_configs.json"), "w"), indent=1)
