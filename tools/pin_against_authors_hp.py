"""End-to-end anchor of the GP / marginal-likelihood restatement against the AUTHORS' OWN botorch outputs.

Experiment3-2-1 (Scripts/Experiment3-2-1.py:254-330) is fully seeded: the Wildcat Wells function has seed 0, the
training sets come from the ranked-DPP sampler (seed = 2 * dataseed, 95th percentile) plus `optgammarand` points, and
the stored result is the raw parameter vector fit_gpytorch_model reached on each of the 200 growing sets
(tests/golden/authors_hp_extract.json, copied from Data/experiment3-2-1/.../data<seed>.pkl). This script rebuilds
those sets with the engine's bit-exact data_gen / objectives, fits them with csrc/gp_fit_large.cu and prints both.

    python tools/pin_against_authors_hp.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200"))
sys.path.insert(0, ROOT)
import data_gen  # noqa: E402
import objectives  # noqa: E402
from b200bo import hpextract  # noqa: E402

g = json.load(open(os.path.join(ROOT, "tests", "golden", "authors_hp_extract.json")))
out = []
for case in g["cases"]:
    sm, ra, dataseed = case["smoothness"], case["rug_amp"], case["dataseed"]
    f_gen = objectives.objectives()
    f_gen.bounds = [(0, 100), (0, 100)]
    f_gen.seed = 0
    f_gen.args = {"N": 1, "Smoothness": sm, "rug_freq": 1, "rug_amp": ra}
    f = f_gen.generate_cont()
    d = data_gen.diverse_data_generator()
    d.options["bounds"] = f_gen.bounds
    d.options["N_samples"] = 10000
    d.options["seed"] = dataseed * 2
    d.gamma = 1e-5
    d.training_size = 10
    d.generate()
    init = d.sample(95)
    rand = d.optgammarand(init, 1200)
    sets = [np.concatenate([init, rand[:m]]) for m in range(1000, 1200)]
    hps = hpextract.hp_extract_batched(sets, f, nu=1.5)
    mine = np.array([[h.hyperparameters[k][0] for k in hpextract.PARAM_NAMES] for h in hps])
    ref = np.array(case["raw"])
    rel = np.abs(mine - ref) / np.maximum(np.abs(ref), 1e-12)
    med = np.median(rel, axis=0)
    print(f"family {sm}/{ra} dataseed {dataseed}: median relative difference (noise, constant, raw_s, raw_l) = {np.round(med, 5)}")
    for i in (0, 99, 199):
        print("   n =", 1010 + i, "engine", np.round(mine[i], 5), "authors", np.round(ref[i], 5))
    out.append({"smoothness": sm, "rug_amp": ra, "dataseed": dataseed, "median_rel_diff": med.tolist(),
                "frac_within_1pct": (rel < 0.01).mean(axis=0).tolist(), "engine_first": mine[0].tolist(), "authors_first": ref[0].tolist()})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pin_authors_hp.json"), "w"), indent=1)
