"""CPU: host-side logic of the product package and the C-ABI surface (no GPU needed: symbols only)."""
import math
import os
import re

import numpy as np
import pytest

from conftest import PKG, ROOT


def test_library_loads_and_exports_every_declared_symbol():
    from b200bo import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "b200bo.h")).read()
    declared = set(re.findall(r"\b(b200bo_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200bo.h but not exported"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert "sm_100a" in _lib.version()
    assert lib.b200bo_wfrag_doubles(8) == 2 * 32 and lib.b200bo_wfrag_doubles(110) == 14 * 15 * 32


def test_column_chunks_per_trial_fill_the_gpu_and_the_last_wave():
    """K-B launches T x chunks CTAs (host rule b200bo::pick_chunks, no CUDA call): a single trial is split until the GPU
    is full, a bench-size batch keeps one CTA per trial, a 1 000-trial batch takes a few chunks so that the last wave
    of CTAs is not two-thirds empty; the workspace is sized for the largest choice."""
    from b200bo import _lib
    lib = _lib.load()
    G = 10000
    chunks = lib.b200bo_grid_acq_chunks
    assert chunks(12500, G, 35, 1) == 1 and chunks(12500, G, 100, 1) == 1 and chunks(100000, G, 60, 1) == 1
    assert chunks(1, G, 35, 1) == (G + 127) // 128                     # as many CTAs as 128-column chunks exist
    c = chunks(1000, G, 35, 1)                                          # 740 resident CTAs: 1 000 CTAs = 2 waves for 1.35
    assert 2 <= c <= 8
    waves = lambda cc: -(-1000 * cc // 740) / cc
    assert waves(c) < 0.8 * waves(1)
    assert chunks(0, G, 35, 1) == 0 and chunks(5, 64, 10, 0) == 1       # empty batch; fewer candidates than one chunk
    for T in (1, 7, 300, 1000, 5000, 12500):
        for n_cap in (10, 64, 65, 128):
            for lattice in (0, 1):
                cc = chunks(T, G, n_cap, lattice)
                assert 1 <= cc <= (G + 127) // 128
                assert lib.b200bo_grid_acq_workspace_bytes(T, G) >= T * cc * 12


def test_built_for_sm_100a_only():
    import subprocess
    from b200bo import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_ops_refuse_cpu_tensors():
    import torch
    from b200bo import _lib, ops
    with pytest.raises(_lib.B200boError):
        ops.dpp_logdet_score(torch.zeros((4, 2), dtype=torch.float64), torch.zeros((1, 2), dtype=torch.int32), 0.1)
    with pytest.raises(_lib.B200boError):
        ops.gp_assemble_chol(torch.zeros((1, 4, 2), dtype=torch.float64), torch.zeros((1, 4), dtype=torch.float64),
                             torch.ones(1, dtype=torch.int32), torch.ones((1, 4), dtype=torch.float64), 3)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), os.path.join(dirpath, f)


def test_host_hash_and_simplex_setup(golden):
    from b200bo import host
    g = golden("simplex_noise.npz")
    assert [host.rand_hash(int(v)) for v in g["hash_in"]] == [int(v) for v in g["hash_out"]]
    for s in g["seeds"]:
        h, vecs = host.simplex_setup(int(s))
        assert h == int(g[f"hseed_{s}"][0]) and np.array_equal(np.array(vecs, dtype=np.int8), g[f"vecs_{s}"])
    assert host.wildcat_peak(1) == (41.7022004702574, 32.664490177209615)


def test_grids_match_reference(golden):
    import data_gen
    import objectives
    o = objectives.objectives()
    o.bounds = [(0, 100), (0, 100)]
    assert np.array_equal(o.construct_X(), golden("wildcat.npz")["construct_X"])
    X, NC = o.construct_X(with_noise_coords=True)
    assert NC.shape == (101, 101, 2) and NC.dtype == np.float64 and NC[0, 1, 0] == 100 / 101
    assert o.dim == 2
    d = data_gen.diverse_data_generator()
    with pytest.raises(Exception):
        d.constructX()
    d.options["bounds"] = [(0, 100), (0, 100)]
    d.constructX()
    assert np.array_equal(d.X, golden("dpp.npz")["pool"])
    d.constructX(budget=40000)
    assert d.X.shape == (40000, 2) and np.array_equal(d.X[:5], golden("dpp.npz")["pool40000_head"])


def test_element_ranks_and_host_unrank(golden):
    from b200bo import host
    from Combination import combination_finder
    g = golden("dpp.npz")
    assert [str(e) for e in host.iid_comb_locs(10000, 10, 2000, 2)] == list(g["gen_s2_elements"])
    cf = combination_finder()
    for row in golden("combinadics.json"):
        assert cf.find_elm(row["n"], row["k"], int(row["m"])) == row["set"]
    assert cf.LargestV(8, 4, 7) == 5 and cf.reduced_comb(5, 2) == 10
    with pytest.raises(Exception):
        cf.find_elm(3, 5, 0)
    with pytest.raises(Exception):
        cf.find_elm(10000, 10, math.comb(10000, 10))


def test_data_gen_api_contract():
    import data_gen
    d = data_gen.diverse_data_generator()
    for attr in ("X", "gamma", "training_size", "options", "data", "key", "acquiredsets", "mean", "sd", "tracker", "rng"):
        assert hasattr(d, attr)
    with pytest.raises(Exception):
        d.key = 5                                   # int seeds are rejected (data_gen.py:115-117)
    d.training_size = 10.0
    assert d.training_size == 10 and d.gamma_correlation_dict is None
    d.options["bounds"] = [(0, 100), (0, 100)]
    d.options["seed"] = 0
    with pytest.raises(Exception):
        d.generate()                                # gamma is None (data_gen.py:327)
    with pytest.raises(Exception):
        d.sample(5)                                 # nothing generated yet (data_gen.py:433-436)
    assert d._window(5, 10000) == (250, 500) and d._window(95, 10000) == (9250, 9500)
    assert d._window(99, 10000) == (9500, 10000) and d._window(1, 10000) == (0, 500)
    r = data_gen.random_data_generator()
    r.options["bounds"] = [(0, 100), (0, 100)]
    r.options["seed"] = 1
    r.training_size = 10
    r.generate()
    s = r.sample()
    assert s.shape == (10, 2) and len({tuple(p) for p in s}) == 10


def test_results_and_model_contract():
    import BO
    import results
    from utils import cumoptgap
    r = results.result_opt()
    r._opt = 100
    r.minstate = False
    r.yall = np.array([50.0, 60.0, 90.0, 90.0])
    assert np.array_equal(r.opt_gap, [50, 40, 10, 10])
    assert r.cum_opt_gap() == np.trapezoid([40, 10, 10]) == cumoptgap(r.opt_gap[1:])
    r.minstate = True
    r._opt = 0
    assert np.array_equal(r.opt_gap, r.yall)
    mll, model = BO.initialize_model(np.zeros((3, 2)), np.zeros((3, 1)), nu=2.5)
    names = [n for n, _ in model.named_parameters()]
    assert names == list(BO.PARAM_NAMES)
    assert [p.item() for _, p in model.named_parameters()] == [2.0, 0.0, 0.0, 0.0]     # gpytorch defaults
    assert np.allclose(model.theta(), [2.0, 0.0, math.log(2), math.log(2)])
    for _, p in model.named_parameters():
        p.copy_(np.float64(1.5))
    h = results.hp_result()
    h.addhp(model)
    assert h.hyperparameters[BO.PARAM_NAMES[3]] == [1.5]
    assert BO.initialize_model(np.zeros((3, 2)), np.zeros((3, 1)), nu=1.5)[1].kernel_id == 2
    assert BO.initialize_model(np.zeros((3, 2)), np.zeros((3, 1)), covar="rbf")[1].kernel_id == 0
    th = BO._fixparam_to_theta({"likelihood.noise_covar.raw_noise": [0.5, 9], "mean_module.constant": [3.0],
                                "covar_module.raw_outputscale": [30.0], "covar_module.base_kernel.raw_lengthscale": [0.0]})
    assert np.allclose(th, [0.5, 3.0, 30.0, math.log(2)])
    import random
    random.seed(4)
    tx, ty, best = BO.generate_initial_data(np.arange(24).reshape(12, 2), np.arange(12, dtype=np.float32), False, n=10)
    assert tx.shape == (10, 2) and ty.shape == (10, 1) and best == ty.max()
    lat = BO._lattice_for([(0, 100), (0, 100)])
    assert (lat.Gx, lat.Gy, lat.h, lat.x0) == (100, 100, 1.0, 0.0)
    lat = BO._lattice_for([(0, 100), (0, 100)], budget=40000)
    assert (lat.Gx, lat.Gy, lat.h) == (200, 200, 0.5)


def test_optimizer_facade_contract():
    import Optimizers
    o = Optimizers.optimizer()
    assert o.max_iter == 100 and o.opt == "BO" and o.options["tol"] == 0.1 and o.options["nu"] == 2.5
    assert o.optima == 100 and o.minimize is False
    with pytest.raises(Exception):
        o.optimize(lambda x: x, method="CG")


def test_shard_ranges_cover_everything():
    from b200bo import dist
    for total in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            spans = [dist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference: the CPU arm (oracle port on the host cores) with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--iters", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "bo_iterations_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_extract_hyperparam_matches_reference_golden(golden):
    """results.hp_series.extract_hyperparam vs the reference's own output (tests/golden/hp_extract.json)."""
    import results
    g = golden("hp_extract.json")
    for seed, case in g.items():
        h = results.hp_series()
        h.data = {k: np.array(v) for k, v in case["data"].items()}
        np.random.seed(int(seed))
        h.extract_hyperparam()
        assert set(h.optdict) == set(case["optdict"])
        for k, v in case["optdict"].items():
            assert np.array_equal(h.optdict[k], np.array(v)), (seed, k)


def test_hp_series_updateres_fills_failed_fits():
    import results
    s = results.hp_series()
    for i, v in enumerate([1.0, None, 3.0, None, None]):
        r = results.hp_result()
        r.hyperparameters = None if v is None else {"a": [v], "b": [10 * v]}
        s.data.append(r)
    s.updateres()
    assert s.data["a"].tolist() == [1.0, 1.0, 3.0, 3.0, 3.0] and s.getkeys() == ["a", "b"]


def test_window_draw_consumes_the_generator_like_the_reference():
    """data_gen._pick draws r = Generator.choice(len(possible), 1) and lets the device find the r-th free rank of the
    window; the reference draws Generator.choice(list(possible), 1) over possible = np.setdiff1d(window, acquired)
    (Src/data_gen.py:445-466). Same stream position, same element — also after many draws and with acquired ranks
    inside and outside the window."""
    a, b = np.random.default_rng(123), np.random.default_rng(123)
    acquired_a, acquired_b = [7, 9999], [7, 9999]
    for pct in [5, 95, 5, 50, 95, 2.5, 97.5, 50, 5] * 6:
        if 2.5 < pct < 97.5:
            lo, hi = int((pct - 2.5) / 100 * 10000), int(pct / 100 * 10000)
        elif pct > 97.5:
            lo, hi = int(95 / 100 * 10000), 10000
        else:
            lo, hi = 0, int(5 / 100 * 10000)
        possible = np.setdiff1d(list(range(lo, hi)), acquired_a)
        loc_ref = a.choice(list(possible), 1)[0]                       # the reference's draw
        acquired_a.append(loc_ref)
        free = (hi - lo) - len({int(x) for x in acquired_b if lo <= x < hi})
        r = int(b.choice(free, 1)[0])                                  # the engine's draw
        loc = np.setdiff1d(np.arange(lo, hi), acquired_b)[r]           # what window_pick_kernel returns for r
        acquired_b.append(loc)
        assert loc == loc_ref
    assert a.integers(0, 1 << 30) == b.integers(0, 1 << 30)            # both generators are at the same position


def test_bench_contract_helpers():
    """bench.py host helpers: both arms print the SAME config (the driver compares them), configs[4] is strong-scaled by
    default (the same 100 000 trials at every N), initial sets are distinct inside a set and a function of the seed."""
    import argparse
    import bench
    args = argparse.Namespace(trials=100000, trials_per_gpu=0, iters=100, n0=10, surfaces=256, gpus=4)
    assert bench.total_trials(args, 4) == (100000, True) and bench.total_trials(args, 1) == (100000, True)
    assert bench.workload_config(args, 4) == bench.workload_config(args, max(1, args.gpus))
    assert bench.workload_config(args, 4)["trials_total"] == 100000
    weak = argparse.Namespace(**{**vars(args), "trials_per_gpu": 12500})
    assert bench.total_trials(weak, 8) == (100000, False)
    a, b = bench.init_sets(500, 10, 3), bench.init_sets(500, 10, 3)
    assert np.array_equal(a, b) and a.shape == (500, 10) and a.min() >= 0 and a.max() < 10000
    assert all(len(set(row)) == 10 for row in a)
    assert not np.array_equal(a, bench.init_sets(500, 10, 4))
