"""Generate tests/golden/*.npz|json from the REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every vector below is an output of the unmodified reference modules
(/root/reference/Src/{SimpleNoise,Combination,objectives,data_gen}.py) imported
through oracle/refshim.py (stubs for plotting/cluster packages + the CPython-3.9
``random.seed(np.int64)`` rule).  The fixtures are what pins the oracle
(tests/test_oracle_*.py) and, on the GPU box where /root/reference does not
exist, what the CUDA kernels are compared with (tests/test_gpu_*.py).
"""
import contextlib
import io
import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402


def quiet():
    return contextlib.redirect_stderr(io.StringIO())


def gen_simplex():
    sn = refshim.load("SimpleNoise")
    out = {}
    seeds = [0, 1, 7, 12345]
    out["seeds"] = np.array(seeds)
    for s in seeds:
        g = sn.SimplexNoise(s, 2)
        rng = np.random.default_rng(1000 + s)
        pts = np.concatenate([rng.uniform(0, 30, size=(480, 2)),
                              np.array([[0, 0], [0.3, 0.7], [1.5, 2.25], [12.34, 56.78], [99 / 19, 99 / 19]]),
                              rng.integers(0, 20, size=(27, 2)).astype(float)])   # exact lattice hits: tie cases
        out[f"pts_{s}"] = pts
        out[f"val_{s}"] = np.array([g.simplexNoise([float(p[0]), float(p[1])]) for p in pts])
        out[f"hseed_{s}"] = np.array([g.seed], dtype=np.uint64)
        out[f"vecs_{s}"] = np.array(g.vecs, dtype=np.int8)
        out[f"consts_{s}"] = np.array([g.f, g.g, g.cornerToFaceSquared, g.valueScaler])
    hv = [0, 1, 255, 256, 65535, 65536, 2 ** 24 - 1, 2 ** 24, 123456789, 2247598438687899 + 99 + (99 << 5)]
    out["hash_in"] = np.array(hv, dtype=np.uint64)
    out["hash_out"] = np.array([sn.randHash(v) for v in hv], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "simplex_noise.npz"), **out)


def gen_wildcat():
    ob = refshim.load("objectives")
    cases = [(0, 0.2, 0.2), (3, 0.8, 0.4), (11, 0.6, 0.8), (5, 0.4, 0.6)]
    out = {"cases": np.array(cases)}
    for i, (seed, sm, ra) in enumerate(cases):
        o = ob.objectives()
        o.bounds = [(0, 100), (0, 100)]
        o.seed = seed
        o.args = {"N": 1, "Smoothness": sm, "rug_freq": 1, "rug_amp": ra}
        with quiet():
            X, surf = o.wildcatwells()
            f = o.generate_cont()
        out[f"surf_{i}"] = surf
        if i == 0:
            out["X"] = X
            q = np.array([[10.5, 20.25], [0, 0], [100, 100], [63, 56], [99.5, 0.5], [33.3, 66.6]])
            out["cont_q"] = q
            out["cont_v"] = np.asarray(f(q), dtype=np.float64)
    o = ob.objectives()
    o.bounds = [(0, 100), (0, 100)]
    out["construct_X"] = o.construct_X()
    np.savez_compressed(os.path.join(HERE, "wildcat.npz"), **out)


def gen_wildcat_cont():
    """generate_cont() (objectives.py:322-351) off the nodes: the reference's own LinearNDInterpolator on the
    200 x 200 half-node lattice of configs[3] (data_gen.constructX(budget=40000)), on random queries, on the
    hull boundary / outside it, plus the Delaunay split its triangulation uses (per-cell diagonal)."""
    from oracle import wildcat as owc
    ob = refshim.load("objectives")
    dg = refshim.load("data_gen")
    d = dg.diverse_data_generator()
    d.options["bounds"] = [(0, 100), (0, 100)]
    d.constructX(budget=40000)
    out = {"lattice200": np.asarray(d.X, dtype=np.float64)}
    cases = [(0, 0.2, 0.2), (3, 0.8, 0.4)]
    out["cases"] = np.array(cases)
    rng = np.random.default_rng(99)
    q = np.concatenate([rng.uniform(0, 100, size=(3000, 2)),
                        np.stack([rng.integers(0, 100, 400) + 0.5, rng.integers(0, 100, 400) + 0.5], axis=1),   # cell centres
                        np.stack([rng.integers(0, 101, 200).astype(float), rng.uniform(0, 100, 200)], axis=1),  # on cell edges
                        rng.integers(0, 400, size=(400, 2)) / 4.0,                                              # quarter nodes
                        np.array([[0, 0], [100, 100], [100, 0], [0, 100], [99.5, 0.5], [10.5, 20.25], [100, 37.25],
                                  [-0.25, 5], [5, 100.5], [101, 101], [50, -1e-9]])])
    out["queries"] = q
    for i, (seed, sm, ra) in enumerate(cases):
        o = ob.objectives()
        o.bounds = [(0, 100), (0, 100)]
        o.seed = seed
        o.args = {"N": 1, "Smoothness": sm, "rug_freq": 1, "rug_amp": ra}
        with quiet():
            f = o.generate_cont()
        out[f"lattice200_v_{i}"] = np.asarray(f(out["lattice200"]), dtype=np.float64)
        out[f"queries_v_{i}"] = np.asarray(f(q), dtype=np.float64)
        mask = owc.diag_mask_from_triangulation(f.tri.points, f.tri.simplices)
        if i == 0:
            out["diag_mask"] = mask
        else:
            assert np.array_equal(mask, out["diag_mask"]), "the Delaunay split must not depend on the surface"
    np.savez_compressed(os.path.join(HERE, "wildcat_cont.npz"), **out)


def gen_combinadics():
    cb = refshim.load("Combination")
    cf = cb.combination_finder()
    r = random.Random(2024)
    rows = []
    for n, k in [(5, 3), (12, 4), (100, 7), (10000, 10), (10000, 5), (10000, 20), (40000, 10), (400, 50)]:
        C = math.comb(n, k)
        ms = list(range(min(C, 10))) if C <= 10 else [0, 1, C - 1, C // 2, C // 3]
        ms += [r.randint(0, C - 1) for _ in range(40)]
        if (n, k) == (10000, 10):
            ms += [10 ** 30, 1234366987659256151596162756710190]
        for m in ms:
            rows.append({"n": n, "k": k, "m": str(m), "set": cf.find_elm(n, k, m)})
    with open(os.path.join(HERE, "combinadics.json"), "w") as fh:
        json.dump(rows, fh)


def gen_dpp():
    dg = refshim.load("data_gen")
    out = {}
    pts = np.array([[0, 0], [10, 0], [0, 10], [50, 50], [99, 99], [3, 77], [60, 20], [25, 25], [80, 45], [45, 80]], dtype=np.float64)
    d = dg.diverse_data_generator()
    gammas = [1e-5, 1e-4, 1e-6]
    out["kat_pts"] = pts
    out["kat_gamma"] = np.array(gammas)
    out["kat_score"] = np.array([d.twoD_score(pts, g) for g in gammas])
    # random sets of several sizes, scored by the reference's twoD_score
    rng = np.random.default_rng(7)
    for k in (2, 3, 5, 10, 16, 25, 50):
        sets = np.stack([rng.choice(10000, size=k, replace=False) for _ in range(64)]).astype(np.int32)
        pool = None
        dd = dg.diverse_data_generator()
        dd.options["bounds"] = [(0, 100), (0, 100)]
        dd.constructX()
        pool = dd.X
        out[f"rand_idx_k{k}"] = sets
        for g in gammas:
            out[f"rand_score_k{k}_g{g:g}"] = np.array([dd.twoD_score(pool[s], g) for s in sets])
    out["pool"] = pool
    dd.constructX(budget=40000)
    out["pool40000_head"] = dd.X[:5]
    out["pool40000_shape"] = np.array(dd.X.shape)
    # full distributions through generate(): un-rank + score + normalise + argsort + sample
    for seed, N, gamma in [(0, 10000, 1e-5), (2, 2000, 1e-5), (4, 2000, 1e-6)]:
        d = dg.diverse_data_generator()
        d.options["bounds"] = [(0, 100), (0, 100)]
        d.options["N_samples"] = N
        d.gamma = gamma
        d.options["seed"] = seed
        d.training_size = 10
        d.generate()
        key = d.key
        tag = f"gen_s{seed}"
        out[tag + "_N"] = np.array([N])
        out[tag + "_gamma"] = np.array([gamma])
        out[tag + "_mean_sd"] = np.array([d.mean, d.sd])
        out[tag + "_scores"] = np.asarray(d.data[key]["scores"])
        out[tag + "_dist"] = np.asarray(d.data[key]["dist"]).astype(np.int16)
        # the element ranks and un-ranked index sets the distribution was built from
        elems = d.iid_comb_locs()
        out[tag + "_elements"] = np.array([str(e) for e in elems])
        out[tag + "_pos"] = np.array([d.combination_sampler.find_elm(10000, 10, e) for e in elems], dtype=np.int32)
        s5 = d.sample(5)
        s95 = d.sample(95)
        out[tag + "_acquired"] = np.array(d.acquiredsets)
        out[tag + "_sample5"] = np.asarray(s5)
        out[tag + "_sample95"] = np.asarray(s95)
        r, z = d.compare_subDPP(key, pts)
        out[tag + "_compare"] = np.array([r, z])
    np.savez_compressed(os.path.join(HERE, "dpp.npz"), **out)


def gen_gammacalc():
    """diverse_data_generator.gammacalc (data_gen.py:758-842): 16 x 16 rank-correlation matrix over gamma = 10^-12..10^3
    (256 generate() calls of ``min_samples`` sets + a re-score each) and the widest correlated pair it returns, from the
    reference's own code. np.random is seeded because gammacorrelation shuffles the order of a NaN batch (:773-774)."""
    import warnings
    dg = refshim.load("data_gen")
    seed, min_samples, np_seed = 5, 60, 123
    d = dg.diverse_data_generator()
    d.options["bounds"] = [(0, 100), (0, 100)]
    d.options["seed"] = seed
    d.training_size = 10
    np.random.seed(np_seed)
    with warnings.catch_warnings(), quiet(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pair = d.gammacalc(min_samples=min_samples)
    grange = list(range(-12, 4))
    matrix = [[float(d.gamma_correlation_dict[a][b]) for b in grange] for a in grange]
    with open(os.path.join(HERE, "gammacalc.json"), "w") as fh:
        json.dump({"seed": seed, "min_samples": min_samples, "np_seed": np_seed, "grange": grange, "matrix": matrix,
                   "pair": [int(pair[0]), int(pair[1])]}, fh)


def gen_author_data():
    """Fixed hyper-parameters the authors' runs used (BOloop_fixparam takes
    fixparam[name][0], BO.py:395).  Values read from
    Data/experiment3-2-1/smoothness/<s>/ruggedness/<r>/data0.pkl with dill."""
    try:
        import dill
        refshim.load("results")
        import results as _  # noqa
    except Exception:
        pass
    rows = {}
    base = "/root/reference/Data/experiment3-2-1/smoothness"
    try:
        import dill
        sys.path.insert(0, refshim.REF_SRC)
        refshim._install_stubs()
        import results  # noqa: F401  (reference's; needed by the unpickler)
        for s in ("0.2", "0.4", "0.6", "0.8"):
            for r in ("0.2", "0.4", "0.6", "0.8"):
                fn = f"{base}/{s}/ruggedness/{r}/data0.pkl"
                if not os.path.exists(fn):
                    continue
                with open(fn, "rb") as fh:
                    hp = dill.load(fh)
                if not getattr(hp, "optdict", None):
                    try:
                        hp.extract_hyperparam()
                    except Exception:
                        pass
                od = getattr(hp, "optdict", None)
                if od:
                    rows[f"{s}/{r}"] = {k: [float(x) for x in np.atleast_1d(v)] for k, v in od.items()}
    except Exception as e:  # pragma: no cover
        rows["error"] = repr(e)
    finally:
        if refshim.REF_SRC in sys.path:
            sys.path.remove(refshim.REF_SRC)
        for k in ("results", "utils"):
            m = sys.modules.get(k)
            if m is not None and getattr(m, "__file__", "").startswith(refshim.REF_SRC):
                del sys.modules[k]
    with open(os.path.join(HERE, "author_hyperparams.json"), "w") as fh:
        json.dump(rows, fh, indent=1)


def gen_bootstrap():
    """utils.bootstrap (Src/utils.py:39-115) on gap-curve shaped inputs, seeded through the global generator."""
    ut = refshim.load("utils")
    rng = np.random.default_rng(5)
    out = {}
    # (N trials, D positions) monotone gap curves like results.opt_gap, and a vector like cum_opt_gap
    curves = np.maximum.accumulate(rng.uniform(0, 100, size=(37, 12)), axis=1)[:, ::-1].copy()
    big = rng.normal(50, 20, size=(300, 5))                       # N > 128: exercises the pairwise split
    vec = rng.gamma(2.0, 100.0, size=41)
    for name, data, trials, seed in (("curves", curves, 200, 11), ("big", big, 64, 12), ("vec", vec, 150, 13)):
        random.seed(seed)
        out[f"{name}_data"] = data
        out[f"{name}_trials"] = np.array([trials, seed])
        out[f"{name}_ci"] = np.asarray(ut.bootstrap(data if name != "curves" else [row for row in data], trials=trials))
    np.savez_compressed(os.path.join(HERE, "bootstrap.npz"), **out)


def gen_hp_extract():
    """hp_series.extract_hyperparam (Src/results.py:342-392) on synthetic hyper-parameter samples; np.random seeded."""
    res = refshim.load("results")
    out = {}
    for seed in (0, 1, 2):
        rng = np.random.default_rng(seed)
        data = {"likelihood.noise_covar.raw_noise": np.abs(rng.normal(2e-3, 5e-4, 400)),
                "mean_module.raw_constant": np.concatenate([rng.normal(40, 1.5, 300), rng.normal(60, 1.0, 280)]),
                "covar_module.raw_outputscale": np.concatenate([rng.normal(190, 8, 350), rng.normal(120, 3, 40)]),
                "covar_module.base_kernel.raw_lengthscale": rng.gamma(20, 0.5, 500)}
        h = res.hp_series.__new__(res.hp_series)          # hp_series.__init__ trips over its setter-less property
        h.data = data
        np.random.seed(seed)
        import types
        had = sys.modules.get("utils")
        sys.modules["utils"] = types.SimpleNamespace(bootstrap=None)     # extract_hyperparam imports it and never uses it
        try:
            with quiet():
                h.extract_hyperparam()
        finally:
            if had is None:
                del sys.modules["utils"]
            else:
                sys.modules["utils"] = had
        out[str(seed)] = {"data": {k: v.tolist() for k, v in data.items()}, "optdict": {k: np.asarray(v).tolist() for k, v in h.optdict.items()}}
    with open(os.path.join(HERE, "hp_extract.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    gen_simplex()
    gen_wildcat()
    gen_wildcat_cont()
    gen_combinadics()
    gen_dpp()
    gen_gammacalc()
    gen_author_data()
    gen_bootstrap()
    gen_hp_extract()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
