"""CPU: the oracle (oracle/*.py) against the golden vectors generated from the reference's own code
(tests/golden/make_golden.py), and — for the GP part, which the reference cannot pin — against
scikit-learn's independent exact GP and finite differences."""
import math

import numpy as np
import pytest

from oracle import dpp as odpp
from oracle import gp as ogp
from oracle import wildcat as owc


def test_simplex_noise_scalar_and_vectorised_bit_exact(golden):
    g = golden("simplex_noise.npz")
    for s in g["seeds"]:
        ng = owc.SimplexNoise2D(int(s))
        assert ng.seed == int(g[f"hseed_{s}"][0])
        assert np.array_equal(np.array(ng.vecs, dtype=np.int8), g[f"vecs_{s}"])
        assert np.array_equal(np.array([ng.f, ng.g, ng.corner_to_face_sq, ng.value_scaler]), g[f"consts_{s}"])
        pts, ref = g[f"pts_{s}"], g[f"val_{s}"]
        assert np.array_equal(ng.noise_grid(pts[:, 0], pts[:, 1]), ref)
        assert np.array_equal(np.array([ng.noise([float(a), float(b)]) for a, b in pts[:100]]), ref[:100])
    assert np.array_equal(np.array([owc.rand_hash(int(v)) for v in g["hash_in"]], dtype=np.uint64), g["hash_out"])
    assert np.array_equal(owc.rand_hash_u64(g["hash_in"]), g["hash_out"])


def test_simplex_known_answers():
    """SURVEY.md Appendix C constants."""
    ng = owc.SimplexNoise2D(0)
    assert ng.seed == 0 and ng.vecs == [[-1, 0], [0, -1], [0, 1], [1, 0]]
    assert ng.noise([0.3, 0.7]) == 0.8976723367630091
    assert ng.noise([12.34, 56.78]) == -0.6020823634749848
    assert owc.SimplexNoise2D(1).seed == 322502465311884


def test_wildcat_surfaces_bit_exact(golden):
    g = golden("wildcat.npz")
    assert np.array_equal(owc.construct_X(), g["construct_X"])
    for i, (seed, sm, ra) in enumerate(g["cases"]):
        X, surf = owc.wildcatwells(int(seed), float(sm), float(ra))
        assert np.array_equal(surf, g[f"surf_{i}"]) and surf.dtype == np.float32
        assert np.array_equal(X, g["X"])
    s0 = g["surf_0"]
    assert s0.max() == 100.0 and np.unravel_index(s0.argmax(), s0.shape) == (63, 56)
    assert abs(float(s0.min()) - (-8.219039)) < 1e-5
    # generate_cont at nodes == node value (f32) exactly; off-node value recorded for information
    q, v = g["cont_q"], g["cont_v"]
    assert v[1] == s0[0, 0] and v[2] == s0[100, 100] and v[3] == s0[56, 63]
    assert np.array_equal(owc.lattice_lookup(s0, q[1:4]), v[1:4].astype(np.float32))
    assert owc.peak_location(0) == (54.88135039273247, 67.781653679623)


def test_generate_cont_off_node_values(golden):
    """objectives.py:349-350: the restated LinearNDInterpolator with the Delaunay split of the reference's own
    triangulation — bit-identical on the 200 x 200 half-node lattice of configs[3], <= 1e-12 on arbitrary queries,
    NaN outside the hull; and the mask the product ships is the mask of the reference's interpolator object."""
    import os
    c, w = golden("wildcat_cont.npz"), golden("wildcat.npz")
    mask = c["diag_mask"]
    assert mask.shape == (100, 100) and set(np.unique(mask)) == {0, 1}
    shipped = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                           "jmd-diversity-in-bayesian-optimization_b200", "b200bo", "data", "qhull_diag_101.bin")
    assert np.array_equal(np.unpackbits(np.fromfile(shipped, dtype=np.uint8))[:10000].reshape(100, 100), mask)
    assert c["lattice200"].shape == (40000, 2) and c["lattice200"][1, 0] == 0.5
    for i in range(len(c["cases"])):
        surf = w[f"surf_{[tuple(x) for x in w['cases']].index(tuple(c['cases'][i]))}"]
        assert np.array_equal(owc.interp_linear(surf, c["lattice200"], mask), c[f"lattice200_v_{i}"])
        got, ref = owc.interp_linear(surf, c["queries"], mask), c[f"queries_v_{i}"]
        assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(ref).sum() == 4
        ok = ~np.isnan(ref)
        assert np.all(np.abs(got[ok] - ref[ok]) <= 1e-12)
        # the split matters: the all-main-diagonal convention of round 1 is off by > 0.4 objective units somewhere
        assert np.nanmax(np.abs(owc.interp_linear(surf, c["queries"], None) - ref)) > 0.4
    q = np.array([[10.5, 20.25]])
    assert abs(owc.interp_linear(w["surf_0"], q, mask)[0] - w["cont_v"][0]) <= 1e-12      # 33.34358835


def test_unrank_matches_reference_find_elm(golden):
    for row in golden("combinadics.json"):
        assert odpp.unrank(row["n"], row["k"], int(row["m"])) == row["set"]
    with pytest.raises(Exception):
        odpp.unrank(10000, 10, math.comb(10000, 10))


def test_dpp_scores_match_reference(golden):
    g = golden("dpp.npz")
    for gamma, ref in zip(g["kat_gamma"], g["kat_score"]):
        assert abs(odpp.score_lu(g["kat_pts"], float(gamma)) - ref) <= 1e-12 * abs(ref)
    assert np.array_equal(odpp.construct_pool(), g["pool"])
    assert np.array_equal(odpp.construct_pool(budget=40000)[:5], g["pool40000_head"])
    assert tuple(odpp.construct_pool(budget=40000).shape) == tuple(g["pool40000_shape"])
    pool = g["pool"]
    for k in (2, 5, 10, 16):
        sets = g[f"rand_idx_k{k}"][:16]
        for gamma in (1e-5, 1e-4):
            ref = g[f"rand_score_k{k}_g{gamma:g}"][:16]
            got = np.array([odpp.score_lu(pool[s], gamma) for s in sets])
            assert np.allclose(got, ref, rtol=1e-11, atol=0)
            chol = np.array([odpp.score_chol(pool[s], gamma) for s in sets])
            bound = np.array([odpp.cond_bound_rel(pool[s], gamma) for s in sets])
            assert np.all(np.abs(chol - ref) <= bound * np.abs(ref))


@pytest.mark.parametrize("seed", [2, 4])
def test_dpp_generate_end_to_end(golden, seed):
    """iid_comb_locs -> un-rank -> score -> normalise -> argsort -> sample, vs the reference's generate()."""
    g = golden("dpp.npz")
    tag = f"gen_s{seed}"
    N, gamma = int(g[tag + "_N"][0]), float(g[tag + "_gamma"][0])
    elems = odpp.iid_comb_locs(10000, 10, N, seed)
    assert [str(e) for e in elems] == list(g[tag + "_elements"])
    out = odpp.generate(g["pool"], 10, N, gamma, seed)
    assert np.array_equal(out["unsorted_idx"], g[tag + "_pos"])
    assert np.allclose([out["mean"], out["sd"]], g[tag + "_mean_sd"], rtol=1e-13)
    assert np.allclose(out["scores"], g[tag + "_scores"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(out["dist"].astype(np.int16), g[tag + "_dist"])
    rng = np.random.default_rng(seed)
    acq = []
    acq.append(odpp.sample_rank(5, N, acq, rng))
    acq.append(odpp.sample_rank(95, N, acq, rng))
    assert acq == list(g[tag + "_acquired"])
    assert np.array_equal(out["dist"][acq[0]], g[tag + "_sample5"])
    assert np.array_equal(out["dist"][acq[1]], g[tag + "_sample95"])
    r, z = odpp.compare_rank(out["scores"], out["mean"], out["sd"], odpp.score_lu(g["kat_pts"], gamma))
    assert r == int(g[tag + "_compare"][0]) and abs(z - g[tag + "_compare"][1]) < 1e-9


def test_sample_windows():
    assert odpp.sample_window(5, 10000) == (250, 500)
    assert odpp.sample_window(95, 10000) == (9250, 9500)
    assert odpp.sample_window(99, 10000) == (9500, 10000)
    assert odpp.sample_window(1, 10000) == (0, 500)
    assert odpp.sample_window(2.5, 10000) == (0, 500)


# ------------------------------------------------------------------ GP (parity unpinned: independent checks)

def _data(golden, n, seed=0):
    rng = np.random.default_rng(seed)
    surf = golden("wildcat.npz")["surf_0"]
    X = rng.integers(0, 100, size=(n, 2)).astype(float)
    X = np.unique(X, axis=0)
    return X, owc.lattice_lookup(surf, X).astype(np.float64)


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_posterior_matches_sklearn(golden, lattice_grid, kid):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel
    X, y = _data(golden, 25)
    theta = (1.762e-3, 41.83, 190.43, 9.8695)
    base = {0: RBF(theta[3]), 1: Matern(theta[3], nu=0.5), 2: Matern(theta[3], nu=1.5), 3: Matern(theta[3], nu=2.5)}[kid]
    gpr = GaussianProcessRegressor(kernel=ConstantKernel(theta[2]) * base + WhiteKernel(theta[0]), optimizer=None, alpha=0.0)
    gpr.fit(X, y - theta[1])
    grid = lattice_grid[::7]
    mu_s, sd_s = gpr.predict(grid, return_std=True)
    L, alpha, _, st = ogp.gp_state(X, y, theta, kid)
    mu, var = ogp.posterior(grid, X, theta, kid, L, alpha)
    assert st == 0
    assert np.allclose(mu, mu_s + theta[1], rtol=1e-9, atol=1e-9)
    assert np.allclose(var, sd_s ** 2 - theta[0], rtol=1e-7, atol=1e-9 * theta[2])
    # sklearn's log marginal likelihood == the data term of the MLL
    raw = np.array([theta[0], theta[1], float(ogp.inv_softplus(theta[2])), float(ogp.inv_softplus(theta[3]))])
    v, g, _ = ogp.mll_value_grad(X, y, raw, kid)
    n = len(y)
    prior = ogp.PRIOR_A * math.log(ogp.PRIOR_RATE) + (ogp.PRIOR_A - 1) * math.log(theta[0]) - ogp.PRIOR_RATE * theta[0] - math.lgamma(ogp.PRIOR_A)
    assert abs(v * n - prior - gpr.log_marginal_likelihood_value_) < 1e-8 * abs(v * n)


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_mll_gradient_finite_differences(golden, kid):
    X, y = _data(golden, 20, seed=3)
    raw = np.array([0.5, 30.0, 5.0, 2.0])
    v, g, st = ogp.mll_value_grad(X, y, raw, kid)
    num = np.zeros(4)
    for i in range(4):
        h = 1e-6 * max(1.0, abs(raw[i]))
        p, m = raw.copy(), raw.copy()
        p[i] += h
        m[i] -= h
        num[i] = (ogp.mll_value_grad(X, y, p, kid)[0] - ogp.mll_value_grad(X, y, m, kid)[0]) / (2 * h)
    assert np.allclose(g, num, rtol=1e-6, atol=1e-8)


def test_ei_properties_and_argmax_rule():
    mu = np.linspace(-3, 3, 50)
    var = np.full(50, 0.5)
    ei = ogp.acquisition(mu, var, 0.0)
    assert np.all(ei >= 0) and np.all(np.diff(ei) > 0)                      # EI >= 0, monotone in mu
    assert abs(ogp.acquisition(np.array([0.0]), np.array([1.0]), 0.0)[0] - ogp.INV_SQRT_2PI) < 1e-15
    assert ogp.norm_cdf(np.array([-30.0]))[0] > 0                           # erfc form keeps the left tail
    v = np.array([1.0, 3.0, 3.0, np.nan, 2.0])
    assert ogp.argmax_masked(v)[0] == 1                                     # first maximum; NaN never wins
    assert ogp.argmax_masked(v, mask=np.array([0, 1, 0, 0, 0]))[0] == 2
    assert ogp.argmax_masked(v, mask=np.ones(5))[0] == -1
    assert ogp.argmax_masked(v, mask=np.ones(5))[2] & ogp.ST_ALL_MASKED


def test_bo_loop_semantics(golden, lattice_grid):
    """yall layout of BOloop (BO.py:146,169,262,267-271): [best0, best after each iteration..., last again, padding]."""
    surf = golden("wildcat.npz")["surf_0"]
    obj = surf[:100, :100].reshape(-1)
    X, y = _data(golden, 10, seed=5)
    theta = ogp.raw_to_theta((0.0017619397258386016, 41.82688735961914, 190.43237411499024, 9.869425163269042))
    out = ogp.bo_loop(lattice_grid, obj, X, y, 30, theta=theta, tol=0.1)
    assert out["yall"].shape == (32,)
    assert out["yall"][0] == y.max()
    k = out["nit"]
    assert np.all(out["yall"][k:] == out["yall"][k])
    assert np.all(np.diff(out["yall"]) >= 0)
    if k < 30:
        assert 100.0 - out["yall"][k] <= 0.1
    assert len({tuple(p) for p in out["X"]}) == len(out["X"])               # never re-queries a point


def test_psd_safe_cholesky_levels():
    K = np.ones((3, 3))
    L, st = ogp.chol_safe(K)
    assert st == ogp.ST_JITTER and np.all(np.isfinite(L))
    L, st = ogp.chol_safe(-np.eye(3))
    assert st == (ogp.ST_JITTER | ogp.ST_CHOL_FAIL)


def test_bootstrap_oracle_matches_reference_vectors(golden):
    """oracle/bootstrap.py vs the reference's own utils.bootstrap (tests/golden/bootstrap.npz): bit-exact."""
    import random
    from oracle import bootstrap as ob
    g = golden("bootstrap.npz")
    for name in ("curves", "big", "vec"):
        trials, seed = (int(v) for v in g[f"{name}_trials"])
        random.seed(seed)
        ci = ob.bootstrap(g[f"{name}_data"], trials)
        assert ci.shape == g[f"{name}_ci"].shape
        assert np.array_equal(ci, g[f"{name}_ci"])
