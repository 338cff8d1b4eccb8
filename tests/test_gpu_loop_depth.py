"""GPU parity of whole BO loops at the depth and on the shapes BASELINE.json's configs run:

* configs[2] (Experiment2: a hyper-parameter fit before every acquisition, Src/BO.py:152-193) — the engine's recorded
  per-iteration hyper-parameters replayed through the oracle loop: picks, curves and nit bit-equal;
* configs[3] / configs[4] (fixed hyper-parameters, 100 iterations, n = 10 -> 110 observations) on the 100 x 100 and the
  200 x 200 (step 0.5, constructX(budget=40000)) lattices against the full oracle loop, both posterior modes;
* BOloop_singleopt (Src/BO.py:284-374);
* initial sets that are not lattice nodes.
"""
import random

import numpy as np
import pytest
import torch

from conftest import AUTHOR_THETA_RAW
from oracle import gp as ogp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_t = lambda a, dt=None: torch.as_tensor(np.ascontiguousarray(a), device=DEV, dtype=dt)


def _inv_softplus(v):
    return v if v > 20.0 else float(np.log(np.expm1(v)))


def _theta_to_raw(th):
    return np.array([th[0], th[1], _inv_softplus(th[2]), _inv_softplus(th[3])])


def _lattice100_case(golden, lattice_grid, T, n0, seed):
    g = golden("wildcat.npz")
    obj = np.stack([g[f"surf_{i}"] for i in range(4)])[:, :100, :100].reshape(4, -1).astype(np.float32)
    rng = np.random.default_rng(seed)
    surf_id = rng.integers(0, 4, T).astype(np.int32)
    idx0 = np.stack([rng.choice(10000, n0, replace=False) for _ in range(T)])
    X0 = lattice_grid[idx0]
    y0 = np.stack([obj[surf_id[t], idx0[t]] for t in range(T)]).astype(np.float64)
    return lattice_grid, obj, surf_id, X0, y0


def _lattice200_case(golden, T, n0, seed):
    """configs[3]: the 200 x 200 half-node lattice; objective values = the reference's own generate_cont() outputs
    (tests/golden/wildcat_cont.npz), cast to fp32 as BO.py:104 does."""
    c = golden("wildcat_cont.npz")
    grid = c["lattice200"]
    obj = np.stack([c["lattice200_v_0"], c["lattice200_v_1"]]).astype(np.float32)
    rng = np.random.default_rng(seed)
    surf_id = rng.integers(0, 2, T).astype(np.int32)
    idx0 = np.stack([rng.choice(40000, n0, replace=False) for _ in range(T)])
    X0 = grid[idx0]
    y0 = np.stack([obj[surf_id[t], idx0[t]] for t in range(T)]).astype(np.float64)
    return grid, obj, surf_id, X0, y0


def test_fit_mode_loop_replays_through_the_oracle(golden, lattice_grid):
    """BOloop (fit before every acquisition, BO.py:144,152-193): run the engine in fit mode, then hand the oracle loop
    the hyper-parameters the engine used at every iteration — picked / yall / nit must be bit-equal. A wrong theta,
    kernel-table or L^-1 hand-off between the fit kernel and the acquisition kernel cannot pass this. Also: every
    recorded theta really is the fit of that iteration's data (re-fitted stand-alone: bit-equal; as good as scipy's
    L-BFGS-B optimum of the oracle MLL on a sample), including the fit after the last append (BO.py:176-189,252)."""
    import BO
    from b200bo import fit
    T, n0, max_iter = 10, 10, 60
    grid, obj, surf_id, X0, y0 = _lattice100_case(golden, lattice_grid, T, n0, seed=21)
    res = BO.BOloop_batched(obj, surf_id, X0, y0, max_iter, theta=None, tol=0.1, device=DEV, groups=1)
    torch.cuda.synchronize()
    nit, picked, yall = res.nit.cpu().numpy(), res.picked.cpu().numpy(), res.yall.cpu().numpy()
    thetas, X, y = res.thetas.cpu().numpy(), res.X.cpu().numpy(), res.y.cpu().numpy()
    assert thetas.shape == (T, max_iter + 1, 4) and res.status.max().item() == 0
    assert nit.max() > 40 and nit.min() < max_iter                  # both early stops (tol) and deep trials are covered
    for i in range(T):
        o = ogp.bo_loop(grid, obj[surf_id[i]], X0[i], y0[i], max_iter, theta=thetas[i, :max(nit[i], 1)], tol=0.1)
        assert nit[i] == o["nit"], (i, nit[i], o["nit"])
        assert np.array_equal(picked[i, :nit[i]], o["idx"]), i
        assert np.array_equal(yall[i], o["yall"]), i
        assert np.array_equal(X[i, :n0 + nit[i]], o["X"]) and np.array_equal(y[i, :n0 + nit[i]], o["y"])
    # every recorded theta — rows 0..nit, the last one being the model re-created after the final append — is the
    # fused fit of exactly that iteration's observations
    rng = np.random.default_rng(0)
    worse = checked = 0
    for i in range(T):
        for k in sorted({0, 1, int(nit[i])} | set(rng.integers(0, nit[i] + 1, 3).tolist())):
            n = n0 + k
            f = fit.LBFGSFitter(1, torch.device(DEV))
            th = torch.zeros((1, 4), dtype=torch.float64, device=DEV)
            f.fit_into(_t(X[i:i + 1, :n]), _t(y[i:i + 1, :n]), _t(np.array([n], dtype=np.int32)), 3, th, n_hint=n)
            assert np.array_equal(th.cpu().numpy()[0], thetas[i, k]), (i, k)
            if checked < 24:
                v_here = ogp.mll_value_grad(X[i, :n], y[i, :n], _theta_to_raw(thetas[i, k]), 3)[0]
                v_ref = ogp.fit_scipy(X[i, :n], y[i, :n], 3)[1]
                worse += v_here < v_ref - 1e-5 * abs(v_ref)
                checked += 1
    assert worse <= 3, worse             # different optimisers may land in different local optima, rarely
    # the hyper-parameter series handed to results.hp_series is the post-append one (BO.py:252): rows 1..nit
    r = BO._pack_result(res, 0, n0, max_iter, False, 100.0, X0[0])
    hp = r.hpdata.hyperparameters
    assert all(len(v) == nit[0] for v in hp.values())
    assert np.allclose(hp["mean_module.raw_constant"], thetas[0, 1:nit[0] + 1, 1], rtol=0, atol=0)
    assert np.allclose(hp["likelihood.noise_covar.raw_noise"], thetas[0, 1:nit[0] + 1, 0], rtol=0, atol=0)


_ORACLE_CACHE = {}


def _oracle_loops(key, grid, obj, surf_id, X0, y0, max_iter, theta, kid):
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = [ogp.bo_loop(grid, obj[surf_id[i]], X0[i], y0[i], max_iter, theta=theta, kernel_id=kid, tol=-1.0)
                              for i in range(len(surf_id))]
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("posterior", ["full", "incremental", "persistent"])
@pytest.mark.parametrize("budget", [10000, 40000])
def test_full_depth_fixed_theta_loop_matches_oracle(golden, lattice_grid, budget, posterior):
    """configs[3] / configs[4] depth: 100 iterations from 10 initial points (n -> 110, every NI / thread-count variant
    of the K-B kernel, rank-1 appends all the way) against the full oracle loop — picked, yall, nit bit-equal."""
    from b200bo import engine, ops
    n0, max_iter, kid = 10, 100, 3
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    if budget == 10000:
        T = 6
        grid, obj, surf_id, X0, y0 = _lattice100_case(golden, lattice_grid, T, n0, seed=31)
    else:
        T = 4
        grid, obj, surf_id, X0, y0 = _lattice200_case(golden, T, n0, seed=32)
        lat = ops.Lattice(200, 200, 0.0, 0.0, 0.5)
        assert np.array_equal(lat.points(DEV).cpu().numpy(), grid)
        # the engine's own objective table on this lattice is the reference interpolator's, bit for bit
        w, c = golden("wildcat.npz"), golden("wildcat_cont.npz")
        cases = [tuple(x) for x in w["cases"]]
        surfs = _t(np.stack([w[f"surf_{cases.index(tuple(x))}"] for x in c["cases"]]))
        sid = torch.arange(2, dtype=torch.int32, device=DEV).repeat_interleave(40000)
        tab = ops.wildcat_eval(surfs, lat.points(DEV).repeat(2, 1), sid).reshape(2, -1).float()
        assert np.array_equal(tab.cpu().numpy(), obj)
    if posterior == "persistent" and budget == 40000:
        # the 200 x 200 table (638 KB) does not fit shared memory: explicit request refused, "auto" = full recomputation
        with pytest.raises(ops._lib.B200boError):
            engine.BatchedBO(_t(grid), _t(obj), _t(surf_id), _t(X0), _t(y0), max_iter, theta=_t(theta), kernel_id=kid,
                             tol=-1.0, posterior="persistent")
        posterior = "auto"
    bo = engine.BatchedBO(_t(grid), _t(obj), _t(surf_id), _t(X0), _t(y0), max_iter, theta=_t(theta), kernel_id=kid,
                          tol=-1.0, posterior=posterior)
    assert bo.persistent == (posterior == "persistent")
    assert bo.lattice is not None and (bo.lattice.Gx, bo.lattice.h) == ((100, 1.0) if budget == 10000 else (200, 0.5))
    res = bo.run()
    torch.cuda.synchronize()
    assert res.status.max().item() == 0 and torch.all(res.nit == max_iter)
    oracle = _oracle_loops(budget, grid, obj, surf_id, X0, y0, max_iter, theta, kid)
    for i, o in enumerate(oracle):
        assert o["nit"] == max_iter
        got = res.picked[i].cpu().numpy()
        assert np.array_equal(got, o["idx"]), (i, np.nonzero(got != o["idx"])[0][:3])
        assert np.array_equal(res.yall[i].cpu().numpy(), o["yall"])
        assert np.array_equal(res.X[i].cpu().numpy(), o["X"]) and np.array_equal(res.y[i].cpu().numpy(), o["y"])


def test_boloop_singleopt_matches_oracle(golden, lattice_grid):
    """BOloop_singleopt (BO.py:284-374): ONE fit on the initial data, then the loop with those hyper-parameters.
    The fitted theta must be (as good as) the optimum scipy's L-BFGS-B finds on the oracle MLL, and the loop the
    oracle loop under that theta: picks, curve, nit bit-equal."""
    import BO
    import objectives
    o = objectives.objectives()
    o.bounds = [(0, 100), (0, 100)]
    o.seed = 3
    o.args = {"N": 1, "Smoothness": 0.8, "rug_freq": 1, "rug_amp": 0.4}
    f = o.generate_cont()
    rng = np.random.default_rng(8)
    X = rng.integers(0, 100, size=(14, 2)).astype(np.float64)          # more than 10 points: only 10 are drawn (BO.py:135-138)
    Y = f(X).astype(np.float32)
    max_iter = 30
    random.seed(41)
    res = BO.BOloop_singleopt(f, max_iter, X, Y, tol=0.1, nu=2.5, device=DEV)
    # the same trial by hand: same permutation of the global generator, one fit, oracle loop
    random.seed(41)
    tx, ty, best = BO.generate_initial_data(X, Y, False, n=10)
    X0 = tx.numpy().astype(np.float64)
    y0 = ty.numpy().reshape(-1).astype(np.float64)
    assert np.array_equal(y0, f(X0).astype(np.float32).astype(np.float64))
    mll, model = BO.initialize_model(tx, ty, nu=2.5)
    BO.fit_gpytorch_model(mll, device=DEV)
    theta = model.theta()
    v_here = ogp.mll_value_grad(X0, y0, model.raw(), 3)[0]
    v_ref = ogp.fit_scipy(X0, y0, 3)[1]
    assert v_here >= v_ref - 1e-6 * abs(v_ref)
    obj = f.grid_table(__import__("b200bo").ops.Lattice(100, 100)).cpu().numpy()
    orc = ogp.bo_loop(lattice_grid, obj, X0, y0, max_iter, theta=theta, kernel_id=3, tol=0.1)
    assert res.nit == orc["nit"] and res.success
    assert np.array_equal(res.yall, orc["yall"]) and res.yall.shape == (max_iter + 2,)
    assert np.array_equal(res.xall.astype(np.float64), orc["xall"])
    assert float(res.yopt) == orc["yall"][-1]
    # nu > 2.5 switches to the RBF kernel (BO.py:133-134): same contract
    random.seed(41)
    res_rbf = BO.BOloop_singleopt(f, 8, X, Y, tol=0.1, nu=3.5, device=DEV)
    random.seed(41)
    tx, ty, _ = BO.generate_initial_data(X, Y, False, n=10)
    mll, model = BO.initialize_model(tx, ty, nu=3.5, covar="rbf")
    assert model.kernel_id == 0
    BO.fit_gpytorch_model(mll, device=DEV)
    orc = ogp.bo_loop(lattice_grid, obj, X0, y0, 8, theta=model.theta(), kernel_id=0, tol=0.1)
    assert res_rbf.nit == orc["nit"] and np.array_equal(res_rbf.yall, orc["yall"])


def test_initial_sets_off_the_lattice(golden, lattice_grid):
    """Initial observations that are not candidate-lattice nodes have no entry in the kernel table: the engine must
    notice on the host and take the generic kernel (same picks as the oracle), never run the lattice kernels on them
    silently; a later reset() onto off-lattice points is flagged per trial (ST_OFF_LATTICE) and reported as failure."""
    import BO
    from b200bo import engine, ops
    T, n0, max_iter = 6, 10, 12
    grid, obj, surf_id, X0, y0 = _lattice100_case(golden, lattice_grid, T, n0, seed=55)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    X_off = X0.copy()
    X_off[::2, 3, 0] += 0.25                                          # every other trial has one off-node observation
    X_off[1, 7] = [99.5, 0.125]
    bo = engine.BatchedBO(_t(grid), _t(obj), _t(surf_id), _t(X_off), _t(y0), max_iter, theta=_t(theta), tol=0.1)
    assert bo.lattice is None
    res = bo.run()
    torch.cuda.synchronize()
    assert res.status.max().item() == 0
    for i in range(T):
        o = ogp.bo_loop(grid, obj[surf_id[i]], X_off[i], y0[i], max_iter, theta=theta, tol=0.1)
        assert int(res.nit[i]) == o["nit"] and np.array_equal(res.picked[i, :o["nit"]].cpu().numpy(), o["idx"])
        assert np.array_equal(res.yall[i].cpu().numpy(), o["yall"])
    with pytest.raises(ops._lib.B200boError):
        engine.BatchedBO(_t(grid), _t(obj), _t(surf_id), _t(X_off), _t(y0), max_iter, theta=_t(theta), posterior="incremental")
    on = engine.BatchedBO(_t(grid), _t(obj), _t(surf_id), _t(X0), _t(y0), max_iter, theta=_t(theta), tol=0.1)
    assert on.lattice is not None
    on.reset(_t(X_off), _t(y0))                                        # resident buffers, no host check: the kernels flag it
    r2 = on.run()
    torch.cuda.synchronize()
    st = r2.status.cpu().numpy()
    assert np.all((st[::2] & ops.ST_OFF_LATTICE) != 0) and (st[1] & ops.ST_OFF_LATTICE) and not (st[3] & ops.ST_OFF_LATTICE)
    assert BO._pack_result(r2, 0, n0, max_iter, False, 100.0, X_off[0]).success is False
    assert BO._pack_result(r2, 3, n0, max_iter, False, 100.0, X_off[3]).success is True
