"""GPU parity: the batched BO loop (BOloop_fixparam semantics on a grid) vs the oracle loop —
selected indices bit-exact, best-so-far curves identical, padding as BO.py:262-271."""
import numpy as np
import pytest
import torch

from conftest import AUTHOR_THETA_RAW
from oracle import gp as ogp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(golden, lattice_grid, T, n0, seed):
    g = golden("wildcat.npz")
    surfs = np.stack([g["surf_0"], g["surf_1"], g["surf_2"], g["surf_3"]])
    obj = surfs[:, :100, :100].reshape(4, -1).astype(np.float32)      # objective at lattice nodes
    rng = np.random.default_rng(seed)
    surf_id = rng.integers(0, 4, T).astype(np.int32)
    idx0 = np.stack([rng.choice(10000, n0, replace=False) for _ in range(T)])
    X0 = lattice_grid[idx0]
    y0 = np.stack([obj[surf_id[t], idx0[t]] for t in range(T)]).astype(np.float64)
    return obj, surf_id, X0, y0


@pytest.mark.parametrize("append", [True, False])
@pytest.mark.parametrize("lattice", ["auto", None])
@pytest.mark.parametrize("kid,acq", [(3, 0), (0, 0), (3, 1)])
def test_fixed_theta_loop_matches_oracle(golden, lattice_grid, kid, acq, lattice, append):
    from b200bo import engine
    T, n0, max_iter = 12, 10, 25
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=kid * 10 + acq)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    t = lambda a, dt=None: torch.as_tensor(a, device=DEV, dtype=dt)
    bo = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), kernel_id=kid,
                          acq_id=acq, acq_param=4.0, tol=0.1, lattice=lattice, append=append)
    assert (bo.lattice is not None) == (lattice == "auto") and bo.append == append
    res = bo.run()
    torch.cuda.synchronize()
    yall = res.yall.cpu().numpy(); picked = res.picked.cpu().numpy(); nit = res.nit.cpu().numpy()
    assert yall.shape == (T, max_iter + 2)
    early = 0
    for i in range(T):
        o = ogp.bo_loop(lattice_grid, obj[surf_id[i]], X0[i], y0[i], max_iter, theta=theta, kernel_id=kid, acq_id=acq,
                        acq_param=4.0, tol=0.1)
        assert nit[i] == o["nit"], (i, nit[i], o["nit"])
        assert np.array_equal(picked[i, :nit[i]], o["idx"]), (i, picked[i, :nit[i]], o["idx"])
        assert np.all(picked[i, nit[i]:] == -1)
        assert np.array_equal(yall[i], o["yall"])
        assert res.n[i].item() == n0 + nit[i]
        assert np.array_equal(res.X[i, :n0 + nit[i]].cpu().numpy(), o["X"])
        early += nit[i] < max_iter
    assert res.status.max().item() == 0
    if acq == 0 and kid == 3:
        assert early > 0          # tol = 0.1 stops trials that reach the optimum (BO.py:152)


def test_tol_disabled_runs_all_iterations_and_never_repeats_a_point(golden, lattice_grid):
    from b200bo import engine
    T, n0, max_iter = 64, 10, 40
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=99)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    t = lambda a: torch.as_tensor(a, device=DEV)
    res = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0).run()
    assert torch.all(res.nit == max_iter)
    X = res.X.cpu().numpy()
    for i in range(T):
        assert len({tuple(p) for p in X[i]}) == n0 + max_iter          # observed candidates are masked
    yall = res.yall.cpu().numpy()
    assert np.all(np.diff(yall, axis=1) >= 0)                          # best-so-far is monotone
    assert np.array_equal(yall[:, -1], yall[:, -2])                    # final append (BO.py:262)


def test_reference_quirk_defaults_after_first_iteration(golden, lattice_grid):
    """BOloop_fixparam as coded (BO.py:392-395 vs :424-429): fixparam only for the first acquisition, gpytorch
    defaults (noise 2, c 0, s = l = log 2) afterwards -> near space-filling search."""
    from b200bo import engine
    T, n0, max_iter = 8, 10, 20
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=5)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    rest = np.array([2.0, 0.0, np.log(2.0), np.log(2.0)])
    t = lambda a: torch.as_tensor(a, device=DEV)
    bo = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), tol=0.1, theta_rest=t(rest))
    for rep in range(2):                                     # second pass exercises reset() with a theta switch
        if rep:
            bo.reset(t(X0), t(y0))
        res = bo.run()
        for i in range(T):
            o = ogp.bo_loop(lattice_grid, obj[surf_id[i]], X0[i], y0[i], max_iter, theta=theta, tol=0.1, theta_rest=rest)
            assert int(res.nit[i]) == o["nit"]
            assert np.array_equal(res.picked[i, :o["nit"]].cpu().numpy(), o["idx"])
            assert np.array_equal(res.yall[i].cpu().numpy(), o["yall"])


@pytest.mark.parametrize("kid,acq", [(3, 0), (0, 0), (3, 1), (2, 2)])
def test_incremental_posterior_loop_matches_oracle(golden, lattice_grid, kid, acq):
    """posterior='incremental' (csrc/grid_inc.cu): same picks, same curves as the oracle's full recomputation."""
    from b200bo import engine
    T, n0, max_iter = 12, 10, 25
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=kid * 10 + acq + 3)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    t = lambda a: torch.as_tensor(a, device=DEV)
    bo = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), kernel_id=kid,
                          acq_id=acq, acq_param=4.0, tol=0.1, posterior="incremental")
    for rep in range(2):
        if rep:
            bo.reset(t(X0), t(y0))
        res = bo.run()
        assert res.status.max().item() == 0
        for i in range(T):
            o = ogp.bo_loop(lattice_grid, obj[surf_id[i]], X0[i], y0[i], max_iter, theta=theta, kernel_id=kid, acq_id=acq,
                            acq_param=4.0, tol=0.1)
            assert int(res.nit[i]) == o["nit"], (i, int(res.nit[i]), o["nit"])
            assert np.array_equal(res.picked[i, :o["nit"]].cpu().numpy(), o["idx"]), i
            assert np.array_equal(res.yall[i].cpu().numpy(), o["yall"])


def test_incremental_equals_full_recompute_on_a_larger_batch(golden, lattice_grid):
    from b200bo import engine, ops
    T, n0, max_iter = 256, 10, 60
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=77)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    t = lambda a: torch.as_tensor(a, device=DEV)
    full = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0).run()
    bo = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0,
                          posterior="incremental")
    inc = bo.run()
    same = (full.picked == inc.picked).all(dim=1)
    # identical posterior up to summation order; an arg-max can only differ on a rounding-level near-tie
    assert same.float().mean().item() >= 0.99, same.float().mean().item()
    assert torch.equal(full.yall[same], inc.yall[same])
    # the state the kernel keeps is the posterior itself: compare with the full kernel's debug fields
    n = inc.n.clone()
    st = ops.gp_assemble_chol(inc.X, inc.y, n, bo.theta, 3)
    dbg = ops.grid_posterior_acq_argmax(t(lattice_grid), inc.X, st["wfrag"], st["alpha"], n, bo.theta, bo.best_f, 3, debug=True)
    # state holds rows 0..n-2 (the last appended point has not been folded in yet): rebuild with one more update
    bo._inc_update(n, False, True)
    mu = bo.theta[:, 1:2] + bo.mu_s
    var = bo.theta[:, 2:3] - bo.ss_s
    assert torch.allclose(mu, dbg["mu"], rtol=1e-9, atol=1e-9 * float(dbg["mu"].abs().max()))
    assert torch.allclose(var, dbg["var"], rtol=1e-8, atol=1e-10 * float(bo.theta[0, 2]))


@pytest.mark.parametrize("fit_mode", [False, True])
def test_cuda_graph_replay_equals_eager_run(golden, lattice_grid, fit_mode):
    """BatchedBO.run_graphed(): the captured launch sequence replayed on new initial sets gives what the eager loop
    gives (fixed theta with rank-1 appends, and fit-before-every-acquisition)."""
    from b200bo import engine, fit, ops
    dev = torch.device("cuda:0")
    surf = golden("wildcat.npz")["surf_2"][:100, :100].reshape(1, -1).astype(np.float32)
    rng = np.random.default_rng(11)
    T, n0, iters = 6, 10, 12
    grid = torch.as_tensor(lattice_grid, device=dev)
    obj = torch.as_tensor(surf, device=dev)
    sid = torch.zeros(T, dtype=torch.int32, device=dev)
    theta = None if fit_mode else torch.as_tensor(ogp.raw_to_theta(AUTHOR_THETA_RAW), device=dev)

    def sets():
        idx = np.stack([rng.choice(10000, n0, replace=False) for _ in range(T)])
        return torch.as_tensor(lattice_grid[idx], device=dev), torch.as_tensor(surf[0][idx].astype(np.float64), device=dev)

    X0, y0 = sets()
    mk = lambda: engine.BatchedBO(grid, obj, sid, X0, y0, iters, theta=theta, tol=-1.0, check_every=0,
                                  fitter=fit.LBFGSFitter(T, dev) if fit_mode else None)
    bo_g, bo_e = mk(), mk()
    for _ in range(3):                                   # first call captures, the next two replay on fresh inputs
        bo_g.reset(X0, y0); bo_e.reset(X0, y0)
        rg = bo_g.run_graphed()
        re_ = bo_e.run()
        torch.cuda.synchronize()
        assert torch.equal(rg.picked, re_.picked) and torch.equal(rg.yall, re_.yall) and torch.equal(rg.nit, re_.nit)
        X0, y0 = sets()


@pytest.mark.parametrize("fit_mode", [False, True])
def test_concurrent_groups_equal_one_batch(golden, lattice_grid, fit_mode):
    """BOloop_batched(groups=3): the batch cut into three groups advanced on separate CUDA streams returns, trial by
    trial, bit for bit what the single batch returns (trials are independent; engine.run_concurrent)."""
    import BO
    T, n0, max_iter = 23, 10, 12
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=5)
    theta = None if fit_mode else ogp.raw_to_theta(AUTHOR_THETA_RAW)
    a = BO.BOloop_batched(obj, surf_id, X0, y0, max_iter, theta=theta, tol=0.1, device=DEV)
    b = BO.BOloop_batched(obj, surf_id, X0, y0, max_iter, theta=theta, tol=0.1, device=DEV, groups=3)
    torch.cuda.synchronize()
    for name in ("yall", "picked", "nit", "status", "X", "y", "n"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    if fit_mode:
        assert torch.equal(a.thetas, b.thetas)


@pytest.mark.parametrize("kid,acq", [(3, 0), (0, 0), (3, 1), (2, 2), (1, 0)])
def test_persistent_loop_matches_oracle(golden, lattice_grid, kid, acq):
    """posterior='persistent' (csrc/bo_persist.cu: every iteration of a trial inside one CTA, one launch per batch):
    same picks, same curves, same observations as the oracle's full recomputation; reset() + second run identical."""
    from b200bo import engine
    T, n0, max_iter = 12, 10, 25
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=kid * 10 + acq + 5)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    t = lambda a: torch.as_tensor(a, device=DEV)
    bo = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), kernel_id=kid,
                          acq_id=acq, acq_param=4.0, tol=0.1, posterior="persistent")
    assert bo.persistent and bo.wfrag is None
    for rep in range(2):
        if rep:
            bo.reset(t(X0), t(y0))
        res = bo.run()
        torch.cuda.synchronize()
        assert res.status.max().item() == 0 and bo.launches == 1
        early = 0
        for i in range(T):
            o = ogp.bo_loop(lattice_grid, obj[surf_id[i]], X0[i], y0[i], max_iter, theta=theta, kernel_id=kid, acq_id=acq,
                            acq_param=4.0, tol=0.1)
            nit = int(res.nit[i])
            assert nit == o["nit"], (i, nit, o["nit"])
            assert np.array_equal(res.picked[i, :nit].cpu().numpy(), o["idx"]), i
            assert np.all(res.picked[i, nit:].cpu().numpy() == -1)
            assert np.array_equal(res.yall[i].cpu().numpy(), o["yall"])
            assert res.n[i].item() == n0 + nit
            assert np.array_equal(res.X[i, :n0 + nit].cpu().numpy(), o["X"])
            assert np.array_equal(res.y[i, :n0 + nit].cpu().numpy(), o["y"])
            early += nit < max_iter
        if acq == 0 and kid == 3:
            assert early > 0


def test_persistent_work_queue_and_table_switches_equal_full_recompute(golden, lattice_grid):
    """More trials than CTAs (the work queue hands every CTA several trials) with three hyper-parameter rows (a CTA
    reloads its shared-memory table when the slot changes): picks equal the full O(n^2) kernel's."""
    from b200bo import engine
    T, n0, max_iter = 400, 10, 40
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=123)
    base = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    rows = np.stack([base, base * np.array([2.0, 1.0, 0.7, 1.5]), base * np.array([0.5, 0.9, 1.3, 0.6])])
    theta = rows[np.random.default_rng(5).integers(0, 3, T)]
    t = lambda a: torch.as_tensor(a, device=DEV)
    full = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0).run()
    bo = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0,
                          posterior="auto")
    assert bo.persistent and bo.order is not None
    per = bo.run()
    torch.cuda.synchronize()
    assert per.status.max().item() == 0 and torch.all(per.nit == max_iter)
    same = (full.picked == per.picked).all(dim=1)
    # identical posterior up to summation order; an arg-max can only differ on a rounding-level near-tie
    assert same.float().mean().item() >= 0.99, same.float().mean().item()
    assert torch.equal(full.yall[same], per.yall[same]) and torch.equal(full.X[same], per.X[same])


def test_persistent_hands_non_positive_pivots_back_to_the_jitter_path(golden, lattice_grid):
    """A duplicated initial point with (almost) no noise makes the appended pivot non-positive: the persistent kernel
    flags the trial (ST_FALLBACK) and the engine re-runs it through the per-iteration kernels, whose Cholesky applies
    psd_safe_cholesky's jitter — same result as posterior='full' for those trials, untouched neighbours unaffected."""
    from b200bo import engine, ops
    T, n0, max_iter = 6, 10, 12
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=9)
    X0[1, 7] = X0[1, 2]; y0[1, 7] = y0[1, 2]
    X0[4, 9] = X0[4, 0]; y0[4, 9] = y0[4, 0]
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    theta = np.array([0.0, theta[1], theta[2], theta[3]])               # no noise: K is exactly singular for trials 1, 4
    t = lambda a: torch.as_tensor(a, device=DEV)
    full = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0).run()
    bo = engine.BatchedBO(t(lattice_grid), t(obj), t(surf_id), t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0,
                          posterior="persistent")
    per = bo.run()
    torch.cuda.synchronize()
    assert bo.fallback_trials == 2
    st = per.status.cpu().numpy()
    # (the full factorisation may see a rounding-level positive pivot where the rank-1 append sees <= 0: no jitter then)
    assert not np.any(st & ops.ST_FALLBACK) and np.any(st[[1, 4]] & ops.ST_JITTER)
    assert torch.equal(per.status[[1, 4]], full.status[[1, 4]])
    assert torch.equal(per.picked[[1, 4]], full.picked[[1, 4]]) and torch.equal(per.yall[[1, 4]], full.yall[[1, 4]])
    keep = [0, 2, 3, 5]
    assert np.all(st[keep] == 0) and torch.equal(per.picked[keep], full.picked[keep])


def test_auto_takes_the_incremental_path_on_a_lattice_too_large_for_shared_memory():
    """200 x 200 lattice, step 0.5 (BASELINE configs[3]): posterior='auto' = the O(n) incremental update with the state in
    HBM; same picks as the full recomputation, and a trial whose rank-1 append meets a non-positive pivot (duplicated
    point, no noise) is re-run through the per-iteration factorisation with psd_safe_cholesky's jitter rule."""
    from b200bo import engine, ops
    lat = ops.Lattice(200, 200, 0.0, 0.0, 0.5)
    pts = lat.points(DEV)
    ph = pts.cpu().numpy()
    obj = (50 + 30 * np.sin(ph[:, 0] / 7.0) * np.cos(ph[:, 1] / 11.0) + 0.1 * ph[:, 0]).astype(np.float32)[None, :]
    rng = np.random.default_rng(2)
    T, n0, max_iter = 6, 10, 14
    idx0 = np.stack([rng.choice(40000, n0, replace=False) for _ in range(T)])
    idx0[3, 9] = idx0[3, 1]                                              # exactly singular K for trial 3
    X0, y0 = ph[idx0], obj[0][idx0].astype(np.float64)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    theta = np.array([0.0, theta[1], theta[2], theta[3]])
    t = lambda a: torch.as_tensor(a, device=DEV)
    sid = torch.zeros(T, dtype=torch.int32, device=DEV)
    full = engine.BatchedBO(pts, t(obj), sid, t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0, lattice=lat).run()
    bo = engine.BatchedBO(pts, t(obj), sid, t(X0), t(y0), max_iter, theta=t(theta), tol=-1.0, lattice=lat, posterior="auto")
    assert bo.incremental and not bo.persistent
    res = bo.run()
    torch.cuda.synchronize()
    assert bo.fallback_trials == 1
    assert torch.equal(res.status, full.status) and torch.equal(res.picked, full.picked) and torch.equal(res.yall, full.yall)
    assert torch.equal(res.nit, full.nit) and torch.equal(res.X, full.X)
