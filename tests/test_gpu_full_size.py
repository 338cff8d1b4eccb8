"""BASELINE.json full sizes through size-independent properties (the oracle needs ~0.5 s per 100-iteration trial, so at
12 500 trials — one GPU's share of configs[4] at N = 8 — parity is checked by invariants every correct run satisfies,
by determinism, by the O(n^2) kernel on a slice and by the oracle on a handful of trials)."""
import numpy as np
import pytest
import torch

from conftest import AUTHOR_THETA_RAW
from oracle import gp as ogp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_configs4_shard_properties(golden, lattice_grid):
    import BO
    from b200bo import engine
    g = golden("wildcat.npz")
    obj = np.stack([g["surf_0"], g["surf_1"], g["surf_2"], g["surf_3"]])[:, :100, :100].reshape(4, -1).astype(np.float32)
    T, n0, iters = 12500, 10, 100
    rng = np.random.default_rng(2024)
    idx0 = rng.integers(0, 10000, size=(T, n0))
    bad = np.nonzero((np.diff(np.sort(idx0, axis=1), axis=1) == 0).any(axis=1))[0]
    for b in bad:
        idx0[b] = rng.choice(10000, n0, replace=False)
    sid = (np.arange(T) % 4).astype(np.int32)
    X0 = lattice_grid[idx0]
    y0 = obj[sid[:, None], idx0].astype(np.float64)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    res = BO.BOloop_batched(obj, sid, X0, y0, iters, theta=theta, tol=-1.0, device=DEV)          # "auto" = K-F
    torch.cuda.synchronize()
    nit, st = res.nit.cpu().numpy(), res.status.cpu().numpy()
    picked, yall = res.picked.cpu().numpy(), res.yall.cpu().numpy()
    assert np.all(st == 0) and np.all(nit == iters) and np.all(res.n.cpu().numpy() == n0 + iters)
    # every pick is a fresh candidate: never an initial point, never picked twice
    allidx = np.concatenate([idx0, picked], axis=1)
    assert np.all(picked >= 0) and np.all(picked < 10000)
    assert np.all(np.diff(np.sort(allidx, axis=1), axis=1) > 0)
    # the curve is recomputable from the picks and the objective table: yall[0] = best initial value,
    # yall[k+1] = max(yall[k], f(pick k)), one padding entry (BO.py:146,168-169,262-271)
    vals = obj[sid[:, None], picked].astype(np.float64)
    ref = np.maximum.accumulate(np.concatenate([y0.max(axis=1, keepdims=True), vals], axis=1), axis=1)
    assert np.array_equal(yall[:, :iters + 1], ref) and np.array_equal(yall[:, iters + 1], yall[:, iters])
    # the observation arrays hold the initial set followed by the picked lattice nodes and their values
    assert np.array_equal(res.X[:, n0:].cpu().numpy(), lattice_grid[picked])
    assert np.array_equal(res.y[:, n0:].cpu().numpy(), vals) and np.array_equal(res.y[:, :n0].cpu().numpy(), y0)
    # determinism: a second run is bit-identical
    again = BO.BOloop_batched(obj, sid, X0, y0, iters, theta=theta, tol=-1.0, device=DEV)
    assert torch.equal(again.picked, res.picked) and torch.equal(again.yall, res.yall)
    # the O(n^2) kernel on a slice: identical pick sequences
    sl = slice(0, 296)
    full = BO.BOloop_batched(obj, sid[sl], X0[sl], y0[sl], iters, theta=theta, tol=-1.0, device=DEV, posterior="full")
    assert torch.equal(full.picked, res.picked[sl]) and torch.equal(full.yall, res.yall[sl])
    # the oracle on a handful
    for i in (0, 4321, 12499):
        o = ogp.bo_loop(lattice_grid, obj[sid[i]], X0[i], y0[i], iters, theta=theta, kernel_id=3, tol=-1.0)
        assert np.array_equal(picked[i], o["idx"]) and np.array_equal(yall[i], o["yall"])


def test_config2_size_dpp_properties():
    """100 000 ten-point sets (configs[1]): idempotence of the scorer, permutation invariance of a set's score within the
    conditioning bound, sortedness and permutation property of the device ranking, z-scores of mean 0 / sd 1."""
    from b200bo import ops
    lat = ops.Lattice(100, 100)
    pool = lat.points(DEV)
    rng = np.random.default_rng(11)
    idx = rng.integers(0, 10000, size=(100000, 10))
    bad = np.nonzero((np.diff(np.sort(idx, axis=1), axis=1) == 0).any(axis=1))[0]
    for b in bad:
        idx[b] = rng.choice(10000, 10, replace=False)
    ti = torch.as_tensor(idx.astype(np.int32), device=DEV)
    ld = ops.dpp_logdet_score(pool, ti, 1e-5)
    assert torch.equal(ld, ops.dpp_logdet_score(pool, ti, 1e-5))
    assert torch.equal(ld, ops.dpp_logdet_score_lattice(lat, ops.dpp_table(lat, 1e-5, DEV), ti))
    perm = torch.as_tensor(np.stack([rng.permutation(10) for _ in range(100000)]), device=DEV)
    ld_p = ops.dpp_logdet_score(pool, torch.gather(ti, 1, perm).contiguous(), 1e-5)
    assert torch.all((ld - ld_p).abs() <= 1e-4 * ld.abs())                  # cond(S) up to ~1e11 at gamma = 1e-5
    stats, z = ops.dpp_zscore(ld)
    assert abs(float(z.mean())) < 1e-9 and abs(float(z.std(unbiased=False)) - 1.0) < 1e-9
    zs, order, ps, dist = ops.dpp_rank(z, ti, pool)
    assert torch.all(zs[1:] >= zs[:-1])
    assert torch.equal(torch.sort(order.long())[0], torch.arange(100000, device=DEV))
    assert torch.equal(zs, z[order.long()]) and torch.equal(ps, ti[order.long()]) and torch.equal(dist, pool[ps.long()])
