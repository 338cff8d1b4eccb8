"""Multi-GPU form of the batched BO entry on ONE GPU: whatever the split — fake ranks, several devices of
one process — the result is the single-batch result, trial for trial and bit for bit."""
import numpy as np
import pytest
import torch

from conftest import AUTHOR_THETA_RAW
from oracle import gp as ogp
from test_gpu_bo_loop import _setup

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FIELDS = ("yall", "picked", "nit", "status", "X", "y", "n")


def _theta_rows(T):
    base = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    rows = np.stack([base, base * np.array([2.0, 1.0, 0.7, 1.5])])
    return rows[np.random.default_rng(1).integers(0, 2, T)]


@pytest.mark.parametrize("posterior", ["auto", "full"])
@pytest.mark.parametrize("world", [2, 3])
def test_fake_rank_shards_equal_the_single_batch(golden, lattice_grid, world, posterior):
    import BO
    from b200bo import dist
    T, n0, max_iter = 23, 10, 18
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=77)
    theta = _theta_rows(T)
    whole = BO.BOloop_batched(obj, surf_id, X0, y0, max_iter, theta=theta, tol=0.1, device=DEV, posterior=posterior)
    parts = []
    for rank in range(world):
        lo, hi = dist.shard_range(T, rank, world)
        parts.append(BO.BOloop_batched(obj, surf_id[lo:hi], X0[lo:hi], y0[lo:hi], max_iter, theta=theta[lo:hi], tol=0.1,
                                       device=DEV, posterior=posterior))
    assert sum(p.nit.shape[0] for p in parts) == T and dist.shard_counts(T, world) == [p.nit.shape[0] for p in parts]
    for f in FIELDS:
        assert torch.equal(torch.cat([getattr(p, f) for p in parts]), getattr(whole, f)), f
    # devices="ranks" in a single process (world 1) is the plain call
    one = BO.BOloop_batched(obj, surf_id, X0, y0, max_iter, theta=theta, tol=0.1, device=DEV, posterior=posterior,
                            devices="ranks")
    for f in FIELDS:
        assert torch.equal(getattr(one, f), getattr(whole, f)), f


@pytest.mark.parametrize("fit_mode", [False, True])
def test_device_list_equals_the_single_batch(golden, lattice_grid, fit_mode):
    """devices=[...]: contiguous ranges driven from one host thread, one stream per entry (here the same GPU twice /
    three times — the plumbing is identical for distinct GPUs)."""
    import BO
    T, n0, max_iter = (9, 10, 6) if fit_mode else (31, 10, 20)
    obj, surf_id, X0, y0 = _setup(golden, lattice_grid, T, n0, seed=5)
    theta = None if fit_mode else _theta_rows(T)
    whole = BO.BOloop_batched(obj, surf_id, X0, y0, max_iter, theta=theta, tol=0.1, device=DEV, groups=1)
    for devs in ([DEV, DEV], [DEV, DEV, DEV]):
        split = BO.BOloop_batched(obj, surf_id, X0, y0, max_iter, theta=theta, tol=0.1, devices=devs, groups=1)
        for f in FIELDS:
            assert torch.equal(getattr(split, f), getattr(whole, f)), f
        if fit_mode:
            assert torch.equal(split.thetas, whole.thetas)
