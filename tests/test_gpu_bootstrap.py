"""utils.bootstrap on the GPU (csrc/bootstrap.cu) vs the reference's own outputs and the numpy oracle: bit-exact."""
import random

import numpy as np
import pytest
import torch

from oracle import bootstrap as ob

pytestmark = pytest.mark.gpu


def test_bootstrap_matches_reference_golden(golden):
    import utils
    g = golden("bootstrap.npz")
    for name in ("curves", "big", "vec"):
        trials, seed = (int(v) for v in g[f"{name}_trials"])
        random.seed(seed)
        data = g[f"{name}_data"]
        ci = utils.bootstrap([row for row in data] if name == "curves" else data, trials=trials)
        assert ci.shape == g[f"{name}_ci"].shape
        assert np.array_equal(ci, g[f"{name}_ci"]), name


def test_bootstrap_full_size_against_oracle():
    """results.py:787-794 shape: 1000 replicates of (trials x 102) gap curves; plus odd sizes around numpy's pairwise
    block boundaries (8, 128) and a constant column (lerp of equal neighbours)."""
    from b200bo import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    for N, C, B in ((1000, 102, 1000), (7, 3, 5), (8, 1, 9), (129, 4, 33), (257, 2, 1)):
        data = np.maximum.accumulate(rng.uniform(0, 100, size=(N, C)), axis=1)[:, ::-1].copy()
        data[:, -1] = 0.25
        random.seed(N)
        idx = ob.draw_indices(N, B)
        out = ops.bootstrap_ci(torch.as_tensor(data, device=dev), torch.as_tensor(idx, device=dev)).cpu().numpy()
        if N * C * B <= 2_000_000:
            ref = ob.bootstrap_from_indices(data, idx)
        else:                                       # full size: vectorised statement of the same sums
            boot = np.stack([np.array([np.sum(np.ascontiguousarray(data[row, c])) for c in range(C)]) / N for row in idx])
            ref = np.stack([np.array([np.sum(np.ascontiguousarray(boot[:, c])) for c in range(C)]) / B,
                            np.percentile(boot, 2.5, axis=0), np.percentile(boot, 97.5, axis=0)])
        assert np.array_equal(out, ref), (N, C, B)


def test_bootstrap_argument_errors():
    from b200bo import ops, _lib
    dev = torch.device("cuda:0")
    data = torch.zeros((4, 2), dtype=torch.float64, device=dev)
    with pytest.raises(ValueError):
        ops.bootstrap_ci(data, torch.zeros((3, 5), dtype=torch.int32, device=dev))
    with pytest.raises(_lib.B200boError):
        ops.bootstrap_ci(data.cpu(), torch.zeros((3, 4), dtype=torch.int32, device=dev))
    with pytest.raises(_lib.B200boError):          # B beyond the in-shared-memory sort
        ops.bootstrap_ci(data, torch.zeros((20000, 4), dtype=torch.int32, device=dev))
    nan = data.clone(); nan[1, 0] = float("nan")
    out = ops.bootstrap_ci(nan, torch.tensor([[0, 1, 2, 3], [0, 0, 0, 0]], dtype=torch.int32, device=dev)).cpu().numpy()
    assert np.isnan(out[:, 0]).all() and np.all(out[:, 1] == 0)
