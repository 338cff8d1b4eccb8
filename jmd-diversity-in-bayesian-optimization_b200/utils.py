"""The helpers of the reference's ``Src/utils.py`` that sit on the result path."""
import random

import numpy as np


def cumoptgap(data):
    """Area under optimality-gap curve(s) (utils.py:31-37: np.trapz, unit spacing); one curve or a list."""
    g = np.asarray(data, dtype=np.float64)
    if g.ndim == 1:
        return float(np.sum((g[1:] + g[:-1]) * 0.5))
    return np.sum((g[..., 1:] + g[..., :-1]) * 0.5, axis=-1)


def bootstrap(data, trials=1000, device="cuda:0"):
    """Bootstrap mean and 2.5 / 97.5 percentile band of an (N, D) matrix of N trials, or of an (N,) vector
    (utils.py:39-115; callers results.py:172-184,787-802).  Returns ``boot_ci`` = [mean, 2.5th, 97.5th percentile] of
    the resampled means, shape (3, D) or (3,), as the reference does.

    The resampled row indices are drawn here from CPython's global generator exactly as ``random.choices`` does
    (``floor(random() * N)``, N draws per replicate, `trials` replicates in order), so after ``random.seed(s)`` the
    replicates are the reference's; the gathered means, the sort and the percentiles run in CUDA kernels
    (csrc/bootstrap.cu) with numpy's summation order, so the result is bit-identical to the reference's.

    Not reproduced: the reference's ``data.ndim > 2`` branch (it raises TypeError at ``range(None)``) and ragged
    object arrays (``np.array`` of ragged lists is an error since numpy 1.24, the version the reference pins)."""
    import torch
    from b200bo import ops
    if isinstance(data, list):
        data = np.array(data)
    data = np.asarray(data)
    if data.ndim > 2:
        raise TypeError("'NoneType' object cannot be interpreted as an integer")      # utils.py:77 with D = None
    if data.dtype == object:
        raise NotImplementedError("ragged inputs are outside the hot path")
    N = len(data)
    rnd = random.random
    idx = np.array([[int(rnd() * N) for _ in range(N)] for _ in range(trials)], dtype=np.int32).reshape(trials, N)
    mat = np.ascontiguousarray(data.reshape(N, -1), dtype=np.float64)
    out = ops.bootstrap_ci(torch.as_tensor(mat, device=device), torch.as_tensor(idx, device=device)).cpu().numpy()
    return out if data.ndim == 2 else out[:, 0]
