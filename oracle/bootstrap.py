"""ORACLE (test infrastructure only — never imported by the product).

CPU restatement of ``utils.bootstrap`` (/root/reference/Src/utils.py:39-115) for equal-length rows:
`trials` replicates, each the position-wise ``np.mean`` of N rows drawn with ``random.choices``; the result is
[np.mean, np.percentile(2.5), np.percentile(97.5)] over the replicates.  Pinned bit-exactly against the reference's
own function (tests/golden/bootstrap.npz from tests/golden/make_golden.py; live in tests/test_reference_live.py).
"""
import random

import numpy as np


def draw_indices(N, trials):
    """random.choices(population, k=N) consumes one random() per draw: floor(random() * N) (Lib/random.py)."""
    return np.array([[int(random.random() * N) for _ in range(N)] for _ in range(trials)], dtype=np.int32).reshape(trials, N)


def bootstrap_from_indices(data, idx):
    data = np.asarray(data, dtype=np.float64)
    vec = data.ndim == 1
    mat = data.reshape(len(data), -1)
    C = mat.shape[1]
    # utils.py:59-66: np.mean over a Python LIST of the drawn values at one position (1-D, pairwise summation)
    boot = np.array([[np.mean(list(mat[row, c])) for c in range(C)] for row in idx])
    ci = np.array([[np.mean(list(boot[:, c])) for c in range(C)],
                   [np.percentile(list(boot[:, c]), 2.5) for c in range(C)],
                   [np.percentile(list(boot[:, c]), 97.5) for c in range(C)]])
    return ci[:, 0] if vec else ci


def bootstrap(data, trials=1000):
    data = np.asarray(data)
    return bootstrap_from_indices(data, draw_indices(len(data), trials))
