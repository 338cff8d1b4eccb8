"""The helpers of the reference's ``Src/utils.py`` that sit on the result path."""
import numpy as np


def cumoptgap(data):
    """Area under optimality-gap curve(s) (utils.py:31-37: np.trapz, unit spacing); one curve or a list."""
    g = np.asarray(data, dtype=np.float64)
    if g.ndim == 1:
        return float(np.sum((g[1:] + g[:-1]) * 0.5))
    return np.sum((g[..., 1:] + g[..., :-1]) * 0.5, axis=-1)
