"""Export the diagonal each unit cell of the 101 x 101 Wildcat Wells node lattice is split along by
scipy's Delaunay triangulation (Qhull) — the triangulation behind the LinearNDInterpolator that the
reference's generate_cont() returns (Src/objectives.py:349-350).

The split depends on the node coordinates only (objectives.construct_X: integers 0..100 per axis), not on
the surface values or the seed, so one 10 000-entry mask covers every surface:
    mask[iy, jx] = 1  -> cell [jx, jx+1] x [iy, iy+1] is cut along (jx, iy)-(jx+1, iy+1)   ("main")
                   0  -> cut along (jx+1, iy)-(jx, iy+1)                                   ("anti")
Written bit-packed (numpy packbits, row-major, 1 250 bytes) to
    jmd-diversity-in-bayesian-optimization_b200/b200bo/data/qhull_diag_101.bin
which the product loads in ops.qhull_diag_mask(); tests/golden/make_golden.py stores the mask derived
from the reference's own interpolator object beside it so the two can be compared.

    python tools/make_diag_mask.py            # needs scipy only (Qhull options 'Qbb Qc Qz Q12', scipy default)
"""
import os
import sys

import numpy as np
from scipy.spatial import Delaunay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200", "b200bo", "data", "qhull_diag_101.bin")


def mask_from_simplices(points: np.ndarray, simplices: np.ndarray) -> np.ndarray:
    """(100,100) uint8 from a triangulation of the 101 x 101 unit lattice."""
    p = points[simplices]                                   # (S,3,2)
    lo = p.min(axis=1)                                      # lower-left corner of the triangle's cell
    assert np.all(p.max(axis=1) - lo == 1), "a triangle spans more than one unit cell"
    has00 = np.any(np.all(p == lo[:, None, :], axis=2), axis=1)
    has11 = np.any(np.all(p == lo[:, None, :] + 1, axis=2), axis=1)
    main = (has00 & has11).astype(np.int8)
    mask = -np.ones((100, 100), dtype=np.int8)
    jx, iy = lo[:, 0].astype(int), lo[:, 1].astype(int)
    for m, i, j in zip(main, iy, jx):
        assert mask[i, j] in (-1, m), "the two triangles of a cell disagree"
        mask[i, j] = m
    assert np.all(mask >= 0) and simplices.shape[0] == 20000
    return mask.astype(np.uint8)


def main():
    ax = np.arange(0, 101, 1).astype(np.float32)            # objectives.construct_X (objectives.py:85-104)
    X = np.stack(np.meshgrid(ax, ax), axis=-1).reshape(-1, 2)
    tri = Delaunay(X.astype(np.float64))
    mask = mask_from_simplices(tri.points, tri.simplices)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.packbits(mask.reshape(-1)).tofile(OUT)
    import scipy
    print(f"scipy {scipy.__version__}: main {int(mask.sum())} / anti {int((1 - mask).sum())} -> {OUT}")


if __name__ == "__main__":
    sys.exit(main())
