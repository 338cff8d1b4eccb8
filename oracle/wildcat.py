"""Oracle: Wildcat Wells surface and N-D simplex noise (d = 2), numpy fp64.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates

* ``SimpleNoise.randHash``            /root/reference/Src/SimpleNoise.py:205-220
* ``SimpleNoise.SimplexNoise.__init__`` SimpleNoise.py:12-81  (constants, shuffled gradient table)
* ``SimplexNoise.simplexNoise``       SimpleNoise.py:85-172
* ``objectives.construct_X``          /root/reference/Src/objectives.py:85-104
* ``objectives.wildcatwells``         objectives.py:107-218 (N = 1 peak, d = 2)
* ``objectives.noise_based_cont``     objectives.py:331-351 (lattice-node look-up only)

The scalar functions mirror the reference statement by statement; the ``*_grid``
functions are vectorised numpy versions of exactly the same arithmetic (same
operation order per element) used to build whole surfaces quickly.
"""
from __future__ import annotations

import math
import random

import numpy as np

# --------------------------------------------------------------------------- hash

def rand_hash(v: int) -> int:
    """SimpleNoise.py:205-220 — Jenkins one-at-a-time over the low three bytes,
    *without* any word-size masking (Python ints).  For 0 <= v the result is
    < 2**58, so uint64 arithmetic reproduces it exactly."""
    h = 255 & v
    h += h << 10
    h ^= h >> 6
    h += 255 & (v >> 8)
    h += h << 10
    h ^= h >> 6
    h += 255 & (v >> 16)
    h += h << 10
    h ^= h >> 6
    h += h << 3
    h ^= h >> 11
    return h + (h << 15)


def rand_hash_u64(v: np.ndarray) -> np.ndarray:
    """Vectorised ``rand_hash`` on uint64 (wrap-around never triggers for v >= 0,
    asserted by tests/test_oracle_wildcat.py against the Python-int version)."""
    v = v.astype(np.uint64)
    m = np.uint64(255)
    h = v & m
    h = h + (h << np.uint64(10))
    h = h ^ (h >> np.uint64(6))
    h = h + ((v >> np.uint64(8)) & m)
    h = h + (h << np.uint64(10))
    h = h ^ (h >> np.uint64(6))
    h = h + ((v >> np.uint64(16)) & m)
    h = h + (h << np.uint64(10))
    h = h ^ (h >> np.uint64(6))
    h = h + (h << np.uint64(3))
    h = h ^ (h >> np.uint64(11))
    return h + (h << np.uint64(15))


# --------------------------------------------------------------------------- noise

class SimplexNoise2D:
    """Constants and gradient table of ``SimplexNoise(seed, 2)`` (SimpleNoise.py:12-81)."""

    def __init__(self, seed: int):
        d = 2
        self.d = d
        self.seed = rand_hash(seed)                       # :20
        self.f = ((d + 1) ** .5 - 1) / d                 # :24
        self.g = self.f / (1 + d * self.f)               # :31
        side = d ** .5 / (d * self.f + 1)                 # :34
        a = (side ** 2 - (side / 2) ** 2) ** .5           # :35
        self.corner_to_face_sq = a ** 2                   # :38,41  (d == 2)
        # :52-59: vf(1) yields [-1],[1]; z=0 -> [0,-1],[0,1]; z=1 -> [-1,0],[1,0]
        vecs = [[0, -1], [0, 1], [-1, 0], [1, 0]]
        self.value_scaler = (d - 1) ** -.5                # :65
        self.value_scaler *= ((d - 1) ** -3.5 * 100 + 13)  # :70  -> 113.0
        r = random.Random()
        r.seed(self.seed)                                  # :81-82
        r.shuffle(vecs)                                    # :83
        self.vecs = vecs
        self.vec_count = len(vecs)

    # -- scalar, statement-by-statement (SimpleNoise.py:85-172) ------------------
    def noise(self, loc) -> float:
        d = self.d
        s = sum(loc) * self.f
        isl = [int(math.floor(v + s)) for v in loc]
        t = sum(isl) * self.g
        cell = [loc[i] - isl[i] + t for i in range(d)]
        order = [v[1] for v in sorted(zip(cell, range(d)))]   # sortWith :222-225
        order.reverse()
        il = list(isl)
        n = 0.0
        skew = 0.0
        for v in [-1] + order:
            if v != -1:
                il[v] += 1
            u = [cell[i] - (il[i] - isl[i]) + skew for i in range(d)]
            t = self.corner_to_face_sq
            for a in u:
                t -= a * a
            if t > 0:
                index = self.seed
                for i in range(d):
                    index += il[i] << ((5 * i) % 16)
                vec = self.vecs[rand_hash(index) % self.vec_count]
                gr = 0
                for i in range(d):
                    gr += vec[i] * u[i]
                t4 = t ** 4
                n += gr * t4
            skew += self.g
        return n * self.value_scaler

    # -- vectorised, identical arithmetic per element ---------------------------
    def noise_grid(self, x: np.ndarray, y: np.ndarray) -> np.ndarray:
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        f, g = self.f, self.g
        s = (x + y) * f                                   # sum(loc) = 0 + x + y
        ix = np.floor(x + s).astype(np.int64)
        iy = np.floor(y + s).astype(np.int64)
        t0 = (ix + iy) * g
        cx = x - ix + t0
        cy = y - iy + t0
        # sorted(zip(cell, idx)) ascending then reversed: first step is the axis
        # with the larger cell value; on equality (cx == cy) tuple order puts
        # (c,0) before (c,1), reversed -> axis 1 first.
        x_first = cx > cy
        vec = np.asarray(self.vecs, dtype=np.float64)
        n = np.zeros_like(x)
        ox = np.zeros_like(ix)
        oy = np.zeros_like(iy)
        skew = 0.0
        for step in range(3):
            if step == 1:
                ox = ox + x_first
                oy = oy + (~x_first)
            elif step == 2:
                ox = ox + (~x_first)
                oy = oy + x_first
            ux = cx - ox + skew
            uy = cy - oy + skew
            t = self.corner_to_face_sq
            t = t - ux * ux
            t = t - uy * uy
            index = np.uint64(self.seed) + (ix + ox).astype(np.uint64) + ((iy + oy).astype(np.uint64) << np.uint64(5))
            vi = (rand_hash_u64(index) % np.uint64(self.vec_count)).astype(np.int64)
            gr = 0 + vec[vi, 0] * ux
            gr = gr + vec[vi, 1] * uy
            t4 = _pow4(t)
            n = np.where(t > 0, n + gr * t4, n)
            skew += g
        return n * self.value_scaler


def _pow4(t: np.ndarray) -> np.ndarray:
    """``t**4`` as CPython evaluates it for floats: C ``pow(t, 4.0)`` (glibc, < 1 ULP
    and in practice correctly rounded) — *not* ``(t*t)*(t*t)`` and not numpy's SIMD
    ``np.power`` (both differ in the last bit for a few % of inputs)."""
    flat = np.asarray(t, dtype=np.float64).ravel()
    out = np.fromiter((math.pow(v, 4.0) for v in flat.tolist()), dtype=np.float64, count=flat.size)
    return out.reshape(np.shape(t))


# --------------------------------------------------------------------------- surface

def construct_X(bounds=((0, 100), (0, 100)), budget=None, with_noise_coords=False):
    """objectives.py:85-104."""
    if budget is None:
        budget = 10000
    nd = len(bounds)
    dim_budget = ((budget ** (1 / nd)) // 2 + 1) * 2
    rr = max([abs(int(np.log(np.ptp(b) / dim_budget))) for b in bounds] + [1])
    res = [np.round(np.ptp(b) / dim_budget, rr) for b in bounds]
    axes = [np.arange(b[0], b[1] + res[i], res[i]).astype(np.float32) for i, b in enumerate(bounds)]
    X = np.stack(np.meshgrid(*axes), axis=-1)
    if with_noise_coords:
        noise_axes = [np.arange(0, 100, 100 / len(axes[i])) for i in range(nd)]
        return X, np.stack(np.meshgrid(*noise_axes), axis=-1)
    return X


def peak_location(seed: int, bounds=((0, 100), (0, 100))):
    """gen_ND_rand + gen_loc_list with N = 1 (objectives.py:122-149,179): the
    first of 100 uniforms drawn per dimension, x block first."""
    st = np.random.RandomState(seed)      # == np.random.seed(seed); np.random.rand
    loc = []
    for b in bounds:
        r = (b[1] - b[0]) * st.rand(100, 1) + b[0]
        loc.append(r[0, 0])
    return tuple(loc)


def wildcat_terms(seed: int, smoothness: float, bounds=((0, 100), (0, 100))):
    """The two fp64 fields the surface is made of: Gaussian pdf and A = (noise+1)·N
    (objectives.py:171-211), N = 1."""
    X, NC = construct_X(bounds, with_noise_coords=True)
    px, py = peak_location(seed, bounds)
    gridsize = np.average(bounds, axis=1)                  # :176  -> [50, 50]
    var = (0.2 * gridsize[0]) ** len(bounds)               # :185  -> 100
    dx = X[..., 0].astype(np.float64) - px
    dy = X[..., 1].astype(np.float64) - py
    # scipy multivariate_normal.pdf with cov = var*I:
    #   exp(-0.5*(d*log(2pi) + log|cov| + maha));  maha = (dx^2+dy^2)/var
    maha = (dx * dx + dy * dy) / var
    gaus = np.exp(-0.5 * (2 * np.log(2 * np.pi) + 2 * np.log(var) + maha))
    ng = SimplexNoise2D(seed)
    fs = 100 * (1 - (1 - smoothness) ** 2)                 # :191
    nx, ny = NC[..., 0], NC[..., 1]
    v = ng.noise_grid(nx / fs, ny / fs) \
        + ng.noise_grid(nx / (fs / 4), ny / (fs / 4)) / 4 \
        + ng.noise_grid(nx / (fs / 8), ny / (fs / 8)) / 8   # :167-168
    A = (v + 1) * 1                                        # :210
    return X, gaus, A


def wildcatwells(seed: int, smoothness: float, rug_amp: float, bounds=((0, 100), (0, 100))):
    """objectives.py:107-218 for args {N:1, Smoothness, rug_freq:any, rug_amp}.
    Returns (X (101,101,2) f32, surf (101,101) f32)."""
    X, gaus, A = wildcat_terms(seed, smoothness, bounds)
    power = -math.log(A.max() / gaus.max()) / math.log(10)   # :213
    surf = gaus + A * 10 ** (power - (1 - rug_amp) ** 5)      # :214
    surf = (surf / surf.flatten().max()) * 100               # :215
    return X, surf.astype(np.float32)


def lattice_lookup(surf: np.ndarray, xy: np.ndarray) -> np.ndarray:
    """Value of ``generate_cont()`` (objectives.py:322-351) at lattice nodes: the
    piecewise-linear interpolant equals the node value there, bit for bit
    (SURVEY.md Appendix C).  ``xy[:,0]`` is the column (x), ``xy[:,1]`` the row."""
    xy = np.asarray(xy)
    ix = np.rint(xy[..., 0]).astype(np.int64)
    iy = np.rint(xy[..., 1]).astype(np.int64)
    if not (np.all(ix == xy[..., 0]) and np.all(iy == xy[..., 1])):
        raise ValueError("lattice_lookup only defined on lattice nodes")
    return surf[iy, ix]


def diag_mask_from_triangulation(points: np.ndarray, simplices: np.ndarray) -> np.ndarray:
    """(100,100) uint8 from the Delaunay triangulation held by the reference's interpolator
    (``generate_cont().tri``, objectives.py:349-350): 1 where cell [jx,jx+1]x[iy,iy+1] is cut along
    (jx,iy)-(jx+1,iy+1), 0 along the other diagonal."""
    p = np.asarray(points)[np.asarray(simplices)]
    lo = p.min(axis=1)
    if not np.all(p.max(axis=1) - lo == 1):
        raise ValueError("not a triangulation of the unit lattice")
    has00 = np.any(np.all(p == lo[:, None, :], axis=2), axis=1)
    has11 = np.any(np.all(p == lo[:, None, :] + 1, axis=2), axis=1)
    mask = np.zeros((100, 100), dtype=np.uint8)
    seen = np.zeros((100, 100), dtype=np.int32)
    jx, iy = lo[:, 0].astype(int), lo[:, 1].astype(int)
    np.add.at(mask, (iy, jx), (has00 & has11).astype(np.uint8))
    np.add.at(seen, (iy, jx), 1)
    if not (np.all(seen == 2) and np.all((mask == 0) | (mask == 2))):
        raise ValueError("every cell must consist of two triangles sharing one diagonal")
    return (mask // 2).astype(np.uint8)


def interp_linear(surf: np.ndarray, xy: np.ndarray, diag_mask=None) -> np.ndarray:
    """``generate_cont()(xy)`` (objectives.py:322-351): scipy LinearNDInterpolator over the 101 x 101 nodes,
    restated — find the unit cell, pick the triangle of the Delaunay split (``diag_mask`` (100,100), 1 = main
    diagonal; None = main everywhere), value = sum of barycentric weight x node value (scipy interpnd.pyx);
    NaN outside [0,100]^2. ``xy[:,0]`` is x (column), ``xy[:,1]`` y (row)."""
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    s = np.asarray(surf, dtype=np.float64)
    x, y = xy[:, 0], xy[:, 1]
    inside = (x >= 0) & (x <= 100) & (y >= 0) & (y <= 100)
    xs, ys = np.where(inside, x, 0.0), np.where(inside, y, 0.0)
    jx = np.minimum(np.floor(xs).astype(np.int64), 99)
    iy = np.minimum(np.floor(ys).astype(np.int64), 99)
    u, v = xs - jx, ys - iy
    v00, v10, v01, v11 = s[iy, jx], s[iy, jx + 1], s[iy + 1, jx], s[iy + 1, jx + 1]
    main = np.ones(len(x), dtype=bool) if diag_mask is None else (np.asarray(diag_mask).reshape(100, 100)[iy, jx] != 0)
    lower = u >= v
    below = u + v <= 1.0
    out = np.where(main,
                   np.where(lower, (1.0 - u) * v00 + (u - v) * v10 + v * v11,
                            (1.0 - v) * v00 + (v - u) * v01 + u * v11),
                   np.where(below, ((1.0 - u) - v) * v00 + u * v10 + v * v01,
                            ((u - 1.0) + v) * v11 + (1.0 - u) * v01 + (1.0 - v) * v10))
    return np.where(inside, out, np.nan)
