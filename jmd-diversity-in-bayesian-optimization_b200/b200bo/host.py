"""Host-side pieces of the reference's algorithms that are *not* arithmetic hot spots but decide
which inputs the kernels see: hashed seeds and shuffled gradient tables of SimplexNoise, the peak
location of a Wildcat Wells instance, and the big-integer RNG stream of the ranked-DPP sampler.
They use CPython's / numpy's own generators because parity of selected sets depends on them.
"""
from __future__ import annotations

import math
import random

import numpy as np


# ---------------------------------------------------------------- SimplexNoise set-up (SimpleNoise.py:12-83,205-220)

def rand_hash(v: int) -> int:
    """One-at-a-time style hash over the low three bytes, unbounded ints (SimpleNoise.py:205-220)."""
    h = v & 255
    for shift in (8, 16, None):
        h += h << 10
        h ^= h >> 6
        if shift is not None:
            h += (v >> shift) & 255
    h += h << 3
    h ^= h >> 11
    return h + (h << 15)


def simplex_setup(seed: int):
    """(hashed seed, 4x2 gradient table) of ``SimplexNoise(seed, 2)``: the table is the four axis
    vectors shuffled by ``random.Random(hashed seed)`` (SimpleNoise.py:52-59,81-83)."""
    hseed = rand_hash(int(seed))
    vecs = [[0, -1], [0, 1], [-1, 0], [1, 0]]
    rnd = random.Random()
    rnd.seed(hseed)
    rnd.shuffle(vecs)
    if hseed >= 1 << 63:
        raise ValueError("hashed seed does not fit the kernel's 64-bit arithmetic")
    return hseed, vecs


def wildcat_peak(seed, bounds=((0, 100), (0, 100))):
    """Peak of the single Gaussian (objectives.py:139-149,179 with N = 1): the first of 100 uniforms
    drawn per dimension after ``np.random.seed(seed)``."""
    st = np.random.RandomState(seed)
    out = []
    for lo, hi in bounds:
        out.append(float(((hi - lo) * st.rand(100, 1) + lo)[0, 0]))
    return tuple(out)


def axis_grid(bounds, budget, endpoint: bool, dtype):
    """Per-dimension node coordinates as objectives.construct_X (objectives.py:85-104, endpoint
    included, float32) and data_gen.constructX (data_gen.py:294-303, endpoint excluded, float64)."""
    nd = len(bounds)
    per_dim = ((budget ** (1 / nd)) // 2 + 1) * 2
    digits = max([abs(int(np.log(np.ptp(b) / per_dim))) for b in bounds] + [1])
    steps = [np.round(np.ptp(b) / per_dim, digits) for b in bounds]
    axes = []
    for (lo, hi), st in zip(bounds, steps):
        stop = hi + st if endpoint else hi
        axes.append(np.arange(lo, stop, st).astype(dtype))
    return axes, steps


# ---------------------------------------------------------------- ranked-DPP sampler RNG (data_gen.py:853-880)

def py39_seed_value(a) -> int:
    """What CPython 3.9 (the reference's interpreter, environment.yml:8) seeds with for a numpy integer:
    non-int seeds are hashed and reinterpreted as an unsigned size_t."""
    h = hash(int(a))
    if h == -1:
        h = -2
    return h if h >= 0 else (1 << 64) + h


def iid_comb_locs(n_pool: int, k: int, n_samples: int, seed):
    """Element ranks in [0, C(n_pool,k)] exactly as data_gen.iid_comb_locs draws them (big branch:
    one ``random.seed(int32 seed); random.randint(0, C)`` per element, seeds from np.random.randint)."""
    total = math.comb(n_pool, k)
    if total <= np.iinfo(np.int64).max:
        # small spaces: random.sample over the global `random` state (data_gen.py:873-879)
        if n_samples / total < 0.6:
            return random.sample(range(0, total), n_samples)
        inverse = set(random.sample(range(0, total), total - n_samples))
        return list(set(range(0, total)) - inverse)
    st = np.random.RandomState(seed) if seed is not None else np.random.RandomState()
    seeds = st.randint(low=np.iinfo(np.int32).min, high=np.iinfo(np.int32).max, size=n_samples)
    rnd = random.Random()
    out = []
    for s in seeds:
        rnd.seed(py39_seed_value(s))
        out.append(rnd.randint(0, total))
    return out


def iid_comb_seeds(n_samples: int, seed):
    """The per-element seeds of data_gen.iid_comb_locs' big branch (data_gen.py:866-867): legacy numpy MT19937."""
    st = np.random.RandomState(seed) if seed is not None else np.random.RandomState()
    return st.randint(low=np.iinfo(np.int32).min, high=np.iinfo(np.int32).max, size=n_samples).astype(np.int64)


def unrank_host(n: int, k: int, m: int):
    """Arbitrary-precision un-ranking for combinations beyond 128 bits (Combination.py:131-170)."""
    if k > n:
        raise Exception("Not possible")
    x = (math.comb(n, k) - 1) - m
    if x < 0:
        raise Exception("check previous")
    a, b = n, k
    out = []
    while b >= 1:
        lo, hi = b - 1, a - 1
        while lo < hi:
            mid = (lo + hi + 1) // 2
            if math.comb(mid, b) <= x:
                lo = mid
            else:
                hi = mid - 1
        out.append((n - 1) - lo)
        x -= math.comb(lo, b)
        a, b = lo, b - 1
    return out
