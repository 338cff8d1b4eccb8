"""Oracle: ranked-DPP ("subDPP") sampler arithmetic, numpy fp64 / Python ints.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates

* ``diverse_data_generator.constructX``      /root/reference/Src/data_gen.py:294-303
* ``diverse_data_generator.twoD_score``      data_gen.py:484-497
* scoring / normalise / argsort block of ``sub_dpp``   data_gen.py:398-415
* ``sub_dpp_sample`` rank windows            data_gen.py:427-470
* ``compare_subDPP``                         data_gen.py:472-482
* ``iid_comb_locs``                          data_gen.py:853-880
* ``combination_finder.find_elm``            /root/reference/Src/Combination.py:131-170
"""
from __future__ import annotations

import math
import random

import numpy as np


def construct_pool(bounds=((0, 100), (0, 100)), budget=10000) -> np.ndarray:
    """data_gen.py:294-303 — (budget_eff, d) f64 lattice, x fastest."""
    dim = len(bounds)
    dim_budget = ((budget ** (1 / dim)) // 2 + 1) * 2
    rr = max([abs(int(np.log(np.ptp(b) / dim_budget))) for b in bounds] + [1])
    res = [np.round(np.ptp(b) / dim_budget, rr) for b in bounds]
    axes = [np.arange(b[0], b[1], res[i]) for i, b in enumerate(bounds)]
    return np.stack(np.meshgrid(*axes), axis=-1).reshape(-1, dim).astype(np.float64)


def kernel_matrix(subset: np.ndarray, gamma: float) -> np.ndarray:
    """S = exp(-gamma * D), D = squareform(pdist(subset)**2): the distance goes
    through a sqrt and is squared again (data_gen.py:489,494)."""
    s = np.asarray(subset, dtype=np.float64)
    diff = s[:, None, :] - s[None, :, :]
    dist = np.sqrt(np.sum(diff * diff, axis=-1))       # scipy pdist 'euclidean'
    D = dist ** 2
    return np.exp(-gamma * D)


def score_lu(subset: np.ndarray, gamma: float) -> float:
    """data_gen.py:484-497: ``slogdet(S)[1]`` (LAPACK LU, sign discarded)."""
    return float(np.linalg.slogdet(kernel_matrix(subset, gamma))[1])


def score_chol(subset: np.ndarray, gamma: float) -> float:
    """Same quantity through Cholesky (S is SPD in exact arithmetic); what the
    CUDA kernel computes: right-looking, column by column, 2*sum(log L_ii)."""
    S = kernel_matrix(subset, gamma)
    k = S.shape[0]
    A = S.copy()
    acc = 0.0
    for j in range(k):
        d = A[j, j]
        if not d > 0.0:
            return float("nan")
        piv = math.sqrt(d)
        acc += math.log(d)
        A[j + 1:, j] /= piv
        for c in range(j + 1, k):
            A[c:, c] -= A[c:, j] * A[c, j]
    return acc


def score_mp(subset, gamma, dps=60) -> float:
    """High-precision logdet (mpmath), for the accuracy report only."""
    import mpmath as mp
    mp.mp.dps = dps
    s = [[mp.mpf(float(v)) for v in row] for row in np.asarray(subset, dtype=np.float64)]
    k = len(s)
    M = mp.matrix(k, k)
    g = mp.mpf(gamma)
    for i in range(k):
        for j in range(k):
            d2 = sum((s[i][c] - s[j][c]) ** 2 for c in range(len(s[i])))
            M[i, j] = mp.e ** (-g * d2)
    return float(mp.log(abs(mp.det(M))))


def cond_bound_rel(subset, gamma) -> float:
    """Relative parity bound for one set: max(1e-8, 8 k cond(S) 2^-53 / |logdet|)
    (SURVEY.md §8(c)); the conditioning, not the kernel, sets the floor."""
    S = kernel_matrix(subset, gamma)
    k = S.shape[0]
    c = np.linalg.cond(S)
    ld = abs(np.linalg.slogdet(S)[1])
    return max(1e-8, 8 * k * c * 2.0 ** -53 / max(ld, 1e-300))


def rank_scores(logdet: np.ndarray):
    """data_gen.py:406-409: population mean / std, z-scores, ascending argsort.
    The reference's ``np.argsort`` is unstable on ties; the engine defines the
    order as stable ascending (score, set index) — identical when no exact ties."""
    logdet = np.asarray(logdet, dtype=np.float64)
    mean = logdet.mean()
    sd = np.std(logdet)
    z = (logdet - mean) / sd
    ind = np.argsort(z, kind="stable")
    return mean, sd, z[ind], ind


def sample_window(percentile: float, n_samples: int):
    """data_gen.py:445-450 / 458-463: half-open rank window for one percentile."""
    if percentile < 97.5 and percentile > 2.5:
        return int((percentile - 2.5) / 100 * n_samples), int(percentile / 100 * n_samples)
    elif percentile > 97.5:
        return int(95 / 100 * n_samples), n_samples
    return 0, int(5 / 100 * n_samples)


def sample_rank(percentile: float, n_samples: int, acquired, rng: np.random.Generator) -> int:
    """data_gen.py:457-468: window minus already-acquired ranks, one uniform choice."""
    lo, hi = sample_window(percentile, n_samples)
    possible = np.setdiff1d(list(range(lo, hi)), acquired)
    return int(rng.choice(list(possible), 1)[0])


def compare_rank(sorted_z: np.ndarray, mean: float, sd: float, logdet: float):
    """data_gen.py:472-482."""
    z = (logdet - mean) / sd
    return int(np.sum(sorted_z < z)), z


# ------------------------------------------------------------------ combinadics

def unrank(n: int, k: int, m: int):
    """Combination.py:131-170 — m-th k-subset of range(n) in lexicographic order.
    ``LargestV`` and ``LargestV_binary`` (Combination.py:11-54,95-117) return the
    same v (largest v < a with C(v,b) <= x); a plain descending scan is used."""
    if k > n:
        raise Exception("Not possible")
    x = (math.comb(n, k) - 1) - m
    a, b = n, k
    ans = []
    for _ in range(k):
        if x < 0:
            # m == C(n,k) (random.randint's inclusive upper end, a 1-in-2.7e33 event
            # for n=10000,k=10): the reference ends in math.comb(-1, b) ->
            # Exception('check previous') (Combination.py:150-153).
            raise Exception("check previous")
        lo, hi = b - 1, a - 1      # C(b-1,b) = 0 <= x always
        while lo < hi:
            mid = (lo + hi + 1) // 2
            if math.comb(mid, b) <= x:
                lo = mid
            else:
                hi = mid - 1
        v = lo
        ans.append(v)
        x -= math.comb(v, b)
        a = v
        b -= 1
    return [(n - 1) - v for v in ans]


def py39_seed_int(a) -> int:
    """CPython 3.9 ``random.seed(np.int64)``: the hash reinterpreted as size_t
    (see oracle/refshim.py:py39_seed)."""
    h = hash(int(a))
    if h == -1:
        h = -2
    return h if h >= 0 else (1 << 64) + h


def iid_comb_locs(n: int, k: int, n_samples: int, seed):
    """data_gen.py:853-880, big-combination branch (C(n,k) > int64 max)."""
    total = math.comb(n, k)
    if total <= np.iinfo(np.int64).max:
        raise NotImplementedError("small-combination branch uses the global `random` state (data_gen.py:873-879)")
    st = np.random.RandomState(seed)
    seeds = st.randint(low=np.iinfo(np.int32).min, high=np.iinfo(np.int32).max, size=n_samples)
    out = []
    r = random.Random()
    for s in seeds:
        r.seed(py39_seed_int(s))
        out.append(r.randint(0, total))
    return out


def generate(pool: np.ndarray, k: int, n_samples: int, gamma: float, seed, scorer=score_lu):
    """sub_dpp (data_gen.py:391-415) end to end: returns dict(mean, sd, scores
    (sorted z), dist (sorted sets (N,k,d)), idx (sorted index sets (N,k)), raw)."""
    elems = iid_comb_locs(pool.shape[0], k, n_samples, seed)
    idx = np.array([unrank(pool.shape[0], k, e) for e in elems], dtype=np.int64)
    raw = np.array([scorer(pool[i], gamma) for i in idx])
    mean, sd, zs, order = rank_scores(raw)
    return dict(mean=mean, sd=sd, scores=zs, dist=pool[idx][order], idx=idx[order], raw=raw, order=order, unsorted_idx=idx)
