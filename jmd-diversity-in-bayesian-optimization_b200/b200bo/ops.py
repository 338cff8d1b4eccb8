"""Tensor-level wrappers of the C ABI: one Python function per entry point.

Every function takes CUDA tensors (contiguous, documented dtype), launches on torch's current
stream and returns tensors.  No CPU implementation exists: CPU tensors raise.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import check

RBF, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3
EI, UCB, PI = 0, 1, 2
ST_JITTER, ST_CHOL_FAIL, ST_NONFINITE, ST_ALL_MASKED, ST_OFF_LATTICE, ST_FALLBACK = 1, 2, 4, 8, 16, 32
NMAX = 128
VAR_FLOOR = 1e-9          # botorch 0.8.0 analytic EI clamp


def _cu(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.B200boError(f"{name}: expected a CUDA tensor (the engine has no CPU path)")
    if t.device.index != torch.cuda.current_device():
        raise _lib.B200boError(f"{name}: tensor lives on {t.device} but the current device is cuda:{torch.cuda.current_device()} "
                               "(the C ABI launches on the current device: wrap the call in torch.cuda.device(...))")
    if t.dtype != dtype:
        raise _lib.B200boError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.B200boError(f"{name}: tensor must be contiguous")
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def kernel_id_from(nu=2.5, covar="matern") -> int:
    """Src/BO.py:133-134 (nu > 2.5 switches to RBF) and :61-64."""
    if covar == "rbf" or nu > 2.5:
        return RBF
    try:
        return {0.5: MATERN12, 1.5: MATERN32, 2.5: MATERN52}[float(nu)]
    except KeyError:
        raise ValueError(f"nu={nu}: gpytorch's MaternKernel supports 0.5, 1.5, 2.5") from None


# ------------------------------------------------------------------ K-D

def dpp_logdet_score(pool: torch.Tensor, idx: torch.Tensor, gamma: float) -> torch.Tensor:
    """log det exp(-gamma D) per index set; pool (P,2) f64, idx (N,k) int32 -> (N,) f64."""
    lib = _lib.load()
    _cu(pool, torch.float64, "pool")
    _cu(idx, torch.int32, "idx")
    if pool.dim() != 2 or pool.shape[1] != 2 or idx.dim() != 2:
        raise _lib.B200boError("dpp_logdet_score: pool must be (P,2), idx (N,k)")
    N, k = idx.shape
    out = torch.empty(N, dtype=torch.float64, device=pool.device)
    if N == 0:
        return out
    check(lib.b200bo_dpp_logdet_score(_ptr(pool), pool.shape[0], _ptr(idx), N, k, float(gamma), _ptr(out), _stream()),
          "dpp_logdet_score")
    return out


def dpp_lattice_of(pool) -> "Lattice | None":
    """The constructX lattice behind a pool, if it is one whose coordinate differences are exact in fp64 (all
    coordinates multiples of 2^-20 below 2^30 — integer or dyadic steps), else None: only then does an L-ensemble
    entry depend on (|dx|, |dy|) alone bit for bit."""
    import numpy as np
    g = pool.detach().cpu().numpy() if isinstance(pool, torch.Tensor) else np.asarray(pool)
    lat = Lattice.from_points(g)
    if lat is None:
        return None
    scaled = g * 2.0 ** 20
    if not (np.all(scaled == np.round(scaled)) and np.abs(g).max() < 2.0 ** 30):
        return None
    return lat


def dpp_table(lat: "Lattice", gamma: float, device) -> torch.Tensor:
    """(Gy,Gx) f64 table of exp(-gamma (sqrt(dx^2+dy^2))^2) over lattice offsets — the entries K-D would compute."""
    out = torch.empty((lat.Gy, lat.Gx), dtype=torch.float64, device=device)
    check(_lib.load().b200bo_dpp_table(lat.Gx, lat.Gy, float(lat.h), float(gamma), _ptr(out), _stream()), "dpp_table")
    return out


def dpp_logdet_score_lattice(lat: "Lattice", table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Lattice variant of dpp_logdet_score: idx (N,k<=12) int32 flat node indices -> (N,) f64, bit-identical values."""
    _cu(table, torch.float64, "table")
    _cu(idx, torch.int32, "idx")
    if idx.dim() != 2 or table.shape != (lat.Gy, lat.Gx):
        raise _lib.B200boError("dpp_logdet_score_lattice: idx must be (N,k), table (Gy,Gx)")
    N, k = idx.shape
    out = torch.empty(N, dtype=torch.float64, device=idx.device)
    if N == 0:
        return out
    check(_lib.load().b200bo_dpp_logdet_score_lattice(_ptr(table), lat.Gx, lat.Gy, _ptr(idx), N, k, _ptr(out), _stream()),
          "dpp_logdet_score_lattice")
    return out


def dpp_zscore(logdet: torch.Tensor):
    """(mean, sd) as a (2,) tensor and z-scores — data_gen.py:406-408."""
    lib = _lib.load()
    _cu(logdet, torch.float64, "logdet")
    N = logdet.numel()
    stats = torch.empty(2, dtype=torch.float64, device=logdet.device)
    z = torch.empty_like(logdet)
    wsb = lib.b200bo_dpp_stats_workspace_bytes(N)
    ws = torch.empty(wsb, dtype=torch.uint8, device=logdet.device)
    check(lib.b200bo_dpp_zscore(_ptr(logdet), N, _ptr(stats), _ptr(z), _ptr(ws), wsb, _stream()), "dpp_zscore")
    return stats, z


def dpp_rank(z: torch.Tensor, pos: torch.Tensor = None, pool: torch.Tensor = None, want_dist=True):
    """Stable ascending rank order of the scores on the device (csrc/dpp_rank.cu radix sort; data_gen.py:409-415):
    returns (z_sorted (N), order (N) int32, pos_sorted (N,k) | None, dist (N,k,2) | None). Ties keep set-index order,
    NaN sorts last."""
    lib = _lib.load()
    _cu(z, torch.float64, "z")
    N = z.numel()
    dev = z.device
    k = 1
    if pos is not None:
        _cu(pos, torch.int32, "pos")
        k = pos.shape[1]
    if pool is not None:
        _cu(pool, torch.float64, "pool")
    z_sorted = torch.empty(N, dtype=torch.float64, device=dev)
    order = torch.empty(N, dtype=torch.int32, device=dev)
    pos_sorted = torch.empty((N, k), dtype=torch.int32, device=dev) if pos is not None else None
    dist = torch.empty((N, k, 2), dtype=torch.float64, device=dev) if (pos is not None and pool is not None and want_dist) else None
    wsb = int(lib.b200bo_dpp_rank_workspace_bytes(N))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    check(lib.b200bo_dpp_rank(_ptr(z), N, _ptr(pos), k, _ptr(pool), 0 if pool is None else pool.shape[0], _ptr(z_sorted),
                              _ptr(order), _ptr(pos_sorted), _ptr(dist), _ptr(ws), ws.numel(), _stream()), "dpp_rank")
    return z_sorted, order, pos_sorted, dist


def dpp_count_less(scores: torch.Tensor, z: float) -> int:
    """np.sum(scores < z) on the device (compare_subDPP, data_gen.py:481). One device -> host read of 8 bytes."""
    lib = _lib.load()
    _cu(scores, torch.float64, "scores")
    out = torch.empty(1, dtype=torch.int64, device=scores.device)
    check(lib.b200bo_dpp_count_less(_ptr(scores), scores.numel(), float(z), _ptr(out), _stream()), "dpp_count_less")
    return int(out.item())


def dpp_window_pick(acquired_mask: torch.Tensor, N: int, lo: int, hi: int, r: int, pos_sorted: torch.Tensor, pool: torch.Tensor):
    """The r-th not-yet-acquired rank of [lo, hi) (marked in acquired_mask, int32 words) and that set's points:
    returns (loc int, set (k,2) f64 host array) — sub_dpp_sample's setdiff1d + look-up (data_gen.py:445-466)."""
    lib = _lib.load()
    _cu(acquired_mask, torch.int32, "acquired_mask"); _cu(pos_sorted, torch.int32, "pos_sorted"); _cu(pool, torch.float64, "pool")
    k = pos_sorted.shape[1]
    out = torch.empty(2 * k + 1, dtype=torch.float64, device=pool.device)       # [set (k,2) | loc as int32 bits]
    loc_view = out[2 * k:].view(torch.int32)
    check(lib.b200bo_dpp_window_pick(_ptr(acquired_mask), int(N), int(lo), int(hi), int(r), _ptr(pos_sorted), k, _ptr(pool),
                                     pool.shape[0], _ptr(loc_view), _ptr(out), _stream()), "dpp_window_pick")
    host = out.cpu()
    return int(host[2 * k:].view(torch.int32)[0]), host[:2 * k].reshape(k, 2).numpy()


def dpp_logdet_score_multigamma(pool: torch.Tensor, idx: torch.Tensor, gammas) -> torch.Tensor:
    """(n_gamma, N) f64 log-dets of the same N sets under every gamma in one launch (gammacalc, data_gen.py:758-842)."""
    lib = _lib.load()
    _cu(pool, torch.float64, "pool"); _cu(idx, torch.int32, "idx")
    g = torch.as_tensor(gammas, dtype=torch.float64, device=pool.device).contiguous()
    N, k = idx.shape
    out = torch.empty((g.numel(), N), dtype=torch.float64, device=pool.device)
    check(lib.b200bo_dpp_logdet_score_multigamma(_ptr(pool), pool.shape[0], _ptr(idx), N, k, _ptr(g), g.numel(), _ptr(out),
                                                 _stream()), "dpp_logdet_score_multigamma")
    return out


_BINOM_CACHE = {}


def _binom_table(n: int, k: int, device):
    """(lo, hi) uint64-as-int64 tensors (k, n+1): C(v,b), b = 1..k, v = 0..n, plus C(n,k)."""
    import math
    import numpy as np
    key = (n, k, str(device))
    if key not in _BINOM_CACHE:
        total = math.comb(n, k)
        if total >= 1 << 128:
            raise _lib.B200boError(f"C({n},{k}) does not fit 128 bits; use the host un-ranker (Combination.py)")
        lo = np.zeros((k, n + 1), dtype=np.uint64)
        hi = np.zeros((k, n + 1), dtype=np.uint64)
        mask = (1 << 64) - 1
        for b in range(1, k + 1):
            c = 0                      # C(v,b) by the recurrence C(v+1,b) = C(v,b) (v+1)/(v+1-b)
            for v in range(n + 1):
                c = 0 if v < b else (1 if v == b else c * v // (v - b))
                if c >> 128:
                    c_store = (1 << 128) - 1        # saturate: never <= x (x < C(n,k) < 2^128)
                else:
                    c_store = c
                lo[b - 1, v] = c_store & mask
                hi[b - 1, v] = c_store >> 64
        _BINOM_CACHE[key] = (torch.as_tensor(lo.view(np.int64), device=device), torch.as_tensor(hi.view(np.int64), device=device), total)
    return _BINOM_CACHE[key]


def unrank(n: int, k: int, ranks, device) -> torch.Tensor:
    """Lexicographic un-ranking of Python-int ranks -> (N,k) int32 index sets (on device)."""
    import numpy as np
    lib = _lib.load()
    lo_t, hi_t, total = _binom_table(n, k, device)
    mask = (1 << 64) - 1
    N = len(ranks)
    m_lo = np.fromiter(((int(r) & mask) for r in ranks), dtype=np.uint64, count=N)
    m_hi = np.fromiter(((int(r) >> 64) for r in ranks), dtype=np.uint64, count=N)
    if any(int(r) < 0 or int(r) >> 128 for r in ranks):
        raise _lib.B200boError("unrank: rank out of the 128-bit range")
    mlo = torch.as_tensor(m_lo.view(np.int64), device=device)
    mhi = torch.as_tensor(m_hi.view(np.int64), device=device)
    out = torch.empty((N, k), dtype=torch.int32, device=device)
    if N:
        check(lib.b200bo_unrank_u128(_ptr(lo_t), _ptr(hi_t), _ptr(mlo), _ptr(mhi), total & mask, total >> 64, n, k, N,
                                     _ptr(out), _stream()), "unrank_u128")
    return out


def seeded_randint_u128(seeds, n: int, device):
    """For every int seed: ``random.seed(np.int64(seed)); random.randint(0, n - 1)`` under CPython 3.9 rules, on the
    device (csrc/mt_rank.cu). seeds: int64 array-like. Returns (lo, hi) int64 tensors holding the uint64 halves."""
    import numpy as np
    if not 0 < n < 1 << 128:
        raise _lib.B200boError("seeded_randint_u128: n must be in (0, 2^128)")
    sd = torch.as_tensor(np.ascontiguousarray(seeds, dtype=np.int64), device=device)
    N = sd.numel()
    lo = torch.empty(N, dtype=torch.int64, device=device)
    hi = torch.empty(N, dtype=torch.int64, device=device)
    fail = torch.zeros(1, dtype=torch.int32, device=device)
    mask = (1 << 64) - 1
    check(_lib.load().b200bo_mt_seeded_randbelow(_ptr(sd), N, n & mask, n >> 64, _ptr(lo), _ptr(hi), _ptr(fail), _stream()),
          "mt_seeded_randbelow")
    if int(fail.item()):
        raise _lib.B200boError("seeded_randint_u128: rejection sampling did not terminate in 56 rounds")
    return lo, hi


def unrank_u128(n: int, k: int, m_lo: torch.Tensor, m_hi: torch.Tensor) -> torch.Tensor:
    """Un-rank device-resident 128-bit ranks (uint64 halves stored as int64) -> (N,k) int32 index sets."""
    lib = _lib.load()
    device = m_lo.device
    lo_t, hi_t, total = _binom_table(n, k, device)
    mask = (1 << 64) - 1
    N = m_lo.numel()
    out = torch.empty((N, k), dtype=torch.int32, device=device)
    if N:
        check(lib.b200bo_unrank_u128(_ptr(lo_t), _ptr(hi_t), _ptr(m_lo), _ptr(m_hi), total & mask, total >> 64, n, k, N,
                                     _ptr(out), _stream()), "unrank_u128")
    return out


# ------------------------------------------------------------------ K-E

def wildcat_surface(hseed: torch.Tensor, vecs: torch.Tensor, peak_xy: torch.Tensor, smooth: torch.Tensor,
                    rug_amp: torch.Tensor) -> torch.Tensor:
    """(S,101,101) f32 surfaces; hseed (S,) int64 (bit pattern of the uint64 hash), vecs (S,4,2) int8,
    peak_xy (S,2) f64, smooth / rug_amp (S,) f64."""
    lib = _lib.load()
    _cu(hseed, torch.int64, "hseed")
    _cu(vecs, torch.int8, "vecs")
    _cu(peak_xy, torch.float64, "peak_xy")
    _cu(smooth, torch.float64, "smooth")
    _cu(rug_amp, torch.float64, "rug_amp")
    S = hseed.numel()
    if vecs.shape != (S, 4, 2) or peak_xy.shape != (S, 2) or smooth.numel() != S or rug_amp.numel() != S:
        raise _lib.B200boError("wildcat_surface: inconsistent shapes")
    surf = torch.empty((S, 101, 101), dtype=torch.float32, device=hseed.device)
    check(lib.b200bo_wildcat_surface(_ptr(hseed), _ptr(vecs), _ptr(peak_xy), _ptr(smooth), _ptr(rug_amp), _ptr(surf), S,
                                     _stream()), "wildcat_surface")
    return surf


_DIAG_MASKS = {}


def qhull_diag_mask(device) -> torch.Tensor:
    """(10000,) uint8 on ``device``: the diagonal scipy's Delaunay (Qhull) triangulation cuts each unit cell of the
    101 x 101 node lattice along (1 = main) — what LinearNDInterpolator in the reference's generate_cont()
    (Src/objectives.py:349-350) interpolates on. Shipped as data/qhull_diag_101.bin (tools/make_diag_mask.py)."""
    import os

    import numpy as np
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _DIAG_MASKS:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "qhull_diag_101.bin")
        bits = np.unpackbits(np.fromfile(path, dtype=np.uint8))[:10000]
        if bits.size != 10000:
            raise _lib.B200boError(f"{path}: expected 10 000 cell bits")
        _DIAG_MASKS[key] = torch.as_tensor(bits.copy(), dtype=torch.uint8, device=dev)
    return _DIAG_MASKS[key]


def wildcat_eval(surf: torch.Tensor, xy: torch.Tensor, surf_id=None, diag_mask="qhull") -> torch.Tensor:
    """f(x) for M points xy (M,2) f64 on surfaces surf (S,101,101) f32; surf_id (M,) int32 or None.
    diag_mask: "qhull" (default: the triangulation scipy builds for the reference), None (main diagonal in
    every cell) or a (10000,) uint8 CUDA tensor."""
    lib = _lib.load()
    _cu(surf, torch.float32, "surf"); _cu(xy, torch.float64, "xy")
    if surf_id is not None:
        _cu(surf_id, torch.int32, "surf_id")
    if isinstance(diag_mask, str):
        if diag_mask != "qhull":
            raise ValueError(f"diag_mask={diag_mask!r}")
        diag_mask = qhull_diag_mask(xy.device)
    if diag_mask is not None:
        _cu(diag_mask, torch.uint8, "diag_mask")
        if diag_mask.numel() != 10000:
            raise _lib.B200boError("diag_mask: expected 100 x 100 cells")
    M = xy.shape[0]
    out = torch.empty(M, dtype=torch.float64, device=xy.device)
    if M:
        check(lib.b200bo_wildcat_eval(_ptr(surf), _ptr(surf_id), _ptr(diag_mask), _ptr(xy), _ptr(out), M, _stream()),
              "wildcat_eval")
    return out


def fp64_fma_peak_tflops(device=None, reps=8) -> float:
    """Measured FP64 FMA throughput of the current GPU in TFLOP/s (best of ``reps``, CUDA events)."""
    lib = _lib.load()
    dev = torch.device(device or "cuda")
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks, iters = sms * 8, 32768
    out = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
    best = 0.0
    for r in range(reps + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = lib.b200bo_fp64_fma_probe(_ptr(out), blocks, iters, _stream())
        e1.record()
        e1.synchronize()
        if flops <= 0:
            raise _lib.B200boError("fp64_fma_probe failed: " + lib.b200bo_last_error_string().decode())
        if r >= 2:
            best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def smem_lds_peak_gbs(device=None, reps=6) -> float:
    """Measured shared-memory read bandwidth of the current GPU in GB/s (conflict-free LDS.64, one 1024-thread CTA per SM;
    best of ``reps``, CUDA events): the roofline denominator of the persistent BO loop (csrc/bo_persist.cu)."""
    lib = _lib.load()
    dev = torch.device(device or "cuda")
    with torch.cuda.device(dev):
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        blocks, iters = sms, 1 << 17
        out = torch.empty(blocks * 1024, dtype=torch.float64, device=dev)
        best = 0.0
        for r in range(reps + 2):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            nbytes = lib.b200bo_smem_lds_probe(_ptr(out), blocks, iters, _stream())
            e1.record()
            e1.synchronize()
            if nbytes <= 0:
                raise _lib.B200boError("smem_lds_probe failed: " + lib.b200bo_last_error_string().decode())
            if r >= 2:
                best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def simplex_noise(hseed: int, vecs, xy: torch.Tensor) -> torch.Tensor:
    """Raw 2-D simplex noise at xy (M,2) f64 for SimplexNoise(seed,2) with hashed seed / table given."""
    import ctypes
    lib = _lib.load()
    _cu(xy, torch.float64, "xy")
    M = xy.shape[0]
    out = torch.empty(M, dtype=torch.float64, device=xy.device)
    flat = [int(v) for row in vecs for v in row]
    arr = (ctypes.c_int8 * 8)(*flat)
    check(lib.b200bo_simplex_noise(int(hseed), ctypes.cast(arr, ctypes.c_void_p), _ptr(xy), _ptr(out), M, _stream()),
          "simplex_noise")
    return out


# ------------------------------------------------------------------ K-A / K-C

def wfrag_doubles(n_max: int) -> int:
    return int(_lib.load().b200bo_wfrag_doubles(int(n_max)))


def gp_assemble_chol(X, y, n, theta, kernel_id, want=("alpha", "wfrag"), status=None, n_hint=0):
    """K-A. X (T,n_max,2) f64, y (T,n_max) f64, n (T,) int32, theta (T,4) f64.
    ``want`` selects outputs among L, Linv, alpha, logdet, wfrag. Returns dict (+ status)."""
    lib = _lib.load()
    _cu(X, torch.float64, "X"); _cu(y, torch.float64, "y"); _cu(n, torch.int32, "n"); _cu(theta, torch.float64, "theta")
    T, n_max = y.shape
    if X.shape != (T, n_max, 2) or n.shape != (T,) or theta.shape != (T, 4):
        raise _lib.B200boError("gp_assemble_chol: inconsistent shapes")
    dev = X.device
    out = {}
    if "L" in want: out["L"] = torch.empty((T, n_max, n_max), dtype=torch.float64, device=dev)
    if "Linv" in want: out["Linv"] = torch.empty((T, n_max, n_max), dtype=torch.float64, device=dev)
    if "alpha" in want: out["alpha"] = torch.empty((T, n_max), dtype=torch.float64, device=dev)
    if "logdet" in want: out["logdet"] = torch.empty((T,), dtype=torch.float64, device=dev)
    if "wfrag" in want: out["wfrag"] = torch.empty((T, wfrag_doubles(n_max)), dtype=torch.float64, device=dev)
    if status is None:
        status = torch.zeros(T, dtype=torch.int32, device=dev)
    check(lib.b200bo_gp_assemble_chol(_ptr(X), _ptr(y), _ptr(n), _ptr(theta), _ptr(out.get("L")), _ptr(out.get("Linv")),
                                      _ptr(out.get("alpha")), _ptr(out.get("logdet")), _ptr(out.get("wfrag")),
                                      _ptr(status), T, n_max, int(n_hint), int(kernel_id), _stream()), "gp_assemble_chol")
    out["status"] = status
    return out


def gp_assemble_chol_into(X, y, n, theta, kernel_id, alpha, wfrag, status, n_hint=0):
    """Allocation-free K-A for the engine loop (alpha / wfrag / status preallocated)."""
    lib = _lib.load()
    T, n_max = y.shape
    check(lib.b200bo_gp_assemble_chol(_ptr(X), _ptr(y), _ptr(n), _ptr(theta), None, None, _ptr(alpha), None, _ptr(wfrag),
                                      _ptr(status), T, n_max, int(n_hint), int(kernel_id), _stream()), "gp_assemble_chol")


def gp_chol_append(X, y, n, theta, kernel_id, wfrag, zvec, alpha, n_redo, status, wrow_out=None):
    """K-A append variant (in place on wfrag / zvec / alpha); n_redo marks trials needing a full refactorisation;
    wrow_out (T,n_max) optionally receives the dense new row of L^-1."""
    lib = _lib.load()
    T, n_max = y.shape
    check(lib.b200bo_gp_chol_append(_ptr(X), _ptr(y), _ptr(n), _ptr(theta), _ptr(wfrag), _ptr(zvec), _ptr(alpha),
                                    _ptr(wrow_out), _ptr(n_redo), _ptr(status), T, n_max, int(kernel_id), _stream()),
          "gp_chol_append")


def kernel_table_signed(theta, kernel_id, lat, out=None):
    """(H, 2Gy-1, 2Gx-1) f64 signed-offset kernel table (input of the incremental K-B)."""
    lib = _lib.load()
    _cu(theta, torch.float64, "theta")
    H = theta.shape[0]
    if out is None:
        out = torch.empty((H, 2 * lat.Gy - 1, 2 * lat.Gx - 1), dtype=torch.float64, device=theta.device)
    check(lib.b200bo_kernel_table_signed(_ptr(theta), H, int(kernel_id), lat.h, lat.Gx, lat.Gy, _ptr(out), _stream()),
          "kernel_table_signed")
    return out


def grid_inc_workspace(T: int, G: int, device) -> torch.Tensor:
    nbytes = int(_lib.load().b200bo_grid_inc_workspace_bytes(int(T), int(G)))
    return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)


def grid_posterior_update_acq_argmax(lat, ktab2, ktab_id, X, wrow, zvec, n, theta, best_f, mu_state, ss_state, obs_mask,
                                     first, acquire, acq_id=EI, acq_param=0.0, var_floor=VAR_FLOOR, extra_mask=None,
                                     arg_idx=None, arg_val=None, acq_out=None, status=None, ws=None):
    """Incremental K-B: fold the new row of L^-1 into (mu_state, ss_state) and, if ``acquire``, arg-max the acquisition."""
    lib = _lib.load()
    T, n_max = zvec.shape
    check(lib.b200bo_grid_posterior_update_acq_argmax(
        lat.Gx, lat.Gy, lat.x0, lat.y0, lat.h, _ptr(ktab2), _ptr(ktab_id), _ptr(X), _ptr(wrow), _ptr(zvec), _ptr(n),
        _ptr(theta), _ptr(best_f), _ptr(mu_state), _ptr(ss_state), _ptr(obs_mask), _ptr(extra_mask), 1 if first else 0,
        1 if acquire else 0, int(acq_id), float(acq_param), float(var_floor), _ptr(arg_idx), _ptr(arg_val), _ptr(acq_out),
        _ptr(status), T, n_max, _ptr(ws), 0 if ws is None else ws.numel(), _stream()), "grid_posterior_update_acq_argmax")


def kernel_table_ysigned(theta, kernel_id, lat, out=None):
    """(H, 2Gy-1, Gx) f64 y-signed kernel table (input of the persistent fixed-theta loop, csrc/bo_persist.cu)."""
    lib = _lib.load()
    _cu(theta, torch.float64, "theta")
    H = theta.shape[0]
    if out is None:
        out = torch.empty((H, 2 * lat.Gy - 1, lat.Gx), dtype=torch.float64, device=theta.device)
    check(lib.b200bo_kernel_table_ysigned(_ptr(theta), H, int(kernel_id), lat.h, lat.Gx, lat.Gy, _ptr(out), _stream()),
          "kernel_table_ysigned")
    return out


def bo_fixed_loop_supported(lat, n_max: int) -> bool:
    """Does the lattice's kernel table (+ n_max rows of L^-1) fit one CTA's shared memory? Host-only."""
    return bool(_lib.load().b200bo_bo_fixed_loop_supported(int(lat.Gx), int(lat.Gy), int(n_max)))


def bo_fixed_loop_workspace(T: int, lat, device) -> torch.Tensor:
    nbytes = int(_lib.load().b200bo_bo_fixed_loop_workspace_bytes(int(T), int(lat.Gx), int(lat.Gy)))
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)


def bo_fixed_loop(lat, ktabY, ktab_id, order, obj, surf_id, X, y, n, theta, best_f, yall, picked, nit, status, max_iter,
                  acq_id=EI, acq_param=0.0, var_floor=VAR_FLOOR, tol=0.1, optimal_y=100.0, ws=None, slot_start=None):
    """K-F: every iteration of every trial in ONE launch (BOloop_fixparam's loop, Src/BO.py:404-446). In place on
    X / y / n / best_f / yall / picked / nit / status; trials flagged ST_FALLBACK come back untouched."""
    lib = _lib.load()
    _cu(ktabY, torch.float64, "ktabY"); _cu(obj, torch.float32, "obj"); _cu(surf_id, torch.int32, "surf_id")
    _cu(X, torch.float64, "X"); _cu(y, torch.float64, "y"); _cu(n, torch.int32, "n"); _cu(theta, torch.float64, "theta")
    _cu(best_f, torch.float64, "best_f"); _cu(yall, torch.float64, "yall"); _cu(picked, torch.int32, "picked")
    _cu(nit, torch.int32, "nit"); _cu(status, torch.int32, "status")
    if ktab_id is not None:
        _cu(ktab_id, torch.int32, "ktab_id")
    if order is not None:
        _cu(order, torch.int32, "order")
    T, n_max = y.shape
    if ws is None:
        ws = bo_fixed_loop_workspace(T, lat, X.device)
    H = ktabY.shape[0]
    if slot_start is not None:
        _cu(slot_start, torch.int32, "slot_start")
    check(lib.b200bo_bo_fixed_loop(lat.Gx, lat.Gy, lat.x0, lat.y0, lat.h, _ptr(ktabY), _ptr(ktab_id), _ptr(order),
                                   _ptr(slot_start), H, _ptr(obj),
                                   _ptr(surf_id), _ptr(X), _ptr(y), _ptr(n), _ptr(theta), _ptr(best_f), _ptr(yall),
                                   _ptr(picked), _ptr(nit), _ptr(status), T, n_max, int(max_iter), int(acq_id),
                                   float(acq_param), float(var_floor), float(tol), float(optimal_y), _ptr(ws), ws.numel(),
                                   _stream()), "bo_fixed_loop")
    return ws


def gp_zvec(y, n, theta, wfrag, zvec):
    lib = _lib.load()
    T, n_max = y.shape
    check(lib.b200bo_gp_zvec(_ptr(y), _ptr(n), _ptr(theta), _ptr(wfrag), _ptr(zvec), T, n_max, _stream()), "gp_zvec")


def mll_value_grad(X, y, n, raw, kernel_id, status=None, n_hint=0, value=None, grad=None):
    """K-C. raw (T,4) f64 = (noise, c, raw_outputscale, raw_lengthscale) -> value (T,), grad (T,4), status."""
    lib = _lib.load()
    _cu(X, torch.float64, "X"); _cu(y, torch.float64, "y"); _cu(n, torch.int32, "n"); _cu(raw, torch.float64, "raw")
    T, n_max = y.shape
    dev = X.device
    if value is None: value = torch.empty(T, dtype=torch.float64, device=dev)
    if grad is None: grad = torch.empty((T, 4), dtype=torch.float64, device=dev)
    if status is None:
        status = torch.zeros(T, dtype=torch.int32, device=dev)
    check(lib.b200bo_mll_value_grad(_ptr(X), _ptr(y), _ptr(n), _ptr(raw), _ptr(value), _ptr(grad), _ptr(status), T, n_max,
                                    int(n_hint), int(kernel_id), _stream()), "mll_value_grad")
    return value, grad, status


# ------------------------------------------------------------------ K-B

def grid_acq_workspace(T: int, G: int, device) -> torch.Tensor:
    nbytes = int(_lib.load().b200bo_grid_acq_workspace_bytes(int(T), int(G)))
    return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)


def grid_posterior_acq_argmax(grid, X, wfrag, alpha, n, theta, best_f, kernel_id, acq_id=EI, acq_param=0.0,
                              var_floor=VAR_FLOOR, mask_observed=True, extra_mask=None, debug=False, status=None,
                              ws=None, arg_idx=None, arg_val=None, n_hint=0):
    """K-B. Returns dict(idx (T,) int32, val (T,) f64, status[, mu, var, acq (T,G)])."""
    lib = _lib.load()
    _cu(grid, torch.float64, "grid"); _cu(X, torch.float64, "X"); _cu(wfrag, torch.float64, "wfrag")
    _cu(alpha, torch.float64, "alpha"); _cu(n, torch.int32, "n"); _cu(theta, torch.float64, "theta")
    _cu(best_f, torch.float64, "best_f")
    G = grid.shape[0]
    T, n_max = alpha.shape
    dev = grid.device
    if extra_mask is not None:
        _cu(extra_mask, torch.uint8, "extra_mask")
    if wfrag.shape != (T, wfrag_doubles(n_max)):
        raise _lib.B200boError("grid_posterior_acq_argmax: wfrag has the wrong shape for n_max")
    if arg_idx is None: arg_idx = torch.empty(T, dtype=torch.int32, device=dev)
    if arg_val is None: arg_val = torch.empty(T, dtype=torch.float64, device=dev)
    if status is None: status = torch.zeros(T, dtype=torch.int32, device=dev)
    if ws is None: ws = grid_acq_workspace(T, G, dev)
    mu = var = acq = None
    if debug:
        mu = torch.empty((T, G), dtype=torch.float64, device=dev)
        var = torch.empty_like(mu)
        acq = torch.empty_like(mu)
    check(lib.b200bo_grid_posterior_acq_argmax(
        _ptr(grid), G, _ptr(X), _ptr(wfrag), _ptr(alpha), _ptr(n), _ptr(theta), _ptr(best_f), _ptr(extra_mask),
        1 if mask_observed else 0, int(acq_id), float(acq_param), float(var_floor), _ptr(arg_idx), _ptr(arg_val),
        _ptr(mu), _ptr(var), _ptr(acq), _ptr(status), T, n_max, int(n_hint), int(kernel_id), _ptr(ws), ws.numel(),
        _stream()),
        "grid_posterior_acq_argmax")
    out = dict(idx=arg_idx, val=arg_val, status=status)
    if debug:
        out.update(mu=mu, var=var, acq=acq)
    return out


class Lattice:
    """Regular candidate lattice: g <-> (g % Gx, g // Gx), x = x0 + h*ix (data_gen.constructX order)."""

    def __init__(self, Gx, Gy, x0=0.0, y0=0.0, h=1.0):
        self.Gx, self.Gy, self.x0, self.y0, self.h = int(Gx), int(Gy), float(x0), float(y0), float(h)

    @property
    def G(self):
        return self.Gx * self.Gy

    def points(self, device):
        ix = torch.arange(self.Gx, dtype=torch.float64, device=device) * self.h + self.x0
        iy = torch.arange(self.Gy, dtype=torch.float64, device=device) * self.h + self.y0
        return torch.stack([ix.repeat(self.Gy), iy.repeat_interleave(self.Gx)], dim=1).contiguous()

    def contains(self, X: torch.Tensor) -> torch.Tensor:
        """Per point of X (...,2): is it a node of this lattice? Same test as the kernels (csrc/grid_acq.cu staging:
        x0 + h*rint((x-x0)/h) == x, unfused, index in range)."""
        fx = torch.round((X[..., 0] - self.x0) / self.h)
        fy = torch.round((X[..., 1] - self.y0) / self.h)
        return ((self.x0 + self.h * fx == X[..., 0]) & (self.y0 + self.h * fy == X[..., 1]) & (fx >= 0) & (fy >= 0)
                & (fx < self.Gx) & (fy < self.Gy))

    @staticmethod
    def from_points(grid):
        """Recognise a constructX-style lattice in a (G,2) point array (host-side, exact comparison)."""
        g = grid.detach().cpu().numpy() if isinstance(grid, torch.Tensor) else grid
        G = g.shape[0]
        if G < 2:
            return None
        x0, y0 = float(g[0, 0]), float(g[0, 1])
        h = float(g[1, 0] - g[0, 0])
        if not h > 0:
            return None
        import numpy as np
        row = np.nonzero(g[:, 1] != y0)[0]
        Gx = int(row[0]) if len(row) else G
        if G % Gx:
            return None
        lat = Lattice(Gx, G // Gx, x0, y0, h)
        ix = np.arange(G) % Gx
        iy = np.arange(G) // Gx
        if np.array_equal(g[:, 0], x0 + h * ix) and np.array_equal(g[:, 1], y0 + h * iy):
            return lat
        return None


def kernel_table(theta, kernel_id, lat: Lattice, out=None):
    """(H,Gy,Gx) f64 table of s*k(h*sqrt(dx^2+dy^2)/l) for H hyper-parameter rows theta (H,4)."""
    lib = _lib.load()
    _cu(theta, torch.float64, "theta")
    H = theta.shape[0]
    if out is None:
        out = torch.empty((H, lat.Gy, lat.Gx), dtype=torch.float64, device=theta.device)
    check(lib.b200bo_kernel_table(_ptr(theta), H, int(kernel_id), lat.h, lat.Gx, lat.Gy, _ptr(out), _stream()), "kernel_table")
    return out


def grid_posterior_acq_argmax_lattice(lat: Lattice, ktab, ktab_id, X, wfrag, alpha, n, theta, best_f, kernel_id,
                                      acq_id=EI, acq_param=0.0, var_floor=VAR_FLOOR, mask_observed=True,
                                      extra_mask=None, debug=False, status=None, ws=None, arg_idx=None, arg_val=None,
                                      n_hint=0):
    """Lattice variant of K-B (kernel values from ``ktab``); same outputs as grid_posterior_acq_argmax."""
    lib = _lib.load()
    _cu(ktab, torch.float64, "ktab"); _cu(X, torch.float64, "X"); _cu(wfrag, torch.float64, "wfrag")
    _cu(alpha, torch.float64, "alpha"); _cu(n, torch.int32, "n"); _cu(theta, torch.float64, "theta")
    _cu(best_f, torch.float64, "best_f")
    if ktab_id is not None:
        _cu(ktab_id, torch.int32, "ktab_id")
    if extra_mask is not None:
        _cu(extra_mask, torch.uint8, "extra_mask")
    G = lat.G
    T, n_max = alpha.shape
    dev = X.device
    if ktab.shape[-2:] != (lat.Gy, lat.Gx):
        raise _lib.B200boError("grid_posterior_acq_argmax_lattice: ktab does not match the lattice")
    if arg_idx is None: arg_idx = torch.empty(T, dtype=torch.int32, device=dev)
    if arg_val is None: arg_val = torch.empty(T, dtype=torch.float64, device=dev)
    if status is None: status = torch.zeros(T, dtype=torch.int32, device=dev)
    if ws is None: ws = grid_acq_workspace(T, G, dev)
    mu = var = acq = None
    if debug:
        mu = torch.empty((T, G), dtype=torch.float64, device=dev)
        var = torch.empty_like(mu)
        acq = torch.empty_like(mu)
    check(lib.b200bo_grid_posterior_acq_argmax_lattice(
        lat.Gx, lat.Gy, lat.x0, lat.y0, lat.h, _ptr(ktab), _ptr(ktab_id), _ptr(X), _ptr(wfrag), _ptr(alpha), _ptr(n),
        _ptr(theta), _ptr(best_f), _ptr(extra_mask), 1 if mask_observed else 0, int(acq_id), float(acq_param),
        float(var_floor), _ptr(arg_idx), _ptr(arg_val), _ptr(mu), _ptr(var), _ptr(acq), _ptr(status), T, n_max,
        int(n_hint), int(kernel_id), _ptr(ws), ws.numel(), _stream()), "grid_posterior_acq_argmax_lattice")
    out = dict(idx=arg_idx, val=arg_val, status=status)
    if debug:
        out.update(mu=mu, var=var, acq=acq)
    return out


# ------------------------------------------------------------------ BO loop bookkeeping

def bo_init(y, n, best_f, yall, picked, nit, max_iter, tol, optimal_y):
    lib = _lib.load()
    T, n_max = y.shape
    check(lib.b200bo_bo_init(_ptr(y), _ptr(n), _ptr(best_f), _ptr(yall), _ptr(picked), _ptr(nit), T, n_max, int(max_iter),
                             float(tol), float(optimal_y), _stream()), "bo_init")


def bo_update(grid, obj, surf_id, arg_idx, X, y, n, best_f, yall, picked, nit, max_iter, it, tol, optimal_y):
    lib = _lib.load()
    T, n_max = y.shape
    check(lib.b200bo_bo_update(_ptr(grid), grid.shape[0], _ptr(obj), _ptr(surf_id), _ptr(arg_idx), _ptr(X), _ptr(y), _ptr(n),
                               _ptr(best_f), _ptr(yall), _ptr(picked), _ptr(nit), T, n_max, int(max_iter), int(it),
                               float(tol), float(optimal_y), _stream()), "bo_update")


def bo_finalize(yall, nit, n, max_iter):
    lib = _lib.load()
    T = nit.numel()
    check(lib.b200bo_bo_finalize(_ptr(yall), _ptr(nit), _ptr(n), T, int(max_iter), _stream()), "bo_finalize")


def bo_fit_counts(n, nit, n_fit, it):
    lib = _lib.load()
    check(lib.b200bo_bo_fit_counts(_ptr(n), _ptr(nit), _ptr(n_fit), n.numel(), int(it), _stream()), "bo_fit_counts")


def bootstrap_ci(data: torch.Tensor, idx: torch.Tensor, percentiles=(2.5, 97.5)) -> torch.Tensor:
    """utils.bootstrap on device: data (N,C) f64, idx (B,N) int32 resampled rows -> (1 + len(percentiles), C) f64 =
    [mean over replicates, percentile...] of the replicate means (numpy summation order / "linear" percentiles)."""
    data = _cu(data, torch.float64, "data")
    idx = _cu(idx, torch.int32, "idx")
    N, C = data.shape
    B = idx.shape[0]
    if idx.shape[1] != N:
        raise ValueError("idx must be (B, N)")
    boot = torch.empty((B, C), dtype=torch.float64, device=data.device)
    check(_lib.load().b200bo_bootstrap_means(_ptr(data), N, C, _ptr(idx), B, _ptr(boot), _stream()), "bootstrap_means")
    q = torch.tensor([float(p) / 100 for p in percentiles], dtype=torch.float64, device=data.device)   # np.true_divide(q, 100)
    out = torch.empty((1 + len(percentiles), C), dtype=torch.float64, device=data.device)
    check(_lib.load().b200bo_column_mean_percentiles(_ptr(boot), B, C, _ptr(q), len(percentiles), _ptr(out), _stream()),
          "column_mean_percentiles")
    return out


# ------------------------------------------------------------------ device scoping
# The C ABI launches on CUDA's *current* device and _stream() returns that device's current stream. Every public
# wrapper above therefore runs with the device of its first CUDA tensor argument selected, so a caller that works on
# cuda:N while another device is current (BO.py / Optimizers.py expose device='cuda:N') launches where its tensors live.

def _device_scoped(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    return fn(*args, **kwargs)
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        for a in kwargs.values():
            if isinstance(a, torch.Tensor) and a.is_cuda and a.device.index != torch.cuda.current_device():
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


for _name, _fn in list(globals().items()):
    if (callable(_fn) and not isinstance(_fn, type) and getattr(_fn, "__module__", None) == __name__
            and not _name.startswith("_") and _name != "check"):
        globals()[_name] = _device_scoped(_fn)
del _name, _fn
