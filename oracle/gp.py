"""Oracle: exact-GP posterior, acquisition, marginal likelihood and the grid BO
loop, numpy fp64.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned**: the
arithmetic restated here lives in botorch 0.8.0 / gpytorch 1.9.0 /
linear_operator 0.2.0 (environment.yml:26,47,71), which are not vendored under
/root/reference and not installed in this image, and the reference holds no
tests or golden vectors for it.  Anchors are the reference's call sites:

* model configuration      /root/reference/Src/BO.py:33-68  (SingleTaskGP, ScaleKernel(Matern|RBF),
                           default GaussianLikelihood: Gamma(1.1,0.05) noise prior, noise >= 1e-4)
* acquisition              BO.py:70,158 (ExpectedImprovement(model, best_f=max y))
* loop semantics           BO.py:123-278 (BOloop), 380-461 (BOloop_fixparam)
* fit                      BO.py:144,189 (fit_gpytorch_model -> scipy L-BFGS-B)

and the published formulas.  An independent check
against scikit-learn's exact GP is in tests/test_oracle_golden.py; mpmath (EI tails), torch autograd
of an independent MLL formulation and the explicit-inverse posterior are in tests/test_oracle_pin.py.

Conventions shared with the CUDA kernels (include/b200bo.h):
  theta = (noise sigma^2, constant c, outputscale s, lengthscale l)  — *actual* values
  raw   = (raw_noise = noise, constant, raw_s, raw_l), s = softplus(raw_s), l = softplus(raw_l)
  kernel_id: 0 RBF, 1 Matern-1/2, 2 Matern-3/2, 3 Matern-5/2
  acq_id:    0 EI, 1 UCB (acq_param = beta), 2 PI
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erfc

RBF, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3
EI, UCB, PI = 0, 1, 2
INV_SQRT_2PI = 0.3989422804014327
INV_SQRT2 = 0.7071067811865476
LOG_2PI = 1.8378770664093453
NOISE_LB = 1e-4          # GreaterThan(1e-4) handed to L-BFGS-B as a bound
PRIOR_A, PRIOR_RATE = 1.1, 0.05
VAR_FLOOR = 1e-9         # botorch 0.8.0 analytic EI: variance.clamp_min(1e-9)

ST_JITTER, ST_CHOL_FAIL, ST_NONFINITE, ST_ALL_MASKED = 1, 2, 4, 8


def softplus(x):
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))   # torch threshold = 20


def inv_softplus(y):
    y = np.asarray(y, dtype=np.float64)
    return np.where(y > 20.0, y, np.log(np.expm1(np.minimum(y, 20.0))))


def sigmoid(x):
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def kernel_id_from(nu=2.5, covar="matern") -> int:
    """BO.py:133-134 (nu > 2.5 => rbf) and BO.py:61-64."""
    if covar == "rbf" or nu > 2.5:
        return RBF
    return {0.5: MATERN12, 1.5: MATERN32, 2.5: MATERN52}[nu]


def sqdist(A, B):
    """Direct-difference squared distance (exact on a lattice)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    dx = A[:, None, 0] - B[None, :, 0]
    dy = A[:, None, 1] - B[None, :, 1]
    return dx * dx + dy * dy


def kernel_from_r2(r2, s, ell, kernel_id):
    """s * k(r / ell); canonical operation order shared with csrc/gp_common.cuh."""
    if kernel_id == RBF:
        return s * np.exp(r2 * (-0.5 / (ell * ell)))
    r = np.sqrt(r2)
    if kernel_id == MATERN12:
        return s * np.exp(-(r / ell))
    if kernel_id == MATERN32:
        a = (math.sqrt(3.0) / ell) * r
        return s * ((1.0 + a) * np.exp(-a))
    if kernel_id == MATERN52:
        a = (math.sqrt(5.0) / ell) * r
        return s * ((1.0 + a + (a * a) / 3.0) * np.exp(-a))
    raise ValueError(kernel_id)


def dkernel_dell_from_r2(r2, s, ell, kernel_id):
    """s * d k(r/ell) / d ell."""
    r = np.sqrt(r2)
    if kernel_id == RBF:
        q = r2 / (ell * ell)
        return s * q * np.exp(-0.5 * q) / ell
    if kernel_id == MATERN12:
        d = r / ell
        return s * d * np.exp(-d) / ell
    if kernel_id == MATERN32:
        a = (math.sqrt(3.0) / ell) * r
        return s * (a * a) * np.exp(-a) / ell
    if kernel_id == MATERN52:
        a = (math.sqrt(5.0) / ell) * r
        return s * ((a * a) / 3.0) * (1.0 + a) * np.exp(-a) / ell
    raise ValueError(kernel_id)


def assemble(X, theta, kernel_id):
    noise, c, s, ell = theta
    K = kernel_from_r2(sqdist(X, X), s, ell, kernel_id)
    K[np.diag_indices_from(K)] += noise
    return K


def cholesky_lower(K):
    """Plain right-looking Cholesky; returns (L, ok)."""
    try:
        return np.linalg.cholesky(K), True
    except np.linalg.LinAlgError:
        return None, False


def chol_safe(K):
    """gpytorch psd_safe_cholesky (fp64): plain, then diagonal jitter 1e-8, 1e-7,
    1e-6 (total, not cumulative).  Returns (L, status bits)."""
    L, ok = cholesky_lower(K)
    if ok:
        return L, 0
    for t in range(3):
        Kj = K.copy()
        Kj[np.diag_indices_from(Kj)] += 1e-8 * 10 ** t
        L, ok = cholesky_lower(Kj)
        if ok:
            return L, ST_JITTER
    return np.full_like(K, np.nan), ST_JITTER | ST_CHOL_FAIL


def gp_state(X, y, theta, kernel_id):
    """K-A: (L, alpha = Khat^-1 (y - c), Linv = L^-1, status)."""
    from scipy.linalg import solve_triangular
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    K = assemble(X, theta, kernel_id)
    L, status = chol_safe(K)
    if status & ST_CHOL_FAIL:
        n = len(y)
        return L, np.full(n, np.nan), L.copy(), status
    z = solve_triangular(L, y - theta[1], lower=True)
    alpha = solve_triangular(L.T, z, lower=False)
    Linv = solve_triangular(L, np.eye(len(y)), lower=True)
    return L, alpha, Linv, status


def posterior(grid, X, theta, kernel_id, L, alpha):
    """A.2: mu = c + k*^T alpha; var = s - ||L^-1 k*||^2 (latent f)."""
    from scipy.linalg import solve_triangular
    noise, c, s, ell = theta
    Ks = kernel_from_r2(sqdist(X, grid), s, ell, kernel_id)      # (n, G)
    mu = c + Ks.T @ alpha
    V = solve_triangular(L, Ks, lower=True)
    var = s - np.sum(V * V, axis=0)
    return mu, var


def norm_pdf(u):
    return np.exp(-0.5 * (u * u)) * INV_SQRT_2PI


def norm_cdf(u):
    """Phi(u) = erfc(-u/sqrt2)/2 — the same function as torch's 0.5(1+erf(u/sqrt2))
    (botorch analytic EI) evaluated without cancellation in the left tail."""
    return 0.5 * erfc(-u * INV_SQRT2)


def acquisition(mu, var, best_f, acq_id=EI, acq_param=0.0, var_floor=VAR_FLOOR):
    sigma = np.sqrt(np.maximum(var, var_floor))
    if acq_id == UCB:
        return mu + math.sqrt(acq_param) * sigma
    u = (mu - best_f) / sigma
    if acq_id == PI:
        return norm_cdf(u)
    return sigma * (norm_pdf(u) + u * norm_cdf(u))


def argmax_masked(values, mask=None):
    """Tie rule: lowest flat index; masked candidates and NaN never win.
    Returns (idx, val, status) with idx = -1 when nothing is selectable."""
    v = np.array(values, dtype=np.float64, copy=True)
    status = 0
    if mask is not None:
        v[np.asarray(mask, dtype=bool)] = -np.inf
    if np.any(np.isnan(v)) or np.any(v == np.inf):
        status |= ST_NONFINITE
    v[np.isnan(v)] = -np.inf
    if not np.any(v > -np.inf):
        return -1, -np.inf, status | ST_ALL_MASKED
    idx = int(np.argmax(v))            # numpy returns the first maximum
    return idx, float(v[idx]), status


def observed_mask(grid, X):
    """Candidates that coincide with an observation (r^2 == 0 exactly)."""
    return np.any(sqdist(X, grid) == 0.0, axis=0)


def grid_step(grid, X, y, theta, kernel_id, acq_id=EI, acq_param=0.0, var_floor=VAR_FLOOR,
              mask_observed=True, extra_mask=None, return_fields=False):
    """One fused K-A + K-B evaluation: returns (idx, acq value, status[, mu, var, acq])."""
    L, alpha, _, st = gp_state(X, y, theta, kernel_id)
    if st & ST_CHOL_FAIL:
        out = (-1, float("nan"), st)
        return out + (None, None, None) if return_fields else out
    mu, var = posterior(grid, X, theta, kernel_id, L, alpha)
    a = acquisition(mu, var, np.max(y), acq_id, acq_param, var_floor)
    mask = observed_mask(grid, X) if mask_observed else np.zeros(len(grid), dtype=bool)
    if extra_mask is not None:
        mask = mask | np.asarray(extra_mask, dtype=bool)
    idx, val, st2 = argmax_masked(a, mask)
    if return_fields:
        return idx, val, st | st2, mu, var, a
    return idx, val, st | st2


# ------------------------------------------------------------------ marginal likelihood

def mll_value_grad(X, y, raw, kernel_id):
    """A.3: value of ExactMarginalLogLikelihood (+ Gamma(1.1,0.05) noise prior),
    divided by n, and its gradient w.r.t. raw = (noise, c, raw_s, raw_l).
    Returns (value, grad(4), status)."""
    from scipy.linalg import solve_triangular
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = len(y)
    noise, c, raw_s, raw_l = [float(v) for v in raw]
    s = float(softplus(raw_s))
    ell = float(softplus(raw_l))
    r2 = sqdist(X, X)
    Kl = kernel_from_r2(r2, 1.0, ell, kernel_id)          # unit-scale kernel K_l
    Ks = s * Kl
    K = Ks.copy()
    K[np.diag_indices_from(K)] += noise
    L, status = chol_safe(K)
    if status & ST_CHOL_FAIL:
        return float("nan"), np.full(4, np.nan), status
    res = y - c
    z = solve_triangular(L, res, lower=True)
    alpha = solve_triangular(L.T, z, lower=False)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    ll = -0.5 * (z @ z) - 0.5 * logdet - 0.5 * n * LOG_2PI
    lp = PRIOR_A * math.log(PRIOR_RATE) + (PRIOR_A - 1.0) * math.log(noise) - PRIOR_RATE * noise - math.lgamma(PRIOR_A)
    value = (ll + lp) / n
    Linv = solve_triangular(L, np.eye(n), lower=True)
    Kinv = Linv.T @ Linv
    A = np.outer(alpha, alpha) - Kinv
    g_noise = 0.5 * np.trace(A) + (PRIOR_A - 1.0) / noise - PRIOR_RATE
    g_c = np.sum(alpha)
    g_s = 0.5 * np.sum(A * Kl) * float(sigmoid(raw_s))                # d(s K_l)/d raw_s = K_l sigmoid(raw_s)
    dK = dkernel_dell_from_r2(r2, s, ell, kernel_id)
    g_l = 0.5 * np.sum(A * dK) * float(sigmoid(raw_l))
    return value, np.array([g_noise, g_c, g_s, g_l]) / n, status


DEFAULT_RAW = (2.0, 0.0, 0.0, 0.0)      # gpytorch defaults: noise 2.0, constant 0, raw_s = raw_l = 0


def fit_scipy(X, y, kernel_id, raw0=DEFAULT_RAW, maxiter=15000):
    """fit_gpytorch_model (BO.py:144,189): scipy L-BFGS-B on -MLL, analytic gradient,
    bound noise >= 1e-4, scipy defaults (m=10, ftol=2.2e-9, gtol=1e-5)."""
    from scipy.optimize import minimize

    def fun(p):
        v, g, st = mll_value_grad(X, y, p, kernel_id)
        if not np.isfinite(v):
            return 1e300, np.zeros(4)
        return -v, -g

    res = minimize(fun, np.array(raw0, dtype=np.float64), jac=True, method="L-BFGS-B",
                   bounds=[(NOISE_LB, None), (None, None), (None, None), (None, None)],
                   options=dict(maxiter=maxiter))
    return res.x, -res.fun, res


def fit_default(X, y, kernel_id):
    """The ``fit`` callable ``bo_loop`` expects: raw parameters only."""
    return fit_scipy(X, y, kernel_id)[0]


def raw_to_theta(raw):
    raw = np.asarray(raw, dtype=np.float64)
    return np.array([raw[0], raw[1], float(softplus(raw[2])), float(softplus(raw[3]))])


# ------------------------------------------------------------------ BO loop on a grid

def bo_loop(grid, obj_values, X0, y0, max_iter, theta=None, kernel_id=MATERN52, acq_id=EI, acq_param=0.0,
            tol=0.1, optimal_y=100.0, var_floor=VAR_FLOOR, fit=None, record=False, theta_rest=None):
    """BOloop / BOloop_fixparam semantics (BO.py:141-278, 380-461) with the
    continuous ``optimize_acqf`` replaced by a dense arg-max over ``grid``.

    grid (G,2) f64 candidates; obj_values (G,) objective at each candidate
    (float32-representable, as ``torch.tensor(obj_func(new_x), dtype=float32)``, BO.py:104).
    theta (4,) fixed for every iteration when given (the *intent* of BOloop_fixparam); theta (k,4): row i is used
    by acquisition i (replay of recorded per-iteration fits);
    when None, ``fit(X, y, kernel_id) -> raw`` is called before every acquisition
    (BOloop: initial fit :144, refit from the default start after every append :176-189).

    Returns dict(yall (max_iter+2,), xall (nit,2), idx (nit,), nit, status, thetas)."""
    X = np.array(X0, dtype=np.float64).reshape(-1, 2)
    y = np.array(y0, dtype=np.float64).reshape(-1)
    yopt = float(np.max(y))
    yall = [yopt]
    picked = []
    thetas = []
    status = 0
    it = 0
    error = 100.0
    while error > tol and it < max_iter:
        it += 1
        if theta is None:
            th = raw_to_theta(fit(X, y, kernel_id))
        elif np.ndim(theta) == 2:
            th = np.asarray(theta)[it - 1]          # replay: the hyper-parameters of acquisition it-1, given per iteration
        else:
            th = theta
        if theta_rest is not None and it > 1:
            th = theta_rest          # BOloop_fixparam as coded: defaults after the first re-created model (BO.py:424-429)
        thetas.append(np.array(th, dtype=np.float64))
        idx, val, st = grid_step(grid, X, y, th, kernel_id, acq_id, acq_param, var_floor)
        status |= st
        if idx < 0:
            it -= 1
            break
        picked.append(idx)
        X = np.vstack([X, grid[idx]])
        y = np.append(y, np.float64(np.float32(obj_values[idx])))
        yopt = float(np.max(y))
        yall.append(yopt)
        error = optimal_y - yopt
    yall.append(yopt)                                           # BO.py:262 / :446
    yall = np.array(yall, dtype=np.float64)
    if len(yall) < max_iter + 2:                                # BO.py:267-271
        yall = np.concatenate([yall, yall[-1] * np.ones(max_iter + 2 - len(yall))])
    out = dict(yall=yall, xall=np.array([grid[i] for i in picked]).reshape(-1, 2), idx=np.array(picked, dtype=np.int64),
               nit=it, status=status, X=X, y=y)
    if record:
        out["thetas"] = np.array(thetas)
    return out
