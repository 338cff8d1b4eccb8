"""Independent formulations of the GP arithmetic the oracle restates (oracle/gp.py; the botorch 0.8.0 / gpytorch 1.9.0
stack is absent, so the oracle stays 'parity unpinned' — these checks bound the risk of a restatement error):

* EI / PI / the normal tail against 50-digit mpmath, far into both tails (the arg-max lives in the left tail late in a run);
* the marginal log-likelihood (value AND gradient) against torch autograd of a formulation that shares no code with the
  oracle: torch.distributions.MultivariateNormal / Gamma log_prob on a covariance built with torch.cdist, softplus through
  torch.nn.functional — the same public building blocks gpytorch's ExactMarginalLogLikelihood reduces to;
* the posterior against the explicit-inverse textbook formulas.
"""
import math

import numpy as np
import pytest
import torch

from oracle import gp as ogp, wildcat as owc


def _data(golden, n, seed=0):
    rng = np.random.default_rng(seed)
    surf = golden("wildcat.npz")["surf_0"]
    X = np.unique(rng.integers(0, 100, size=(n, 2)).astype(float), axis=0)
    return X, owc.lattice_lookup(surf, X).astype(np.float64)


def test_ei_pi_and_normal_tail_against_mpmath():
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.default_rng(0)
    us = np.concatenate([np.linspace(-36.0, 8.0, 177), rng.uniform(-12, 4, 60)])
    for sigma in (1e-4, 0.37, 21.0):
        mu = us * sigma + 3.25
        var = np.full_like(mu, sigma * sigma)
        ei = ogp.acquisition(mu, var, 3.25, ogp.EI)
        pi = ogp.acquisition(mu, var, 3.25, ogp.PI)
        for k, m in enumerate(mu):
            u = (mp.mpf(float(m)) - mp.mpf(3.25)) / mp.sqrt(mp.mpf(float(var[k])))
            Phi = mp.erfc(-u / mp.sqrt(2)) / 2
            phi = mp.exp(-u * u / 2) / mp.sqrt(2 * mp.pi)
            ei_x = mp.sqrt(mp.mpf(float(var[k]))) * (phi + u * Phi)
            # u itself carries ~2 ulps of fp64 rounding, which the tail amplifies by u^2
            assert abs(pi[k] - float(Phi)) <= 4e-16 * (10 + float(u * u)) * float(Phi) + 1e-305, (float(u), pi[k], float(Phi))
            # phi + u Phi cancels in the left tail: the fp64 evaluation keeps |u|^2 ulps of the two terms
            tol = 4e-16 * float(mp.sqrt(mp.mpf(float(var[k])))) * float(phi) * (4 + float(u * u)) + 1e-305
            assert abs(ei[k] - float(ei_x)) <= tol + 4e-16 * (10 + float(u * u)) * float(ei_x), (float(u), ei[k], float(ei_x))


def _torch_mll(X, y, raw, kid):
    """(MLL value, gradient wrt raw) by autograd; raw = (noise, constant, raw_outputscale, raw_lengthscale)."""
    Xt, yt = torch.as_tensor(X, dtype=torch.float64), torch.as_tensor(y, dtype=torch.float64)
    r = torch.tensor(raw, dtype=torch.float64, requires_grad=True)
    noise, c = r[0], r[1]
    s, ell = torch.nn.functional.softplus(r[2]), torch.nn.functional.softplus(r[3])
    d = torch.cdist(Xt / ell, Xt / ell, compute_mode="donot_use_mm_for_euclid_dist")
    d = d - torch.diag(torch.diag(d))                                    # exact zeros on the diagonal
    if kid == ogp.RBF:
        K = torch.exp(-0.5 * d ** 2)
    elif kid == ogp.MATERN12:
        K = torch.exp(-d)
    elif kid == ogp.MATERN32:
        K = (1 + math.sqrt(3) * d) * torch.exp(-math.sqrt(3) * d)
    else:
        K = (1 + math.sqrt(5) * d + 5.0 / 3.0 * d ** 2) * torch.exp(-math.sqrt(5) * d)
    cov = s * K + noise * torch.eye(len(y), dtype=torch.float64)
    ll = torch.distributions.MultivariateNormal(c.expand(len(y)), covariance_matrix=cov).log_prob(yt)
    lp = torch.distributions.Gamma(torch.tensor(1.1, dtype=torch.float64), torch.tensor(0.05, dtype=torch.float64)).log_prob(noise)
    val = (ll + lp) / len(y)
    val.backward()
    return float(val.detach()), r.grad.numpy()


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
@pytest.mark.parametrize("n,raw", [(12, (0.3, 30.0, 3.0, 2.0)), (40, (2.0, 0.0, 0.0, 0.0)), (60, (1.762e-3, 41.83, 190.43, 9.8695)),
                                   (25, (1e-4, 55.0, 25.0, -1.5))])
def test_mll_value_and_gradient_against_torch_autograd(golden, kid, n, raw):
    X, y = _data(golden, n, seed=n)
    v, g, st = ogp.mll_value_grad(X, y, np.array(raw), kid)
    vt, gt = _torch_mll(X, y, raw, kid)
    assert st == 0
    assert abs(v - vt) <= 1e-9 * abs(vt)
    # the analytic gradient and autograd agree to the conditioning of K (noise 1e-3 with outputscale 190: cond ~ 1e7)
    assert np.allclose(g, gt, rtol=2e-6, atol=1e-9), (g, gt)


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_posterior_against_the_explicit_inverse(golden, lattice_grid, kid):
    X, y = _data(golden, 30, seed=7)
    theta = (0.05, 41.83, 150.0, 7.5)
    L, alpha, _, st = ogp.gp_state(X, y, theta, kid)
    grid = lattice_grid[::11]
    mu, var = ogp.posterior(grid, X, theta, kid, L, alpha)
    d = lambda A, B: np.sqrt(((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)) / theta[3]
    k = {0: lambda r: np.exp(-0.5 * r * r), 1: lambda r: np.exp(-r), 2: lambda r: (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r),
         3: lambda r: (1 + np.sqrt(5) * r + 5 / 3 * r * r) * np.exp(-np.sqrt(5) * r)}[kid]
    Kinv = np.linalg.inv(theta[2] * k(d(X, X)) + theta[0] * np.eye(len(y)))
    ks = theta[2] * k(d(grid, X))
    assert st == 0
    assert np.allclose(mu, theta[1] + ks @ Kinv @ (y - theta[1]), rtol=1e-8, atol=1e-8)
    assert np.allclose(var, theta[2] - np.einsum("gi,ij,gj->g", ks, Kinv, ks), rtol=1e-6, atol=1e-7 * theta[2])
