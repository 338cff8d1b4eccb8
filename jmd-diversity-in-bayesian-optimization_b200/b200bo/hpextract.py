"""Batched hyper-parameter extraction: the ``HPextract`` fan-out of Scripts/Experiment3-2-1.py:63-87,296-320.

The reference maps ``HPextract(X, f)`` over 200 growing training sets (1 010 ... 1 209 points) on ipyparallel
engines: evaluate the objective, build ``SingleTaskGP`` with a Matern-3/2 kernel, ``fit_gpytorch_model`` once,
record the four raw parameters in a ``results.hp_result``.  Here every set is one CTA of one launch
(csrc/gp_fit_large.cu, or the in-register kernel below 128 points).
"""
from __future__ import annotations

import numpy as np
import torch

from . import fit, ops

PARAM_NAMES = ("likelihood.noise_covar.raw_noise", "mean_module.raw_constant", "covar_module.raw_outputscale",
               "covar_module.base_kernel.raw_lengthscale")


def hp_extract_batched(train_sets, f, nu=1.5, covar="matern", device="cuda:0"):
    """train_sets: list of (n_i, 2) arrays; f: objective callable (objectives.generate_cont()).  Returns a list of
    ``results.hp_result`` (``hyperparameters`` is None where the fit failed, as HPextract's last resort does).

    The reference shuffles each set with ``random.sample`` before fitting (BO.generate_initial_data); the order of
    the observations does not change the likelihood, so the sets are used as given."""
    import results
    dev = torch.device(device)
    T = len(train_sets)
    n = np.array([len(s) for s in train_sets], dtype=np.int32)
    n_max = int(n.max())
    X = np.zeros((T, n_max, 2))
    y = np.zeros((T, n_max))
    for t, s in enumerate(train_sets):
        s = np.asarray(s, dtype=np.float64)
        X[t, :n[t]] = s
        y[t, :n[t]] = np.asarray(f(s), dtype=np.float32).astype(np.float64)     # Y is cast to float32 (BO.py:26-27)
    kid = ops.kernel_id_from(nu, covar)
    Xd, yd, nd = (torch.as_tensor(a, device=dev) for a in (X, y, n))
    status = torch.zeros(T, dtype=torch.int32, device=dev)
    if n_max >= fit.LARGE_N:
        r = fit.fit_large(Xd, yd, nd, kid, status=status)
        raw, mll = r["raw"], r["mll"]
    else:
        fz = fit.LBFGSFitter(T, dev)
        theta = torch.zeros((T, 4), dtype=torch.float64, device=dev)
        fz.fit_into(Xd, yd, nd, kid, theta, status)
        raw, mll = fz.raw, fz.mll
    raw = raw.cpu().numpy()
    okay = torch.isfinite(mll).cpu().numpy() & ((status.cpu().numpy() & ops.ST_CHOL_FAIL) == 0)
    out = []
    for t in range(T):
        h = results.hp_result()
        h.hyperparameters = {k: [float(raw[t, i])] for i, k in enumerate(PARAM_NAMES)} if okay[t] else None
        out.append(h)
    return out
