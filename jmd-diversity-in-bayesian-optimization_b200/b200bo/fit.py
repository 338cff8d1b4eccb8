"""Batched GP hyper-parameter fitting on device: K-C (value + gradient of the marginal likelihood)
driven by one projected L-BFGS state machine per trial (csrc/lbfgs.cu).

Host role: enqueue `mll_value_grad` + `lbfgs_step` rounds; a device-side flag array tells which trials
still need evaluations, and the host polls it every ``poll`` rounds only to stop enqueueing early.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops
from ._lib import check

DEFAULT_RAW = (2.0, 0.0, 0.0, 0.0)   # fresh gpytorch defaults every fit (BO.py:176-181): noise 2, c 0, raw_s = raw_l = 0
NOISE_LB = 1e-4                      # botorch default GaussianLikelihood: GreaterThan(1e-4) handed to L-BFGS-B


class LBFGSFitter:
    def __init__(self, T: int, device, raw0=DEFAULT_RAW, gtol=1e-5, ftol=2.220446049250313e-9, max_iter=200,
                 max_ls=30, max_rounds=600, poll=8, fused=True):
        lib = _lib.load()
        self.T = int(T)
        dev = torch.device(device)
        self.state = torch.zeros(max(int(lib.b200bo_lbfgs_state_bytes(self.T)), 8), dtype=torch.uint8, device=dev)
        self.xt = torch.empty((self.T, 4), dtype=torch.float64, device=dev)
        self.value = torch.empty(self.T, dtype=torch.float64, device=dev)
        self.grad = torch.empty((self.T, 4), dtype=torch.float64, device=dev)
        self.n_eval = torch.empty(self.T, dtype=torch.int32, device=dev)
        self.raw = torch.empty((self.T, 4), dtype=torch.float64, device=dev)
        self.mll = torch.empty(self.T, dtype=torch.float64, device=dev)
        self.iters = torch.zeros((self.T, 2), dtype=torch.int32, device=dev)
        self.kc_status = torch.zeros(self.T, dtype=torch.int32, device=dev)
        r0 = torch.as_tensor(raw0, dtype=torch.float64, device=dev)
        self.raw0 = r0.contiguous()
        self.raw0_per_trial = 1 if r0.dim() == 2 else 0
        self.gtol, self.ftol, self.max_iter, self.max_ls = float(gtol), float(ftol), int(max_iter), int(max_ls)
        self.max_rounds, self.poll = int(max_rounds), int(poll)
        self.last_launches = 0
        self.last_rounds = 0
        self.fused = bool(fused)
        self.lpt = os.environ.get("B200BO_FIT_LPT", "1") != "0"     # launch the fused fit longest-expected-first
        # optional accounting (bench.py): algorithmic flops of the evaluations run so far, n^3 + 6 n^2 each
        self.account = False
        self.flops = torch.zeros(1, dtype=torch.float64, device=dev)
        self.evals = torch.zeros(1, dtype=torch.float64, device=dev)

    def fit_into(self, X, y, n, kernel_id, theta_out, status=None, n_hint=0):
        """Fit every trial with n[t] > 0 from the default start; writes ACTUAL theta (T,4) into theta_out."""
        with torch.cuda.device(X.device):
            return self._fit_into(X, y, n, kernel_id, theta_out, status, n_hint)

    def _fit_into(self, X, y, n, kernel_id, theta_out, status, n_hint):
        lib = _lib.load()
        p, s = ops._ptr, ops._stream()
        T = self.T
        if self.fused:
            # one launch: every trial runs its own L-BFGS to convergence inside its CTA (csrc/gp_fit.cu:gp_fit_sweep_kernel)
            # longest expected fit first (the evaluation counts of this trial's previous fit predict the next one's): the
            # launch then ends with its shortest fits instead of idling on a long one that started late
            order = None
            if self.lpt:
                order = torch.argsort(torch.where(n > 0, self.iters[:, 1], torch.zeros_like(self.iters[:, 1])),
                                      descending=True, stable=True).to(torch.int32)
            check(lib.b200bo_gp_fit_ordered(p(X), p(y), p(n), p(self.raw0), self.raw0_per_trial, p(theta_out), p(self.raw),
                                            p(self.mll), p(self.iters), p(status if status is not None else self.kc_status),
                                            T, y.shape[1], int(n_hint), int(kernel_id), NOISE_LB, self.gtol, self.ftol,
                                            self.max_iter, self.max_ls, self.max_rounds, p(order), s), "gp_fit")
            self.last_launches = 1 + (2 if order is not None else 0)
            self.last_rounds = 0
            if self.account:
                nf = n.clamp(min=0).double()
                ev = torch.where(n > 0, self.iters[:, 1], torch.zeros_like(self.iters[:, 1])).double()
                self.flops += (ev * (nf ** 3 + 6.0 * nf ** 2)).sum()
                self.evals += ev.sum()
            return theta_out
        check(lib.b200bo_lbfgs_init(p(self.state), p(self.xt), p(self.raw0), self.raw0_per_trial, p(self.n_eval), p(n), T,
                                    NOISE_LB, s), "lbfgs_init")
        launches = 1
        rounds = 0
        n_max = y.shape[1]
        while rounds < self.max_rounds:
            check(lib.b200bo_mll_value_grad(p(X), p(y), p(self.n_eval), p(self.xt), p(self.value), p(self.grad),
                                            p(self.kc_status), T, n_max, int(n_hint), int(kernel_id), s), "mll_value_grad")
            check(lib.b200bo_lbfgs_step(p(self.state), p(self.xt), p(self.value), p(self.grad), p(self.n_eval), p(n),
                                        p(theta_out), p(self.raw), p(self.mll), p(self.iters), T, NOISE_LB, self.gtol,
                                        self.ftol, self.max_iter, self.max_ls, s), "lbfgs_step")
            launches += 2
            rounds += 1
            if rounds % self.poll == 0 and not bool((self.n_eval > 0).any()):
                break
        else:
            # round budget exhausted: one more evaluation and a forced finish so every trial reports
            check(lib.b200bo_mll_value_grad(p(X), p(y), p(self.n_eval), p(self.xt), p(self.value), p(self.grad),
                                            p(self.kc_status), T, n_max, int(n_hint), int(kernel_id), s), "mll_value_grad")
            check(lib.b200bo_lbfgs_step(p(self.state), p(self.xt), p(self.value), p(self.grad), p(self.n_eval), p(n),
                                        p(theta_out), p(self.raw), p(self.mll), p(self.iters), T, NOISE_LB, self.gtol,
                                        self.ftol, self.max_iter, -1, s), "lbfgs_step")
            launches += 2
        self.last_launches = launches
        self.last_rounds = rounds
        return theta_out


LARGE_N = 128          # observation counts above the in-register K-C kernels (B200BO_NMAX) take the blocked path


def mll_value_grad_large(X, y, n, raw, kernel_id, status=None):
    """K-C for n up to 4000 (csrc/gp_fit_large.cu): X (T,n_max,2), y (T,n_max), n (T) i32, raw (T,4) ->
    value (T), grad (T,4)."""
    with torch.cuda.device(X.device):
        return _mll_value_grad_large(X, y, n, raw, kernel_id, status)


def _mll_value_grad_large(X, y, n, raw, kernel_id, status):
    lib = _lib.load()
    p, s = ops._ptr, ops._stream()
    T, n_max = y.shape
    ws = torch.empty(int(lib.b200bo_gp_fit_large_workspace_bytes(T, n_max)), dtype=torch.uint8, device=X.device)
    value = torch.empty(T, dtype=torch.float64, device=X.device)
    grad = torch.empty((T, 4), dtype=torch.float64, device=X.device)
    check(lib.b200bo_mll_value_grad_large(p(X), p(y), p(n), p(raw), p(value), p(grad), p(status), T, n_max, int(kernel_id),
                                          p(ws), ws.numel(), s), "mll_value_grad_large")
    return value, grad


def fit_large(X, y, n, kernel_id, raw0=DEFAULT_RAW, gtol=1e-5, ftol=2.220446049250313e-9, max_iter=200, max_ls=30,
              max_rounds=600, status=None):
    """Fit T large models (one CTA each) from ``raw0``; returns dict(theta, raw, mll, iters) of device tensors.
    The batched counterpart of Experiment3-2-1.py's HPextract fan-out (one fit per ipyparallel engine there)."""
    with torch.cuda.device(X.device):
        return _fit_large(X, y, n, kernel_id, raw0, gtol, ftol, max_iter, max_ls, max_rounds, status)


def _fit_large(X, y, n, kernel_id, raw0, gtol, ftol, max_iter, max_ls, max_rounds, status):
    lib = _lib.load()
    p, s = ops._ptr, ops._stream()
    dev = X.device
    T, n_max = y.shape
    ws = torch.empty(int(lib.b200bo_gp_fit_large_workspace_bytes(T, n_max)), dtype=torch.uint8, device=dev)
    r0 = torch.as_tensor(raw0, dtype=torch.float64, device=dev).contiguous()
    theta = torch.zeros((T, 4), dtype=torch.float64, device=dev)
    raw = torch.zeros((T, 4), dtype=torch.float64, device=dev)
    mll = torch.zeros(T, dtype=torch.float64, device=dev)
    iters = torch.zeros((T, 2), dtype=torch.int32, device=dev)
    check(lib.b200bo_gp_fit_large(p(X), p(y), p(n), p(r0), 1 if r0.dim() == 2 else 0, p(theta), p(raw), p(mll), p(iters),
                                  p(status), T, n_max, int(kernel_id), NOISE_LB, float(gtol), float(ftol), int(max_iter),
                                  int(max_ls), int(max_rounds), p(ws), ws.numel(), s), "gp_fit_large")
    return {"theta": theta, "raw": raw, "mll": mll, "iters": iters}
