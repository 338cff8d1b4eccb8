"""Batched BO engine: T independent trials advanced in lock-step on one GPU.

This is the batched counterpart of ``BO.BOloop`` / ``BO.BOloop_fixparam`` (Src/BO.py:123-278,
380-461): the reference runs one trial per process (ipyparallel / ray fan-out,
Scripts/Experiment2.py:75-92); here the trial is the data-parallel axis of every kernel launch.
All state lives in HBM for the whole run; the host only enqueues launches.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass

import torch

from . import ops


@dataclass
class BOResult:
    yall: torch.Tensor      # (T, max_iter+2) f64 best-so-far curve, padded as BO.py:262-271
    picked: torch.Tensor    # (T, max_iter) int32 flat candidate index per iteration, -1 = none
    nit: torch.Tensor       # (T,) int32 iterations run
    status: torch.Tensor    # (T,) int32 B200BO_ST_* bits
    X: torch.Tensor         # (T, n_max, 2) f64 observations (initial set first)
    y: torch.Tensor         # (T, n_max) f64
    n: torch.Tensor         # (T,) int32 observations at the end
    thetas: torch.Tensor | None = None   # (T, max_iter+1, 4), fit mode: [:, k] = the fit on n0+k observations — used by
    #                                       acquisition k (k < nit); [:, 1:nit+1] = what the reference records after every
    #                                       append (res_hp.addhp, Src/BO.py:252)


def _on_device(fn):
    """Run a BatchedBO method with the batch's GPU as torch's current device: the C ABI launches on the current device
    and ops._stream() hands it that device's current stream, whatever device the caller had selected."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapper


class BatchedBO:
    """grid (G,2) f64 candidates; obj (S,G) f32 objective at every candidate for S surfaces; surf_id (T,)
    which surface a trial optimises; X0 (T,n0,2) f64 and y0 (T,n0) f64 the initial sets; theta (T,4) or (4,)
    ACTUAL hyper-parameters held fixed, or None to refit before every acquisition (``fitter``)."""

    def __init__(self, grid, obj, surf_id, X0, y0, max_iter, **kwargs):
        if grid.device.type != "cuda":
            raise ops._lib.B200boError("BatchedBO needs CUDA tensors: the engine has no CPU path")
        self.device = grid.device
        with torch.cuda.device(self.device):
            self._setup(grid, obj, surf_id, X0, y0, max_iter, **kwargs)

    def _setup(self, grid, obj, surf_id, X0, y0, max_iter, theta=None, kernel_id=ops.MATERN52, acq_id=ops.EI,
               acq_param=0.0, tol=0.1, optimal_y=100.0, var_floor=ops.VAR_FLOOR, fitter=None, check_every=8,
               lattice="auto", append=True, theta_rest=None, posterior="full"):
        dev = grid.device
        self.grid = grid.contiguous()
        self.obj = obj.contiguous()
        self.T, self.n0 = y0.shape
        self.G = grid.shape[0]
        self.max_iter = int(max_iter)
        self.n_max = self.n0 + self.max_iter
        if self.n_max > ops.NMAX:
            raise ops._lib.B200boError(f"n0 + max_iter = {self.n_max} exceeds the engine limit {ops.NMAX}")
        T, n_max = self.T, self.n_max
        f64, i32 = torch.float64, torch.int32
        self.surf_id = surf_id.to(device=dev, dtype=i32).contiguous()
        self.X = torch.zeros((T, n_max, 2), dtype=f64, device=dev)
        self.y = torch.zeros((T, n_max), dtype=f64, device=dev)
        self.X[:, :self.n0] = X0
        self.y[:, :self.n0] = y0
        self.n = torch.full((T,), self.n0, dtype=i32, device=dev)
        self.fixed = theta is not None
        if self.fixed:
            th = torch.as_tensor(theta, dtype=f64, device=dev)
            self.theta = (th.expand(T, 4) if th.dim() == 1 else th).contiguous().clone()
        else:
            if fitter is None:
                raise ValueError("theta=None needs a fitter")
            self.theta = torch.zeros((T, 4), dtype=f64, device=dev)
        self.fitter = fitter
        # reference_quirk: BOloop_fixparam as coded copies fixparam into the first model only; every model
        # re-created after an append keeps gpytorch's defaults (Src/BO.py:392-395 vs :424-429). theta_rest, when
        # given, replaces theta from the second iteration on.
        self.theta_rest = None
        if theta_rest is not None:
            tr = torch.as_tensor(theta_rest, dtype=f64, device=dev)
            self.theta_rest = (tr.expand(T, 4) if tr.dim() == 1 else tr).contiguous().clone()
            self.theta_first = self.theta.clone()
        self.kernel_id, self.acq_id, self.acq_param = int(kernel_id), int(acq_id), float(acq_param)
        self.tol, self.optimal_y, self.var_floor = float(tol), float(optimal_y), float(var_floor)
        # candidates on a regular lattice (data_gen.constructX) -> kernel values come from a table
        self.lattice = ops.Lattice.from_points(self.grid) if isinstance(lattice, str) and lattice == "auto" else lattice
        if self.lattice is not None and not bool(self.lattice.contains(self.X[:, :self.n0]).all()):
            # an initial observation off the candidate lattice has no entry in the kernel table (the lattice kernels
            # would only flag it, ST_OFF_LATTICE): such batches take the generic sqrt / exp kernel
            if posterior in ("incremental", "persistent"):
                raise ops._lib.B200boError("initial observations are not nodes of the candidate lattice")
            self.lattice = None
        # fixed theta: L^-1 grows by one row per iteration (rank-1 append) instead of being refactorised
        self.append = bool(append) and self.fixed
        # persistent: the whole loop of a trial in one CTA (csrc/bo_persist.cu) — fixed theta, lattice candidates whose
        # kernel table fits shared memory; "auto" takes it whenever it applies, else the full recomputation
        can_persist = (self.fixed and self.lattice is not None and self.append and theta_rest is None and self.max_iter > 0
                       and ops.bo_fixed_loop_supported(self.lattice, n_max))
        if posterior == "persistent" and not can_persist:
            raise ops._lib.B200boError("posterior='persistent' needs fixed theta, append=True, no theta_rest and a candidate "
                                       "lattice whose kernel table fits shared memory (100x100 with n0+max_iter <= 116)")
        # incremental: the same O(n) update per candidate with the (mu, var) state in HBM and the table in L1 / L2 — for
        # fixed theta on lattices whose table does not fit shared memory (the 200 x 200 lattice of BASELINE configs[3]:
        # 4 x the full recomputation's throughput); "auto" takes it when the state (16 B per trial and candidate) fits
        can_inc = (self.fixed and self.lattice is not None and self.append and theta_rest is None and self.max_iter > 0
                   and 16 * T * self.G <= 48 << 30)
        self.auto_fallback = posterior == "auto"               # auto-selected O(n) paths re-run failed trials through "full"
        if posterior == "auto":
            posterior = "persistent" if can_persist else ("incremental" if can_inc else "full")
        self.persistent = posterior == "persistent"
        self.fallback_trials = 0
        self._kwargs = dict(kernel_id=kernel_id, acq_id=acq_id, acq_param=acq_param, tol=tol, optimal_y=optimal_y,
                            var_floor=var_floor, check_every=check_every, append=append)
        classic = not self.persistent                    # the per-iteration kernels' state (0.7 GB at T = 12 500)
        self.alpha = torch.zeros((T, n_max), dtype=f64, device=dev) if classic else None
        self.wfrag = torch.zeros((T, ops.wfrag_doubles(n_max)), dtype=f64, device=dev) if classic else None
        self.zvec = torch.zeros((T, n_max), dtype=f64, device=dev) if self.append and classic else None
        self.n_redo = torch.zeros(T, dtype=i32, device=dev) if self.append and classic else None
        self.best_f = torch.empty(T, dtype=f64, device=dev)
        self.yall = torch.zeros((T, self.max_iter + 2), dtype=f64, device=dev)
        self.picked = torch.empty((T, max(self.max_iter, 1)), dtype=i32, device=dev)
        self.nit = torch.zeros(T, dtype=i32, device=dev)
        self.status = torch.zeros(T, dtype=i32, device=dev)
        self.arg_idx = torch.empty(T, dtype=i32, device=dev)
        self.arg_val = torch.empty(T, dtype=f64, device=dev)
        self.ws = ops.grid_acq_workspace(T, self.G, dev) if classic else None
        self.check_every = int(check_every)
        self.posterior = posterior
        self.ktab = self.ktab_id = None
        if self.persistent:
            uniq, inv = torch.unique(self.theta, dim=0, return_inverse=True)
            self.ktabY = ops.kernel_table_ysigned(uniq.contiguous(), self.kernel_id, self.lattice)
            self.ktab_id = inv.to(i32).contiguous()
            # trials that share a table run back to back on a CTA (the table is reloaded only when the slot changes)
            self.order = self.slot_start = None
            if uniq.shape[0] > 1:
                self.order = torch.sort(self.ktab_id, stable=True)[1].to(i32).contiguous()
                counts = torch.bincount(self.ktab_id.long(), minlength=uniq.shape[0])
                self.slot_start = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)]).to(i32).contiguous()
            self.ws_loop = ops.bo_fixed_loop_workspace(T, self.lattice, dev)
        elif self.lattice is not None:
            if self.fixed:
                uniq, inv = torch.unique(self.theta, dim=0, return_inverse=True)
                self.ktab = ops.kernel_table(uniq.contiguous(), self.kernel_id, self.lattice)
                self.ktab_id = inv.to(i32).contiguous()
            elif T * self.G * 8 <= 8 << 30:
                self.ktab = torch.empty((T, self.lattice.Gy, self.lattice.Gx), dtype=f64, device=dev)
                self.ktab_id = torch.arange(T, dtype=i32, device=dev)
            else:
                self.lattice = None
        # incremental posterior (fixed theta on a lattice): per-candidate (mu - c, s - var) persist in HBM and every
        # iteration folds in the one new row of L^-1 (csrc/grid_inc.cu) instead of recomputing the O(n^2) product
        self.incremental = posterior == "incremental"
        if self.incremental:
            if not (self.fixed and self.lattice is not None and self.append and self.theta_rest is None):
                raise ops._lib.B200boError("posterior='incremental' needs fixed theta, lattice candidates and append=True")
            G = self.G
            self.mu_s = torch.empty((T, G), dtype=f64, device=dev)
            self.ss_s = torch.empty((T, G), dtype=f64, device=dev)
            self.obs_mask = torch.zeros((T, (G + 31) // 32), dtype=i32, device=dev)
            self.wrow = torch.zeros((T, n_max), dtype=f64, device=dev)
            self.n_ramp = torch.zeros(T, dtype=i32, device=dev)
            uniq, _ = torch.unique(self.theta, dim=0, return_inverse=True)
            self.ktab2 = ops.kernel_table_signed(uniq.contiguous(), self.kernel_id, self.lattice)
            self.ws_inc = ops.grid_inc_workspace(T, G, dev)
        self.thetas = None if self.fixed else torch.zeros((T, self.max_iter + 1, 4), dtype=f64, device=dev)
        self.n_fit = None if self.fixed else torch.zeros(T, dtype=i32, device=dev)
        self.iters_done = 0
        self.launches = 0
        self.profile = None        # set to a list to collect (name, n, start_event, end_event) per launch

    @_on_device
    def reset(self, X0, y0, surf_id=None):
        """Load a new batch of initial sets into the resident buffers (no allocation)."""
        self.X.zero_()
        self.y.zero_()
        self.X[:, :self.n0].copy_(X0, non_blocking=True)
        self.y[:, :self.n0].copy_(y0, non_blocking=True)
        if surf_id is not None:
            self.surf_id.copy_(surf_id, non_blocking=True)
        self.n.fill_(self.n0)
        self.status.zero_()
        self.nit.zero_()
        if self.alpha is not None:
            self.alpha.zero_()
        if self.append and self.wfrag is not None:
            self.wfrag.zero_()
        if self.theta_rest is not None:
            self._switch_theta(self.theta_first)
        if self.incremental:
            self.obs_mask.zero_()
            self.zvec.zero_()
        self.launches = 0
        self.iters_done = 0

    def _timed(self, name, n, fn):
        if self.profile is None:
            return fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.profile.append((name, n, e0, e1))

    def _switch_theta(self, th):
        self.theta.copy_(th)
        if self.lattice is not None and self.fixed:
            uniq, inv = torch.unique(self.theta, dim=0, return_inverse=True)
            self.ktab = ops.kernel_table(uniq.contiguous(), self.kernel_id, self.lattice)
            self.ktab_id = inv.to(torch.int32).contiguous()

    @_on_device
    def step(self, it: int):
        n_hint = self.n0 + it
        if self.theta_rest is not None and it == 1:
            self._switch_theta(self.theta_rest)
        if not self.fixed:
            self._timed("fit", n_hint, lambda: self._fit(it, n_hint))
        self._timed("gp_chol", n_hint, lambda: self._factorise(it, n_hint))
        self._timed("grid_acq", n_hint, lambda: self._acquire(n_hint))
        ops.bo_update(self.grid, self.obj, self.surf_id, self.arg_idx, self.X, self.y, self.n, self.best_f, self.yall,
                      self.picked, self.nit, self.max_iter, it, self.tol, self.optimal_y)
        self.launches += 4

    def _fit(self, it, n_hint):
        """Hyper-parameter fit on the n0+it observations of every trial that ran iteration it-1: the live ones (their
        theta drives acquisition ``it``) and the ones that just stopped (the reference refits and records after every
        append, Src/BO.py:176-189,252)."""
        ops.bo_fit_counts(self.n, self.nit, self.n_fit, it)
        self.fitter.fit_into(self.X, self.y, self.n_fit, self.kernel_id, self.theta, self.status, n_hint)
        self.thetas[:, it] = self.theta
        self.launches += self.fitter.last_launches + 1

    def _factorise(self, it, n_hint):
        if not self.append or it == 0 or (it == 1 and self.theta_rest is not None):
            ops.gp_assemble_chol_into(self.X, self.y, self.n, self.theta, self.kernel_id, self.alpha, self.wfrag,
                                      self.status, n_hint=n_hint)
            if self.append:
                ops.gp_zvec(self.y, self.n, self.theta, self.wfrag, self.zvec)
                self.launches += 1
            return
        ops.gp_chol_append(self.X, self.y, self.n, self.theta, self.kernel_id, self.wfrag, self.zvec, self.alpha,
                           self.n_redo, self.status)
        # trials whose new pivot was not positive (none in practice) are refactorised with the jitter rule
        ops.gp_assemble_chol_into(self.X, self.y, self.n_redo, self.theta, self.kernel_id, self.alpha, self.wfrag,
                                  self.status, n_hint=n_hint)
        ops.gp_zvec(self.y, self.n_redo, self.theta, self.wfrag, self.zvec)
        self.launches += 2

    def _acquire(self, n_hint):
        if self.lattice is not None:
            if not self.fixed:
                ops.kernel_table(self.theta, self.kernel_id, self.lattice, out=self.ktab)
                self.launches += 1
            ops.grid_posterior_acq_argmax_lattice(self.lattice, self.ktab, self.ktab_id, self.X, self.wfrag, self.alpha,
                                                  self.n, self.theta, self.best_f, self.kernel_id, self.acq_id,
                                                  self.acq_param, self.var_floor, status=self.status, ws=self.ws,
                                                  arg_idx=self.arg_idx, arg_val=self.arg_val, n_hint=n_hint)
        else:
            ops.grid_posterior_acq_argmax(self.grid, self.X, self.wfrag, self.alpha, self.n, self.theta, self.best_f,
                                          self.kernel_id, self.acq_id, self.acq_param, self.var_floor,
                                          status=self.status, ws=self.ws, arg_idx=self.arg_idx, arg_val=self.arg_val,
                                          n_hint=n_hint)

    @_on_device
    def _inc_update(self, n_arr, first, acquire):
        ops.gp_chol_append(self.X, self.y, n_arr, self.theta, self.kernel_id, self.wfrag, self.zvec, self.alpha, None,
                           self.status, wrow_out=self.wrow)
        ops.grid_posterior_update_acq_argmax(self.lattice, self.ktab2, self.ktab_id, self.X, self.wrow, self.zvec, n_arr,
                                             self.theta, self.best_f, self.mu_s, self.ss_s, self.obs_mask, first, acquire,
                                             self.acq_id, self.acq_param, self.var_floor, arg_idx=self.arg_idx,
                                             arg_val=self.arg_val, status=self.status, ws=self.ws_inc)
        self.launches += 3 if acquire else 2

    def _step_incremental(self, it: int):
        if it == 0:
            # feed the initial set one observation at a time; the last one also acquires
            for i in range(1, self.n0 + 1):
                self.n_ramp.fill_(i)
                self._timed("grid_inc", i, lambda: self._inc_update(self.n_ramp, i == 1, i == self.n0))
        else:
            self._timed("grid_inc", self.n0 + it, lambda: self._inc_update(self.n, False, True))
        ops.bo_update(self.grid, self.obj, self.surf_id, self.arg_idx, self.X, self.y, self.n, self.best_f, self.yall,
                      self.picked, self.nit, self.max_iter, it, self.tol, self.optimal_y)
        self.launches += 1

    @_on_device
    def run_graphed(self) -> BOResult:
        """The whole run as ONE CUDA graph: the first call runs eagerly (warm-up: lazy module loads, attribute
        settings) and captures the launch sequence of every iteration (no host synchronisation exists inside when
        check_every == 0); later calls — after ``reset()`` with new initial sets of the same shape — replay it with
        one host call instead of ~6 launches per iteration. Latency lever for small batches (a single trial is
        launch-bound); identical results (same kernels, same order)."""
        if self.persistent:
            return self.run()                                  # already one launch for the whole run
        if self.check_every or self.theta_rest is not None:
            raise ops._lib.B200boError("run_graphed needs check_every=0 and no theta_rest (host logic inside the loop)")
        if getattr(self, "_graph", None) is None:
            X0, y0 = self.X[:, :self.n0].clone(), self.y[:, :self.n0].clone()
            self.run()                                         # eager warm-up on the current stream
            torch.cuda.synchronize()
            self.reset(X0, y0)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._graph_result = self.run()
            self._graph = g
        self._graph.replay()
        return self._graph_result

    @_on_device
    def _begin(self):
        if self.persistent:
            return self._launch_persistent()
        ops.bo_init(self.y, self.n, self.best_f, self.yall, self.picked, self.nit, self.max_iter, self.tol, self.optimal_y)
        self.launches += 1

    @_on_device
    def _advance(self, it: int):
        if self.persistent:
            return                                             # every iteration already runs inside the one launch
        if self.incremental:
            self._step_incremental(it)
        else:
            self.step(it)
        self.iters_done = it + 1

    @_on_device
    def _end(self) -> BOResult:
        if self.persistent:
            return self._finish_persistent()
        if not self.fixed and self.iters_done > 0:
            # the model re-created after the LAST append is fitted and recorded too (Src/BO.py:176-189,252)
            self._timed("fit", self.n0 + self.iters_done, lambda: self._fit(self.iters_done, self.n0 + self.iters_done))
        ops.bo_finalize(self.yall, self.nit, self.n, self.max_iter)
        self.launches += 1
        if self.incremental and self.auto_fallback:
            # a rank-1 append that met a non-positive pivot has no jitter rule (ST_CHOL_FAIL): such trials (none in
            # practice) are re-run through the per-iteration factorisation, as the persistent path does
            self._rerun_full((self.status & ops.ST_CHOL_FAIL) != 0)
        return BOResult(self.yall, self.picked[:, :self.max_iter], self.nit, self.status, self.X, self.y, self.n,
                        self.thetas)

    def _launch_persistent(self):
        """One launch for the whole batch (csrc/bo_persist.cu); asynchronous, no host look at the results."""
        self._timed("bo_fixed_loop", self.n_max, lambda: ops.bo_fixed_loop(
            self.lattice, self.ktabY, self.ktab_id, self.order, self.obj, self.surf_id, self.X, self.y, self.n, self.theta,
            self.best_f, self.yall, self.picked, self.nit, self.status, self.max_iter, self.acq_id, self.acq_param,
            self.var_floor, self.tol, self.optimal_y, ws=self.ws_loop, slot_start=self.slot_start))
        self.launches += 1
        self.iters_done = self.max_iter

    def _finish_persistent(self) -> BOResult:
        """Trials the kernel handed back (ST_FALLBACK: an appended pivot was not positive, i.e. psd_safe_cholesky would
        have added jitter) are re-run through the per-iteration kernels, which implement the jitter rule; that needs one
        host look at the status words (the only synchronisation of the persistent path)."""
        self._rerun_full((self.status & ops.ST_FALLBACK) != 0)
        return BOResult(self.yall, self.picked[:, :self.max_iter], self.nit, self.status, self.X, self.y, self.n, None)

    def _rerun_full(self, fb):
        if bool(fb.any()):
            idx = torch.nonzero(fb).flatten()
            self.fallback_trials += int(idx.numel())
            sub = BatchedBO(self.grid, self.obj, self.surf_id[idx], self.X[idx, :self.n0], self.y[idx, :self.n0],
                            self.max_iter, theta=self.theta[idx], lattice=self.lattice, posterior="full", **self._kwargs)
            r = sub.run()
            for dst, src in ((self.yall, r.yall), (self.picked, sub.picked), (self.nit, r.nit), (self.X, r.X),
                             (self.y, r.y), (self.n, r.n), (self.best_f, sub.best_f), (self.status, r.status)):
                dst[idx] = src
            self.launches += sub.launches

    @_on_device
    def run(self) -> BOResult:
        self._begin()
        for it in range(0 if self.persistent else self.max_iter):
            self._advance(it)
            if self.check_every and (it + 1) % self.check_every == 0 and it + 1 < self.max_iter:
                if not bool((self.n > 0).any()):        # every trial reached tol: stop enqueueing
                    break
        return self._end()


def run_concurrent(bos, streams=None):
    """Advance several independent BatchedBO batches from one host thread, each on its own CUDA stream.

    Trials never exchange data (the reference runs them as separate processes, Scripts/Experiment2.py:83-91), so a batch
    may be cut anywhere. With a fit before every acquisition the per-iteration kernels of ONE batch end with a long
    tail — the fused fit runs until its slowest trial has converged while most CTAs have already left — and on
    separate streams the tail of one group overlaps the bulk of the next. Results are those of the single-batch run,
    trial by trial (tests/test_gpu_bo_loop.py::test_concurrent_groups_equal_one_batch). Returns one BOResult per batch."""
    if not bos:
        return []
    curs = [torch.cuda.current_stream(bo.device) for bo in bos]     # batches may live on different GPUs
    if streams is None:
        streams = [torch.cuda.Stream(bo.device) for bo in bos]
    for s, cur in zip(streams, curs):
        s.wait_stream(cur)                              # buffers were initialised on the caller's stream
    max_iter = max(bo.max_iter for bo in bos)
    check_every = min((bo.check_every for bo in bos if bo.check_every), default=0)
    for bo, s in zip(bos, streams):
        with torch.cuda.stream(s):
            bo._begin()
    for it in range(max_iter):
        for bo, s in zip(bos, streams):
            if it < bo.max_iter:
                with torch.cuda.stream(s):
                    bo._advance(it)
        if check_every and (it + 1) % check_every == 0 and it + 1 < max_iter:
            alive = False
            for bo, s in zip(bos, streams):
                with torch.cuda.stream(s):
                    alive = alive or bool((bo.n > 0).any())
            if not alive:
                break
    out = []
    for bo, s, cur in zip(bos, streams, curs):
        with torch.cuda.stream(s):
            out.append(bo._end())
        cur.wait_stream(s)
    return out


def concat_results(results) -> BOResult:
    """One BOResult from the per-group results of run_concurrent (same max_iter and n_max in every group)."""
    cat = lambda name: torch.cat([getattr(r, name) for r in results], dim=0)
    thetas = None if results[0].thetas is None else cat("thetas")
    return BOResult(cat("yall"), cat("picked"), cat("nit"), cat("status"), cat("X"), cat("y"), cat("n"), thetas)
