"""Drop-in for the reference's ``Src/BO.py``: BO loops on an exact GP with analytic EI.

Signatures are the reference's (BO.py:19,33,72,123,284,380). Underneath, the botorch/gpytorch model,
``fit_gpytorch_model`` and ``optimize_acqf`` are replaced by the B200 engine (b200bo.engine.BatchedBO):
the acquisition is evaluated densely on the candidate lattice and arg-maxed (the north-star redesign,
SURVEY.md fact 5), hyper-parameters are fitted by a batched L-BFGS on device. A single-trial call is a
batch of one; ``BOloop_batched`` is the native entry for many trials.

Deliberate differences from the reference as written, all documented in DESIGN.md:
* fp64 arithmetic (the reference trains in fp32);
* ``BOloop_fixparam`` keeps the given hyper-parameters for *every* iteration (the reference's intent; as
  coded it falls back to gpytorch defaults after the first model re-creation, BO.py:424-429);
  ``reference_quirk=True`` reproduces the as-coded behaviour;
* candidates that coincide with an observation are masked instead of being jittered (BO.py:90-92);
* numerical failures are per-trial status bits, not exceptions with roll-back (BO.py:187-239);
* ``minstate=True`` (the objective is passed negated, Optimizers.py:48-52): the reference seeds ``yopt`` / ``yall[0]``
  and the first EI's ``best_f`` with ``train_obj.min()`` of the NEGATED values, i.e. the worst initial point
  (BO.py:27-28,141,146); here the incumbent is the best initial point from the start (max of the negated values), as
  it is from the second iteration on in the reference too (``train_obj_ei.max()``, BO.py:168). ``generate_initial_data``
  itself still returns the reference's value. The Wildcat Wells experiments all maximise.
"""
import random

import numpy as np
import torch

import results
from b200bo import engine, fit, ops

PARAM_NAMES = ("likelihood.noise_covar.raw_noise", "mean_module.raw_constant", "covar_module.raw_outputscale",
               "covar_module.base_kernel.raw_lengthscale")
_ALIASES = {"mean_module.constant": "mean_module.raw_constant"}      # older gpytorch naming in the authors' pickles

DEFAULT_GRID_BUDGET = 10000


class _Param:
    """Minimal stand-in for a torch Parameter: ``.item()`` and in-place ``copy_`` (BO.py:394-395, results.py:238-242)."""

    def __init__(self, value):
        self.value = float(value)

    def item(self):
        return self.value

    def copy_(self, other):
        self.value = float(other.item() if hasattr(other, "item") else other)
        return self

    def __float__(self):
        return self.value


class GPModel:
    """What ``initialize_model`` returns in place of a botorch SingleTaskGP: training data, kernel family and
    the four raw parameters under gpytorch's names (defaults: noise 2.0, constant 0, raw scales 0)."""

    def __init__(self, train_x, train_obj, nu=1.5, covar="matern"):
        self.train_x = torch.as_tensor(train_x, dtype=torch.float64).reshape(-1, 2).cpu()
        self.train_obj = torch.as_tensor(train_obj, dtype=torch.float64).reshape(-1).cpu()
        self.kernel_id = ops.kernel_id_from(nu, covar)
        self._params = {n: _Param(v) for n, v in zip(PARAM_NAMES, fit.DEFAULT_RAW)}

    def named_parameters(self):
        return list(self._params.items())

    def raw(self):
        return np.array([self._params[n].value for n in PARAM_NAMES], dtype=np.float64)

    def set_raw(self, raw):
        for n, v in zip(PARAM_NAMES, raw):
            self._params[n].value = float(v)

    def theta(self):
        r = self.raw()
        sp = lambda v: v if v > 20.0 else float(np.log1p(np.exp(v)))
        return np.array([r[0], r[1], sp(r[2]), sp(r[3])])


class MarginalLogLikelihood:
    """Stand-in for gpytorch's ExactMarginalLogLikelihood: carries the model for ``fit_gpytorch_model``."""

    def __init__(self, model):
        self.model = model


def generate_initial_data(X, Y, minstate, n=10):
    """Random permutation of the initial points (BO.py:19-31; uses Python's global ``random`` like the reference)."""
    if len(X) < n:
        n = len(X)
    picks = np.array(random.sample(range(len(X)), n))
    train_x = torch.from_numpy(np.asarray(X)[picks]).type(torch.FloatTensor)
    train_obj = torch.from_numpy(np.asarray(Y)[picks])
    best = train_obj.min() if minstate is True else train_obj.max()
    return train_x, train_obj.reshape(-1, 1), best


def initialize_model(train_x, train_obj, nu=1.5, covar="matern"):
    """(mll, model) with default hyper-parameters (BO.py:33-68)."""
    model = GPModel(train_x, train_obj, nu=nu, covar=covar)
    return MarginalLogLikelihood(model), model


def fit_gpytorch_model(mll, device="cuda:0"):
    """Maximise the marginal likelihood of ``mll.model`` from its current raw parameters (BO.py:144,189)."""
    m = mll.model
    dev = torch.device(device)
    n = m.train_x.shape[0]
    with torch.cuda.device(dev):
        X = m.train_x.to(dev).reshape(1, n, 2).contiguous()
        y = m.train_obj.to(dev).reshape(1, n).contiguous()
        nn = torch.tensor([n], dtype=torch.int32, device=dev)
        if n >= fit.LARGE_N:      # Experiment3-2-1.py:63-72 fits 1 010 ... 1 209 points: blocked K-C (csrc/gp_fit_large.cu)
            r = fit.fit_large(X, y, nn, m.kernel_id, raw0=tuple(m.raw()))
            m.set_raw(r["raw"][0].cpu().numpy())
            return mll
        f = fit.LBFGSFitter(1, dev, raw0=tuple(m.raw()))
        theta = torch.empty((1, 4), dtype=torch.float64, device=dev)
        f.fit_into(X, y, nn, m.kernel_id, theta)
        m.set_raw(f.raw[0].cpu().numpy())
    return mll


class ExpectedImprovement:
    """Analytic EI bound to a model and incumbent (BO.py:70,158)."""
    acq_id = ops.EI

    def __init__(self, model, best_f, maximize=True):
        self.model = model
        self.best_f = float(best_f)


def _lattice_for(bound, budget=DEFAULT_GRID_BUDGET):
    """Candidate lattice = data_gen.constructX(budget) over the box the reference hands to optimize_acqf,
    [lo, hi-1] per dimension (BO.py:78): 100x100 integer nodes for budget 10 000, 200x200 step 0.5 for 40 000."""
    from b200bo import host
    axes, steps = host.axis_grid([tuple(b) for b in bound], budget, endpoint=False, dtype=np.float64)
    return ops.Lattice(len(axes[0]), len(axes[1]), float(axes[0][0]), float(axes[1][0]), float(steps[0]))


def _objective_table(obj_func, lat, dev):
    if hasattr(obj_func, "grid_table"):
        return obj_func.grid_table(lat).reshape(1, -1).contiguous()
    pts = lat.points(dev)
    vals = np.asarray(obj_func(pts.cpu().numpy()), dtype=np.float32).reshape(1, -1)
    return torch.as_tensor(vals, device=dev).contiguous()


def optimize_acqf_and_get_observation(acq_func, train_x, obj_func, raw, bound, restarts=1, budget=DEFAULT_GRID_BUDGET,
                                      device="cuda:0"):
    """Arg-max of the acquisition over the candidate lattice and the objective there (BO.py:72-105).
    ``raw`` / ``restarts`` (Sobol samples / L-BFGS restarts of optimize_acqf) have no meaning on a grid."""
    with torch.cuda.device(torch.device(device)):
        return _optimize_acqf_on_lattice(acq_func, obj_func, bound, budget, torch.device(device))


def _optimize_acqf_on_lattice(acq_func, obj_func, bound, budget, dev):
    m = acq_func.model
    lat = _lattice_for(bound, budget)
    n = m.train_x.shape[0]
    X = m.train_x.to(dev).reshape(1, n, 2).contiguous()
    y = m.train_obj.to(dev).reshape(1, n).contiguous()
    nn = torch.tensor([n], dtype=torch.int32, device=dev)
    th = torch.as_tensor(m.theta(), device=dev).reshape(1, 4).contiguous()
    st = ops.gp_assemble_chol(X, y, nn, th, m.kernel_id)
    best = torch.tensor([acq_func.best_f], dtype=torch.float64, device=dev)
    ktab = ops.kernel_table(th, m.kernel_id, lat)
    out = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, X, st["wfrag"], st["alpha"], nn, th, best, m.kernel_id,
                                                acq_id=acq_func.acq_id)
    if int(out["status"][0]) & ops.ST_OFF_LATTICE:
        out = ops.grid_posterior_acq_argmax(lat.points(dev), X, st["wfrag"], st["alpha"], nn, th, best, m.kernel_id,
                                            acq_id=acq_func.acq_id)
    idx = int(out["idx"][0])
    if idx < 0:
        raise RuntimeError("no selectable candidate (status %d)" % int(out["status"][0]))
    new_x = lat.points(dev)[idx].to(torch.float32).reshape(1, 2).cpu()
    new_obj = torch.tensor(np.asarray(obj_func(new_x.numpy().astype(np.float64))), dtype=torch.float32).reshape(1, 1)
    return new_x, new_obj


def _pack_result(res, i, nt, max_iter, minstate, optimal_y, X_train, hp_names=None):
    r = results.result_opt()
    nit = int(res.nit[i])
    yall = res.yall[i].cpu().numpy().copy()
    r.yall = -yall if minstate else yall                     # BO.py:258-263
    Xall = res.X[i, :nt + nit].cpu().numpy()
    yobs = res.y[i, :nt + nit].cpu().numpy()
    r.xall = Xall[nt:].astype(np.float32)
    r.yopt = torch.tensor(float(yobs.max()), dtype=torch.float32)
    r.xopt = torch.from_numpy(Xall[yobs == yobs.max()].astype(np.float32))
    r.nit = nit
    r.success = not bool(int(res.status[i]) & (ops.ST_CHOL_FAIL | ops.ST_ALL_MASKED | ops.ST_OFF_LATTICE))
    r.status = int(res.status[i])
    r._opt = optimal_y
    r.minstate = minstate
    r.traindata = X_train
    if res.thetas is not None:
        series = results.hp_series()
        # the model fitted after append k is what the reference records at iteration k (BO.py:176-189,252):
        # thetas[:, k] is the fit on n0+k observations, so iterations 1..nit -> rows 1..nit
        th = res.thetas[i, 1:nit + 1].cpu().numpy()
        for row in th:
            h = results.hp_result()
            inv_sp = lambda v: v if v > 20.0 else float(np.log(np.expm1(v)))
            h.hyperparameters = {PARAM_NAMES[0]: [row[0]], PARAM_NAMES[1]: [row[1]], PARAM_NAMES[2]: [inv_sp(row[2])],
                                 PARAM_NAMES[3]: [inv_sp(row[3])]}
            series.data.append(h)
        series.updateres()
        r.hpdata = series
    return r


def _fixparam_to_theta(fixparam):
    fp = {_ALIASES.get(k, k): v for k, v in fixparam.items()}
    raw = [float(np.atleast_1d(fp[n])[0]) for n in PARAM_NAMES]      # fixparam[name][0], BO.py:395
    sp = lambda v: v if v > 20.0 else float(np.log1p(np.exp(v)))
    return np.array([raw[0], raw[1], sp(raw[2]), sp(raw[3])])


def BOloop_batched(obj_tables, surf_id, init_X, init_y, max_iter, theta=None, grid=None, acq="EI", acq_param=0.0,
                   tol=0.1, optimal_y=100.0, nu=2.5, covar="matern", device="cuda:0", lattice="auto", theta_rest=None,
                   posterior="auto", groups="auto", devices=None):
    """Native batched entry: T independent trials. obj_tables (S,G) f32 objective at every candidate, surf_id (T,),
    init_X (T,n0,2), init_y (T,n0); theta (4,) / (T,4) ACTUAL hyper-parameters or None (= fit before every
    acquisition). grid: (G,2) candidates or an ops.Lattice (default: the 100x100 lattice of data_gen.constructX).
    posterior: "auto" = the one-launch persistent loop (csrc/bo_persist.cu) when theta is fixed and the lattice's kernel
    table fits shared memory, else "full" (O(n^2) fused recomputation per iteration); "incremental": O(n) update per
    candidate with the state in HBM. groups > 1 cuts the batch into that many contiguous groups advanced concurrently on
    separate CUDA streams (pays when theta is None and the per-iteration fit has a long tail:
    b200bo.engine.run_concurrent); "auto" = 3 groups for fitted batches of at least 384 trials, else 1.
    devices — where the trials run (they are independent; the reference fans them out over ipyparallel engines,
    Scripts/Experiment2.py:75-92,110-141):
      None             all T trials on ``device``;
      ["cuda:0", ...]  contiguous trial ranges on several GPUs of this process, driven from this host thread (one stream
                       per GPU); the result is concatenated on the first device;
      "ranks"          one process per GPU under torchrun: this rank runs its range (b200bo.dist.shard_range) on
                       ``device`` and every rank returns the all-gathered result of the whole batch.
    Whatever the split, the result equals the single-device result trial for trial.
    Returns b200bo.engine.BOResult (device tensors)."""
    common = dict(max_iter=max_iter, grid=grid, acq=acq, acq_param=acq_param, tol=tol, optimal_y=optimal_y, nu=nu,
                  covar=covar, lattice=lattice, posterior=posterior, groups=groups)
    T = int(np.shape(init_y)[0])
    cut = lambda a, lo, hi: a if a is None or (np.ndim(a) == 1 and np.shape(a)[0] == 4) else a[lo:hi]

    def shard(lo, hi, dev, **kw):
        with torch.cuda.device(dev):
            return _boloop_batched(obj_tables, surf_id[lo:hi], init_X[lo:hi], init_y[lo:hi], theta=cut(theta, lo, hi),
                                   theta_rest=cut(theta_rest, lo, hi), dev=dev, **common, **kw)

    fields = ("yall", "picked", "nit", "status", "X", "y", "n", "thetas")
    if isinstance(devices, str) and devices == "ranks":
        from b200bo import dist as bdist
        dev = torch.device(device)

        def run_local(lo, hi):
            r = shard(lo, hi, dev)[0]
            return {f: getattr(r, f) for f in fields}
        return engine.BOResult(**bdist.run_sharded(T, run_local))
    if devices is not None:
        from b200bo import dist as bdist
        devs = [torch.device(d) for d in devices]
        bos = []
        for k, dev in enumerate(devs):
            lo, hi = bdist.shard_range(T, k, len(devs))
            if hi > lo:
                bos += shard(lo, hi, dev, build_only=True)
        res = engine.run_concurrent(bos)
        first = devs[0]
        out = engine.concat_results([engine.BOResult(**{f: (None if getattr(r, f) is None else getattr(r, f).to(first))
                                                        for f in fields}) for r in res])
        for dev in devs:
            torch.cuda.synchronize(dev)
        return out
    return shard(0, T, torch.device(device))[0]


def _boloop_batched(obj_tables, surf_id, init_X, init_y, max_iter, theta, grid, acq, acq_param, tol, optimal_y, nu, covar,
                    lattice, theta_rest, posterior, groups, dev, build_only=False):
    """One device. Returns [BOResult], or with build_only the list of BatchedBO groups (not yet run)."""
    t = lambda a, dt: torch.as_tensor(a, dtype=dt).to(dev, non_blocking=True).contiguous()
    if grid is None:
        grid = ops.Lattice(100, 100)
    lat = grid if isinstance(grid, ops.Lattice) else lattice
    pts = grid.points(dev) if isinstance(grid, ops.Lattice) else t(grid, torch.float64)
    T = np.shape(init_y)[0]
    acq_id = {"EI": ops.EI, "UCB": ops.UCB, "PI": ops.PI}[acq]
    obj_d, sid_d = t(obj_tables, torch.float32), t(surf_id, torch.int32)
    X_d, y_d = t(init_X, torch.float64), t(init_y, torch.float64)
    th_d = None if theta is None else t(theta, torch.float64)
    rest_d = None if theta_rest is None else t(theta_rest, torch.float64)

    def make(lo, hi):
        sub = lambda a: a if a is None or a.dim() == 1 and a.shape[0] == 4 else a[lo:hi]
        fitter = None if theta is not None else fit.LBFGSFitter(hi - lo, dev)
        return engine.BatchedBO(pts, obj_d, sid_d[lo:hi], X_d[lo:hi], y_d[lo:hi], max_iter, theta=sub(th_d),
                                kernel_id=ops.kernel_id_from(nu, covar), acq_id=acq_id, acq_param=acq_param, tol=tol,
                                optimal_y=optimal_y, fitter=fitter, lattice=lat, theta_rest=sub(rest_d),
                                posterior=posterior)

    if groups == "auto":                                   # measured on 1 000 / 4 000 trials: +8..22 % with a fit per iteration
        groups = 3 if (theta is None and T >= 384) else 1
    groups = max(1, min(int(groups), T))
    cuts = [T * g // groups for g in range(groups + 1)]
    bos = [make(cuts[g], cuts[g + 1]) for g in range(groups)]
    if build_only:
        return bos
    if groups == 1:
        return [bos[0].run()]
    return [engine.concat_results(engine.run_concurrent(bos))]


def _single(obj_func, max_iter, X, Y, theta, bounds, tol, nu, covar, minstate, optimal_y, budget, device,
            theta_rest=None):
    if nu > 2.5:
        covar = "rbf"                                        # BO.py:133-134
    nt = 10 if len(X) > 10 else len(X)
    train_x, train_obj, _ = generate_initial_data(X, Y, minstate, n=nt)
    dev = torch.device(device)
    lat = _lattice_for(bounds, budget)
    with torch.cuda.device(dev):
        table = _objective_table(obj_func, lat, dev)
    # minimisation = maximise -f (Optimizers.py:48-52): obj_func already is f_c there, Y = f_c(X)
    X0 = train_x.double().reshape(1, nt, 2)
    y0 = train_obj.double().reshape(1, nt)
    # error = optimal_y - yopt on the (possibly negated) objective, exactly as BO.py:241 computes it
    res = BOloop_batched(table, np.zeros(1, dtype=np.int32), X0, y0, max_iter, theta=theta, grid=lat, tol=tol,
                         optimal_y=optimal_y, nu=nu, covar=covar, device=device, theta_rest=theta_rest)
    torch.cuda.synchronize(dev)
    r = _pack_result(res, 0, nt, max_iter, minstate, optimal_y, X)
    return r


def BOloop(obj_func, max_iter, X, Y, bounds=[(0, 100), (0, 100)], verbose=False, tol=1, nu=2.5, covar="matern",
           minstate=False, optimal_y=100, budget=DEFAULT_GRID_BUDGET, device="cuda:0"):
    """BO with a hyper-parameter fit before every acquisition (BO.py:123-278)."""
    return _single(obj_func, max_iter, X, Y, None, bounds, tol, nu, covar, minstate, optimal_y, budget, device)


def BOloop_fixparam(obj_func, max_iter, X, Y, fixparam, bounds=[(0, 100), (0, 100)], verbose=False, tol=1, nu=2.5,
                    covar="matern", minstate=False, optimal_y=100, budget=DEFAULT_GRID_BUDGET, device="cuda:0",
                    reference_quirk=False):
    """BO with fixed hyper-parameters ``fixparam[name][0]`` (BO.py:380-461)."""
    theta = _fixparam_to_theta(fixparam)
    # as coded, only the first model carries fixparam; every re-created model keeps gpytorch's defaults
    # (noise 2, constant 0, outputscale = lengthscale = softplus(0)), BO.py:392-395 vs :424-429
    rest = np.array([2.0, 0.0, np.log(2.0), np.log(2.0)]) if reference_quirk else None
    return _single(obj_func, max_iter, X, Y, theta, bounds, tol, nu, covar, minstate, optimal_y, budget, device,
                   theta_rest=rest)


def BOloop_singleopt(obj_func, max_iter, X, Y, bounds=[(0, 100), (0, 100)], verbose=False, tol=1, nu=2.5, covar="matern",
                     minstate=False, optimal_y=100, budget=DEFAULT_GRID_BUDGET, device="cuda:0"):
    """Fit once on the initial data, then loop with those hyper-parameters (the intent of BO.py:284-374)."""
    if nu > 2.5:
        covar = "rbf"
    nt = 10 if len(X) > 10 else len(X)
    state = random.getstate()
    train_x, train_obj, _ = generate_initial_data(X, Y, minstate, n=nt)
    random.setstate(state)                                   # _single draws the same permutation again
    mll, model = initialize_model(train_x, train_obj, nu=nu, covar=covar)
    fit_gpytorch_model(mll, device=device)
    return _single(obj_func, max_iter, X, Y, model.theta(), bounds, tol, nu, covar, minstate, optimal_y, budget, device)
