"""Randomised differential test of the whole batched BO loop against the oracle: random lattice shapes / offsets /
steps, kernels, acquisitions, hyper-parameters, initial-set sizes and iteration counts (fixed seeds). Picks, curves and
iteration counts must be bit-identical; every posterior mode (full with / without rank-1 append, incremental, and the
persistent one-launch loop) is run on the same case."""
import numpy as np
import pytest
import torch

from oracle import gp as ogp

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _case(seed):
    rng = np.random.default_rng(seed)
    Gx, Gy = int(rng.integers(9, 60)), int(rng.integers(9, 60))
    h = float(rng.choice([1.0, 0.5, 2.0, 0.25]))
    x0, y0 = float(rng.integers(-5, 6)), float(rng.integers(-5, 6))
    ax, ay = x0 + h * np.arange(Gx), y0 + h * np.arange(Gy)
    grid = np.stack(np.meshgrid(ax, ay), axis=-1).reshape(-1, 2)
    G = Gx * Gy
    S = 3
    cx, cy = rng.uniform(ax[0], ax[-1], size=(S, 1)), rng.uniform(ay[0], ay[-1], size=(S, 1))
    w = rng.uniform(3, 12, size=(S, 1)) * h
    obj = (100 * np.exp(-((grid[None, :, 0] - cx) ** 2 + (grid[None, :, 1] - cy) ** 2) / (2 * w ** 2))
           + 8 * np.sin(grid[None, :, 0] / (1.7 * h) + rng.uniform(0, 6, size=(S, 1))) * np.cos(grid[None, :, 1] / (2.3 * h)))
    obj = obj.astype(np.float32)
    T = int(rng.integers(3, 9))
    n0 = int(rng.integers(2, 14))
    iters = int(rng.integers(4, 22))
    kid = int(rng.integers(0, 4))
    acq = int(rng.integers(0, 3))
    theta = np.array([10 ** rng.uniform(-3, 0.5), rng.uniform(-20, 60), 10 ** rng.uniform(0.5, 2.7), h * rng.uniform(1.5, 15)])
    sid = rng.integers(0, S, size=T).astype(np.int32)
    idx0 = np.stack([rng.choice(G, n0, replace=False) for _ in range(T)])
    X0 = grid[idx0]
    y0v = np.stack([obj[sid[t], idx0[t]] for t in range(T)]).astype(np.float64)
    tol = float(rng.choice([-1.0, 0.1, 5.0]))
    return dict(grid=grid, obj=obj, sid=sid, X0=X0, y0=y0v, iters=iters, kid=kid, acq=acq, theta=theta, tol=tol,
                beta=float(rng.uniform(0.5, 3.0)))


@pytest.mark.parametrize("seed", range(16))
def test_random_bo_loops_match_oracle(seed):
    from b200bo import engine
    c = _case(seed)
    t = lambda a: torch.as_tensor(a, device=DEV)
    ref = [ogp.bo_loop(c["grid"], c["obj"][c["sid"][i]], c["X0"][i], c["y0"][i], c["iters"], theta=c["theta"], kernel_id=c["kid"],
                       acq_id=c["acq"], acq_param=c["beta"], tol=c["tol"], optimal_y=float(c["obj"].max()))
           for i in range(len(c["sid"]))]
    for mode in ("full-append", "full-refactor", "incremental", "persistent"):
        bo = engine.BatchedBO(t(c["grid"]), t(c["obj"]), t(c["sid"]), t(c["X0"]), t(c["y0"]), c["iters"], theta=t(c["theta"]),
                              kernel_id=c["kid"], acq_id=c["acq"], acq_param=c["beta"], tol=c["tol"],
                              optimal_y=float(c["obj"].max()), append=(mode != "full-refactor"),
                              posterior=mode if mode in ("incremental", "persistent") else "full")
        res = bo.run()
        torch.cuda.synchronize()
        nit = res.nit.cpu().numpy(); picked = res.picked.cpu().numpy(); yall = res.yall.cpu().numpy()
        ties = 0
        for i, o in enumerate(ref):
            k = min(nit[i], o["nit"])
            same = picked[i, :k] == o["idx"][:k]
            if nit[i] == o["nit"] and same.all():
                assert np.array_equal(yall[i], o["yall"]), (seed, mode, i)
                continue
            # first divergence must be a numerical tie of the acquisition (flat / underflowing EI far from the data:
            # exp and erfc of the two libraries differ in the last bits there) — the documented tie rule then picks
            # different, equally good candidates and the trajectories separate
            j = int(np.argmin(same)) if not same.all() else k
            assert j < k, (seed, mode, i, nit[i], o["nit"])
            n0 = c["X0"].shape[1]
            _, _, _, _, _, a = ogp.grid_step(c["grid"], o["X"][:n0 + j], o["y"][:n0 + j], c["theta"], c["kid"], c["acq"], c["beta"],
                                             return_fields=True)
            gap = abs(a[picked[i, j]] - a[o["idx"][j]])
            assert gap <= 1e-9 * np.nanmax(np.abs(a)) + 1e-290, (seed, mode, i, j, gap, np.nanmax(np.abs(a)))
            ties += 1
        assert ties <= 2, (seed, mode, ties)
        assert int(res.status.max()) & ~1 == 0            # at most the jitter bit
