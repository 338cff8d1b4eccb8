"""GPU parity: K-A (assemble + Cholesky + L^-1 + alpha), K-B (fused grid posterior + acquisition +
arg-max) and K-C (marginal likelihood value and gradient) vs the fp64 oracle."""
import numpy as np
import pytest
import torch

from conftest import AUTHOR_THETA_RAW
from oracle import gp as ogp

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _theta():
    return ogp.raw_to_theta(AUTHOR_THETA_RAW)


def _make_batch(golden, lattice_grid, T, n_list, n_max, seed=0, theta=None):
    rng = np.random.default_rng(seed)
    surf = golden("wildcat.npz")["surf_0"]
    X = np.zeros((T, n_max, 2))
    y = np.zeros((T, n_max))
    n = np.zeros(T, dtype=np.int32)
    th = np.zeros((T, 4))
    for t in range(T):
        nt = n_list[t % len(n_list)]
        idx = rng.choice(10000, nt, replace=False)
        X[t, :nt] = lattice_grid[idx]
        y[t, :nt] = surf[:100, :100].reshape(-1)[idx]
        n[t] = nt
        th[t] = theta if theta is not None else _theta() * rng.uniform(0.7, 1.3, 4)
    return X, y, n, th


def _to(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV) if dtype is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_assemble_chol_alpha_linv(golden, lattice_grid, kid):
    from b200bo import ops
    n_max = 64
    X, y, n, th = _make_batch(golden, lattice_grid, 12, [1, 2, 7, 10, 33, 64], n_max, seed=kid)
    out = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), kid, want=("L", "Linv", "alpha", "logdet", "wfrag"))
    torch.cuda.synchronize()
    for t in range(len(n)):
        nt = n[t]
        L, alpha, Linv, st = ogp.gp_state(X[t, :nt], y[t, :nt], th[t], kid)
        assert out["status"][t].item() == st == 0
        Lg = out["L"][t].cpu().numpy()
        assert np.allclose(Lg[:nt, :nt], L, rtol=1e-9, atol=1e-12 * th[t, 2])
        assert np.all(Lg[nt:] == 0) and np.all(np.triu(Lg, 1) == 0)
        Wg = out["Linv"][t].cpu().numpy()
        scale = np.abs(Linv).max()
        assert np.allclose(Wg[:nt, :nt], Linv, rtol=1e-7, atol=1e-9 * scale)
        ag = out["alpha"][t].cpu().numpy()
        assert np.allclose(ag[:nt], alpha, rtol=1e-7, atol=1e-9 * np.abs(alpha).max())
        assert np.all(ag[nt:] == 0)
        ld = 2 * np.sum(np.log(np.diag(L)))
        assert abs(out["logdet"][t].item() - ld) <= 1e-10 * abs(ld) + 1e-12
        # fragment layout: block (I,J) at I(I+1)+J, lane l -> W[8I + l/4][4J + l%4]
        wf = out["wfrag"][t].cpu().numpy()
        nI = (nt + 7) // 8                      # only the prefix for the trial's own n is written
        for I in range(nI):
            for J in range(2 * I + 2):
                blk = wf[(I * (I + 1) + J) * 32:(I * (I + 1) + J + 1) * 32]
                for lane in (0, 5, 31):
                    r, c = 8 * I + lane // 4, 4 * J + lane % 4
                    assert blk[lane] == (Wg[r, c] if (r < n_max and c < n_max) else 0.0)


def test_chol_jitter_and_failure(lattice_grid):
    """Duplicate observations with ~zero noise make Khat singular: jitter path (status bit 0);
    a negative 'noise' cannot be repaired: failure bit, NaN alpha, batch continues."""
    from b200bo import ops
    n_max = 8
    X = np.zeros((3, n_max, 2)); y = np.zeros((3, n_max)); n = np.array([4, 4, 4], dtype=np.int32)
    pts = np.array([[1.0, 1.0], [1.0, 1.0], [5.0, 9.0], [20.0, 3.0]])
    X[:, :4] = pts
    y[:, :4] = [1.0, 1.0, 2.0, 3.0]
    th = np.array([[1e-3, 0.0, 1.0, 5.0], [0.0, 0.0, 1.0, 5.0], [-10.0, 0.0, 1.0, 5.0]])
    out = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 3, want=("alpha", "L"))
    st = out["status"].cpu().numpy()
    for t in range(3):
        _, alpha, _, st_o = ogp.gp_state(X[t, :4], y[t, :4], th[t], 3)
        assert st[t] == st_o, (t, st[t], st_o)
    assert st[0] == 0 and st[1] == ogp.ST_JITTER and st[2] == (ogp.ST_JITTER | ogp.ST_CHOL_FAIL)
    assert np.all(np.isnan(out["alpha"][2, :4].cpu().numpy()))
    assert np.all(np.isfinite(out["alpha"][1, :4].cpu().numpy()))


@pytest.mark.parametrize("kid,acq", [(3, 0), (2, 0), (1, 0), (0, 0), (3, 1), (3, 2)])
def test_grid_posterior_acq_argmax(golden, lattice_grid, kid, acq):
    from b200bo import ops
    n_max = 40
    n_list = [1, 3, 10, 11, 17, 25, 32, 40]
    T = 8
    theta = _theta()
    X, y, n, th = _make_batch(golden, lattice_grid, T, n_list, n_max, seed=10 + kid, theta=theta)
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), kid)
    best = np.array([y[t, :n[t]].max() for t in range(T)])
    beta = 4.0
    out = ops.grid_posterior_acq_argmax(_to(lattice_grid), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best),
                                        kid, acq_id=acq, acq_param=beta, debug=True)
    torch.cuda.synchronize()
    for t in range(T):
        nt = n[t]
        idx, val, sto, mu, var, a = ogp.grid_step(lattice_grid, X[t, :nt], y[t, :nt], th[t], kid, acq, beta, return_fields=True)
        mu_g = out["mu"][t].cpu().numpy(); var_g = out["var"][t].cpu().numpy(); a_g = out["acq"][t].cpu().numpy()
        # tolerance stated by north_star: 1e-8 relative in fp64 (mean, variance); variance is a
        # difference s - |v|^2, so its error scales with s
        assert np.allclose(mu_g, mu, rtol=1e-8, atol=1e-8 * np.abs(mu).max()), (t, np.abs(mu_g - mu).max())
        assert np.allclose(var_g, var, rtol=1e-8, atol=1e-10 * th[t, 2]), (t, np.abs(var_g - var).max())
        good = np.abs(a_g - a) <= 1e-7 * np.abs(a) + 1e-9 * np.abs(a).max()
        assert good.mean() > 0.999, (t, good.mean())
        # next-query arg-max: bit-exact index
        assert out["idx"][t].item() == idx, (t, out["idx"][t].item(), idx, a[idx], a[out["idx"][t].item()])
        assert abs(out["val"][t].item() - val) <= 1e-7 * abs(val) + 1e-300
        assert out["status"][t].item() == sto


def test_argmax_tie_and_mask_rules(lattice_grid):
    """Far from every observation EI underflows/flattens: ties resolve to the lowest flat index;
    observed candidates never win; extra_mask removes candidates; all masked -> idx -1 + status."""
    from b200bo import ops
    n_max = 8
    X = np.zeros((1, n_max, 2)); y = np.zeros((1, n_max)); n = np.array([2], dtype=np.int32)
    X[0, :2] = [[0.0, 0.0], [1.0, 0.0]]
    y[0, :2] = [5.0, 5.0]
    th = np.array([[1e-2, 0.0, 1.0, 0.05]])        # tiny lengthscale: prior everywhere off the data
    grid = lattice_grid[:64].copy()
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 0)
    best = np.array([5.0])
    out = ops.grid_posterior_acq_argmax(_to(grid), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 0, debug=True)
    a = out["acq"][0].cpu().numpy()
    assert np.all(a[2:] == a[2])                   # exact ties off the data
    assert out["idx"][0].item() == 2               # lowest non-observed index
    mask = np.zeros(64, dtype=np.uint8); mask[2:10] = 1
    out = ops.grid_posterior_acq_argmax(_to(grid), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 0,
                                        extra_mask=_to(mask))
    assert out["idx"][0].item() == 10
    out = ops.grid_posterior_acq_argmax(_to(grid), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 0,
                                        extra_mask=_to(np.ones(64, dtype=np.uint8)))
    assert out["idx"][0].item() == -1 and (out["status"][0].item() & ogp.ST_ALL_MASKED)
    out = ops.grid_posterior_acq_argmax(_to(grid), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 0,
                                        mask_observed=False, debug=True)
    oi, _, _ = ogp.grid_step(grid, X[0, :2], y[0, :2], th[0], 0, mask_observed=False)
    assert out["idx"][0].item() == oi


@pytest.mark.parametrize("n_max", [16, 64, 112, 128])
def test_large_n_and_all_template_sizes(golden, lattice_grid, n_max):
    from b200bo import ops
    theta = _theta()
    X, y, n, th = _make_batch(golden, lattice_grid, 3, [n_max, n_max - 1, max(1, n_max - 9)], n_max, seed=n_max, theta=theta)
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 3)
    best = np.array([y[t, :n[t]].max() for t in range(3)])
    sub = lattice_grid[::3][:2000].copy()          # ragged tail: 2000 is not a multiple of 256
    out = ops.grid_posterior_acq_argmax(_to(sub), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3, debug=True)
    for t in range(3):
        nt = n[t]
        idx, val, sto, mu, var, a = ogp.grid_step(sub, X[t, :nt], y[t, :nt], th[t], 3, return_fields=True)
        assert np.allclose(out["mu"][t].cpu().numpy(), mu, rtol=1e-8, atol=1e-8 * np.abs(mu).max())
        assert np.allclose(out["var"][t].cpu().numpy(), var, rtol=1e-8, atol=1e-9 * th[t, 2])
        assert out["idx"][t].item() == idx


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_mll_value_and_gradient(golden, lattice_grid, kid):
    from b200bo import ops
    n_max = 48
    X, y, n, _ = _make_batch(golden, lattice_grid, 10, [2, 5, 10, 20, 48], n_max, seed=kid + 50)
    rng = np.random.default_rng(kid)
    raw = np.stack([np.array([rng.uniform(1e-3, 3.0), rng.uniform(0, 50), rng.uniform(-2, 200), rng.uniform(-2, 15)]) for _ in range(10)])
    raw[0] = ogp.DEFAULT_RAW
    v, g, st = ops.mll_value_grad(_to(X), _to(y), _to(n), _to(raw), kid)
    v = v.cpu().numpy(); g = g.cpu().numpy()
    for t in range(10):
        vo, go, so = ogp.mll_value_grad(X[t, :n[t]], y[t, :n[t]], raw[t], kid)
        assert st[t].item() == so
        assert abs(v[t] - vo) <= 1e-8 * abs(vo), (t, v[t], vo)
        assert np.allclose(g[t], go, rtol=1e-7, atol=1e-9 * np.abs(go).max()), (t, g[t], go)


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_lattice_variant_is_bit_identical_to_generic(golden, lattice_grid, kid):
    """Table look-ups replace G*n kernel evaluations on a lattice; values must not change at all."""
    from b200bo import ops
    n_max = 72
    T = 6
    rng = np.random.default_rng(kid)
    X, y, n, th = _make_batch(golden, lattice_grid, T, [5, 10, 33, 64, 72, 1], n_max, seed=kid + 7)
    th[3:] = th[0]                                     # trials 0,3,4,5 share a table slot
    uniq, inv = np.unique(th, axis=0, return_inverse=True)
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), kid)
    best = np.array([y[t, :n[t]].max() for t in range(T)])
    grid = _to(lattice_grid)
    lat = ops.Lattice.from_points(grid)
    assert (lat.Gx, lat.Gy, lat.x0, lat.y0, lat.h) == (100, 100, 0.0, 0.0, 1.0)
    assert torch.equal(lat.points(grid.device), grid)
    ktab = ops.kernel_table(_to(uniq), kid, lat)
    a = ops.grid_posterior_acq_argmax(grid, _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), kid, debug=True)
    b = ops.grid_posterior_acq_argmax_lattice(lat, ktab, _to(inv.astype(np.int32)), _to(X), st["wfrag"], st["alpha"], _to(n),
                                              _to(th), _to(best), kid, debug=True)
    for key in ("mu", "var", "acq", "idx", "val", "status"):
        assert torch.equal(a[key], b[key]), key


@pytest.mark.parametrize("shape", [(8, 40), (13, 37), (255, 20), (100, 256), (300, 12), (7, 50)])
def test_lattice_tiles_on_odd_shapes_equal_generic(shape):
    """Byte-coordinate tile (8 <= Gx <= 255, Gy <= 256: packed |dx|,|dy| look-ups, observed-point bit mask, x stepped by 8
    per tile with row wrap) and the general tile (other shapes) against the generic kernel: bit-identical fields, index and
    status, including masked observed candidates, several column chunks per trial and a ragged last tile."""
    from b200bo import ops
    Gx, Gy = shape
    h, x0, y0 = 0.5, -3.0, 2.0
    gx, gy = np.meshgrid(x0 + h * np.arange(Gx), y0 + h * np.arange(Gy))
    grid_h = np.stack([gx, gy], axis=-1).reshape(-1, 2)
    G = Gx * Gy
    rng = np.random.default_rng(Gx * 1000 + Gy)
    n_list = [3, 17, 40]
    T, n_max = len(n_list), 40
    X = np.zeros((T, n_max, 2)); y = np.zeros((T, n_max)); n = np.array(n_list, dtype=np.int32)
    for t in range(T):
        pick = rng.choice(G, n[t], replace=False)
        pick[0] = G - 1                                  # an observed point in the ragged last tile
        X[t, :n[t]] = grid_h[pick]
        y[t, :n[t]] = rng.uniform(0, 100, n[t])
    th = np.tile(np.array([0.05, 40.0, 150.0, 4.0]), (T, 1))
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 3)
    best = np.array([y[t, :n[t]].max() for t in range(T)])
    grid = _to(grid_h)
    lat = ops.Lattice.from_points(grid)
    assert (lat.Gx, lat.Gy) == (Gx, Gy)
    ktab = ops.kernel_table(_to(th[:1]), 3, lat)
    a = ops.grid_posterior_acq_argmax(grid, _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3, debug=True)
    b = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3,
                                              debug=True)
    for key in ("mu", "var", "acq", "idx", "val", "status"):
        assert torch.equal(a[key], b[key]), key
    # no debug fields: the pruned production path picks the same candidate
    c = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3)
    assert torch.equal(c["idx"], a["idx"])
    # an observed point must never be picked
    for t in range(T):
        assert not np.any(np.all(X[t, :n[t]] == grid_h[int(a["idx"][t])], axis=1))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 10, 11, 12, 13, 34, 44, 65, 68, 97, 100, 105, 121, 124])
def test_lattice_tail_rows_off_the_mma_path(lattice_grid, n):
    """Uniform batches with a ragged last row block. 1..2 live rows: the byte-coordinate tile takes them off the DMMA path
    (per-lane DFMA chains), so var may differ from the generic kernel in the last bits (<= 1e-13 s). 3..7 live rows and
    n > 32: alpha rides in the spare row 7 and the mean comes out of the MMAs, so mu differs in the last bits
    (<= 1e-13 max|mu|). Otherwise bit-identical. The pick must be the same unless the two best values tie to 1e-12."""
    from b200bo import ops
    T = 4
    rng = np.random.default_rng(n)
    G = lattice_grid.shape[0]
    X = np.zeros((T, n, 2)); y = np.zeros((T, n))
    for t in range(T):
        X[t] = lattice_grid[rng.choice(G, n, replace=False)]
        y[t] = 40 + 25 * np.sin(X[t, :, 0] / 9.0) * np.cos(X[t, :, 1] / 13.0) + rng.normal(0, 0.5, n)
    nn = np.full(T, n, dtype=np.int32)
    th = np.tile(_theta(), (T, 1))
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(nn), _to(th), 3)
    best = y.max(axis=1)
    grid = _to(lattice_grid)
    lat = ops.Lattice.from_points(grid)
    ktab = ops.kernel_table(_to(th[:1]), 3, lat)
    a = ops.grid_posterior_acq_argmax(grid, _to(X), st["wfrag"], st["alpha"], _to(nn), _to(th), _to(best), 3, debug=True)
    b = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, _to(X), st["wfrag"], st["alpha"], _to(nn), _to(th), _to(best), 3,
                                              debug=True)
    c = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, _to(X), st["wfrag"], st["alpha"], _to(nn), _to(th), _to(best), 3)
    assert torch.equal(a["status"], b["status"])
    s = th[0, 2]
    live = (n - 1) % 8 + 1                               # live rows of the last row block
    mean_on_mma = n > 32 and 3 <= live <= 7              # alpha rides in the spare row 7 (NI >= 5, no tail rows)
    if mean_on_mma:                                      # mean summed in MMA order: last-bit differences only
        assert float((a["mu"] - b["mu"]).abs().max()) <= 1e-13 * float(a["mu"].abs().max())
    else:
        assert torch.equal(a["mu"], b["mu"])
    assert float((a["var"] - b["var"]).abs().max()) <= 1e-13 * s
    if live >= 3:                                        # no tail rows: every row of V stays on the MMA path
        assert torch.equal(a["var"], b["var"])
        if not mean_on_mma:
            assert torch.equal(a["acq"], b["acq"])
    for t in range(T):
        ia, ib, ic = int(a["idx"][t]), int(b["idx"][t]), int(c["idx"][t])
        acq = a["acq"][t]
        assert ib == ic
        if ia != ib:
            assert abs(float(acq[ia] - acq[ib])) <= 1e-12 * abs(float(acq[ia]))
        idx, val, sto, mu, var, aq = ogp.grid_step(lattice_grid, X[t], y[t], th[t], 3, return_fields=True)
        assert np.allclose(b["var"][t].cpu().numpy(), var, rtol=1e-8, atol=1e-10 * s)
        assert np.allclose(a["var"][t].cpu().numpy(), var, rtol=1e-8, atol=1e-10 * s)
        assert np.allclose(a["mu"][t].cpu().numpy(), mu, rtol=1e-8, atol=1e-8 * np.abs(mu).max())
        # the arg-max of the production tail-row / mean-row variants is the ORACLE's arg-max; a different index is
        # only acceptable on a rounding-level tie of the oracle's own acquisition values
        assert ib == idx or abs(aq[ib] - aq[idx]) <= 1e-12 * abs(aq[idx]), (t, ib, idx)
        assert ia == idx or abs(aq[ia] - aq[idx]) <= 1e-12 * abs(aq[idx]), (t, ia, idx)


def test_lattice_half_step_grid_and_off_lattice_flag(golden):
    """constructX(budget=40000): 200x200 nodes, step 0.5 (BASELINE config 4 candidates)."""
    from b200bo import ops
    ax = np.arange(0, 100, 0.5)
    grid = np.stack(np.meshgrid(ax, ax), axis=-1).reshape(-1, 2)
    lat = ops.Lattice.from_points(grid)
    assert (lat.Gx, lat.Gy, lat.h) == (200, 200, 0.5)
    rng = np.random.default_rng(3)
    n_max = 24
    X = np.zeros((2, n_max, 2)); y = np.zeros((2, n_max)); n = np.array([20, 24], dtype=np.int32)
    for t in range(2):
        X[t, :n[t]] = grid[rng.choice(40000, n[t], replace=False)]
        y[t, :n[t]] = rng.uniform(0, 100, n[t])
    th = np.tile(_theta(), (2, 1))
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 3)
    best = np.array([y[t, :n[t]].max() for t in range(2)])
    ktab = ops.kernel_table(_to(th[:1]), 3, lat)
    b = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3, debug=True)
    for t in range(2):
        idx, val, sto, mu, var, a = ogp.grid_step(grid, X[t, :n[t]], y[t, :n[t]], th[t], 3, return_fields=True)
        assert np.allclose(b["mu"][t].cpu().numpy(), mu, rtol=1e-8, atol=1e-8 * np.abs(mu).max())
        assert np.allclose(b["var"][t].cpu().numpy(), var, rtol=1e-8, atol=1e-10 * th[t, 2])
        assert b["idx"][t].item() == idx and b["status"][t].item() == 0
    X[1, 3] += 0.25                                     # not a node any more
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 3)
    b = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3)
    assert b["status"][0].item() == 0 and (b["status"][1].item() & ops.ST_OFF_LATTICE)


@pytest.mark.parametrize("kid", [0, 3])
def test_rank1_append_matches_refactorisation(golden, lattice_grid, kid):
    """K-A append variant: growing L^-1 / z / alpha one observation at a time equals a fresh factorisation."""
    from b200bo import ops
    n_max, T, n0 = 40, 6, 7
    X, y, n_full, th = _make_batch(golden, lattice_grid, T, [n_max], n_max, seed=kid + 21, theta=_theta())
    Xd, yd, thd = _to(X), _to(y), _to(th)
    n = torch.full((T,), n0, dtype=torch.int32, device=DEV)
    wf = torch.zeros((T, ops.wfrag_doubles(n_max)), dtype=torch.float64, device=DEV)
    al = torch.zeros((T, n_max), dtype=torch.float64, device=DEV)
    zv = torch.zeros((T, n_max), dtype=torch.float64, device=DEV)
    st = torch.zeros(T, dtype=torch.int32, device=DEV)
    redo = torch.zeros(T, dtype=torch.int32, device=DEV)
    ops.gp_assemble_chol_into(Xd, yd, n, thd, kid, al, wf, st)
    ops.gp_zvec(yd, n, thd, wf, zv)
    for m in range(n0 + 1, n_max + 1):
        n.fill_(m)
        ops.gp_chol_append(Xd, yd, n, thd, kid, wf, zv, al, redo, st)
        assert int(redo.max()) == 0 and int(st.max()) == 0
        if m in (n0 + 1, 16, 17, 24, 33, n_max):
            ref = ops.gp_assemble_chol(Xd, yd, n, thd, kid, want=("alpha", "wfrag", "Linv"))
            nb = ((m + 7) // 8) * (((m + 7) // 8) + 1) * 32
            a, b = wf[:, :nb].cpu().numpy(), ref["wfrag"][:, :nb].cpu().numpy()
            assert np.allclose(a, b, rtol=1e-9, atol=1e-11 * np.abs(b).max()), (m, np.abs(a - b).max())
            assert np.all(a[b == 0] == 0)
            aa, ab = al[:, :m].cpu().numpy(), ref["alpha"][:, :m].cpu().numpy()
            assert np.allclose(aa, ab, rtol=1e-8, atol=1e-10 * np.abs(ab).max()), (m, np.abs(aa - ab).max())
            for t in range(T):
                z_o = ref["Linv"][t, :m, :m].cpu().numpy() @ (y[t, :m] - th[t, 1])
                assert np.allclose(zv[t, :m].cpu().numpy(), z_o, rtol=1e-8, atol=1e-10 * np.abs(z_o).max())


def test_rank1_append_failure_is_handed_to_full_kernel():
    """Appending a duplicate of an observation with zero noise makes the new pivot non-positive: n_redo flags the
    trial and the full kernel repairs it with psd_safe_cholesky's jitter (status bit 0)."""
    from b200bo import ops
    n_max = 8
    X = np.zeros((2, n_max, 2)); y = np.zeros((2, n_max))
    X[:, :4] = [[1.0, 1.0], [5.0, 9.0], [20.0, 3.0], [1.0, 1.0]]
    y[:, :4] = [1.0, 2.0, 3.0, 1.0]
    th = np.array([[0.0, 0.0, 1.0, 5.0], [1e-2, 0.0, 1.0, 5.0]])
    Xd, yd, thd = _to(X), _to(y), _to(th)
    n = torch.full((2,), 3, dtype=torch.int32, device=DEV)
    wf = torch.zeros((2, ops.wfrag_doubles(n_max)), dtype=torch.float64, device=DEV)
    al = torch.zeros((2, n_max), dtype=torch.float64, device=DEV); zv = torch.zeros_like(al)
    st = torch.zeros(2, dtype=torch.int32, device=DEV); redo = torch.zeros(2, dtype=torch.int32, device=DEV)
    ops.gp_assemble_chol_into(Xd, yd, n, thd, 3, al, wf, st)
    ops.gp_zvec(yd, n, thd, wf, zv)
    n.fill_(4)
    ops.gp_chol_append(Xd, yd, n, thd, 3, wf, zv, al, redo, st)
    assert redo.tolist() == [4, 0] and st.tolist() == [0, 0]
    ops.gp_assemble_chol_into(Xd, yd, redo, thd, 3, al, wf, st)
    ops.gp_zvec(yd, redo, thd, wf, zv)
    assert st.tolist() == [ogp.ST_JITTER, 0]
    for t in range(2):
        _, alpha, _, so = ogp.gp_state(X[t, :4], y[t, :4], th[t], 3)
        assert np.allclose(al[t, :4].cpu().numpy(), alpha, rtol=1e-6, atol=1e-8 * np.abs(alpha).max())


def test_edge_cases_empty_batch_inactive_trials_and_nonfinite(golden, lattice_grid):
    from b200bo import ops, _lib
    # empty batch: every entry is a no-op
    e = lambda *shape, dt=torch.float64: torch.empty(shape, dtype=dt, device=DEV)
    out = ops.gp_assemble_chol(e(0, 8, 2), e(0, 8), e(0, dt=torch.int32), e(0, 4), 3)
    assert out["alpha"].shape == (0, 8)
    r = ops.grid_posterior_acq_argmax(_to(lattice_grid[:64].copy()), e(0, 8, 2), e(0, ops.wfrag_doubles(8)), e(0, 8),
                                      e(0, dt=torch.int32), e(0, 4), e(0), 3)
    assert r["idx"].numel() == 0
    # inactive trial (n = 0) between two live ones: skipped, idx -1, no status, neighbours untouched
    n_max = 16
    X, y, n, th = _make_batch(golden, lattice_grid, 3, [12], n_max, seed=4, theta=_theta())
    n[1] = 0
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 3)
    best = np.array([y[t, :12].max() for t in range(3)])
    out = ops.grid_posterior_acq_argmax(_to(lattice_grid), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3)
    assert out["idx"][1].item() == -1 and out["status"][1].item() == 0
    for t in (0, 2):
        idx, _, _ = ogp.grid_step(lattice_grid, X[t, :12], y[t, :12], th[t], 3)
        assert out["idx"][t].item() == idx
    # a NaN observation poisons alpha -> every acquisition value is NaN: never wins, status says why
    y2 = y.copy(); y2[0, 3] = np.nan
    st = ops.gp_assemble_chol(_to(X), _to(y2), _to(n), _to(th), 3)
    out = ops.grid_posterior_acq_argmax(_to(lattice_grid), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3)
    assert out["idx"][0].item() == -1
    assert out["status"][0].item() & ogp.ST_NONFINITE and out["status"][0].item() & ogp.ST_ALL_MASKED
    assert out["idx"][2].item() >= 0 and out["status"][2].item() == 0
    # argument errors are reported, nothing is launched
    with pytest.raises(_lib.B200boError):
        ops.gp_assemble_chol(e(1, 200, 2), e(1, 200), torch.ones(1, dtype=torch.int32, device=DEV), e(1, 4), 3)   # n_max > 128
    with pytest.raises(_lib.B200boError):
        ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 9)                                                    # bad kernel id
    with pytest.raises(_lib.B200boError):
        ops.grid_posterior_acq_argmax(_to(lattice_grid), _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3,
                                      acq_id=ops.UCB, acq_param=-1.0)
    assert "UCB" in _lib.load().b200bo_last_error_string().decode()


def test_ragged_n_inside_one_launch(golden, lattice_grid):
    """Trials with very different n in one launch (lattice kernel: template sized by the largest, ragged body for the rest)."""
    from b200bo import ops
    n_max = 104
    X, y, n, th = _make_batch(golden, lattice_grid, 6, [104, 9, 57, 1, 33, 96], n_max, seed=31, theta=_theta())
    st = ops.gp_assemble_chol(_to(X), _to(y), _to(n), _to(th), 3)
    best = np.array([y[t, :n[t]].max() for t in range(6)])
    grid = _to(lattice_grid)
    lat = ops.Lattice.from_points(grid)
    ktab = ops.kernel_table(_to(th[:1]), 3, lat)
    out = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3, debug=True)
    for t in range(6):
        idx, val, sto, mu, var, a = ogp.grid_step(lattice_grid, X[t, :n[t]], y[t, :n[t]], th[t], 3, return_fields=True)
        assert np.allclose(out["mu"][t].cpu().numpy(), mu, rtol=1e-8, atol=1e-8 * np.abs(mu).max())
        assert np.allclose(out["var"][t].cpu().numpy(), var, rtol=1e-8, atol=1e-10 * th[t, 2])
        assert out["idx"][t].item() == idx
    # a too-small n_hint must not corrupt anything: flagged, idx -1
    out = ops.grid_posterior_acq_argmax_lattice(lat, ktab, None, _to(X), st["wfrag"], st["alpha"], _to(n), _to(th), _to(best), 3, n_hint=40)
    assert out["idx"][0].item() == -1 and out["idx"][1].item() >= 0 and out["idx"][4].item() >= 0
