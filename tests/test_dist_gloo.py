"""CPU, world_size 2 over gloo: the multi-GPU path's host logic — contiguous trial shards and the single
final gather of gap curves — must reproduce the single-process result bit for bit."""
import os
import subprocess
import sys

import numpy as np

from conftest import PKG, ROOT

WORKER = r"""
import os, sys
import numpy as np, torch
sys.path.insert(0, {pkg!r})
from b200bo import dist
rank, world, _ = dist.init(backend="gloo")
total, width = {total}, 12
lo, hi = dist.shard_range(total, rank, world)
full = torch.arange(total * width, dtype=torch.float64).reshape(total, width) * 0.5     # stands for the gap curves
local = full[lo:hi].clone()
dist.barrier()
out = dist.gather_curves(local)
assert out.shape == full.shape and torch.equal(out, full), (rank, out.shape)
t = dist.max_over_ranks(1.0 + rank, "cpu")
assert t == float(world)
if rank == 0:
    np.save({out!r}, out.numpy())
torch.distributed.destroy_process_group()
"""


def _run(total, tmp_path, port):
    out = str(tmp_path / f"gathered_{total}.npy")
    script = tmp_path / f"worker_{total}.py"
    script.write_text(WORKER.format(pkg=PKG, total=total, out=out))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.load(out)


def test_even_shards(tmp_path):
    got = _run(10, tmp_path, 29611)
    assert np.array_equal(got, np.arange(120, dtype=np.float64).reshape(10, 12) * 0.5)


def test_ragged_shards(tmp_path):
    got = _run(7, tmp_path, 29612)
    assert np.array_equal(got, np.arange(84, dtype=np.float64).reshape(7, 12) * 0.5)


BO_WORKER = r"""
import os, sys
import numpy as np, torch
sys.path.insert(0, {pkg!r}); sys.path.insert(0, {root!r})
from b200bo import dist
from oracle import gp as ogp, dpp as odpp
rank, world, _ = dist.init(backend="gloo")
d = np.load({inp!r})
grid = odpp.construct_pool()
theta = d["theta"]
T, iters = d["X0"].shape[0], int(d["iters"])

def run_local(lo, hi):
    # the rank's trials through the CPU oracle (the product runs this range on the rank's GPU: BO.BOloop_batched(devices="ranks"))
    rows = [ogp.bo_loop(grid, d["obj"][d["sid"][i]], d["X0"][i], d["y0"][i], iters, theta=theta, kernel_id=3, tol=0.1) for i in range(lo, hi)]
    picked = np.full((hi - lo, iters), -1, dtype=np.int32)
    for k, r in enumerate(rows):
        picked[k, :r["nit"]] = r["idx"]
    return dict(yall=torch.from_numpy(np.stack([r["yall"] for r in rows]).reshape(hi - lo, iters + 2)),
                nit=torch.tensor([r["nit"] for r in rows], dtype=torch.int32), picked=torch.from_numpy(picked), thetas=None)

out = dist.run_sharded(T, run_local)
assert out["thetas"] is None and out["yall"].shape == (T, iters + 2) and out["nit"].shape == (T,)
np.savez({out!r} + f".rank{{rank}}.npz", yall=out["yall"].numpy(), nit=out["nit"].numpy(), picked=out["picked"].numpy())
torch.distributed.destroy_process_group()
"""


def test_sharded_bo_trials_equal_the_single_process_run(tmp_path, golden):
    """dist.run_sharded (the host logic of BO.BOloop_batched(devices="ranks")) over two gloo ranks, ragged shards (5 trials):
    every rank ends with the gap curves, iteration counts and picks of ALL trials, bit for bit what one process computes."""
    from oracle import gp as ogp, dpp as odpp
    from conftest import AUTHOR_THETA_RAW
    g = golden("wildcat.npz")
    obj = np.stack([g["surf_0"], g["surf_1"]])[:, :100, :100].reshape(2, -1).astype(np.float32)
    rng = np.random.default_rng(3)
    T, n0, iters = 5, 10, 6
    sid = rng.integers(0, 2, T).astype(np.int32)
    grid = odpp.construct_pool()
    idx0 = np.stack([rng.choice(10000, n0, replace=False) for _ in range(T)])
    X0, y0 = grid[idx0], np.stack([obj[sid[i], idx0[i]] for i in range(T)]).astype(np.float64)
    theta = ogp.raw_to_theta(AUTHOR_THETA_RAW)
    inp, out = str(tmp_path / "in.npz"), str(tmp_path / "out")
    np.savez(inp, obj=obj, sid=sid, X0=X0, y0=y0, theta=theta, iters=iters)
    script = tmp_path / "bo_worker.py"
    script.write_text(BO_WORKER.format(pkg=PKG, root=ROOT, inp=inp, out=out))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ref = [ogp.bo_loop(grid, obj[sid[i]], X0[i], y0[i], iters, theta=theta, kernel_id=3, tol=0.1) for i in range(T)]
    for rank in (0, 1):
        got = np.load(out + f".rank{rank}.npz")
        for i, o in enumerate(ref):
            assert got["nit"][i] == o["nit"]
            assert np.array_equal(got["yall"][i], o["yall"])
            assert np.array_equal(got["picked"][i, :o["nit"]], o["idx"]) and np.all(got["picked"][i, o["nit"]:] == -1)
