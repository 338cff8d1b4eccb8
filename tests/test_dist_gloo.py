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
