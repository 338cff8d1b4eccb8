"""Multi-GPU plumbing: trials are independent, so they shard across ranks with no data-path
collective; the only exchange is one gather of the optimality-gap curves at the end.
One process per GPU (torchrun); NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(total: int, rank: int, world: int):
    """Contiguous trial range of ``rank``: sizes differ by at most one, earlier ranks take the extras."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def gather_curves(local: torch.Tensor, counts=None) -> torch.Tensor:
    """All-gather per-rank (T_r, W) tensors into the global (sum T_r, W) tensor, rank order = trial order.
    Equal shards use one all_gather_into_tensor; ragged shards are padded to the largest and trimmed."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        c = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        allc = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(allc, c)
        counts = [int(v.item()) for v in allc]
    mx = max(counts)
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad])
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    if all(cn == mx for cn in counts):
        return out
    return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(world)])


def shard_counts(total: int, world: int):
    """Shard sizes in rank order (a pure function of (total, world): no collective needed to know them)."""
    return [hi - lo for lo, hi in (shard_range(total, r, world) for r in range(world))]


def run_sharded(total: int, run_local, rank=None, world=None):
    """The multi-GPU form of a batch of independent trials (replaces the ipyparallel fan-out of
    Scripts/Experiment2.py:75-92,110-141): this rank runs ``run_local(lo, hi)`` on its contiguous trial range — a dict
    name -> tensor whose leading dimension is hi - lo — and every tensor is all-gathered in rank order, so each rank
    returns the result of the whole batch, trial for trial what one process computes on all of it. No collective on
    the data path; one gather per result tensor at the end."""
    if rank is None or world is None:
        r, w, _ = env_rank_world()
        if dist.is_initialized():
            r, w = dist.get_rank(), dist.get_world_size()
        rank, world = (r if rank is None else rank), (w if world is None else world)
    lo, hi = shard_range(total, rank, world)
    local = run_local(lo, hi)
    if world == 1:
        return local
    counts = shard_counts(total, world)
    return {k: (None if v is None else gather_curves(v, counts)) for k, v in local.items()}


def max_over_ranks(value: float, device) -> float:
    """Max of a per-rank scalar (timings are reported as the slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    """Sum of a per-rank scalar (e.g. the iterations every rank ran)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(device=None):
    if dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None and torch.device(device).type == "cuda":
            dist.barrier(device_ids=[torch.device(device).index or 0])
        else:
            dist.barrier()
