"""Multi-GPU plumbing for the hot path: one process per GPU, batch-sharded, no data-path collective.

Every (sample, dimension) slot is independent given h, and h is per-sample, so the forward pass and
`compute_ll` shard over the batch with ZERO exchange.  The only collective is the
training gradient all-reduce (NCCL over NVLink on the GPUs, gloo in the CPU tests): one flat bucket,
sum then divide by the world size.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of `n` samples owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_bounds(t.shape[0], rank, world)
    return t[b:e]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: int = None, weight: float = 1.0) -> None:
    """Sum the gradients of `params` over all ranks in ONE flat all-reduce and scale by weight/world.

    Parameters without a gradient on this rank contribute zeros (every rank must pass the same list).
    With per-rank losses that are means over the local shard, pass weight = local_n * world / global_n
    beforehand (or use equal shards) so the result is the gradient of the global mean.
    """
    if not dist.is_initialized():
        return
    world = dist.get_world_size() if world is None else world
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) * weight for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar (e.g. a CUDA-event time in ms) over all ranks."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
