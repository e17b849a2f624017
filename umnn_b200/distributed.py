"""Multi-GPU plumbing for the hot path: one process per GPU, batch-sharded, no data-path collective.

Every (sample, dimension) slot is independent given h, and h is per-sample, so the forward pass and
`compute_ll` shard over the batch with ZERO exchange.  The only collective is the
training gradient all-reduce (NCCL over NVLink/NVSwitch on the GPUs, gloo in the CPU tests).  The reference
trains on one device (UCIExperiments.py:129-144, MNISTExperiment.py:160-167); this is the data-parallel
version of that loop's `loss.backward(); optimizer.step()`:

    reducer = BucketedGradientAllReduce(model.parameters())
    for x in shards:
        reducer.zero_grad()                 # instead of optimizer.zero_grad(): keeps the flat gradient views
        loss = -model.compute_ll(x)[0].mean()
        loss.backward()                     # buckets are all-reduced while the rest of the backward still runs
        reducer.finish()                    # waits for the collectives; .grad now holds the global mean gradient
        optimizer.step()

Gradients live in a few flat buckets and every `p.grad` is a VIEW into its bucket, so nothing is concatenated or
copied back around the collective (the round-1 helper did both: 2 x 540 MB per step at the MNIST shape).  A
bucket's all-reduce is launched asynchronously from the gradient hook of the last parameter that fills it, in
bucket order on every rank, on the process group's own stream; `finish()` makes the compute stream wait for them
and reports how long it had to (the exposed communication time).
"""
from __future__ import annotations

import contextlib
from typing import Iterable, List, Optional, Tuple

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


class _Bucket:
    __slots__ = ("flat", "params", "offsets", "pending", "work", "launched")

    def __init__(self, flat, params, offsets):
        self.flat, self.params, self.offsets = flat, params, offsets
        self.pending = len(params)
        self.work = None
        self.launched = False


class BucketedGradientAllReduce:
    """Data-parallel gradient averaging: flat gradient buckets, all-reduced asynchronously from gradient hooks.

    params        the parameters to synchronise (every rank must pass the same list in the same order); those with
                  requires_grad=False are ignored.
    bucket_bytes  target bucket size.  Buckets are filled in REVERSE parameter order -- the order in which a
                  backward pass produces gradients -- so the first bucket is complete (and on the wire) while the
                  conditioner of the earlier flow blocks is still back-propagating.
    weight        factor applied to this rank's gradient before the reduction (e.g. local_n * world / global_n for
                  unequal shards of a mean loss); the result is sum_r weight_r * grad_r / world.
    group         process group (default: the world).  Without an initialised process group everything degrades to a
                  single-rank no-op, so the same training loop runs on one GPU.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 4 << 20, weight: float = 1.0,
                 group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.weight = float(weight)
        self.active = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.active else 1
        self.buckets: List[_Bucket] = []
        self._bucket_of = {}
        self._hooks = []
        self._next = 0
        self._sync = True
        self._events = None
        cur: List[torch.nn.Parameter] = []
        cur_bytes = 0
        key = None
        for p in reversed(self.params):
            k = (p.device, p.dtype)
            nbytes = p.numel() * p.element_size()
            if cur and (k != key or cur_bytes + nbytes > bucket_bytes):
                self._close(cur)
                cur, cur_bytes = [], 0
            key = k
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ----- construction
    def _close(self, plist):
        total = sum(p.numel() for p in plist)
        flat = torch.zeros(total, dtype=plist[0].dtype, device=plist[0].device)
        offsets, off = [], 0
        for p in plist:
            offsets.append(off)
            off += p.numel()
        b = _Bucket(flat, list(plist), offsets)
        idx = len(self.buckets)
        self.buckets.append(b)
        for i, p in enumerate(plist):
            self._bucket_of[id(p)] = (idx, i)
            p.grad = flat[offsets[i]:offsets[i] + p.numel()].view_as(p)

    @contextlib.contextmanager
    def no_sync(self):
        """Backward passes inside this context only accumulate into the buckets (gradient accumulation over
        sub-batches, MNISTExperiment.py:160-164); the LAST backward of a step runs outside it and reduces."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    # ----- per step
    def zero_grad(self) -> None:
        """Zero the flat buckets and re-point every .grad at its view (use instead of optimizer.zero_grad())."""
        self._next = 0
        for b in self.buckets:
            b.flat.zero_()
            b.pending = len(b.params)
            b.work = None
            b.launched = False
            for i, p in enumerate(b.params):
                v = b.flat[b.offsets[i]:b.offsets[i] + p.numel()].view_as(p)
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    p.grad = v

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi, i = self._bucket_of[id(p)]
        b = self.buckets[bi]
        v = b.flat[b.offsets[i]:b.offsets[i] + p.numel()].view_as(p)
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
            # someone reset .grad (optimizer.zero_grad(set_to_none=True)): adopt the fresh tensor's values
            v.copy_(p.grad)
            p.grad = v
        if not self._sync:
            return
        b.pending -= 1
        self._launch_ready()

    def _launch_ready(self) -> None:
        # strictly in bucket order, so every rank issues the same sequence of collectives
        while self._next < len(self.buckets) and self.buckets[self._next].pending <= 0:
            self._launch(self.buckets[self._next])
            self._next += 1

    def _launch(self, b: _Bucket) -> None:
        b.launched = True
        if not self.active or self.world == 1:
            return
        if self.weight != 1.0:
            b.flat.mul_(self.weight)
        b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        """Launch whatever has not been launched (parameters that received no gradient contribute zeros), wait for
        every collective and divide by the world size.  On CUDA the wait is a stream dependency, not a host block;
        `exposed_ms()` (after a synchronize) says how long the compute stream stalled in here."""
        cuda = self.params and self.params[0].is_cuda
        ev = None
        if cuda and self.active and self.world > 1:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        for b in self.buckets[self._next:]:
            self._launch(b)
        self._next = len(self.buckets)
        if self.active and self.world > 1:
            for b in self.buckets:
                if b.work is not None:
                    b.work.wait()
                    b.work = None
            for b in self.buckets:
                b.flat.div_(self.world)
        for b in self.buckets:      # parameters that received no gradient this step: .grad = their (reduced) zeros
            for i, p in enumerate(b.params):
                if p.grad is None:
                    p.grad = b.flat[b.offsets[i]:b.offsets[i] + p.numel()].view_as(p)
        if ev is not None:
            ev[1].record()
        self._events = ev

    def exposed_ms(self) -> Optional[float]:
        """Time the compute stream spent waiting inside the last finish() (call after torch.cuda.synchronize())."""
        if self._events is None:
            return None
        return self._events[0].elapsed_time(self._events[1])

    def close(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def total_bytes(self) -> int:
        return sum(b.flat.numel() * b.flat.element_size() for b in self.buckets)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: int = None, weight: float = 1.0) -> None:
    """One-shot form (no hooks, no overlap): average the existing .grad tensors of `params` over all ranks.

    Kept for loops that cannot hold a reducer object; it reduces each gradient tensor in place (one asynchronous
    collective per tensor, then one wait), without building a flat copy.  Parameters without a gradient on this
    rank contribute zeros (every rank must pass the same list).  Prefer BucketedGradientAllReduce.
    """
    if not dist.is_initialized():
        return
    world = dist.get_world_size() if world is None else world
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    grads = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        if weight != 1.0:
            p.grad.mul_(weight)
        grads.append(p.grad)
    works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True) for g in grads]
    for w in works:
        w.wait()
    for g in grads:
        g.div_(world)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar (e.g. a CUDA-event time in ms) over all ranks."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
