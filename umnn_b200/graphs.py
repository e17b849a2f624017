"""CUDA-graph replay of a flow's log-likelihood for a fixed batch shape.

At small per-GPU batches (config 5 sharded over 8 GPUs is 12-13 samples per GPU) UMNNMAFFlow.compute_ll
(models/UMNN/UMNNMAFFlow.py:109-119) is launch-bound: per block a handful of masked-linear GEMMs, the fused
integral launch and ~10 elementwise kernels.  Everything on that path is stream-ordered and free of host
synchronisation, so the whole stack of blocks is captured once and replayed as ONE graph launch.

The parameter packing launches are captured too (kernel.repack_every_call), so a replay always reads the live
parameter values: in-place updates (optimizer steps, load_state_dict) need no re-capture; replacing a parameter
tensor, changing nb_steps or the batch shape does.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernel


class GraphedLogLikelihood:
    """ll, z = graphed(x[, context]) == model.compute_ll(x[, context]) for x of the captured shape, no autograd."""

    def __init__(self, model, batch_size: int, context_size: int = 0, warmup: int = 2):
        first = next(model.parameters())
        if not first.is_cuda:
            raise ValueError("GraphedLogLikelihood needs a model on a CUDA device")
        self.model = model
        dev = first.device
        n_in = model.nets[0].input_size
        self.x = torch.zeros(batch_size, n_in, device=dev)
        self.context = torch.zeros(batch_size, context_size, device=dev) if context_size else None
        was_training = model.training
        model.eval()
        try:
            with torch.no_grad(), kernel.repack_every_call():
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(max(1, warmup)):       # loads the CC tables, sizes the allocator pools
                        model.compute_ll(self.x, context=self.context)
                torch.cuda.current_stream(dev).wait_stream(side)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.ll, self.z = model.compute_ll(self.x, context=self.context)
        finally:
            model.train(was_training)

    def __call__(self, x: torch.Tensor, context: Optional[torch.Tensor] = None):
        if x.shape != self.x.shape:
            raise ValueError(f"captured for x of shape {tuple(self.x.shape)}, got {tuple(x.shape)}")
        self.x.copy_(x)
        if self.context is not None:
            if context is None or context.shape != self.context.shape:
                raise ValueError("captured with a context tensor of shape " + str(tuple(self.context.shape)))
            self.context.copy_(context)
        self.graph.replay()
        return self.ll, self.z
