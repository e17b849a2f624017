"""world_size-2 gloo tests of the N>1 host logic: batch sharding with no data-path collective in the
forward, and the single flat gradient all-reduce of training."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from umnn_b200 import UMNNMAFFlow
from umnn_b200.distributed import allreduce_gradients, max_over_ranks, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 65536, 100):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(3)
    return UMNNMAFFlow(nb_flow=2, nb_in=4, hidden_derivative=[24, 24], hidden_embedding=[32, 32], embedding_s=5,
                       nb_steps=15, solver="CCParallel")


def _worker(rank, world, port, x_all, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = _model()
    b, e = shard_bounds(x_all.shape[0], rank, world)
    x = x_all[b:e]
    # forward: no collective -- every rank computes its own shard
    ll, z = model.compute_ll(x)
    # training: local sum of log-likelihoods, then ONE flat all-reduce of the gradients
    (-ll.sum() / x_all.shape[0] * world).backward()
    allreduce_gradients(model.parameters())
    t = max_over_ranks(float(rank + 1))
    torch.save({"ll": ll.detach(), "z": z.detach(), "t": t,
                "grads": {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_sharded_step_matches_single_process(tmp_path):
    world = 2
    torch.manual_seed(0)
    x_all = torch.randn(11, 4)          # odd batch: ragged shards (6 + 5)
    mp.spawn(_worker, args=(world, _free_port(), x_all, str(tmp_path)), nprocs=world, join=True)
    model = _model()
    ll, z = model.compute_ll(x_all)
    (-ll.sum() / x_all.shape[0]).backward()
    parts = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    ll_cat = torch.cat([p["ll"] for p in parts])
    z_cat = torch.cat([p["z"] for p in parts])
    assert torch.allclose(ll_cat, ll.detach(), rtol=1e-6, atol=1e-6)
    assert torch.allclose(z_cat, z.detach(), rtol=1e-6, atol=1e-6)
    assert all(p["t"] == 2.0 for p in parts)
    for r in range(world):
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            g = parts[r]["grads"][k]
            assert torch.allclose(g, p.grad, rtol=2e-4, atol=1e-6), (r, k)
