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
from umnn_b200.distributed import BucketedGradientAllReduce, allreduce_gradients, max_over_ranks, shard_bounds


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


# ---- bucketed, hook-launched all-reduce ---------------------------------------------------------------------------
def _bucket_worker(rank, world, port, x_all, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = _model()
    extra = torch.nn.Parameter(torch.ones(3))                 # never used in the loss: contributes zeros.  It sits in
    params = [extra] + list(model.parameters())               # the LAST bucket (buckets fill in reverse order and are
                                                              # launched in order, so an unused parameter holds back
                                                              # only the buckets after its own)
    reducer = BucketedGradientAllReduce(params, bucket_bytes=2048)      # small buckets: several collectives in flight
    assert len(reducer.buckets) >= 4 and reducer.world == world
    b, e = shard_bounds(x_all.shape[0], rank, world)
    x = x_all[b:e]
    results = {}
    # step 1: plain step.  step 2: the caller resets .grad to None (optimizer.zero_grad default) -- the hook adopts the
    # fresh gradient tensors.  step 3: two accumulated sub-batches, the first under no_sync().
    for step in (1, 2, 3):
        reducer.zero_grad()
        if step == 2:
            for p in params:
                p.grad = None
        if step == 3:
            half = x.shape[0] // 2
            with reducer.no_sync():
                (-model.compute_ll(x[:half])[0].sum() / x_all.shape[0] * world).backward()
            (-model.compute_ll(x[half:])[0].sum() / x_all.shape[0] * world).backward()
        else:
            (-model.compute_ll(x)[0].sum() / x_all.shape[0] * world).backward()
        launched_before_finish = reducer._next
        reducer.finish()
        for p in reducer.params:      # every .grad is (again) a view into its bucket
            bi, i = reducer._bucket_of[id(p)]
            assert p.grad.data_ptr() == reducer.buckets[bi].flat[reducer.buckets[bi].offsets[i]:].data_ptr()
        results[step] = {"grads": {k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad},
                         "extra": extra.grad.clone(), "launched_before_finish": launched_before_finish}
    torch.save(results, os.path.join(out_dir, f"bucket_rank{rank}.pt"))
    reducer.close()
    dist.destroy_process_group()


def test_bucketed_allreduce_overlaps_and_matches_single_process(tmp_path):
    world = 2
    torch.manual_seed(0)
    x_all = torch.randn(11, 4)
    mp.spawn(_bucket_worker, args=(world, _free_port(), x_all, str(tmp_path)), nprocs=world, join=True)
    model = _model()
    (-model.compute_ll(x_all)[0].sum() / x_all.shape[0]).backward()
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"bucket_rank{r}.pt"))
        for step in (1, 2, 3):
            # buckets were on the wire before finish(): launched from the gradient hooks during the backward
            assert res[step]["launched_before_finish"] >= 1
            assert torch.count_nonzero(res[step]["extra"]) == 0
            for k, p in model.named_parameters():
                if p.requires_grad and p.grad is not None:
                    assert torch.allclose(res[step]["grads"][k], p.grad, rtol=2e-4, atol=1e-6), (r, step, k)


def test_bucketed_allreduce_is_a_noop_without_a_process_group():
    model = _model()
    reducer = BucketedGradientAllReduce(model.parameters(), bucket_bytes=4096)
    reducer.zero_grad()
    x = torch.randn(5, 4)
    (-model.compute_ll(x)[0].mean()).backward()
    reducer.finish()
    ref = _model()
    (-ref.compute_ll(x)[0].mean()).backward()
    for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        if q.grad is not None:
            assert torch.allclose(p.grad, q.grad, rtol=1e-6, atol=1e-7), k
    assert reducer.exposed_ms() is None and reducer.total_bytes > 0
