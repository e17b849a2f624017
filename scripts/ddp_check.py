"""NCCL check of the data-parallel training step (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_check.py

Every rank trains the same UMNNMAFFlow for a few Adam steps on ITS shard of a global batch with
BucketedGradientAllReduce (buckets all-reduced on NCCL from the gradient hooks); rank 0 also trains a copy on the
whole batch alone.  The sharded run must follow the single-GPU run (same mean-loss gradient up to summation order),
all ranks must end with identical parameters, and at least one bucket must have been launched before finish().
"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import torch.distributed as dist

def main():
    from umnn_b200 import UMNNMAFFlow
    from umnn_b200.distributed import BucketedGradientAllReduce, shard_bounds
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)

    def build():
        torch.manual_seed(0)
        return UMNNMAFFlow(nb_flow=3, nb_in=6, hidden_derivative=[200, 200, 200], hidden_embedding=[256, 256], embedding_s=30,
                           nb_steps=50, solver="CCParallel", device=dev).to(dev)
    gen = torch.Generator().manual_seed(1)
    B = 4001                                   # odd: ragged shards
    x_all = torch.randn(B, 6, generator=gen)
    x_all[:, 1] = 0.5 * x_all[:, 0] ** 2 + 0.3 * x_all[:, 1]
    b, e = shard_bounds(B, rank, world)
    x = x_all[b:e].to(dev)

    model = build()
    opt = torch.optim.Adam(model.parameters(), 1e-3)
    red = BucketedGradientAllReduce(model.parameters(), bucket_bytes=256 << 10, weight=(e - b) * world / B)
    early, losses = [], []
    grad0 = None
    for it in range(4):
        red.zero_grad()
        ll, _ = model.compute_ll(x)
        loss = -ll.mean()
        loss.backward()
        early.append(red._next)
        red.finish()
        if it == 0:       # the all-reduced mean-loss gradient before any update: compared with the single-GPU gradient below
            torch.cuda.synchronize()
            grad0 = torch.cat([p.grad.detach().reshape(-1).clone() for p in model.parameters() if p.requires_grad])
        opt.step()
        t = torch.tensor([float(loss.detach()) * (e - b) / B], device=dev)
        dist.all_reduce(t)
        losses.append(float(t))
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters() if p.requires_grad])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        ref = build()
        ropt = torch.optim.Adam(ref.parameters(), 1e-3)
        rl = []
        xa = x_all.to(dev)
        rgrad0 = None
        for it in range(4):
            ropt.zero_grad()
            ll, _ = ref.compute_ll(xa)
            loss = -ll.mean()
            loss.backward()
            if it == 0:
                rgrad0 = torch.cat([p.grad.detach().reshape(-1).clone() for p in ref.parameters() if p.requires_grad])
            ropt.step()
            rl.append(float(loss.detach()))
        rflat = torch.cat([p.detach().reshape(-1) for p in ref.parameters() if p.requires_grad])
        dparam = float((flat - rflat).abs().max())
        # Adam divides by sqrt(v): an entry whose gradient is ~0 moves by up to lr per step whatever its sign, so the parameter
        # drift after 4 steps is bounded by 4 * lr, not by the gradient error.  The gradient itself is the tight check.
        dgrad = float((grad0 - rgrad0).norm() / rgrad0.norm())
        dloss = max(abs(a - b2) for a, b2 in zip(losses, rl))
        print(f"ddp_check world={world}: buckets={len(red.buckets)} launched-before-finish={early} exposed_ms={red.exposed_ms():.3f} "
              f"ranks identical={same} | losses sharded {['%.5f' % v for v in losses]} single {['%.5f' % v for v in rl]} "
              f"max|dloss|={dloss:.2e} |dgrad|/|grad|={dgrad:.2e} max|dparam|={dparam:.2e} (Adam, 4 steps of lr 1e-3)", flush=True)
        ok = same and min(early) >= 1 and dloss < 2e-3 and dgrad < 2e-3 and dparam <= 4.5e-3
        print("DDP_CHECK_OK" if ok else "DDP_CHECK_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
