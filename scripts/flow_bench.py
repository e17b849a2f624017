"""Flow-level timings: UMNNMAFFlow.compute_ll forward, a full training step (forward + backward +
Adam) and invert, for flows shaped like the reference's drivers, on the kernel route and -- where memory allows --
inside `umnn_b200.torch_route()` (the reference's algorithm as torch-CUDA ops in the same process).

  python scripts/flow_bench.py [toy|power|bsds|mnist ...] [--no-torch]
"""
import contextlib, io, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch

# name: (D, E, Q, integrand hidden, MADE hidden, blocks, B for ll / train, B for invert, torch-route batch or None)
FLOWS = {
    "toy":   dict(D=2,   E=10, Q=50,  hid=[100] * 4,            made=[100] * 4,   blocks=5, B=10000, B_inv=1000, B_torch=10000),
    "power": dict(D=6,   E=30, Q=50,  hid=[200] * 3,            made=[512, 512],  blocks=5, B=10000, B_inv=1000, B_torch=10000),
    "bsds":  dict(D=63,  E=30, Q=100, hid=[200] * 3,            made=[1024, 1024], blocks=5, B=8192,  B_inv=64,   B_torch=256),
    "mnist": dict(D=784, E=30, Q=50,  hid=[100, 50, 50, 50, 50], made=[1024] * 3,  blocks=5, B=100,   B_inv=16,   B_torch=100),
}


def timed(fn, warmup, reps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def build(cfg, blocks=None):
    from umnn_b200 import UMNNMAFFlow
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    return UMNNMAFFlow(nb_flow=blocks or cfg["blocks"], nb_in=cfg["D"], hidden_derivative=cfg["hid"],
                       hidden_embedding=cfg["made"], embedding_s=cfg["E"], nb_steps=cfg["Q"], solver="CCParallel",
                       device=dev).to(dev)


def bench_route(name, cfg, route, B, do_invert):
    from umnn_b200 import torch_route
    os.environ["UMNN_B200_INVERT"] = "torch" if route == "torch" else ""
    with (torch_route() if route == "torch" else contextlib.nullcontext()):
        _bench_route(name, cfg, route, B, do_invert)


def _bench_route(name, cfg, route, B, do_invert):
    dev = torch.device("cuda:0")
    out = {"flow": name, "route": route, "B": B, "D": cfg["D"], "Q": cfg["Q"], "blocks": cfg["blocks"]}
    model = build(cfg)
    x = torch.randn(B, cfg["D"], device=dev)
    model.eval()
    with torch.no_grad():
        out["compute_ll_ms"] = round(timed(lambda: model.compute_ll(x), 2, 5), 3)
    if route == "kernel":
        from umnn_b200 import GraphedLogLikelihood
        graphed = GraphedLogLikelihood(model, B)
        out["compute_ll_graph_ms"] = round(timed(lambda: graphed(x), 2, 5), 3)
        small = min(B, 16)
        xs = x[:small].contiguous()
        graphed_s = GraphedLogLikelihood(model, small)
        with torch.no_grad():
            out[f"compute_ll_B{small}_ms"] = round(timed(lambda: model.compute_ll(xs), 3, 20), 3)
        out[f"compute_ll_B{small}_graph_ms"] = round(timed(lambda: graphed_s(xs), 3, 20), 3)
        del graphed, graphed_s
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        opt.zero_grad()
        ll, _ = model.compute_ll(x)
        (-ll.mean()).backward()
        opt.step()
    torch.cuda.reset_peak_memory_stats()
    out["train_step_ms"] = round(timed(step, 2, 3), 3)
    out["train_peak_GiB"] = round(torch.cuda.max_memory_allocated() / 2**30, 2)
    if do_invert:
        one = build(cfg, blocks=1).eval()
        zi = torch.randn(cfg["B_inv"], cfg["D"], device=dev)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            out["invert_1block_ms"] = round(timed(lambda: one.invert(zi, iter=10), 1, 1), 1)
        out["B_invert"] = cfg["B_inv"]
    del model, opt
    torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name in (args or list(FLOWS)):
        cfg = FLOWS[name]
        bench_route(name, cfg, "kernel", cfg["B"], True)
        if "--no-torch" not in sys.argv and cfg["B_torch"]:
            try:
                bench_route(name, cfg, "torch", cfg["B_torch"], True)
            except torch.cuda.OutOfMemoryError as ex:
                print(json.dumps({"flow": name, "route": "torch", "error": "out of memory"}), flush=True)
                torch.cuda.empty_cache()
