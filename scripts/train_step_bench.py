"""End-to-end training step of a UCI-POWER-shaped flow (config 3: 5 blocks, MADE [512,512], integrand [200]^3,
E=30, Q=50, B=10000): compute_ll forward + backward + Adam step, per backward path."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch

def run(mode, B=10000, steps=5):
    os.environ["UMNN_B200_BACKWARD"] = mode
    from umnn_b200 import UMNNMAFFlow
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    model = UMNNMAFFlow(nb_flow=5, nb_in=6, hidden_derivative=[200, 200, 200], hidden_embedding=[512, 512],
                        embedding_s=30, nb_steps=50, solver="CCParallel", device=dev).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x = torch.randn(B, 6, device=dev)
    def step():
        opt.zero_grad()
        ll, _ = model.compute_ll(x)
        loss = -ll.mean()
        loss.backward()
        opt.step()
        return loss
    for _ in range(2): step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps): loss = step()
    e.record(); torch.cuda.synchronize()
    print(f"backward={mode}: {s.elapsed_time(e) / steps:.1f} ms per training step (B={B}, 5 blocks), loss {float(loss.detach()):.4f}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)

if __name__ == "__main__":
    for mode in (sys.argv[1:] or ["auto", "fp32", "torch"]):
        torch.cuda.reset_peak_memory_stats()
        run(mode)
