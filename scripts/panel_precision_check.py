"""Gradient of a flow's mean log-likelihood with the default backward (hi-only operand panels from one chunk of rows on),
hi + lo panels and a_hilo against the FP32 backward, normwise, per parameter group.  One GPU.

    python scripts/panel_precision_check.py [B]
"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch


def grads(mode, B):
    from umnn_b200 import UMNNMAFFlow
    for k in ("UMNN_B200_BWD_PANELS", "UMNN_B200_BACKWARD"):
        os.environ.pop(k, None)
    if mode == "fp32":
        os.environ["UMNN_B200_BACKWARD"] = "fp32"
    elif mode != "default":
        os.environ["UMNN_B200_BWD_PANELS"] = mode
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = UMNNMAFFlow(nb_flow=3, nb_in=6, hidden_derivative=[200, 200, 200], hidden_embedding=[256, 256], embedding_s=30,
                        nb_steps=50, solver="CCParallel", device=dev).to(dev)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, 6, generator=gen)
    x[:, 1] = 0.5 * x[:, 0] ** 2 + 0.3 * x[:, 1]
    ll, _ = model.compute_ll(x.to(dev))
    (-ll.mean()).backward()
    torch.cuda.synchronize()
    integ, cond = [], []
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        (integ if "integrand" in n or "parallel_nets" in n or ".net.net" in n and "made" not in n.lower() else cond).append((n, p.grad.detach().double().reshape(-1).clone()))
    return dict(model.named_parameters()), {n: p.grad.detach().double().reshape(-1).clone() for n, p in model.named_parameters() if p.grad is not None}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4001
    _, ref = grads("fp32", B)
    names = list(ref)
    groups = {"all": names, "integrand": [n for n in names if "integrand" in n.lower() or "derivative" in n.lower()],
              "conditioner": [n for n in names if not ("integrand" in n.lower() or "derivative" in n.lower())]}
    print("parameter names (first 6):", names[:6])
    if "--per-tensor" in sys.argv:
        _, g = grads(os.environ.get("UMNN_B200_BWD_PANELS", "hi"), B)
        tot = float(torch.cat([ref[n] for n in names]).norm())
        for n in names:
            d = float((g[n] - ref[n]).norm())
            print(f"  {n:44s} numel {ref[n].numel():7d} |ref| {float(ref[n].norm()):.3e} |err| {d:.3e} rel {d / max(float(ref[n].norm()), 1e-30):.2e} share-of-total {d / tot:.2e}")
        return
    for mode in ("default", "hi_head", "hi", "hilo"):
        _, g = grads(mode, B)
        line = [f"B={B} rows/block={B * 6 * 53} mode={mode:8s}"]
        for gname, ns in groups.items():
            if not ns:
                continue
            a = torch.cat([g[n] for n in ns]); r = torch.cat([ref[n] for n in ns])
            line.append(f"{gname}: |d|/|ref| {float((a - r).norm() / r.norm()):.2e} (|ref| {float(r.norm()):.3e})")
        print("  ".join(line), flush=True)


if __name__ == "__main__":
    main()
