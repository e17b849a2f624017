"""Per-layer error of the tensor-core backward's weight gradient with hi-only / hi+lo operand panels against the FP32 backward,
for several cotangent structures (one integrand call, config-3 shape).  One GPU."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from oracle import umnn_oracle as orc
from umnn_b200 import IntegrandNetwork, kernel, _native

B, D, E, hidden, Q = int(sys.argv[1]) if len(sys.argv) > 1 else 4000, 6, 30, [200, 200, 200], 50
spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]))
flat = orc.synth_params(spec, 0)
net = IntegrandNetwork(D, 1 + E, hidden, 1)
off = 0
with torch.no_grad():
    for p in net.parameters():
        p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p)); off += p.numel()
dev = torch.device("cuda:0"); net.to(dev).eval()
g = torch.Generator(device=dev).manual_seed(1)
x = 2 * torch.randn(B, D, device=dev, generator=g); h = torch.randn(B, E * D, device=dev, generator=g)
x0 = torch.zeros_like(x)
ks = net.kernel_spec()
sizes = [p.numel() for p in net.parameters()]
names = [n for n, _ in net.named_parameters()]
cases = {
    "g random, no fx": (torch.randn(B, D, device=dev, generator=g), None),
    "g positive, no fx": (torch.rand(B, D, device=dev, generator=g) + 0.5, None),
    "g = 0, fx random": (torch.zeros(B, D, device=dev), torch.randn(B, D, device=dev, generator=g)),
    "g = 0, fx = -1": (torch.zeros(B, D, device=dev), -torch.ones(B, D, device=dev)),
    "g positive, fx = -1 (likelihood-like)": (torch.rand(B, D, device=dev, generator=g) + 0.5, -torch.ones(B, D, device=dev)),
}
for cname, (go, gfx) in cases.items():
    res = {}
    for mode in ("fp32", "hi", "hi_head", "hilo"):
        os.environ.pop("UMNN_B200_BWD_PANELS", None)
        prec = _native.PREC_FP32 if mode == "fp32" else _native.PREC_FP16X3
        if mode != "fp32":
            os.environ["UMNN_B200_BWD_PANELS"] = mode
        res[mode] = kernel.cc_backward(ks, x0, x, h, go, Q, grad_fx=gfx, precision=prec)[2].double()
        torch.cuda.synchronize()
    out = [f"{cname:40s}"]
    for mode in ("hi", "hi_head", "hilo"):
        o = 0
        parts = []
        for n, sz in zip(names, sizes):
            a, r = res[mode][o:o + sz], res["fp32"][o:o + sz]
            parts.append(f"{float((a - r).norm() / r.norm()):.1e}")
            o += sz
        out.append(f"{mode}: " + " ".join(parts))
    print(" | ".join(out), flush=True)
print("columns:", " ".join(names))
