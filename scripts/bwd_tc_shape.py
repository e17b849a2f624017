"""Run the default tensor-core backward of a named shape a few times (for ncu launch lists): python scripts/bwd_tc_shape.py cfg5"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from oracle import umnn_oracle as orc
from umnn_b200 import IntegrandNetwork, kernel, _native
SHAPES = {"cfg3": (10000, 6, 30, [200, 200, 200], 50), "cfg2": (10000, 2, 10, [100] * 4, 50),
          "cfg5": (100, 784, 30, [100, 50, 50, 50, 50], 50), "cfg4s": (1024, 63, 30, [200, 200, 200], 100)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
B, D, E, hidden, Q = SHAPES[name]
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
go = torch.randn(B, D, device=dev, generator=g); x0 = torch.zeros_like(x)
ks = net.kernel_spec()
for _ in range(3):
    kernel.cc_backward(ks, x0, x, h, go, Q, precision=_native.PREC_AUTO)
torch.cuda.synchronize()
