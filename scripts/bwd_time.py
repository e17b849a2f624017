"""Time the default tensor-core backward on a named shape (one line): python scripts/bwd_time.py cfg3 [reps]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from oracle import umnn_oracle as orc
from umnn_b200 import IntegrandNetwork, kernel, _native
SHAPES = {"cfg3": (10000, 6, 30, [200, 200, 200], 50), "cfg2": (10000, 2, 10, [100] * 4, 50),
          "cfg5": (100, 784, 30, [100, 50, 50, 50, 50], 50), "cfg4s": (1024, 63, 30, [200, 200, 200], 100),
          "cfg4m": (8192, 63, 30, [200, 200, 200], 100), "cfg3x4": (40000, 6, 30, [200, 200, 200], 50)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
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
fn = lambda: kernel.cc_backward(ks, x0, x, h, go, Q, precision=_native.PREC_AUTO)
fn(); fn(); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps): fn()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
rows = B * D * (Q + 3)
fpe = 2 * sum(a * b for a, b in zip(spec.widths[:-1], spec.widths[1:]))
print(f"{name}: backward {ms:.3f} ms, {rows} rows, {ms * 1e6 / rows:.3f} ns/row, algorithmic {3 * fpe * rows / (ms * 1e-3) / 1e12:.1f} TFLOP/s "
      f"[panels={os.environ.get('UMNN_B200_BWD_PANELS', 'auto')} kbs={os.environ.get('UMNN_B200_WGRAD_KBS', 'auto')} "
      f"tiles={os.environ.get('UMNN_B200_BWD_TILES', '96')}]", flush=True)
