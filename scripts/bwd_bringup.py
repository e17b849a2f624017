"""Timing of the fused backward against the torch route on the device (one process)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from oracle import umnn_oracle as orc
from umnn_b200 import IntegrandNetwork, kernel
from umnn_b200.integral import _integrate_grads_chunked

CASES = [(1000, 6, 30, [200, 200, 200], 50), (10000, 6, 30, [200, 200, 200], 50),
         (10000, 2, 10, [100] * 4, 50), (100, 784, 30, [100, 50, 50, 50, 50], 50)]
if len(sys.argv) > 1:
    CASES = [CASES[int(sys.argv[1])]]
for (B, D, E, hidden, Q) in CASES:
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]))
    flat = orc.synth_params(spec, 0)
    x0, x, h, g = orc.synth_inputs(B, D, E * D, 1)
    net = IntegrandNetwork(D, 1 + E, hidden, 1)
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p)); off += p.numel()
    dev = torch.device("cuda:0"); net.to(dev).eval()
    t = [torch.from_numpy(a).to(dev) for a in (x0, x, h, g)]
    ks = net.kernel_spec()
    def native(): return kernel.cc_backward(ks, t[0], t[1], t[2], t[3], Q)
    def torch_route(): return _integrate_grads_chunked(t[0], t[1], net, t[2], Q, t[3], False)
    res = {}
    for name, fn, reps in (("native", native, 3), ("torch", torch_route, 2)):
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps): out = fn()
        e.record(); torch.cuda.synchronize()
        res[name] = (s.elapsed_time(e) / reps, out)
    dflat_n, dflat_t = res["native"][1][2], res["torch"][1][0].detach()
    err = float((dflat_n - dflat_t).abs().max() / dflat_t.abs().max())
    print(f"B={B} D={D} hidden={hidden}: native {res['native'][0]:.2f} ms, torch route {res['torch'][0]:.2f} ms, "
          f"speedup {res['torch'][0] / res['native'][0]:.2f}x, dflat rel-to-max diff {err:.2e}", flush=True)
