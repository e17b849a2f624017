"""Tiny forward + backward through every kernel family (run under compute-sanitizer)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from oracle import umnn_oracle as orc
from umnn_b200 import IntegrandNetwork, kernel, _native, cc_integrate

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B, D, E, hidden, Q = 24, 3, 6, [40, 24, 24], 30
if which == "multitile":
    # several tiles per CTA in every tensor-core kernel (narrow forward: > 296 x 128 rows; wide passes F / D: > 148 x 128):
    # the tile-to-tile hand-overs (prep buffers, "partials full / empty", "d0 full / empty", the straddling slot's carry)
    # only exist from the second tile on
    B = 700
spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]))
flat = orc.synth_params(spec, 0, 1.5)
x0, x, h, g = orc.synth_inputs(B, D, E * D, 1, x0_zero=False)
net = IntegrandNetwork(D, 1 + E, hidden, 1)
off = 0
with torch.no_grad():
    for p in net.parameters():
        p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p)); off += p.numel()
dev = torch.device("cuda:0"); net.to(dev).eval()
t = [torch.from_numpy(a).to(dev) for a in (x0, x, h, g)]
ks = net.kernel_spec()
ref = orc.integrate_parallel(spec, flat, x0, x, h, Q)
# [40, 24, 24] takes the narrow (two CTAs per SM) forward shape; the backward's pass F is always the wide shape.
# fp16x3 adds the guarded bf16 re-runs (no-op launches) of the forward and of the three backward passes.
# fp16x3_hi: the same with hi-only operand panels and 3 blocks per pass-W stage (the large-batch configuration);
# overflow: inputs that leave the fp16 range, so the guarded FP32 re-runs (forward and backward) really execute.
for name, prec in (("fp32", _native.PREC_FP32), ("bf16x3", _native.PREC_BF16X3), ("fp16x3", _native.PREC_FP16X3),
                   ("fp16x3_hi", _native.PREC_FP16X3), ("overflow", _native.PREC_FP16X3), ("multitile", _native.PREC_FP16X3)):
    if which not in ("all", name) or (which == "all" and name == "multitile"):
        continue
    if name == "fp16x3_hi":
        os.environ["UMNN_B200_BWD_PANELS"] = "hi"
        os.environ["UMNN_B200_WGRAD_KBS"] = "3"
    xs = t[1] * 3.0e6 if name == "overflow" else t[1]
    out, fx, fx0 = cc_integrate(net, t[0], xs, t[2], Q, want_fx=True, want_fx0=True, precision=prec)
    grads = kernel.cc_backward(ks, t[0], xs, t[2], t[3], Q, grad_fx=t[3], precision=prec)
    torch.cuda.synchronize()
    os.environ.pop("UMNN_B200_BWD_PANELS", None)
    os.environ.pop("UMNN_B200_WGRAD_KBS", None)
    if name == "overflow":
        print(f"{name}: finite {bool(torch.isfinite(out).all())}, |d_params| {float(grads[2].abs().sum()):.4e}", flush=True)
        continue
    err = float(np.max(np.abs(out.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1e-6)))
    print(f"{name}: forward rel-err {err:.2e}, |d_params| {float(grads[2].abs().sum()):.4f}", flush=True)
# sampling direction: one dimension through umnn_invert_dimension
if which in ("all", "invert"):
    from umnn_b200 import UMNNMAFFlow
    import contextlib, io
    torch.manual_seed(0)
    flow = UMNNMAFFlow(nb_flow=1, nb_in=3, hidden_derivative=[40, 24], hidden_embedding=[32, 32], embedding_s=6, nb_steps=20,
                       solver="CCParallel", device=dev).to(dev).eval()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        xb = flow.invert(torch.randn(5, 3, device=dev), iter=3)
    torch.cuda.synchronize()
    print(f"invert: finite {bool(torch.isfinite(xb).all())}", flush=True)
