import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from oracle import umnn_oracle as orc
from umnn_b200 import IntegrandNetwork, kernel, _native
B, D, E, hidden, Q = 256, 6, 30, [200, 200, 200], 50
spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]))
flat = orc.synth_params(spec, 0)
x0, x, h, g = orc.synth_inputs(B, D, E * D, 1, x0_zero=False)
net = IntegrandNetwork(D, 1 + E, hidden, 1)
off = 0
with torch.no_grad():
    for p in net.parameters():
        p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p)); off += p.numel()
dev = torch.device("cuda:0"); net.to(dev).eval()
t = [torch.from_numpy(a).to(dev) for a in (x0, x, h, g)]
ks = net.kernel_spec()
for jac in (False, True):
    gfx = np.random.RandomState(3).standard_normal(x.shape).astype(np.float32) if jac else None
    tg = None if gfx is None else torch.from_numpy(gfx).to(dev)
    a = [o.cpu().numpy() for o in kernel.cc_backward(ks, t[0], t[1], t[2], t[3], Q, grad_fx=tg, precision=_native.PREC_BF16X3)]
    b = [o.cpu().numpy() for o in kernel.cc_backward(ks, t[0], t[1], t[2], t[3], Q, grad_fx=tg, precision=_native.PREC_FP32)]
    dh_a = a[3].reshape(B, E, D); dh_b = b[3].reshape(B, E, D)
    err = np.abs(dh_a - dh_b).max(axis=1)          # [B, D] per slot
    scale = np.abs(dh_b).max()
    flat_err = err.reshape(-1) / scale
    worst = np.argsort(-flat_err)[:12]
    print("jac", jac, "dh rel-to-max", flat_err.max(), "worst slots", worst.tolist(), [f"{flat_err[i]:.1e}" for i in worst])
    print("   median slot err", np.median(flat_err), " frac > 1e-4:", float((flat_err > 1e-4).mean()))
    ex = np.abs(a[1] - b[1]).reshape(-1) / np.abs(b[1]).max()
    print("   dx worst", np.argsort(-ex)[:8].tolist(), [f"{v:.1e}" for v in np.sort(-ex)[:8] * -1])
    # per-layer dflat error
    o = 0
    for l, (i_, o_) in enumerate(zip(spec.widths[:-1], spec.widths[1:])):
        nW = i_ * o_
        eW = np.abs(a[2][o:o + nW] - b[2][o:o + nW]).max() / np.abs(b[2][o:o + nW]).max(); o += nW
        eb = np.abs(a[2][o:o + o_] - b[2][o:o + o_]).max() / max(np.abs(b[2][o:o + o_]).max(), 1e-30); o += o_
        print(f"   layer {l}: dW rel-to-max {eW:.2e}  db {eb:.2e}")

# structure of the dW2 error (last run: jac True)
o = 31 * 200 + 200
W2a = a[2][o:o + 40000].reshape(200, 200); W2b = b[2][o:o + 40000].reshape(200, 200)
err = np.abs(W2a - W2b) / np.abs(W2b).max()
print("dW2 err by row block (n):", [f"{err[i:i+25].max():.1e}" for i in range(0, 200, 25)])
print("dW2 err by col block (k):", [f"{err[:, i:i+25].max():.1e}" for i in range(0, 200, 25)])
print("dW2 normwise rel err:", np.linalg.norm(W2a - W2b) / np.linalg.norm(W2b), " dh normwise:", np.linalg.norm(a[3] - b[3]) / np.linalg.norm(b[3]))
print("dW2 max|.|", np.abs(W2b).max(), "median|.|", np.median(np.abs(W2b)))
idx = np.unravel_index(np.argsort(-err.reshape(-1))[:10], err.shape)
print("worst (n,k):", list(zip(idx[0].tolist(), idx[1].tolist())), [f"{err[i,j]:.1e}" for i, j in zip(*idx)])
