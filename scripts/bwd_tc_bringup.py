"""Bring-up of the tensor-core backward: one case per process, compared with the FFMA backward."""
import os, subprocess, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
CASES = {
    "tiny": (8, 6, 30, [200, 200, 200], 50, False),
    "cfg3s": (256, 6, 30, [200, 200, 200], 50, True),
    "cfg3": (10000, 6, 30, [200, 200, 200], 50, True),
    "cfg2": (10000, 2, 10, [100] * 4, 50, True),
    "cfg5": (100, 784, 30, [100, 50, 50, 50, 50], 50, True),
    "odd": (37, 3, 1, [20, 20], 40, True),
    "cfg4s": (1024, 63, 30, [200, 200, 200], 100, True),
    "cfg3_trained": (10000, 6, 30, [200, 200, 200], 50, True, 2.5),
}

def run_case(name):
    import numpy as np, torch
    from oracle import umnn_oracle as orc
    from umnn_b200 import IntegrandNetwork, kernel, _native
    B, D, E, hidden, Q, jac = CASES[name][:6]
    gain = CASES[name][6] if len(CASES[name]) > 6 else 1.0
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]))
    flat = orc.synth_params(spec, 0, gain)
    x0, x, h, g = orc.synth_inputs(B, D, E * D, 1, x0_zero=False)
    gfx = np.random.RandomState(3).standard_normal(x.shape).astype(np.float32) if jac else None
    net = IntegrandNetwork(D, 1 + E, hidden, 1)
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p)); off += p.numel()
    dev = torch.device("cuda:0"); net.to(dev).eval()
    t = [torch.from_numpy(a).to(dev) for a in (x0, x, h, g)]
    tg = None if gfx is None else torch.from_numpy(gfx).to(dev)
    ks = net.kernel_spec()
    res = {}
    for pname, prec in (("fp32", _native.PREC_FP32), ("bf16x3", _native.PREC_BF16X3), ("fp16x3", _native.PREC_FP16X3)):
        fn = lambda: kernel.cc_backward(ks, t[0], t[1], t[2], t[3], Q, grad_fx=tg, precision=prec)
        out = fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3): fn()
        e.record(); torch.cuda.synchronize()
        res[pname] = (s.elapsed_time(e) / 3, [o.cpu().numpy() for o in out])
    def rtm(a, b): return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
    def nrm(a, b): return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))
    b = res["fp32"][1]
    line = f"{name}: fp32 {res['fp32'][0]:.2f} ms"
    for pname in ("bf16x3", "fp16x3"):
        a = res[pname][1]
        line += (f" | {pname} {res[pname][0]:.2f} ms ({res['fp32'][0] / res[pname][0]:.1f}x) rel-to-max vs fp32: dx0={rtm(a[0], b[0]):.2e} "
                 f"dx={rtm(a[1], b[1]):.2e} dflat={rtm(a[2], b[2]):.2e} dh={rtm(a[3], b[3]):.2e} normwise dflat={nrm(a[2], b[2]):.2e} "
                 f"dh={nrm(a[3], b[3]):.2e}")
    print(line, flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for name in CASES:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=180)
                tail = (r.stdout + r.stderr).strip().splitlines()[-8:]
                print(f"--- {name} rc={r.returncode} ({time.time() - t0:.1f}s)"); print("\n".join(tail), flush=True)
            except subprocess.TimeoutExpired:
                print(f"--- {name} TIMEOUT", flush=True)
