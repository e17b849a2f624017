"""Bring-up check of the BF16x3 tensor-core kernel: one configuration per process (a trap cannot take
the others down).  python scripts/tc_bringup.py <case>   or no argument to run every case under timeout."""
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CASES = {
    # name: (B, D, E, hidden, Q, layout, gain, x0_zero, want_f)
    "tiny200": (4, 6, 30, [200, 200, 200], 50, "strided", 1.0, True, False),
    "cfg3s": (64, 6, 30, [200, 200, 200], 50, "strided", 1.0, False, True),
    "cfg3s_trained": (64, 6, 30, [200, 200, 200], 50, "strided", 2.5, False, True),
    "cfg3": (10000, 6, 30, [200, 200, 200], 50, "strided", 1.0, True, True),
    "cfg4s": (2048, 63, 30, [200, 200, 200], 100, "strided", 1.0, True, True),
    "cfg2": (10000, 2, 10, [100, 100, 100, 100], 50, "strided", 1.0, True, True),
    "cfg5": (100, 784, 30, [100, 50, 50, 50, 50], 50, "strided", 1.0, True, True),
    "cfg1": (100, 1, 2, [64, 64, 64], 50, "contig", 1.0, True, True),
    "odd1": (37, 3, 1, [20, 20], 40, "strided", 1.5, False, True),
    "q200": (9, 5, 4, [48, 32, 16], 200, "strided", 1.5, False, True),
    "cfg2_trained": (10000, 2, 10, [100, 100, 100, 100], 50, "strided", 2.5, False, True),
    "cfg5_trained": (100, 784, 30, [100, 50, 50, 50, 50], 50, "strided", 2.5, False, True),
    "cfg3_trained": (10000, 6, 30, [200, 200, 200], 50, "strided", 2.5, False, True),
}


def run_case(name):
    import numpy as np
    import torch
    from oracle import c_binding, umnn_oracle as orc
    from umnn_b200 import IntegrandNN, IntegrandNetwork, _native, cc_integrate
    B, D, E, hidden, Q, layout, gain, x0_zero, want_f = CASES[name]
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]), orc.HIDDEN_LEAKY if layout == "strided" else orc.HIDDEN_RELU)
    flat = orc.synth_params(spec, 0, gain)
    Hh = E * D if layout == "strided" else E
    x0, x, h, _ = orc.synth_inputs(B, D, Hh, 1, x0_zero)
    net = IntegrandNetwork(D, 1 + E, hidden, 1) if layout == "strided" else IntegrandNN(1 + E, hidden)
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p))
            off += p.numel()
    dev = torch.device("cuda:0")
    net.to(dev).eval()   # packed parameters cached: the timing loop is kernel launches only
    xd, hd, x0d = torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev), torch.from_numpy(x0).to(dev)
    res = {}
    variants = (("fp32", _native.PREC_FP32, "1"), ("bf16x3", _native.PREC_BF16X3, "1"),
                ("bf16x3-wide", _native.PREC_BF16X3, "0"), ("fp16x3", _native.PREC_FP16X3, "1"),
                ("fp16x3-wide", _native.PREC_FP16X3, "0"))
    for pname, prec, narrow in variants:
        os.environ["UMNN_B200_TC_NARROW"] = narrow     # read by the launcher on every call
        out, fx, fx0 = cc_integrate(net, x0d, xd, hd, Q, want_fx=want_f, want_fx0=want_f, precision=prec)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        s.record()
        for _ in range(reps):
            cc_integrate(net, x0d, xd, hd, Q, want_fx=want_f, want_fx0=want_f, precision=prec)
        e.record()
        torch.cuda.synchronize()
        res[pname] = (out.cpu().numpy(), None if fx is None else fx.cpu().numpy(),
                      None if fx0 is None else fx0.cpu().numpy(), s.elapsed_time(e) / reps)
    n_chk = min(B, max(4, 200000 // (D * (Q + 2) * 64)))
    ref, rfx, rfx0 = c_binding.cc_forward(spec, flat, x0[:n_chk], x[:n_chk], h[:n_chk], Q, layout)

    def rel(a, b):
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)))
    line = f"{name}: "
    for pname, _, _ in variants:
        o, fx, fx0, ms = res[pname]
        line += f"[{pname} {ms:.3f} ms rel_int={rel(o[:n_chk], ref):.2e}"
        if fx is not None:
            line += f" rel_fx={rel(fx[:n_chk], rfx):.2e} rel_fx0={rel(fx0[:n_chk], rfx0):.2e}"
        line += "] "
    line += (f"bf16x3_vs_fp32_all={rel(res['bf16x3'][0], res['fp32'][0]):.2e} "
             f"fp16x3_vs_fp32_all={rel(res['fp16x3'][0], res['fp32'][0]):.2e} "
             f"speedup={res['fp32'][3] / res['bf16x3'][3]:.2f}x narrow_gain={res['bf16x3-wide'][3] / res['bf16x3'][3]:.2f}x")
    print(line, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for name in CASES:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                                   timeout=180)
                tail = (r.stdout + r.stderr).strip().splitlines()[-6:]
                print(f"--- {name} rc={r.returncode} ({time.time() - t0:.1f}s)")
                print("\n".join(tail), flush=True)
            except subprocess.TimeoutExpired:
                print(f"--- {name} TIMEOUT", flush=True)
