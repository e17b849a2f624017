"""Where do the tensor-core backward's gradient errors come from?  

For a few shapes: the float64 oracle's gradients and per-slot kink margins (oracle.kink_margins), then every native
backward (FP32, FP16x3 with hi-only / A-hi+lo / hi+lo operand panels, BF16x3) measured against float64 with the
flip-aware metric of tests/conftest.py: how many slots deviate by more than 1e-4 of the largest entry, whether each of
them has a hidden unit within the implementation's rounding band of its kink (explained) or not, the largest error
among the slots that do not deviate, and normwise / max errors of d_params.  Also the mutual agreement of the paths.

    python scripts/grad_flip_analysis.py [case ...]      (writes one line per case and path)
"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np

CASES = {
    # name: (B, D, E, hidden, Q, gain)
    "cfg3s": (256, 6, 30, [200, 200, 200], 50, 1.0),
    "cfg3s_trained": (256, 6, 30, [200, 200, 200], 50, 2.5),
    "cfg2s": (512, 2, 10, [100, 100, 100, 100], 50, 1.0),
    "cfg5s": (4, 784, 30, [100, 50, 50, 50, 50], 50, 1.0),
    "cfg4s": (32, 63, 30, [200, 200, 200], 100, 1.0),
}


def main(names):
    import torch
    from conftest import KINK_BAND, dh_to_slots, norm_err, rel_to_max, slot_grad_report
    from oracle import umnn_oracle as orc
    from umnn_b200 import IntegrandNetwork, kernel, _native
    dev = torch.device("cuda:0")
    for name in names:
        B, D, E, hidden, Q, gain = CASES[name]
        spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]))
        flat = orc.synth_params(spec, 0, gain)
        x0, x, h, g = orc.synth_inputs(B, D, E * D, 1, x0_zero=False)
        f64 = lambda a: a.astype(np.float64)
        t0 = time.time()
        r_x0, r_x, r_flat, r_h = orc.integral_backward(spec, f64(flat), f64(x0), f64(x), f64(h), f64(g), Q, "strided", chunk=64)
        margins = orc.kink_margins(spec, flat, x0, x, h, Q, "strided")
        t_or = time.time() - t0
        net = IntegrandNetwork(D, 1 + E, hidden, 1)
        off = 0
        with torch.no_grad():
            for p in net.parameters():
                p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p)); off += p.numel()
        net.to(dev).eval()
        t = [torch.from_numpy(a).to(dev) for a in (x0, x, h, g)]
        ks = net.kernel_spec()
        print(f"== {name}: B={B} D={D} E={E} hidden={hidden} Q={Q} gain={gain}; slots={B * D}, "
              f"kink-ambiguous at 5e-6: {int((margins < 5e-6).sum())}, at 1.5e-4: {int((margins < 1.5e-4).sum())} "
              f"(float64 oracle {t_or:.1f} s)", flush=True)
        got = {}
        for label, prec, panels, band in (("fp32", _native.PREC_FP32, None, "fp32"),
                                          ("fp16x3 panels=hi", _native.PREC_FP16X3, "hi", "fp16x3"),
                                          ("fp16x3 panels=a_hilo", _native.PREC_FP16X3, "a_hilo", "fp16x3"),
                                          ("fp16x3 panels=hilo", _native.PREC_FP16X3, "hilo", "fp16x3"),
                                          ("bf16x3 panels=hilo", _native.PREC_BF16X3, "hilo", "bf16x3")):
            if panels:
                os.environ["UMNN_B200_BWD_PANELS"] = panels
            if _native.lib().umnn_workspace_bytes(kernel.make_desc(ks, t[1], Q, prec), 1) == 0:
                continue
            out = kernel.cc_backward(ks, t[0], t[1], t[2], t[3], Q, precision=prec)
            torch.cuda.synchronize()
            d_x0, d_x, d_flat, d_h = [o.cpu().numpy() for o in out]
            got[label] = (d_flat, d_h)
            rep = slot_grad_report(dh_to_slots(d_h, B, D, "strided"), dh_to_slots(r_h, B, D, "strided"), margins,
                                   KINK_BAND[band], (B, D, -1))
            print(f"  {label:22s} d_h: flipped slots {rep['n_flipped']:4d} (unexplained {rep['n_flipped_unexplained']}), "
                  f"max unflipped {rep['max_unflipped']:.2e}, max all {rep['max_all']:.2e}, normwise {rep['normwise']:.2e} | "
                  f"d_params: normwise {norm_err(d_flat, r_flat):.2e}, max {rel_to_max(d_flat.astype(np.float64), r_flat):.2e} | "
                  f"d_x {rel_to_max(d_x.astype(np.float64), r_x):.1e} d_x0 {rel_to_max(d_x0.astype(np.float64), r_x0):.1e}", flush=True)
        os.environ.pop("UMNN_B200_BWD_PANELS", None)
        base = got.get("fp16x3 panels=hilo")
        if base is not None:
            for label in ("fp32", "fp16x3 panels=hi", "fp16x3 panels=a_hilo", "bf16x3 panels=hilo"):
                if label in got:
                    print(f"  {label:22s} vs fp16x3 panels=hilo: d_params normwise {norm_err(got[label][0], base[0].astype(np.float64)):.2e}, "
                          f"d_h normwise {norm_err(got[label][1], base[1].astype(np.float64)):.2e}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
