#!/usr/bin/env python
"""Generate the golden vectors that pin ``oracle/`` (and, through it, the CUDA path).

Run ONLY in the build container, where the upstream reference is mounted read-only
at /root/reference:

    python tests/golden/make_golden.py

It imports the *unmodified* reference package (``models.UMNN`` from /root/reference),
feeds it the seeded synthetic inputs of ``oracle.umnn_oracle.synth_*`` and writes the
reference's own outputs to ``tests/golden/*.npz``.  Only outputs, shapes and seeds are
stored (inputs are regenerated from the frozen ``numpy.random.RandomState`` stream; an
input checksum guards against drift).  Nothing here runs on the GPU box.
"""
import hashlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("UMNN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)          # reference `models` package must win over the repo's shim
sys.path.append(REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models.UMNN.MonotonicNN import IntegrandNN  # noqa: E402
from models.UMNN.NeuralIntegral import NeuralIntegral  # noqa: E402
from models.UMNN.ParallelNeuralIntegral import ParallelNeuralIntegral  # noqa: E402
from models.UMNN.UMNNMAF import IntegrandNetwork  # noqa: E402
from models.UMNN.UMNNMAFFlow import UMNNMAFFlow  # noqa: E402

assert os.path.realpath(sys.modules["models"].__file__).startswith(os.path.realpath(REF)), "wrong `models` imported"

from oracle import umnn_oracle as orc  # noqa: E402

torch.set_num_threads(8)


def checksum(*arrays):
    m = hashlib.sha256()
    for a in arrays:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()[:16]


def load_flat(module, flat):
    off = 0
    with torch.no_grad():
        for p in module.parameters():
            n = p.numel()
            p.copy_(torch.from_numpy(flat[off:off + n].copy()).view_as(p))
            off += n
    assert off == flat.size


def flat_of(module):
    return torch.cat([p.contiguous().view(-1) for p in module.parameters()])


# name, layout, B, Dx, E, hidden, Q, gain, x0_zero, param_seed, data_seed, stride for d_flat
CASES = [
    ("cfg1_monotonic", "contig", 100, 1, 2, [64, 64, 64], 50, 1.0, True, 0, 1, 1),
    ("cfg1_monotonic_x0", "contig", 37, 1, 2, [64, 64, 64], 20, 2.0, False, 2, 3, 1),
    ("cfg2_toy", "strided", 64, 2, 10, [100, 100, 100, 100], 50, 1.0, True, 0, 1, 1),
    ("cfg3_power", "strided", 32, 6, 30, [200, 200, 200], 50, 1.0, True, 0, 1, 7),
    ("cfg3_power_trained", "strided", 32, 6, 30, [200, 200, 200], 50, 2.5, False, 4, 5, 7),
    ("cfg4_bsds", "strided", 4, 63, 30, [200, 200, 200], 100, 1.0, True, 0, 1, 7),
    ("cfg5_mnist", "strided", 2, 784, 30, [100, 50, 50, 50, 50], 50, 1.0, True, 0, 1, 1),
    ("small_odd", "strided", 5, 3, 1, [20, 20], 5, 1.5, False, 6, 7, 1),
    ("jit_shape", "strided", 10, 5, 1, [50, 50], 20, 1.0, True, 8, 9, 1),
]


def run_case(name, layout, B, Dx, E, hidden, Q, gain, x0_zero, pseed, dseed, stride):
    if layout == "strided":
        spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]), orc.HIDDEN_LEAKY, orc.OUT_ELU_PLUS_1)
        net = IntegrandNetwork(Dx, 1 + E, hidden, 1)
        Hh = E * Dx
    else:
        spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]), orc.HIDDEN_RELU, orc.OUT_ELU_PLUS_1)
        net = IntegrandNN(1 + E, hidden)
        Hh = E
    flat = orc.synth_params(spec, pseed, gain)
    load_flat(net, flat)
    x0n, xn, hn, gn = orc.synth_inputs(B, Dx, Hh, dseed, x0_zero)
    out = {"meta_layout": layout, "meta_shape": np.array([B, Dx, E, Q]), "meta_hidden": np.array(hidden),
           "meta_gain": np.float64(gain), "meta_seeds": np.array([pseed, dseed]), "meta_x0_zero": np.array(x0_zero),
           "meta_dflat_stride": np.array(stride), "input_checksum": checksum(flat, x0n, xn, hn, gn)}

    for tag, fn in (("par", ParallelNeuralIntegral), ("seq", NeuralIntegral)):
        if tag == "seq" and B * Dx * (Q + 1) > 200000:
            continue
        x0 = torch.from_numpy(x0n.copy()).requires_grad_(True)
        x = torch.from_numpy(xn.copy()).requires_grad_(True)
        h = torch.from_numpy(hn.copy()).requires_grad_(True)
        for p in net.parameters():
            p.grad = None
        z = fn.apply(x0, x, net, flat_of(net), h, Q)
        z.backward(torch.from_numpy(gn.copy()))
        out[f"{tag}_integral"] = z.detach().numpy()
        out[f"{tag}_dx0"] = x0.grad.numpy()
        out[f"{tag}_dx"] = x.grad.numpy()
        out[f"{tag}_dh"] = h.grad.numpy()
        dflat = torch.cat([p.grad.contiguous().view(-1) for p in net.parameters()]).numpy()
        out[f"{tag}_dflat"] = dflat[::stride].copy()
    with torch.no_grad():
        out["f_at_x"] = net(torch.from_numpy(xn.copy()), torch.from_numpy(hn.copy())).numpy()
        out["f_at_x0"] = net(torch.from_numpy(x0n.copy()), torch.from_numpy(hn.copy())).numpy()

    # fp64 run of the same reference code: the "truth" used for error budgeting
    net64 = net.double()
    P = sys.modules["models.UMNN.ParallelNeuralIntegral"]
    with torch.no_grad():
        # float64 weights/nodes: recompute in float64 instead of the float32-rounded tables
        idx = np.arange(Q + 1)
        t64 = torch.from_numpy(np.cos(idx * np.pi / Q)).view(-1, 1)
        lam = np.cos(np.outer(idx, idx) * np.pi / Q)
        lam[:, 0] = .5
        lam[:, -1] = .5 * lam[:, -1]
        lam = lam * 2 / Q
        Wm = np.zeros(Q + 1)
        ev = idx[idx % 2 == 0]
        Wm[ev] = 2 / (1 - ev.astype(np.float64) ** 2)
        Wm[0] = 1
        w64 = torch.from_numpy(lam.T @ Wm).view(-1, 1)
        x0d, xd, hd = (torch.from_numpy(a.astype(np.float64)) for a in (x0n, xn, hn))
        z64 = P.integrate(x0d, Q, (xd - x0d) / Q, net64, hd, False, None, False, w64, t64)
    out["fp64_integral"] = z64.numpy()
    net.float()
    load_flat(net, flat)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: integral[0,:3]={out['par_integral'][0, :3]}  fp32-vs-fp64 max rel "
          f"{np.max(np.abs(out['par_integral'] - out['fp64_integral']) / np.maximum(np.abs(out['fp64_integral']), 1e-6)):.2e}")


def run_cc():
    P = sys.modules["models.UMNN.ParallelNeuralIntegral"]
    out = {}
    for Q in (1, 2, 5, 20, 30, 50, 100, 200):
        w, t = P.compute_cc_weights(Q)
        out[f"w_{Q}"] = w.numpy().reshape(-1)
        out[f"t_{Q}"] = t.numpy().reshape(-1)
    np.savez_compressed(os.path.join(HERE, "cc_weights.npz"), **out)
    print("cc_weights: done")


def run_flow():
    """UMNNMAFFlow.compute_ll (2 blocks, CCParallel) with weights drawn from RandomState."""
    torch.manual_seed(0)
    D, E, Q, B = 6, 8, 20, 16
    hid_int, hid_emb = [50, 50], [64, 64]
    model = UMNNMAFFlow(nb_flow=2, nb_in=D, hidden_derivative=hid_int, hidden_embedding=hid_emb,
                        embedding_s=E, nb_steps=Q, solver="CCParallel")
    rng = np.random.RandomState(11)
    sd = model.state_dict()
    stored = {}
    for k, v in sd.items():
        if k.endswith(".weight") or k.endswith(".bias"):
            fan_in = v.shape[1] if v.dim() == 2 else sd[k.replace(".bias", ".weight")].shape[1]
            arr = rng.uniform(-1, 1, size=tuple(v.shape)).astype(np.float32) * (1.5 / np.sqrt(fan_in))
            sd[k] = torch.from_numpy(arr)
    model.load_state_dict(sd)
    xn = rng.standard_normal((B, D)).astype(np.float32)
    x = torch.from_numpy(xn.copy()).requires_grad_(True)
    ll, z = model.compute_ll(x)
    ll.sum().backward()
    stored["ll"] = ll.detach().numpy()
    stored["z"] = z.detach().numpy()
    stored["dx"] = x.grad.numpy()
    g = {k: p.grad.numpy() for k, p in model.named_parameters() if p.grad is not None}
    stored["grad_keys"] = np.array(sorted(g.keys()))
    for k in g:
        stored["grad/" + k] = g[k]
    stored["state_keys"] = np.array(list(sd.keys()))
    stored["state_shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])
    stored["meta"] = np.array([D, E, Q, B])
    model.eval()
    with torch.no_grad():
        z_eval = model.forward(torch.from_numpy(xn.copy()))
        lj = model.compute_log_jac(torch.from_numpy(xn.copy()))
    stored["z_forward_eval"] = z_eval.numpy()
    stored["log_jac_eval"] = lj.numpy()
    # sampling direction: UMNNMAFFlow.invert on the first 4 latent vectors (5 refinement rounds)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        x_back = model.invert(torch.from_numpy(stored["z"][:4].copy()), iter=5)
    stored["invert_x"] = x_back.numpy()
    stored["invert_x_true"] = xn[:4]
    np.savez_compressed(os.path.join(HERE, "flow_ll.npz"), **stored)
    print("flow_ll: ll[:3] =", stored["ll"][:3])


if __name__ == "__main__":
    run_cc()
    for c in CASES:
        run_case(*c)
    run_flow()
