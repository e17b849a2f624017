#!/usr/bin/env python
"""Golden case from a flow that was ACTUALLY TRAINED ("re-run with trained weights before freezing the
precision choice"; the other "trained-like" cases are random weights scaled up).

Run ONLY in the build container (reference mounted at /root/reference):

    python tests/golden/make_trained_golden.py

Trains the UNMODIFIED reference UMNNMAFFlow (one block, POWER-shaped: D = 6, E = 30, integrand [200, 200, 200],
MADE [64, 64], Q = 20, CCParallel) for 300 Adam steps on a seeded synthetic density with the reference's own training
loop shape (UCIExperiments.py:129-144: -ll.mean(), clip_grad_value_ 1, Adam 1e-3), then records, for a held-out batch,
what the reference computes with the trained integrand at Q = 50: the conditioner output h (a real MADE output, not
N(0,1)), the integral, f(x), f(x0), the Leibniz gradients, and the float64 integral.  Unlike the seeded cases the
trained weights and the inputs are STORED in the file (they cannot be regenerated from a seed).
"""
import hashlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("UMNN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.append(REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models.UMNN.ParallelNeuralIntegral import ParallelNeuralIntegral  # noqa: E402
from models.UMNN.UMNNMAFFlow import UMNNMAFFlow  # noqa: E402

assert os.path.realpath(sys.modules["models"].__file__).startswith(os.path.realpath(REF)), "wrong `models` imported"
torch.set_num_threads(8)


def data(n, rng):
    """A 6-D density with nonlinear dependencies and unequal scales (what a UCI table looks like after whitening)."""
    z = rng.standard_normal((n, 6)).astype(np.float32)
    x = np.empty_like(z)
    x[:, 0] = z[:, 0]
    x[:, 1] = 0.6 * z[:, 0] ** 2 - 0.6 + 0.5 * z[:, 1]
    x[:, 2] = np.sin(2.0 * x[:, 0]) + 0.3 * z[:, 2]
    x[:, 3] = 0.8 * x[:, 1] * z[:, 3]
    x[:, 4] = np.tanh(x[:, 2] + x[:, 3]) + 0.2 * z[:, 4]
    x[:, 5] = 0.5 * (x[:, 0] + x[:, 4]) + np.where(z[:, 5] > 0, 1.0, -1.0) * 0.7 + 0.15 * z[:, 5]
    return x


def main(steps=300, B=128, Q_train=20, Q_eval=50, B_eval=32):
    torch.manual_seed(0)
    rng = np.random.RandomState(123)
    D, E, hidden = 6, 30, [200, 200, 200]
    model = UMNNMAFFlow(nb_flow=1, nb_in=D, hidden_derivative=hidden, hidden_embedding=[64, 64], embedding_s=E,
                        nb_steps=Q_train, solver="CCParallel")
    opt = torch.optim.Adam(model.parameters(), 1e-3)
    losses = []
    for it in range(steps):
        x = torch.from_numpy(data(B, rng))
        opt.zero_grad()
        ll, _ = model.compute_ll(x)
        loss = -ll.mean()
        loss.backward()
        torch.nn.utils.clip_grad.clip_grad_value_(model.parameters(), 1.)
        opt.step()
        losses.append(float(loss.detach()))
        if it % 50 == 0 or it == steps - 1:
            print(f"step {it}: loss {losses[-1]:.4f}", flush=True)
    assert losses[-1] < losses[0] - 1.0, "the flow did not train"

    blk = model.nets[0]
    net = blk.net.parallel_nets
    flat = torch.cat([p.detach().contiguous().view(-1) for p in net.parameters()]).numpy().astype(np.float32)
    xn = data(B_eval, rng)
    with torch.no_grad():
        hn = blk.net.make_embeding(torch.from_numpy(xn)).numpy().astype(np.float32)      # the trained conditioner's output
    x0n = (0.25 * rng.standard_normal((B_eval, D))).astype(np.float32)
    gn = rng.standard_normal((B_eval, D)).astype(np.float32)

    out = {"meta_layout": "strided", "meta_shape": np.array([B_eval, D, E, Q_eval]), "meta_hidden": np.array(hidden),
           "meta_gain": np.float64(1.0), "meta_seeds": np.array([-1, -1]), "meta_x0_zero": np.array(False),
           "meta_dflat_stride": np.array(7), "stored_flat": flat, "stored_x0": x0n, "stored_x": xn, "stored_h": hn,
           "stored_grad_out": gn, "train_losses": np.array(losses, np.float32)}
    m = hashlib.sha256()
    for a in (flat, x0n, xn, hn, gn):
        m.update(np.ascontiguousarray(a).tobytes())
    out["input_checksum"] = m.hexdigest()[:16]

    x0 = torch.from_numpy(x0n.copy()).requires_grad_(True)
    x = torch.from_numpy(xn.copy()).requires_grad_(True)
    h = torch.from_numpy(hn.copy()).requires_grad_(True)
    for p in net.parameters():
        p.grad = None
    flat_t = torch.cat([p.contiguous().view(-1) for p in net.parameters()])
    z = ParallelNeuralIntegral.apply(x0, x, net, flat_t, h, Q_eval)
    z.backward(torch.from_numpy(gn.copy()))
    out["par_integral"] = z.detach().numpy()
    out["par_dx0"], out["par_dx"], out["par_dh"] = x0.grad.numpy(), x.grad.numpy(), h.grad.numpy()
    out["par_dflat"] = torch.cat([p.grad.contiguous().view(-1) for p in net.parameters()]).numpy()[::7].copy()
    with torch.no_grad():
        out["f_at_x"] = net(torch.from_numpy(xn.copy()), torch.from_numpy(hn.copy())).numpy()
        out["f_at_x0"] = net(torch.from_numpy(x0n.copy()), torch.from_numpy(hn.copy())).numpy()
    # float64 run of the same reference code with float64 tables
    P = sys.modules["models.UMNN.ParallelNeuralIntegral"]
    net64 = net.double()
    idx = np.arange(Q_eval + 1)
    t64 = torch.from_numpy(np.cos(idx * np.pi / Q_eval)).view(-1, 1)
    lam = np.cos(np.outer(idx, idx) * np.pi / Q_eval)
    lam[:, 0] = .5
    lam[:, -1] = .5 * lam[:, -1]
    lam = lam * 2 / Q_eval
    Wm = np.zeros(Q_eval + 1)
    ev = idx[idx % 2 == 0]
    Wm[ev] = 2 / (1 - ev.astype(np.float64) ** 2)
    Wm[0] = 1
    w64 = torch.from_numpy(lam.T @ Wm).view(-1, 1)
    with torch.no_grad():
        x0d, xd, hd = (torch.from_numpy(a.astype(np.float64)) for a in (x0n, xn, hn))
        out["fp64_integral"] = P.integrate(x0d, Q_eval, (xd - x0d) / Q_eval, net64, hd, False, None, False, w64, t64).numpy()
    np.savez_compressed(os.path.join(HERE, "cfg3_power_adam.npz"), **out)
    rel = np.max(np.abs(out["par_integral"] - out["fp64_integral"]) / np.maximum(np.abs(out["fp64_integral"]), 1e-6))
    print(f"cfg3_power_adam: loss {losses[0]:.3f} -> {losses[-1]:.3f}; |W| max {np.abs(flat).max():.3f}; "
          f"fp32-vs-fp64 integral max rel {rel:.2e}; f(x) range [{out['f_at_x'].min():.3e}, {out['f_at_x'].max():.3e}]")


if __name__ == "__main__":
    main()
