"""API-level tests of the reference-shaped surface on CPU (torch route).

Mirrors the reference's own tests (tests/test_numerical_validation.py: analytic known answers,
finite-difference gradients, the y = x^3 fit; tests/test_jit.py: backward smoke, full model,
TorchScript of IntegrandNetwork) and adds parity against the golden vectors recorded from the
reference (tests/golden/).
"""
import contextlib
import io
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import GOLDEN_CASES, GOLDEN_DIR, load_golden_case, rel_err, rel_to_max

import models
from models.UMNN import UMNNMAFFlow, MonotonicNN, IntegrandNN, IntegrandNetwork, UMNNMAF, MADE
from models.UMNN.NeuralIntegral import NeuralIntegral
from models.UMNN.ParallelNeuralIntegral import ParallelNeuralIntegral, integrate, compute_cc_weights
from models.UMNN.UMNNMAF import EmbeddingNetwork


class _Analytic(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, h):
        return self.fn(x)


# ---- reference tests/test_numerical_validation.py:18-97, 319-402 -------------------------------------
@pytest.mark.parametrize("fn,lo,hi,exact", [
    (lambda x: 1 + x ** 2, 0.0, 2.0, 14.0 / 3.0),
    (lambda x: 2 * torch.ones_like(x), 0.0, 3.0, 6.0),
    (lambda x: x, 0.0, 2.0, 2.0),
    (lambda x: x ** 2, 1.0, 3.0, 26.0 / 3.0),
    (lambda x: torch.exp(x), 0.0, 1.0, float(np.e - 1)),
])
@pytest.mark.parametrize("fn_cls", [ParallelNeuralIntegral, NeuralIntegral])
def test_known_integrals(fn, lo, hi, exact, fn_cls):
    integrand = _Analytic(fn)
    x0 = torch.full((1, 1), lo)
    x = torch.full((1, 1), hi)
    h = torch.zeros(1, 1, requires_grad=True)
    for Q, tol in ((20, 1e-3), (100, 1e-4), (200, 1e-4)):
        out = fn_cls.apply(x0, x, integrand, torch.tensor([]), h, Q)
        assert abs(out.item() - exact) < tol * max(1.0, abs(exact))


def test_module_paths_and_aliases():
    assert models.UMNNMAFFlow is UMNNMAFFlow
    import models.UMNN as U
    assert U.ParallelNeuralIntegral is ParallelNeuralIntegral and U.NeuralIntegral is NeuralIntegral
    w, t = compute_cc_weights(20)
    assert w.shape == (21, 1) and t.shape == (21, 1) and w.dtype == torch.float32
    m = UMNNMAFFlow(nb_flow=1, nb_in=3, hidden_derivative=[8], hidden_embedding=[8], embedding_s=2, nb_steps=5)
    for name in ("computell", "forcei_lpschitz", "forceLipshitz", "computeLipshitz"):
        assert callable(getattr(m, name))
    assert callable(m.nets[0].computeLL) and callable(m.nets[0].net.parallel_nets.computeLipshitz)
    with pytest.raises(IndexError):
        m.nets[1]
    with pytest.raises(ValueError):
        m.nets.append("not a module")
    assert UMNNMAF(EmbeddingNetwork(3, [8], [8], 2), 3, 5, solver="nope").forward(torch.zeros(2, 3)) is None


# ---- parity with the recorded reference outputs -------------------------------------------------------
def _build(inp):
    if inp["layout"] == "strided":
        return IntegrandNetwork(inp["Dx"], 1 + inp["E"], inp["hidden"], 1)
    return IntegrandNN(1 + inp["E"], inp["hidden"])


def _load(net, flat):
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p))
            off += p.numel()


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("tag", ["par", "seq"])
def test_autograd_functions_match_reference(name, tag):
    spec, flat, inp, g = load_golden_case(name)
    if f"{tag}_integral" not in g:
        pytest.skip("not recorded")
    net = _build(inp)
    _load(net, flat)
    x0 = torch.from_numpy(inp["x0"].copy()).requires_grad_(True)
    x = torch.from_numpy(inp["x"].copy()).requires_grad_(True)
    h = torch.from_numpy(inp["h"].copy()).requires_grad_(True)
    fn = ParallelNeuralIntegral if tag == "par" else NeuralIntegral
    flat_t = torch.cat([p.contiguous().view(-1) for p in net.parameters()])
    z = fn.apply(x0, x, net, flat_t, h, inp["Q"])
    z.backward(torch.from_numpy(inp["grad_out"].copy()))
    assert rel_err(z.detach().numpy(), g[f"{tag}_integral"]) < 3e-6
    assert rel_to_max(x.grad.numpy(), g[f"{tag}_dx"]) < 1e-5
    assert rel_to_max(x0.grad.numpy(), g[f"{tag}_dx0"]) < 1e-5
    assert rel_to_max(h.grad.numpy(), g[f"{tag}_dh"]) < 5e-4
    dflat = torch.cat([p.grad.contiguous().view(-1) for p in net.parameters()]).numpy()
    assert rel_to_max(dflat[::int(g["meta_dflat_stride"])], g[f"{tag}_dflat"]) < 5e-4


def _flow_from_golden():
    g = dict(np.load(os.path.join(GOLDEN_DIR, "flow_ll.npz"), allow_pickle=False))
    D, E, Q, B = (int(v) for v in g["meta"])
    model = UMNNMAFFlow(nb_flow=2, nb_in=D, hidden_derivative=[50, 50], hidden_embedding=[64, 64],
                        embedding_s=E, nb_steps=Q, solver="CCParallel")
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["state_keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["state_shapes"]]
    rng = np.random.RandomState(11)
    for k, v in sd.items():
        if k.endswith(".weight") or k.endswith(".bias"):
            fan_in = v.shape[1] if v.dim() == 2 else sd[k.replace(".bias", ".weight")].shape[1]
            sd[k] = torch.from_numpy(rng.uniform(-1, 1, size=tuple(v.shape)).astype(np.float32) * (1.5 / np.sqrt(fan_in)))
    model.load_state_dict(sd)
    x = rng.standard_normal((B, D)).astype(np.float32)
    return model, x, g


def test_flow_state_dict_and_compute_ll_match_reference():
    model, xn, g = _flow_from_golden()
    x = torch.from_numpy(xn.copy()).requires_grad_(True)
    ll, z = model.compute_ll(x)
    ll.sum().backward()
    assert np.max(np.abs(ll.detach().numpy() - g["ll"])) < 2e-5 * np.max(np.abs(g["ll"]))
    assert np.max(np.abs(z.detach().numpy() - g["z"])) < 2e-5 * max(1.0, np.max(np.abs(g["z"])))
    assert rel_to_max(x.grad.numpy(), g["dx"]) < 2e-4
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert sorted(grads.keys()) == [str(k) for k in g["grad_keys"]]
    for k, v in grads.items():
        assert rel_to_max(v.numpy(), g["grad/" + k]) < 1e-3, k
    # the drivers' misspelt entry point gives the same numbers
    ll2, _ = model.computell(torch.from_numpy(xn.copy()))
    assert torch.allclose(ll2, ll.detach(), rtol=1e-6, atol=1e-6)


def test_flow_eval_paths_and_invert_match_reference():
    model, xn, g = _flow_from_golden()
    model.eval()
    with torch.no_grad():
        z = model.forward(torch.from_numpy(xn.copy()))
        lj = model.compute_log_jac(torch.from_numpy(xn.copy()))
    assert np.max(np.abs(z.numpy() - g["z_forward_eval"])) < 2e-5 * max(1.0, np.max(np.abs(g["z_forward_eval"])))
    assert np.max(np.abs(lj.numpy() - g["log_jac_eval"])) < 2e-5
    with contextlib.redirect_stdout(io.StringIO()):
        x_back = model.invert(torch.from_numpy(g["z"][:4].copy()), iter=5)
    # bracket refinement takes discrete decisions: identical unless a comparison flips on rounding noise
    assert np.max(np.abs(x_back.numpy() - g["invert_x"])) < 1e-3
    assert np.max(np.abs(x_back.numpy() - g["invert_x_true"])) < 5e-2


def test_set_steps_nb_rebuilds_tables():
    model, xn, _ = _flow_from_golden()
    model.set_steps_nb(30)
    assert model.nets[0].cc_weights.shape == (31, 1) and model.nets[0].nb_steps == 30
    model.eval()
    with torch.no_grad():
        z30 = model.forward(torch.from_numpy(xn.copy()))
    assert torch.isfinite(z30).all()


# ---- reference tests/test_jit.py ----------------------------------------------------------------------
def test_backward_pass_smoke():
    integrand = IntegrandNetwork(nnets=5, nin=2, hidden_sizes=[50, 50], nout=1)
    flat = torch.cat([p.contiguous().view(-1) for p in integrand.parameters()])
    for fn in (NeuralIntegral, ParallelNeuralIntegral):
        x0 = torch.zeros(10, 5, requires_grad=True)
        x = torch.randn(10, 5, requires_grad=True)
        h = torch.randn(10, 5, requires_grad=True)
        out = fn.apply(x0, x, integrand, flat, h, 20)
        out.sum().backward()
        assert out.shape == (10, 5) and x.grad is not None and h.grad is not None and x0.grad is not None


def test_full_model_smoke():
    emb = EmbeddingNetwork(in_d=10, hiddens_embedding=[50, 50], hiddens_integrand=[50, 50], out_made=1, cond_in=0)
    model = UMNNMAF(net=emb, input_size=10, nb_steps=20, solver="CCParallel")
    x = torch.randn(32, 10, requires_grad=True)
    z = model(x)
    z.sum().backward()
    n_with_grad = sum(1 for p in model.parameters() if p.grad is not None)
    assert n_with_grad == len(list(model.parameters())) - 1      # `scaling` is frozen
    ll, _ = model.compute_ll(torch.randn(32, 10))
    assert ll.shape == (32,)


def test_integrand_network_torchscript(tmp_path):
    net = IntegrandNetwork(nnets=5, nin=2, hidden_sizes=[50, 50], nout=1)
    net.eval()
    x, h = torch.randn(10, 5), torch.randn(10, 5)
    want = net(x, h)
    scripted = torch.jit.script(net)
    assert torch.allclose(scripted(x, h), want, rtol=1e-5)
    traced = torch.jit.trace(net, (x, h))
    x2, h2 = torch.randn(20, 5), torch.randn(20, 5)
    assert torch.allclose(traced(x2, h2), net(x2, h2), rtol=1e-5)
    path = os.path.join(tmp_path, "integrand_traced.pt")
    torch.jit.save(traced, path)
    assert torch.allclose(torch.jit.load(path)(x, h), want, rtol=1e-5)


def test_finite_difference_gradient_of_limits():
    torch.manual_seed(0)
    net = IntegrandNetwork(3, 2, [20, 20], 1).double()
    x0 = torch.randn(10, 3, dtype=torch.float64, requires_grad=True)
    x = (x0.detach() + torch.rand(10, 3, dtype=torch.float64)).requires_grad_(True)
    h = torch.randn(10, 3, dtype=torch.float64, requires_grad=True)
    w, t = compute_cc_weights(20)
    out = ParallelNeuralIntegral.apply(x0, x, net, torch.cat([p.view(-1) for p in net.parameters()]), h, 20)
    out.sum().backward()
    eps = 1e-6
    with torch.no_grad():
        base = integrate(x0, 20, (x - x0) / 20, net, h).sum()
        xp = x0.clone()
        xp[0, 0] += eps
        fd = (integrate(xp, 20, (x - xp) / 20, net, h).sum() - base) / eps
    assert abs(fd.item() - x0.grad[0, 0].item()) < 1e-3 * max(1.0, abs(fd.item()))


def test_monotonic_fit_cubic():
    torch.manual_seed(0)
    model = MonotonicNN(2, [32, 32], nb_steps=30)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    x = torch.linspace(0, 2, 64).view(-1, 1)
    hh = torch.zeros(64, 1)
    y = x ** 3
    for _ in range(150):
        opt.zero_grad()
        loss = ((model(x, hh) - y) ** 2).mean()
        loss.backward()
        opt.step()
    assert loss.item() < 0.1
    pred = model(x, hh).detach().view(-1)
    assert torch.all(pred[1:] >= pred[:-1] - 1e-5)      # monotone in x


def test_made_is_autoregressive():
    made = MADE(5, [16, 16], 15, natural_ordering=True)
    x = torch.randn(1, 5, requires_grad=True)
    out = made(x)
    for k in range(15):
        gx, = torch.autograd.grad(out[0, k], x, retain_graph=True)
        assert torch.all(gx[0, (k % 5):] == 0)         # output k%5 sees only inputs < k%5


def test_prepare_integral_needs_cuda_and_a_recognised_integrand():
    """prepare_integral is a kernel-route entry: no CPU fallback, no anonymous callables."""
    import umnn_b200
    net = umnn_b200.IntegrandNN(3, [16, 16])
    with pytest.raises(ValueError):
        umnn_b200.prepare_integral(net, 8, 10)                       # parameters live on the CPU
    with pytest.raises(ValueError):
        umnn_b200.prepare_integral(lambda x, h: x, 8, 10, device="cuda:0")
