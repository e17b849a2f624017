"""Pin the numpy oracle (oracle/umnn_oracle.py) to the reference's own outputs.

The golden vectors under tests/golden/ were produced by tests/golden/make_golden.py, which
imports the unmodified reference from /root/reference.  Tolerances are the reference's own
noise floors (parallel-vs-sequential 5e-7, fp32-vs-fp64 2.2e-7 on the integral;
1e-5..1e-4 rel-to-max on parameter / context gradients because of LeakyReLU kink flips).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_DIR, load_golden_case, rel_err, rel_to_max
from oracle import umnn_oracle as orc

INTEGRAL_TOL = 3e-6     # fp32 summation-order noise between two fp32 implementations
GRAD_TOL = 5e-4         # rel-to-max, kink flips included


def test_cc_weights_match_reference_bitwise():
    g = np.load(os.path.join(GOLDEN_DIR, "cc_weights.npz"))
    for Q in (1, 2, 5, 20, 30, 50, 100, 200):
        w, t = orc.cc_nodes_weights(Q)
        assert w.dtype == np.float32 and t.dtype == np.float32
        np.testing.assert_array_equal(t, g[f"t_{Q}"])
        # float64 contraction order may differ between BLAS builds: allow 1 ulp of the largest weight
        assert np.max(np.abs(w - g[f"w_{Q}"])) <= np.spacing(np.float32(np.max(np.abs(w))))
        assert abs(float(w.astype(np.float64).sum()) - 2.0) < 1e-6
        assert t[0] == 1.0 and t[-1] == -1.0


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_parallel(name):
    spec, flat, inp, g = load_golden_case(name)
    z = orc.integrate_parallel(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], inp["layout"])
    assert rel_err(z, g["par_integral"]) < INTEGRAL_TOL
    # and against the fp64 run of the reference
    assert rel_err(z, g["fp64_integral"].astype(np.float32)) < INTEGRAL_TOL


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if n not in ("cfg4_bsds",)])
def test_forward_sequential(name):
    spec, flat, inp, g = load_golden_case(name)
    if "seq_integral" not in g:
        pytest.skip("sequential variant not recorded for this case")
    z = orc.integrate_sequential(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], inp["layout"])
    assert rel_err(z, g["seq_integral"]) < INTEGRAL_TOL


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_point_evaluations(name):
    spec, flat, inp, g = load_golden_case(name)
    fx = orc._integrand(spec, flat, inp["x"], inp["h"], inp["layout"])
    fx0 = orc._integrand(spec, flat, inp["x0"], inp["h"], inp["layout"])
    # ELU+1 is quantised at 2^-24 near zero: allow two quanta absolute
    assert np.all(np.abs(fx - g["f_at_x"]) <= 2e-6 * np.abs(g["f_at_x"]) + 1.2e-7)
    assert np.all(np.abs(fx0 - g["f_at_x0"]) <= 2e-6 * np.abs(g["f_at_x0"]) + 1.2e-7)
    assert np.all(fx >= 0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_backward(name):
    spec, flat, inp, g = load_golden_case(name)
    dx0, dx, dflat, dh = orc.integral_backward(spec, flat, inp["x0"], inp["x"], inp["h"], inp["grad_out"],
                                               inp["Q"], inp["layout"])
    stride = int(g["meta_dflat_stride"])
    assert rel_to_max(dx, g["par_dx"]) < 1e-5
    assert rel_to_max(dx0, g["par_dx0"]) < 1e-5
    assert rel_to_max(dh, g["par_dh"]) < GRAD_TOL
    assert rel_to_max(dflat[::stride], g["par_dflat"]) < GRAD_TOL


def test_backward_fp64_matches_finite_differences():
    """The analytic VJP in float64 against central differences (independent of the reference)."""
    spec = orc.MLPSpec((3, 7, 5, 1), orc.HIDDEN_LEAKY, orc.OUT_ELU_PLUS_1)
    flat = orc.synth_params(spec, 3, 2.0).astype(np.float64)
    x0, x, h, g = (a.astype(np.float64) for a in orc.synth_inputs(4, 3, 6, 5, False))
    Q = 12

    def loss(flat_, x0_, x_, h_):
        return float((orc.integrate_parallel(spec, flat_, x0_, x_, h_, Q) * g).sum())

    dx0, dx, dflat, dh = orc.integral_backward(spec, flat, x0, x, h, g, Q)
    eps = 1e-6
    rng = np.random.RandomState(0)
    for _ in range(10):
        i = rng.randint(flat.size)
        e = np.zeros_like(flat)
        e[i] = eps
        fd = (loss(flat + e, x0, x, h) - loss(flat - e, x0, x, h)) / (2 * eps)
        assert abs(fd - dflat[i]) < 1e-6 * max(1.0, abs(fd))
    for _ in range(10):
        i = tuple(rng.randint(s) for s in h.shape)
        e = np.zeros_like(h)
        e[i] = eps
        fd = (loss(flat, x0, x, h + e) - loss(flat, x0, x, h - e)) / (2 * eps)
        assert abs(fd - dh[i]) < 1e-6 * max(1.0, abs(fd))


def _golden_flow():
    g = dict(np.load(os.path.join(GOLDEN_DIR, "flow_ll.npz"), allow_pickle=False))
    D, E, Q, B = (int(v) for v in g["meta"])
    rng = np.random.RandomState(11)
    state = {}
    keys = [str(k) for k in g["state_keys"]]
    shapes = [eval(str(s)) for s in g["state_shapes"]]
    shp = dict(zip(keys, shapes))
    for k in keys:
        if k.endswith(".weight") or k.endswith(".bias"):
            fan_in = shp[k][1] if len(shp[k]) == 2 else shp[k.replace(".bias", ".weight")][1]
            state[k] = rng.uniform(-1, 1, size=shp[k]).astype(np.float32) * (1.5 / np.sqrt(fan_in))
    x = rng.standard_normal((B, D)).astype(np.float32)
    blocks = []
    for i in range(2):
        made_w = [(state[f"Flow{i}.net.made.net.{j}.weight"], state[f"Flow{i}.net.made.net.{j}.bias"]) for j in (0, 2, 4)]
        hid = [made_w[0][0].shape[0], made_w[1][0].shape[0]]
        flat = np.concatenate([np.concatenate([state[f"Flow{i}.net.parallel_nets.net.{j}.weight"].reshape(-1),
                                               state[f"Flow{i}.net.parallel_nets.net.{j}.bias"].reshape(-1)])
                               for j in (0, 2, 4)])
        spec = orc.MLPSpec((1 + E, 50, 50, 1))
        blocks.append(dict(spec=spec, flat=flat, made=(D, hid, D * E, made_w)))
    return g, blocks, x, Q


def test_flow_compute_ll_matches_reference():
    g, blocks, x, Q = _golden_flow()
    ll, z = orc.flow_compute_ll(blocks, x, Q)
    assert np.max(np.abs(z - g["z"])) < 2e-5 * max(1.0, float(np.max(np.abs(g["z"]))))
    assert np.max(np.abs(ll - g["ll"])) < 2e-5 * max(1.0, float(np.max(np.abs(g["ll"]))))


def test_flow_invert_matches_reference():
    """Sampling direction (UMNNMAFFlow.invert, 5 refinement rounds) against the reference's own output."""
    g, blocks, x, Q = _golden_flow()
    x_back = orc.flow_invert(blocks, g["z"][:4].copy(), Q, n_iter=5)
    assert np.max(np.abs(x_back - g["invert_x"])) < 1e-3
    assert np.max(np.abs(x_back - g["invert_x_true"])) < 5e-2


def test_invert_bracket_step_properties():
    """The bracket update keeps the closest grid point and its neighbours (UMNNMAF.py:218-230)."""
    rng = np.random.RandomState(3)
    G, B = 10, 7
    grid = orc.invert_grid(G)
    left = np.full(B, -50, np.float32)
    right = np.full(B, 50, np.float32)
    x_cur = grid[:, None] * (right - left)[None, :] + left[None, :]
    integ = np.tanh(x_cur / 20).astype(np.float32)             # monotone stand-in for the integral
    target = rng.uniform(-0.7, 0.7, B).astype(np.float32)
    off = np.zeros(B, np.float32)
    l, r, x_next, x_mid = orc.invert_bracket_step(integ, x_cur, grid, off, 1.0, target)
    true_x = 20 * np.arctanh(target)
    assert np.all(l <= true_x + 1e-4) and np.all(true_x <= r + 1e-4)
    assert np.allclose(r - l, 100.0 / 9, rtol=1e-5)
    assert np.allclose(x_next[0], l) and np.allclose(x_next[-1], r, rtol=1e-6)
    assert np.all(np.abs(x_mid - true_x) <= 100.0 / 9)
