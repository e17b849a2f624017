"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the golden
vectors recorded from the reference.  Run on the B200 box with `pytest -m gpu`.

Tolerances: integral max rel-err <= 1e-5 (north star) for EVERY precision the product selects, on every network
(the FP32 kernel is expected near 5e-7, the noise floor between two fp32 summation orders; FP16x3 measures
3e-7..2.5e-6); point evaluations within 1e-5 relative + two ELU+1 quanta (2^-23) absolute.  Gradients use the
flip-aware metric of conftest.py: against the float64 oracle, every slot whose hidden pre-activations stay clear of
their kinks must agree to 1e-4 of the largest entry, slots that deviate must be explained by a near-kink unit and are
counted, and everything together is held to 2e-3 normwise (d_params additionally to 5e-3 of its largest entry).
"""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from conftest import (GOLDEN_CASES, GOLDEN_DIR, assert_slot_grad_ok, assert_sum_grad_ok, dh_to_slots, load_golden_case,
                      norm_err, rel_err, rel_to_max)
from oracle import c_binding, umnn_oracle as orc

pytestmark = pytest.mark.gpu

INTEGRAL_TOL = 1e-5

# "auto" resolves to the FP16x3 tensor-core split (fp16 hi + lo operands, 22 bits, guarded FP32 re-run) wherever the
# tensor-core kernel serves the shape, else to the FP32 kernel.  Explicit "fp16x3" / "bf16x3" raise on shapes the
# tensor-core kernel cannot serve; "bf16x3" is a diagnostic precision (include/umnn_b200.h) that AUTO never selects.
PRECISIONS = ["fp32", "auto"]


def _prec(name):
    from umnn_b200 import _native
    return {"fp32": _native.PREC_FP32, "bf16x3": _native.PREC_BF16X3, "auto": _native.PREC_AUTO,
            "fp16x3": _native.PREC_FP16X3}[name]


def _dev():
    return torch.device("cuda:0")


def _point_ok(got, want, rel=1e-5):
    return np.all(np.abs(got - want) <= rel * np.abs(want) + 2.4e-7)


def _net_for(spec, flat, layout, Dx):
    from umnn_b200 import IntegrandNN, IntegrandNetwork
    hidden = list(spec.widths[1:-1])
    if layout == "strided":
        net = IntegrandNetwork(Dx, spec.widths[0], hidden, 1,
                               act_func="ELU" if spec.out_act == orc.OUT_ELU_PLUS_1 else "Sigmoid")
    else:
        net = IntegrandNN(spec.widths[0], hidden)
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p))
            off += p.numel()
    return net.to(_dev())


def _run_kernel(spec, flat, x0, x, h, Q, layout, want_f=True, x0_none=False, precision="auto"):
    from umnn_b200 import cc_integrate
    net = _net_for(spec, flat, layout, x.shape[1])
    d = _dev()
    out, fx, fx0 = cc_integrate(net, None if x0_none else torch.from_numpy(x0).to(d), torch.from_numpy(x).to(d),
                                torch.from_numpy(h).to(d), Q, want_fx=want_f, want_fx0=want_f, precision=_prec(precision))
    torch.cuda.synchronize()
    return out.cpu().numpy(), None if fx is None else fx.cpu().numpy(), None if fx0 is None else fx0.cpu().numpy()


def test_native_library_is_the_loaded_path():
    from umnn_b200 import _native
    lib = _native.lib()
    assert lib.umnn_abi_version() == 1
    loaded = open("/proc/self/maps").read()
    assert "libumnn_b200.so" in loaded


@pytest.mark.parametrize("hidden,narrow", [([100, 100, 100, 100], 1), ([100, 50, 50, 50, 50], 1), ([126, 126], 1),
                                           ([127, 127], 0), ([200, 200, 200], 0)])
def test_narrow_shape_selection(hidden, narrow):
    """Integrands whose padded widths fit 128 tensor-memory columns get the narrow kernel shape (two CTAs per SM)."""
    import ctypes
    from umnn_b200 import _native
    desc = _native.make_desc(_native.LAYOUT_STRIDED_D, 1000, 6, 30, [31] + hidden + [1], _native.ACT_LEAKY_RELU,
                             _native.OUT_ELU_PLUS_1, 50, _native.PREC_AUTO)
    is_narrow, n = ctypes.c_int32(-1), ctypes.c_int32(-1)
    _native.check(_native.lib().umnn_tc_forward_occupancy(desc, 1, ctypes.byref(is_narrow), ctypes.byref(n)))
    assert is_narrow.value == narrow and n.value >= 1


def test_narrow_shape_really_overlaps_two_tiles_per_sm():
    """The occupancy calculators (CUDA runtime, ncu) report one cluster per SM pair for the narrow shape, yet the
    hardware co-schedules two (ncu: 23.6 of 64 warps active = 2 x 12): the evidence that counts is the speed-up over
    the same kernel with one CTA per SM (measured 1.40x on config 5; an 11-warp variant that did NOT fit twice ran at
    0.70x).  Both shapes give the same integral up to summation order."""
    from umnn_b200 import cc_integrate
    spec = orc.MLPSpec((31, 100, 50, 50, 50, 50, 1))
    flat = orc.synth_params(spec, 0, 1.0)
    net = _net_for(spec, flat, "strided", 784).eval()
    d = _dev()
    g = torch.Generator(device=d).manual_seed(4)
    x = 2 * torch.randn(100, 784, device=d, generator=g)
    h = torch.randn(100, 30 * 784, device=d, generator=g)
    res = {}
    prev = os.environ.get("UMNN_B200_TC_NARROW")
    try:
        for narrow in ("1", "0"):
            os.environ["UMNN_B200_TC_NARROW"] = narrow
            for _ in range(3):
                out = cc_integrate(net, None, x, h, 50, want_fx=True)[0]
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                cc_integrate(net, None, x, h, 50, want_fx=True)
            e.record()
            torch.cuda.synchronize()
            res[narrow] = (s.elapsed_time(e) / 20, out.cpu().numpy())
    finally:
        if prev is None:
            os.environ.pop("UMNN_B200_TC_NARROW", None)
        else:
            os.environ["UMNN_B200_TC_NARROW"] = prev
    assert rel_err(res["1"][1], res["0"][1]) < 2e-6
    assert res["0"][0] / res["1"][0] > 1.08, (res["1"][0], res["0"][0])       # 0.7 .. 1.0 if only one CTA fits per SM


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_matches_golden_and_oracle(name, precision):
    spec, flat, inp, g = load_golden_case(name)
    tol = INTEGRAL_TOL
    out, fx, fx0 = _run_kernel(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], inp["layout"], precision=precision)
    assert rel_err(out, g["par_integral"]) < tol
    assert rel_err(out, g["fp64_integral"].astype(np.float32)) < tol
    ref = orc.integrate_parallel(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], inp["layout"])
    assert rel_err(out, ref) < tol
    assert _point_ok(fx, g["f_at_x"], tol) and _point_ok(fx0, g["f_at_x0"], tol)
    # integral only (no extra rows): same integral up to summation-order noise
    out2, _, _ = _run_kernel(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], inp["layout"], want_f=False,
                             precision=precision)
    assert rel_err(out2, out) < 2e-6


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("fn_name", ["ParallelNeuralIntegral", "NeuralIntegral"])
def test_autograd_function_on_cuda_matches_golden(name, fn_name):
    import umnn_b200
    spec, flat, inp, g = load_golden_case(name)
    fn = getattr(umnn_b200, fn_name)
    net = _net_for(spec, flat, inp["layout"], inp["Dx"])
    d = _dev()
    x0 = torch.from_numpy(inp["x0"]).to(d).requires_grad_(True)
    x = torch.from_numpy(inp["x"]).to(d).requires_grad_(True)
    h = torch.from_numpy(inp["h"]).to(d).requires_grad_(True)
    z = fn.apply(x0, x, net, torch.cat([p.view(-1) for p in net.parameters()]), h, inp["Q"])
    z.backward(torch.from_numpy(inp["grad_out"]).to(d))
    tol = INTEGRAL_TOL
    assert rel_err(z.detach().cpu().numpy(), g["par_integral"]) < tol
    assert rel_to_max(x.grad.cpu().numpy(), g["par_dx"]) < tol
    assert rel_to_max(x0.grad.cpu().numpy(), g["par_dx0"]) < tol
    # gradients against the float64 oracle with the flip-aware metric, and the reference's own fp32 vectors within
    # the plain bounds (the reference flips kinks against float64 too)
    B, Dx, layout = inp["B"], inp["Dx"], inp["layout"]
    margins = orc.kink_margins(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], layout)
    _, _, r_flat, r_h = _oracle_backward_with_jac(spec, flat, inp, None)
    assert_slot_grad_ok(dh_to_slots(h.grad.cpu().numpy(), B, Dx, layout), dh_to_slots(r_h, B, Dx, layout), margins, "auto",
                        (B, Dx, -1), what=f"{name} d_h")
    dflat = torch.cat([p.grad.view(-1) for p in net.parameters()]).cpu().numpy()
    assert_sum_grad_ok(dflat, r_flat, what=f"{name} d_params")
    assert_sum_grad_ok(dflat[::int(g["meta_dflat_stride"])], g["par_dflat"], what=f"{name} d_params vs reference")
    assert norm_err(h.grad.cpu().numpy(), g["par_dh"]) < 2e-3


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("Q", [1, 2, 3, 7, 31, 127, 128, 200, 1024])
def test_node_count_edges(Q, precision):
    """Q+1(+2) rows per slot below, at and above the 128-row tile; slots straddle tiles."""
    spec = orc.MLPSpec((4, 24, 16, 1))
    flat = orc.synth_params(spec, Q, 2.0)
    x0, x, h, _ = orc.synth_inputs(13, 5, 15, Q + 1, x0_zero=False)
    out, fx, fx0 = _run_kernel(spec, flat, x0, x, h, Q, "strided", precision=precision)
    ref, rfx, rfx0 = c_binding.cc_forward(spec, flat, x0, x, h, Q)
    tol = INTEGRAL_TOL
    assert rel_err(out, ref) < tol
    assert _point_ok(fx, rfx, tol) and _point_ok(fx0, rfx0, tol)


@pytest.mark.parametrize("hidden", [[8], [20, 20], [50, 50, 50, 50], [64, 64, 64], [100, 50, 50, 50, 50],
                                    [256, 256], [17, 33, 65], [12] * 7])
@pytest.mark.parametrize("acts", [(orc.HIDDEN_LEAKY, orc.OUT_ELU_PLUS_1), (orc.HIDDEN_LEAKY, orc.OUT_SIGMOID)])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_network_shapes_strided(hidden, acts, precision):
    E, D, B, Q = 6, 7, 19, 20
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]), acts[0], acts[1])
    flat = orc.synth_params(spec, len(hidden), 1.7)
    x0, x, h, _ = orc.synth_inputs(B, D, E * D, 3, x0_zero=False)
    out, fx, fx0 = _run_kernel(spec, flat, x0, x, h, Q, "strided", precision=precision)
    ref, rfx, rfx0 = c_binding.cc_forward(spec, flat, x0, x, h, Q)
    tol = INTEGRAL_TOL
    assert rel_err(out, ref) < tol
    assert _point_ok(fx, rfx, tol) and _point_ok(fx0, rfx0, tol)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("E", [0, 1, 2, 30, 255])
def test_contiguous_layout_and_context_sizes(E, precision):
    spec = orc.MLPSpec((1 + E, 32, 32, 1), orc.HIDDEN_RELU, orc.OUT_ELU_PLUS_1)
    flat = orc.synth_params(spec, E, 1.5)
    x0, x, h, _ = orc.synth_inputs(41, 1, E, 5, x0_zero=False)
    out, fx, fx0 = _run_kernel(spec, flat, x0, x, h, 25, "contig", precision=precision)
    ref, rfx, rfx0 = c_binding.cc_forward(spec, flat, x0, x, h, 25, layout="contig")
    tol = INTEGRAL_TOL
    assert rel_err(out, ref) < tol
    assert _point_ok(fx, rfx, tol) and _point_ok(fx0, rfx0, tol)


def test_batch_edges_and_null_x0():
    spec = orc.MLPSpec((3, 16, 1))
    flat = orc.synth_params(spec, 0, 1.0)
    for B in (0, 1, 2, 129, 1000):
        x0, x, h, _ = orc.synth_inputs(B, 2, 4, B, x0_zero=True)
        out, fx, _ = _run_kernel(spec, flat, x0, x, h, 10, "strided", x0_none=True)
        assert out.shape == (B, 2)
        if B:
            ref, rfx, _ = c_binding.cc_forward(spec, flat, x0, x, h, 10)
            assert rel_err(out, ref) < INTEGRAL_TOL and _point_ok(fx, rfx)


def test_degenerate_limits_and_antisymmetry():
    spec = orc.MLPSpec((31, 200, 200, 200, 1))
    flat = orc.synth_params(spec, 0, 2.0)
    x0, x, h, _ = orc.synth_inputs(64, 6, 180, 2, x0_zero=False)
    same, _, _ = _run_kernel(spec, flat, x, x, h, 50, "strided", want_f=False)
    assert np.all(same == 0.0)
    fwd, _, _ = _run_kernel(spec, flat, x0, x, h, 50, "strided", want_f=False)
    bwd, _, _ = _run_kernel(spec, flat, x, x0, h, 50, "strided", want_f=False)
    assert rel_err(-bwd, fwd, floor=1e-4) < 1e-5
    fwd, _, _ = _run_kernel(spec, flat, x0, x, h, 50, "strided", want_f=False, precision="fp32")
    bwd, _, _ = _run_kernel(spec, flat, x, x0, h, 50, "strided", want_f=False, precision="fp32")
    assert rel_err(-bwd, fwd, floor=1e-4) < 1e-5


@pytest.mark.parametrize("precision", PRECISIONS)
def test_deterministic_and_batch_invariant(precision):
    spec = orc.MLPSpec((31, 200, 200, 200, 1))
    flat = orc.synth_params(spec, 0, 1.0)
    x0, x, h, _ = orc.synth_inputs(300, 6, 180, 2, x0_zero=False)
    a, _, _ = _run_kernel(spec, flat, x0, x, h, 50, "strided", precision=precision)
    b, _, _ = _run_kernel(spec, flat, x0, x, h, 50, "strided", precision=precision)
    np.testing.assert_array_equal(a, b)
    c, _, _ = _run_kernel(spec, flat, x0[100:200], x[100:200], h[100:200], 50, "strided", precision=precision)
    assert rel_err(c, a[100:200]) < 2e-6


@pytest.mark.parametrize("layout", ["strided", "contig"])
def test_fp16x3_overflow_is_caught_by_the_guarded_rerun(layout):
    """Activations beyond the fp16 range (|a| > 65504) would become inf in the fp16 operands; the kernel tracks the
    largest value it converts (ReLU) or sees the NaN that the inf/-inf pair produces downstream (LeakyReLU) and raises
    a device flag, and the second launch -- the FP32 kernel, a no-op otherwise -- recomputes the call: the result is
    bit-identical to UMNN_PREC_FP32, i.e. the rare overflow case gets the parity anchor's arithmetic and meets the
    same 1e-5 bar as everything else.  The ReLU network (contig) would swallow a NaN in max(v, 0), which is why the
    flag does not rely on NaN propagation there."""
    if layout == "strided":
        spec = orc.MLPSpec((31, 200, 200, 200, 1))
        x0, x, h, _ = orc.synth_inputs(64, 6, 180, 2, x0_zero=True)
    else:
        spec = orc.MLPSpec((3, 64, 64, 64, 1), orc.HIDDEN_RELU, orc.OUT_ELU_PLUS_1)
        x0, x, h, _ = orc.synth_inputs(300, 1, 2, 2, x0_zero=True)
    flat = orc.synth_params(spec, 0, 1.0)
    big = (x * 3.0e6).astype(np.float32)
    for prec in ("fp16x3", "auto"):
        a, fa, _ = _run_kernel(spec, flat, x0, big, h, 50, layout, precision=prec)
        b, fb, _ = _run_kernel(spec, flat, x0, big, h, 50, layout, precision="fp32")
        assert np.all(np.isfinite(a)) and np.all(np.isfinite(fa))
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(fa, fb)
    ref, _, _ = c_binding.cc_forward(spec, flat, x0, big, h, 50, layout)
    assert rel_err(a, ref) < INTEGRAL_TOL
    # in range: the re-run stays a no-op and the fp16 result stands (it differs from fp32 in the last bits)
    c, _, _ = _run_kernel(spec, flat, x0, x, h, 50, layout, precision="fp16x3")
    d, _, _ = _run_kernel(spec, flat, x0, x, h, 50, layout, precision="fp32")
    assert not np.array_equal(c, d)
    ref, _, _ = c_binding.cc_forward(spec, flat, x0, x, h, 50, layout)
    assert rel_err(c, ref) < 2e-6


def test_fp16x3_backward_overflow_reruns_in_fp32():
    """Same guard in the backward: an overflowing re-evaluation makes the second sequence of launches -- the FP32
    backward, gated on the device flag -- recompute every gradient; in range they stay no-ops.  The per-slot outputs
    then equal the FP32 backward's up to the order in which a slot that straddles two tiles is summed; d_params is a
    sum over differently sized chunks (the re-run borrows the tensor-core workspace), so it agrees to rounding."""
    from umnn_b200 import kernel
    spec = orc.MLPSpec((31, 200, 200, 200, 1))
    flat = orc.synth_params(spec, 0, 1.0)
    x0, x, h, g = orc.synth_inputs(3000, 6, 180, 2, x0_zero=True)      # two chunks of the scratch
    net = _net_for(spec, flat, "strided", 6)
    ks = net.kernel_spec()
    d = _dev()
    for scale, overflow in ((3.0e6, True), (1.0, False)):
        xb = x.copy()
        xb[-100:] *= scale                                             # only the second chunk leaves the fp16 range
        xs = torch.from_numpy(xb.astype(np.float32)).to(d)
        args = (ks, None, xs, torch.from_numpy(h).to(d), torch.from_numpy(g).to(d), 50)
        a = kernel.cc_backward(*args, precision=_prec("fp16x3"))
        b = kernel.cc_backward(*args, precision=_prec("fp32"))
        torch.cuda.synchronize()
        for ta in a[1:]:
            assert torch.isfinite(ta).all()
        if overflow:
            assert torch.equal(a[1], b[1])                                                    # d_x = f(x) * g per slot
            assert norm_err(a[3].cpu().numpy(), b[3].cpu().numpy()) < 1e-6                    # d_h
            assert norm_err(a[2].cpu().numpy(), b[2].cpu().numpy()) < 1e-5                    # d_params
        else:
            assert not torch.equal(a[1], b[1])                                                # the fp16 result stands


def test_bf16x3_is_a_diagnostic_precision_below_the_bar():
    """UMNN_PREC_BF16X3 (never selected by AUTO) is kept for A/B measurements of the operand formats: it meets the
    1e-5 bar on default-initialised networks but its ~17-bit operands measure 1e-5..2.2e-5 on trained-scale weights,
    where FP16X3 holds 3e-6.  This test documents both facts; nothing in the product depends on BF16X3."""
    spec = orc.MLPSpec((31, 200, 200, 200, 1))
    x0, x, h, _ = orc.synth_inputs(64, 6, 180, 2, x0_zero=False)
    for gain, bf16_bound in ((1.0, INTEGRAL_TOL), (2.5, 1e-4)):
        flat = orc.synth_params(spec, 0, gain)
        ref, _, _ = c_binding.cc_forward(spec, flat, x0, x, h, 50)
        b, _, _ = _run_kernel(spec, flat, x0, x, h, 50, "strided", precision="bf16x3")
        a, _, _ = _run_kernel(spec, flat, x0, x, h, 50, "strided", precision="fp16x3")
        assert rel_err(b, ref) < bf16_bound
        assert rel_err(a, ref) < INTEGRAL_TOL
        if gain > 1.0:
            assert rel_err(a, ref) < rel_err(b, ref)


def test_packed_parameter_cache_tracks_updates():
    """eval-mode networks reuse their packed parameters until a parameter changes in place."""
    from umnn_b200 import cc_integrate
    spec = orc.MLPSpec((31, 200, 200, 200, 1))
    flat = orc.synth_params(spec, 0, 1.0)
    x0, x, h, _ = orc.synth_inputs(16, 6, 180, 2, x0_zero=True)
    net = _net_for(spec, flat, "strided", 6).eval()
    d = _dev()
    xd, hd = torch.from_numpy(x).to(d), torch.from_numpy(h).to(d)
    a, _, _ = cc_integrate(net, None, xd, hd, 50)
    a2, _, _ = cc_integrate(net, None, xd, hd, 50)
    assert torch.equal(a, a2) and "_umnn_packed" in net.net[0].__dict__
    with torch.no_grad():
        net.net[6].weight.mul_(1.25)
    b, _, _ = cc_integrate(net, None, xd, hd, 50)
    flat2 = flat.copy()
    flat2[-201:-1] *= 1.25
    ref = orc.integrate_parallel(spec, flat2, x0, x, h, 50)
    assert rel_err(b.cpu().numpy(), ref) < INTEGRAL_TOL
    assert not torch.equal(a, b)


def test_host_buffer_entry():
    from umnn_b200 import _native, kernel
    spec = orc.MLPSpec((11, 100, 100, 100, 100, 1))
    flat = orc.synth_params(spec, 0, 1.0)
    x0, x, h, _ = orc.synth_inputs(50, 2, 20, 1, x0_zero=False)
    out, fx, fx0 = kernel.cc_forward_host(spec.widths, _native.LAYOUT_STRIDED_D, _native.ACT_LEAKY_RELU,
                                          _native.OUT_ELU_PLUS_1, flat, x0, x, h, 50, True, True)
    ref, rfx, rfx0 = c_binding.cc_forward(spec, flat, x0, x, h, 50)
    assert rel_err(out, ref) < INTEGRAL_TOL and _point_ok(fx, rfx) and _point_ok(fx0, rfx0)


def test_native_errors_surface_as_exceptions():
    from umnn_b200 import IntegrandNetwork, cc_integrate
    net = IntegrandNetwork(3, 3, [16], 1).to(_dev())
    x = torch.randn(4, 3, device=_dev())
    with pytest.raises(ValueError):
        cc_integrate(net, None, x, torch.randn(4, 5, device=_dev()), 10)           # wrong h width
    with pytest.raises(ValueError):
        cc_integrate(net, None, x.cpu(), torch.randn(4, 6), 10)                    # CPU tensors
    from umnn_b200 import _native
    with pytest.raises(_native.NativeError):
        cc_integrate(net, None, x, torch.randn(4, 6, device=_dev()), 5000)         # Q out of range


# ---- fused backward (umnn_cc_backward) ---------------------------------------------------------------
def _oracle_backward_with_jac(spec, flat, inp, grad_fx):
    """Oracle gradients of  sum(integral * g) + sum(f(x,h) * grad_fx)  (float64 arithmetic)."""
    f64 = lambda a: a.astype(np.float64)
    x0, x, h, g = f64(inp["x0"]), f64(inp["x"]), f64(inp["h"]), f64(inp["grad_out"])
    flat64 = f64(flat)
    d_x0, d_x, d_flat, d_h = orc.integral_backward(spec, flat64, x0, x, h, g, inp["Q"], inp["layout"])
    if grad_fx is not None:
        if inp["layout"] == "strided":
            rows = orc.slot_inputs_strided(x, h)
        else:
            rows = np.concatenate([x, h], axis=1)
        gp, d_rows = orc.mlp_rows_vjp(spec, flat64, rows, f64(grad_fx).reshape(-1))
        d_flat = d_flat + gp
        B, Dx = x.shape
        E = spec.widths[0] - 1
        d_x = d_x + d_rows[:, 0].reshape(B, Dx)
        if inp["layout"] == "strided":
            d_h = d_h + d_rows[:, 1:].reshape(B, Dx, E).transpose(0, 2, 1).reshape(B, E * Dx)
        else:
            d_h = d_h + d_rows[:, 1:]
    return d_x0, d_x, d_flat, d_h


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("with_jac", [False, True])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_native_backward_matches_oracle(name, with_jac, precision):
    from umnn_b200 import kernel, _native
    spec, flat, inp, g = load_golden_case(name)
    net = _net_for(spec, flat, inp["layout"], inp["Dx"])
    kspec = net.kernel_spec()
    d = _dev()
    grad_fx = np.random.RandomState(5).standard_normal(inp["x"].shape).astype(np.float32) if with_jac else None
    x0, x, h, go = (torch.from_numpy(inp[k]).to(d) for k in ("x0", "x", "h", "grad_out"))
    prec = _prec(precision)
    if _native.lib().umnn_workspace_bytes(kernel.make_desc(kspec, x, inp["Q"], prec), 1) == 0:
        pytest.skip("this backward does not serve the shape")
    tgfx = None if grad_fx is None else torch.from_numpy(grad_fx).to(d)
    d_x0, d_x, d_flat, d_h = kernel.cc_backward(kspec, x0, x, h, go, inp["Q"], grad_fx=tgfx, precision=prec)
    torch.cuda.synchronize()
    r_x0, r_x, r_flat, r_h = _oracle_backward_with_jac(spec, flat, inp, grad_fx)
    B, Dx, layout = inp["B"], inp["Dx"], inp["layout"]
    margins = orc.kink_margins(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], layout)
    assert rel_to_max(d_x0.cpu().numpy(), r_x0) < INTEGRAL_TOL
    if with_jac:     # d_x carries grad_fx * df/dx(x, h): one more per-slot quantity that a kink flip at x changes
        assert_slot_grad_ok(d_x.cpu().numpy(), r_x, margins, precision, (B, Dx, 1), what=f"{name} d_x")
    else:
        assert rel_to_max(d_x.cpu().numpy(), r_x) < INTEGRAL_TOL
    assert_slot_grad_ok(dh_to_slots(d_h.cpu().numpy(), B, Dx, layout), dh_to_slots(r_h, B, Dx, layout), margins, precision,
                        (B, Dx, -1), what=f"{name} d_h")
    assert_sum_grad_ok(d_flat.cpu().numpy(), r_flat, what=f"{name} d_params")
    if not with_jac:
        stride = int(g["meta_dflat_stride"])
        assert_sum_grad_ok(d_flat.cpu().numpy()[::stride], g["par_dflat"], what=f"{name} d_params vs reference")
        assert norm_err(d_h.cpu().numpy(), g["par_dh"]) < 2e-3
    # deterministic, and partial outputs may be skipped
    again = kernel.cc_backward(kspec, x0, x, h, go, inp["Q"], grad_fx=tgfx, precision=prec)
    assert torch.equal(again[2], d_flat) and torch.equal(again[3], d_h)
    only_h = kernel.cc_backward(kspec, x0, x, h, go, inp["Q"], need_x0=False, need_x=False, need_params=False,
                                precision=prec)
    assert only_h[0] is None and only_h[2] is None
    assert torch.equal(only_h[3], kernel.cc_backward(kspec, x0, x, h, go, inp["Q"], precision=prec)[3])


@pytest.mark.parametrize("B,D,E,hidden,Q,layout", [
    (3000, 6, 30, [200, 200, 200], 50, "strided"),      # several chunks of the scratch, straddling slots
    (257, 3, 4, [24, 16], 200, "strided"),              # slots longer than three tiles
    (1, 1, 1, [8], 1, "contig"),                        # tiny everything
    (50, 1, 255, [256, 256], 7, "contig"),              # widest input / hidden layers
    (700, 3, 40, [64, 48], 20, "strided"),              # 32 < 1 + E <= 64: the input gradient spans two 32-column pairs (two
                                                        # column groups park it for the d_h reduction), several tiles per CTA
])
@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_native_backward_shapes(B, D, E, hidden, Q, layout, precision):
    from umnn_b200 import kernel, _native
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]), orc.HIDDEN_LEAKY if layout == "strided" else orc.HIDDEN_RELU)
    flat = orc.synth_params(spec, 3, 1.5)
    Hh = E * D if layout == "strided" else E
    x0, x, h, go = orc.synth_inputs(B, D, Hh, 4, x0_zero=False)
    net = _net_for(spec, flat, layout, D)
    kspec = net.kernel_spec()
    d = _dev()
    xd = torch.from_numpy(x).to(d)
    prec = _prec(precision)
    if _native.lib().umnn_workspace_bytes(kernel.make_desc(kspec, xd, Q, prec), 1) == 0:
        pytest.skip("this backward does not serve the shape")
    got = kernel.cc_backward(kspec, torch.from_numpy(x0).to(d), xd, torch.from_numpy(h).to(d), torch.from_numpy(go).to(d), Q,
                             precision=prec)
    n_chk = min(B, 40)
    inp = dict(x0=x0[:n_chk], x=x[:n_chk], h=h[:n_chk], grad_out=go[:n_chk], Q=Q, layout=layout)
    r_x0, r_x, _, r_h = _oracle_backward_with_jac(spec, flat, inp, None)
    tol = INTEGRAL_TOL
    assert rel_to_max(got[0][:n_chk].cpu().numpy(), r_x0) < tol
    assert rel_to_max(got[1][:n_chk].cpu().numpy(), r_x) < tol
    margins = orc.kink_margins(spec, flat, x0[:n_chk], x[:n_chk], h[:n_chk], Q, layout)
    assert_slot_grad_ok(dh_to_slots(got[3][:n_chk].cpu().numpy(), n_chk, D, layout), dh_to_slots(r_h, n_chk, D, layout), margins,
                        precision, (n_chk, D, -1), what="d_h")
    # parameter gradient: the same batch through the torch route on the device
    from umnn_b200.integral import _integrate_grads_chunked
    ref_flat, _ = _integrate_grads_chunked(torch.from_numpy(x0).to(d), xd, net, torch.from_numpy(h).to(d), Q,
                                           torch.from_numpy(go).to(d), False)
    assert_sum_grad_ok(got[2].cpu().numpy(), ref_flat.detach().cpu().numpy(), what="d_params vs torch ops")


def test_weight_gradient_with_coherent_cotangents():
    """A likelihood's cotangents are coherent: the Jacobian rows carry the same value (-1/(B f), f nearly constant at
    initialisation), so the rank-1 head of the dgrad chain, dz_J = dv * w_out (.) act', holds the SAME number in every such
    row and a hi-only bf16 panel rounds them all the same way -- the error does not average out over the rows (9e-3 of the last
    hidden layer's weight gradient on a POWER-shaped flow, 1.3e-3 here with all-hi panels).  The default panels of a call this
    large keep the lo part of the two head panels: every layer's weight gradient within 5e-4 of the FP32 backward's."""
    from umnn_b200 import kernel, _native
    B, D, E, hidden, Q = 4000, 6, 30, [200, 200, 200], 50          # 1.27 M rows: above the hi-only threshold
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]))
    flat = orc.synth_params(spec, 0)
    net = _net_for(spec, flat, "strided", D)
    kspec = net.kernel_spec()
    d = _dev()
    g = torch.Generator(device=d).manual_seed(1)
    x = 2 * torch.randn(B, D, device=d, generator=g)
    h = torch.randn(B, E * D, device=d, generator=g)
    x0 = torch.zeros_like(x)
    go = torch.rand(B, D, device=d, generator=g) + 0.5
    gfx = -torch.ones(B, D, device=d)
    ref = kernel.cc_backward(kspec, x0, x, h, go, Q, grad_fx=gfx, precision=_native.PREC_FP32)[2].double()
    got = kernel.cc_backward(kspec, x0, x, h, go, Q, grad_fx=gfx, precision=_native.PREC_FP16X3)[2].double()
    off = 0
    for p in net.parameters():
        n = p.numel()
        err = float((got[off:off + n] - ref[off:off + n]).norm() / ref[off:off + n].norm())
        assert err < 5e-4, f"parameter of shape {tuple(p.shape)}: {err:.2e}"
        off += n


@pytest.mark.parametrize("layout,B,D,E,hidden,Q", [("contig", 100, 1, 2, [64, 64, 64], 50), ("strided", 37, 6, 30, [200, 200, 200], 50),
                                                  ("strided", 5, 3, 4, [24, 16], 7)])
def test_prepared_integral_matches_generic_entry(layout, B, D, E, hidden, Q):
    """prepare_integral(...)(x, h) is the same launch as cc_integrate with the per-call host work resolved once: same
    bits; parameters are snapshotted until refresh()."""
    import umnn_b200
    from umnn_b200 import cc_integrate
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]), orc.HIDDEN_LEAKY if layout == "strided" else orc.HIDDEN_RELU)
    flat = orc.synth_params(spec, 2, 1.3)
    net = _net_for(spec, flat, layout, D)
    d = _dev()
    g = torch.Generator(device=d).manual_seed(9)
    x = 2 * torch.randn(B, D, device=d, generator=g)
    x0 = 0.3 * torch.randn(B, D, device=d, generator=g)
    h = torch.randn(B, (E * D) if layout == "strided" else E, device=d, generator=g)
    prep = umnn_b200.prepare_integral(net, B, Q, want_fx=True, want_fx0=True)
    for lo in (None, x0):
        got = prep(x, h, lo)
        ref = cc_integrate(net, lo, x, h, Q, want_fx=True, want_fx0=True)
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
    with pytest.raises(ValueError):
        prep(x[:-1], h[:-1])
    with pytest.raises(ValueError):
        prep(x.double(), h)
    # snapshot semantics
    before = prep(x, h)[0].clone()
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.01)
    assert torch.equal(prep(x, h)[0], before)
    prep.refresh()
    after = prep(x, h)[0]
    # (same extra rows per slot as the prepared launch: the rows-to-tiles cut decides where a slot's node sum is split)
    assert torch.equal(after, cc_integrate(net, None, x, h, Q, want_fx=True, want_fx0=True)[0]) and not torch.equal(after, before)


# ---- full-size configurations (BASELINE.json): sampled oracle checks + size-independent properties ----
def _full_size_check(B, D, E, hidden, Q, n_check, seed):
    from umnn_b200 import IntegrandNetwork, cc_integrate
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]))
    flat = orc.synth_params(spec, 0, 1.0)
    net = _net_for(spec, flat, "strided", D)
    d = _dev()
    g = torch.Generator(device=d).manual_seed(seed)
    x = 2 * torch.randn(B, D, device=d, generator=g)
    h = torch.randn(B, E * D, device=d, generator=g)
    out, fx, _ = cc_integrate(net, None, x, h, Q, want_fx=True)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and torch.isfinite(fx).all() and (fx >= 0).all()
    # (1) samples against the C oracle: contiguous runs at the start of the batch (first CTA's slot range), at its end
    #     (last CTA, the ragged tail) and across the middle CTA boundary, plus a random subset
    run = max(1, min(B, n_check) // 4)
    mid = B // 2
    picks = torch.cat([torch.arange(0, run), torch.arange(B - run, B), torch.arange(max(0, mid - run // 2), min(B, mid + run - run // 2)),
                       torch.randperm(B, generator=torch.Generator().manual_seed(seed))[:run]])
    idx = torch.unique(picks)
    xs, hs = x[idx.to(d)].cpu().numpy(), h[idx.to(d)].cpu().numpy()
    ref, rfx, _ = c_binding.cc_forward(spec, flat, np.zeros_like(xs), xs, hs, Q)
    assert rel_err(out[idx.to(d)].cpu().numpy(), ref) < INTEGRAL_TOL
    assert _point_ok(fx[idx.to(d)].cpu().numpy(), rfx)
    # (2) checksum of checksums: the same samples recomputed in 7 uneven chunks give the same totals
    bounds = [0] + sorted(np.random.RandomState(seed).choice(np.arange(1, B), 6, replace=False).tolist()) + [B]
    total = 0.0
    for a, b in zip(bounds[:-1], bounds[1:]):
        part, _, _ = cc_integrate(net, None, x[a:b], h[a:b], Q, want_fx=True)
        total += float(part.double().sum())
        assert rel_err(part.cpu().numpy(), out[a:b].cpu().numpy()) < 2e-6
    assert abs(total - float(out.double().sum())) < 1e-6 * float(out.double().abs().sum())
    # (3) sign: f > 0 so the integral from 0 has the sign of x
    assert torch.all((out * x) >= 0)


def test_config3_full_size():
    _full_size_check(10000, 6, 30, [200, 200, 200], 50, n_check=512, seed=3)


def test_config2_full_size():
    _full_size_check(10000, 2, 10, [100, 100, 100, 100], 50, n_check=256, seed=2)


def test_config5_full_size():
    _full_size_check(100, 784, 30, [100, 50, 50, 50, 50], 50, n_check=24, seed=5)


def test_config4_full_size():
    _full_size_check(65536, 63, 30, [200, 200, 200], 100, n_check=384, seed=4)


# ---- flows and the monotone regressor on the device ----------------------------------------------------
def _flow_from_golden(device):
    from umnn_b200 import UMNNMAFFlow
    g = dict(np.load(os.path.join(GOLDEN_DIR, "flow_ll.npz"), allow_pickle=False))
    D, E, Q, B = (int(v) for v in g["meta"])
    model = UMNNMAFFlow(nb_flow=2, nb_in=D, hidden_derivative=[50, 50], hidden_embedding=[64, 64],
                        embedding_s=E, nb_steps=Q, solver="CCParallel")
    sd = model.state_dict()
    rng = np.random.RandomState(11)
    for k, v in sd.items():
        if k.endswith(".weight") or k.endswith(".bias"):
            fan_in = v.shape[1] if v.dim() == 2 else sd[k.replace(".bias", ".weight")].shape[1]
            sd[k] = torch.from_numpy(rng.uniform(-1, 1, size=tuple(v.shape)).astype(np.float32) * (1.5 / np.sqrt(fan_in)))
    model.load_state_dict(sd)
    x = rng.standard_normal((B, D)).astype(np.float32)
    return model.to(device), x, g


def test_flow_compute_ll_on_cuda_matches_reference():
    model, xn, g = _flow_from_golden(_dev())
    x = torch.from_numpy(xn).to(_dev()).requires_grad_(True)
    ll, z = model.compute_ll(x)
    ll.sum().backward()
    assert np.max(np.abs(ll.detach().cpu().numpy() - g["ll"])) < 2e-5 * np.max(np.abs(g["ll"]))
    assert np.max(np.abs(z.detach().cpu().numpy() - g["z"])) < 2e-5 * max(1.0, np.max(np.abs(g["z"])))
    assert norm_err(x.grad.cpu().numpy(), g["dx"]) < 2e-3 and rel_to_max(x.grad.cpu().numpy(), g["dx"]) < 2e-2
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert_sum_grad_ok(p.grad.cpu().numpy(), g["grad/" + k], what=k)
    model.eval()
    with torch.no_grad():
        z_eval = model.forward(torch.from_numpy(xn).to(_dev()))
        with contextlib.redirect_stdout(io.StringIO()):
            x_back = model.invert(torch.from_numpy(g["z"][:4].copy()).to(_dev()), iter=5)
    assert np.max(np.abs(z_eval.cpu().numpy() - g["z_forward_eval"])) < 2e-5 * max(1.0, np.max(np.abs(g["z_forward_eval"])))
    assert np.max(np.abs(x_back.cpu().numpy() - g["invert_x"])) < 1e-3


def test_invert_bracket_step_bit_exact_vs_oracle():
    """umnn_invert_bracket_step against the numpy restatement of UMNNMAF.py:213-231: bit-exact, including the
    reference's wrap-around neighbour reads at grid points 0 and G-1."""
    from umnn_b200 import kernel
    rng = np.random.RandomState(5)
    for B, G in ((1, 10), (7, 10), (333, 10), (64, 4)):
        grid = orc.invert_grid(G)
        left = rng.uniform(-50, -1, B).astype(np.float32)
        right = rng.uniform(1, 50, B).astype(np.float32)
        x_cur = (grid[:, None] * (right - left)[None, :] + left[None, :]).astype(np.float32)
        integ = np.tanh(x_cur / 10).astype(np.float32) * rng.uniform(0.5, 2, B).astype(np.float32)[None, :]
        target = rng.uniform(-2.5, 2.5, B).astype(np.float32)      # some targets fall outside the bracket image
        offset = rng.standard_normal(B).astype(np.float32) * 0.1
        scale = np.float32(1.25)
        l_ref, r_ref, xn_ref, xm_ref = orc.invert_bracket_step(integ, x_cur, grid, offset, scale, target)
        dev = _dev()
        D = 3                                                       # strided columns like left[:, j]
        lt = torch.zeros(B, D, device=dev); rt = torch.zeros(B, D, device=dev); xm = torch.zeros(B, D, device=dev)
        tg = torch.zeros(B, D, device=dev); tg[:, 1] = torch.from_numpy(target).to(dev)
        lt[:, 1] = torch.from_numpy(left).to(dev); rt[:, 1] = torch.from_numpy(right).to(dev)
        g_t = torch.from_numpy(grid).to(dev)
        x0_t = torch.empty(G, B, device=dev)
        kernel.invert_bracket_step(None, None, g_t, None, None, None, lt[:, 1], rt[:, 1], x0_t, None)
        assert np.array_equal(x0_t.cpu().numpy(), x_cur)
        x1_t = torch.empty(G, B, device=dev)
        kernel.invert_bracket_step(torch.from_numpy(integ).to(dev), x0_t, g_t, torch.from_numpy(offset).to(dev),
                                   torch.tensor([scale], device=dev), tg[:, 1], lt[:, 1], rt[:, 1], x1_t, xm[:, 1])
        assert np.array_equal(lt[:, 1].cpu().numpy(), l_ref)
        assert np.array_equal(rt[:, 1].cpu().numpy(), r_ref)
        assert np.array_equal(x1_t.cpu().numpy(), xn_ref)
        assert np.array_equal(xm[:, 1].cpu().numpy(), xm_ref)
        assert float(lt[:, 0].abs().sum() + lt[:, 2].abs().sum() + xm[:, 0].abs().sum()) == 0.0


def test_invert_native_matches_op_by_op_loop(monkeypatch):
    """The two-launch-per-round invert and the reference-shaped torch loop (same integrals through the same kernel)
    agree exactly, and both recover x from z = forward(x)."""
    model, xn, g = _flow_from_golden(_dev())
    model.eval()
    z = torch.from_numpy(g["z"][:16].copy()).to(_dev())
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        x_native = model.invert(z, iter=6)
        monkeypatch.setenv("UMNN_B200_INVERT", "torch")
        x_loop = model.invert(z, iter=6)
    assert torch.equal(x_native, x_loop)
    assert np.max(np.abs(x_native.cpu().numpy() - xn[:16])) < 5e-2


@pytest.mark.parametrize("iters", [1, 6, 7])
def test_invert_dimension_call_matches_round_by_round(monkeypatch, iters):
    """umnn_invert_dimension (one native call per dimension: context replication, bracket reset and all refinement
    rounds enqueued from C, result written straight into x[:, j]) against the same launches issued round by round from
    Python: identical results, also after a parameter update and for odd round counts (grid ping-pong)."""
    from umnn_b200 import UMNNMAFFlow
    torch.manual_seed(0)
    model = UMNNMAFFlow(nb_flow=1, nb_in=24, hidden_derivative=[50, 50, 50], hidden_embedding=[64, 64], embedding_s=10,
                        nb_steps=20, solver="CCParallel", device=_dev()).to(_dev())
    model.eval()
    z = torch.randn(16, 24, device=_dev())
    for attempt in range(2):
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            x_call = model.invert(z, iter=iters)
            monkeypatch.setenv("UMNN_B200_INVERT", "rounds")
            x_rounds = model.invert(z, iter=iters)
            monkeypatch.delenv("UMNN_B200_INVERT")
        assert torch.equal(x_call, x_rounds)
        with torch.no_grad():
            for p in model.parameters():
                if p.requires_grad:
                    p.mul_(0.97)


def test_invert_conditional_flow_and_partial_conditioner_outputs():
    """EmbeddingNetwork.embedding_of_dim (the E rows of the last masked layer one dimension needs) against the full
    conditioner pass, for plain and conditional MADE, on the device; and invert() of a conditional flow."""
    from umnn_b200 import EmbeddingNetwork, UMNNMAFFlow
    torch.manual_seed(1)
    for cond in (0, 3):
        emb = EmbeddingNetwork(7, [64, 64], [32, 32], 6, cond_in=cond, device=_dev())
        x = torch.randn(9, 7, device=_dev())
        ctx = torch.randn(9, cond, device=_dev()) if cond else None
        with torch.no_grad():
            full = emb.make_embeding(x, ctx)
            for j in range(7):
                assert torch.allclose(emb.embedding_of_dim(x, j, ctx), full[:, j::7], rtol=1e-5, atol=1e-6)
    model = UMNNMAFFlow(nb_flow=1, nb_in=5, hidden_derivative=[50, 50], hidden_embedding=[64, 64], embedding_s=8, nb_steps=20,
                        solver="CCParallel", cond_in=2, device=_dev()).to(_dev())
    model.eval()
    x = 0.5 * torch.randn(12, 5, device=_dev())
    ctx = torch.randn(12, 2, device=_dev())
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        z = model.forward(x, context=ctx)
        x_back = model.invert(z, iter=8, context=ctx)
    assert float((x_back - x).abs().max()) < 5e-2


def test_monotonic_nn_on_cuda():
    from umnn_b200 import MonotonicNN
    torch.manual_seed(0)
    model = MonotonicNN(3, [64, 64, 64], nb_steps=50, dev=_dev()).to(_dev())
    x = torch.randn(100, 1, device=_dev())
    h = torch.randn(100, 2, device=_dev())
    y = model(x, h)
    y.sum().backward()
    cpu = MonotonicNN(3, [64, 64, 64], nb_steps=50)
    cpu.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    y_cpu = cpu(x.cpu(), h.cpu())
    assert torch.allclose(y.detach().cpu(), y_cpu.detach(), rtol=1e-5, atol=1e-5)
    xs = torch.linspace(-3, 3, 200, device=_dev()).view(-1, 1)
    ys = model(xs, h[:1].expand(200, -1)).view(-1)
    assert torch.all(ys[1:] >= ys[:-1] - 1e-5)


def test_training_curves_agree_between_backward_paths(monkeypatch):
    """Training-level gate for the BF16x3 backward: the same flow trained for 25 Adam steps
    with the tensor-core backward and with the FP32 backward follows the same loss curve."""
    from umnn_b200 import UMNNMAFFlow

    def train(mode):
        monkeypatch.setenv("UMNN_B200_BACKWARD", mode)
        torch.manual_seed(0)
        model = UMNNMAFFlow(nb_flow=2, nb_in=6, hidden_derivative=[200, 200, 200], hidden_embedding=[128, 128],
                            embedding_s=30, nb_steps=50, solver="CCParallel", device=_dev()).to(_dev())
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        gen = torch.Generator(device="cpu").manual_seed(1)
        data = torch.randn(512, 6, generator=gen)
        data[:, 1] = 0.5 * data[:, 0] ** 2 + 0.3 * data[:, 1]
        data = data.to(_dev())
        losses = []
        for _ in range(25):
            opt.zero_grad()
            ll, _ = model.compute_ll(data)
            loss = -ll.mean()
            loss.backward()
            opt.step()
            losses.append(float(loss))
        return np.array(losses)

    ref = train("fp32")
    assert ref[-1] < ref[0] - 0.1                                  # it actually trains
    for mode in ("auto",):                                         # auto = fp16x3 re-evaluation, guarded
        tc = train(mode)
        assert np.max(np.abs(tc - ref)) < 2e-3 * max(1.0, np.max(np.abs(ref))), mode


def test_fused_flow_block_matches_unfused():
    """UMNNMAF.forward_and_log_jac on the kernel route (integral + Jacobian point from one launch, Jacobian
    cotangent folded into the fused backward) against forward() + the torch evaluation of the Jacobian."""
    from umnn_b200 import EmbeddingNetwork, UMNNMAF
    torch.manual_seed(0)
    emb = EmbeddingNetwork(6, [64, 64], [200, 200, 200], 30, device=_dev())
    blk = UMNNMAF(emb, 6, 50, _dev(), solver="CCParallel").to(_dev())
    x = torch.randn(128, 6, device=_dev())
    outs = []
    for fused in (True, False):
        for p in blk.parameters():
            p.grad = None
        xx = x.clone().requires_grad_(True)
        z, lj = blk.forward_and_log_jac(xx) if fused else blk.compute_log_jac_bis(xx)
        (z.pow(2).sum() + lj.sum()).backward()
        outs.append((z.detach(), lj.detach(), xx.grad.clone(),
                     torch.cat([p.grad.reshape(-1) for p in blk.parameters() if p.grad is not None])))
    (z1, l1, gx1, gp1), (z2, l2, gx2, gp2) = outs
    assert torch.allclose(z1, z2, rtol=1e-5, atol=1e-5) and torch.allclose(l1, l2, rtol=1e-5, atol=1e-5)
    assert norm_err(gx1.cpu().numpy(), gx2.cpu().numpy()) < 2e-3
    assert norm_err(gp1.cpu().numpy(), gp2.cpu().numpy()) < 2e-3


def test_cuda_graph_capture_and_replay():
    """The fused launch is stream-ordered and allocation-light: an eval-mode forward (integral + Jacobian
    point) can be captured in a CUDA graph and replayed on new inputs (the latency path for small batches)."""
    from umnn_b200 import cc_integrate
    spec = orc.MLPSpec((31, 100, 50, 50, 50, 50, 1))
    flat = orc.synth_params(spec, 0, 1.0)
    net = _net_for(spec, flat, "strided", 8).eval()
    d = _dev()
    gen = torch.Generator(device=d).manual_seed(0)
    x = torch.randn(12, 8, device=d, generator=gen)
    h = torch.randn(12, 240, device=d, generator=gen)
    cc_integrate(net, None, x, h, 50, want_fx=True)              # warm-up: packs parameters, loads tables
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        cc_integrate(net, None, x, h, 50, want_fx=True)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out, fx, _ = cc_integrate(net, None, x, h, 50, want_fx=True)
    x.copy_(torch.randn(12, 8, device=d, generator=gen))
    h.copy_(torch.randn(12, 240, device=d, generator=gen))
    graph.replay()
    torch.cuda.synchronize()
    want, want_fx, _ = cc_integrate(net, None, x, h, 50, want_fx=True)
    assert torch.equal(out, want) and torch.equal(fx, want_fx)


def test_graphed_log_likelihood_replays_and_tracks_parameter_updates():
    """UMNNMAFFlow.compute_ll captured as one CUDA graph: replays equal the eager call on new
    inputs, and -- because the parameter packing is captured too -- keep doing so after an in-place update."""
    from umnn_b200 import GraphedLogLikelihood
    model, xn, g = _flow_from_golden(_dev())
    model.eval()
    B = 8
    graphed = GraphedLogLikelihood(model, B)
    x = torch.from_numpy(xn[:B].copy()).to(_dev())
    ll, z = graphed(x)
    torch.cuda.synchronize()
    with torch.no_grad():
        ll_e, z_e = model.compute_ll(x)
    assert torch.equal(ll, ll_e) and torch.equal(z, z_e)
    ll_first = ll.clone()                       # the graph's outputs are static tensors, rewritten by every replay
    assert np.max(np.abs(ll.cpu().numpy() - g["ll"][:B])) < 2e-5 * np.max(np.abs(g["ll"]))
    with torch.no_grad():
        for p in model.parameters():
            if p.requires_grad:
                p.mul_(0.9)
    x2 = torch.from_numpy(xn[B:2 * B].copy()).to(_dev())
    ll2, z2 = graphed(x2)
    torch.cuda.synchronize()
    with torch.no_grad():
        ll2_e, z2_e = model.compute_ll(x2)
    assert torch.equal(ll2, ll2_e) and torch.equal(z2, z2_e)
    assert not torch.equal(ll2, ll_first)
    with pytest.raises(ValueError):
        graphed(x2[:5])


@pytest.mark.parametrize("chunks", [None, 1, 3, 64])
def test_host_buffer_entry_is_pipelined_and_exact(chunks):
    """cc_integrate_host (pinned host tensors in and out, copies overlapped with the launches over batch chunks)
    returns exactly what the resident-input call returns."""
    from umnn_b200 import cc_integrate, cc_integrate_host
    spec = orc.MLPSpec((31, 200, 200, 200, 1))
    net = _net_for(spec, orc.synth_params(spec, 0, 1.0), "strided", 6).eval()
    gen = torch.Generator(device="cpu").manual_seed(3)
    x_host = (2 * torch.randn(50, 6, generator=gen)).pin_memory()
    h_host = torch.randn(50, 180, generator=gen).pin_memory()
    out, fx = cc_integrate_host(net, x_host, h_host, 50, want_fx=True, chunks=chunks)
    torch.cuda.synchronize()
    want, want_fx, _ = cc_integrate(net, None, x_host.to(_dev()), h_host.to(_dev()), 50, want_fx=True)
    # point evaluations are per row: exact.  The per-slot sums see a different tile partition when the batch is cut
    # (a slot that straddles two tiles is summed in two pieces): summation-order noise only.
    assert torch.equal(fx, want_fx.cpu())
    assert rel_err(out.numpy(), want.cpu().numpy()) < 2e-6
    if chunks in (None, 1):
        assert torch.equal(out, want.cpu())
    out2, none = cc_integrate_host(net, x_host, h_host, 50, chunks=chunks)
    torch.cuda.synchronize()
    want2 = cc_integrate(net, None, x_host.to(_dev()), h_host.to(_dev()), 50)[0]   # rows per slot differ without f(x)
    assert none is None and rel_err(out2.numpy(), want2.cpu().numpy()) < 2e-6
