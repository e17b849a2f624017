"""CPU check of the precision design of the tensor-core kernels: a numpy emulation of the hi/lo operand split with
exact accumulation (oracle/operand_split.py) against the reference's own float64 results (tests/golden).

What it pins, without a GPU: the FP16x3 split (default) stays inside the north-star 1e-5 on every golden network,
including the "trained-like" stress network, and is several times more accurate there than the BF16x3 split, whose
error is the representation error of its operands (the GPU kernels measure the same levels)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden_case, rel_err
from oracle import operand_split as ops


def test_rounding_helpers_match_ieee():
    rs = np.random.RandomState(0)
    a = (rs.standard_normal(4096) * np.exp(rs.uniform(-8, 8, 4096))).astype(np.float32)
    b = ops.round_bf16(a)
    assert np.all((b.view(np.uint32) & 0xFFFF) == 0)                      # 16 significant bits left
    assert np.all(np.abs(b - a) <= np.abs(a) * 2.0 ** -8)                  # half an ulp of an 8-bit significand
    ties = np.array([1.0 + 2.0 ** -8, 1.0 + 3 * 2.0 ** -8], np.float32)    # exactly between two bf16 values
    np.testing.assert_array_equal(ops.round_bf16(ties), np.array([1.0, 1.0 + 2.0 ** -6], np.float32))   # ties to even
    hi, lo = ops.split(a, "bf16")
    assert np.all(np.abs((hi.astype(np.float64) + lo) - a) <= np.abs(a) * 2.0 ** -16)
    a16 = a[(np.abs(a) > 2.0 ** -3) & (np.abs(a) < 6.0e4)]
    hi, lo = ops.split(a16, "fp16")
    assert np.all(np.abs((hi.astype(np.float64) + lo) - a16) <= np.abs(a16) * 2.0 ** -22)
    assert np.isinf(ops.round_fp16(np.float32(7.0e4)))                     # the range the kernel's guard watches


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_split_error_levels_against_reference_fp64(name):
    spec, flat, inp, g = load_golden_case(name)
    if len(spec.widths) < 4:
        pytest.skip("one hidden layer: no hidden-to-hidden product, the kernels use the FP32 path")
    want = g["fp64_integral"].astype(np.float32)
    err = {fmt: rel_err(ops.integrate_parallel_split(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], inp["layout"], fmt),
                        want) for fmt in ("fp16", "bf16")}
    # emulation: fp16 1.3e-7 .. 8.8e-7, bf16 1.6e-7 .. 1.2e-5 (stress network cfg3_power_trained: 8.8e-7 vs 1.19e-5;
    # the B200 kernels, whose fp32 accumulation order adds ~3e-7, measure 2.1e-6 vs 1.26e-5 on the same network)
    assert err["fp16"] < 3e-6
    assert err["bf16"] < 3e-5
    if float(g["meta_gain"]) != 1.0:
        assert err["fp16"] * 3 < err["bf16"]       # the stress networks are where the extra 5 bits show


def test_fp16_reevaluation_flips_fewer_activation_kinks():
    """The backward re-evaluates the network and takes act'(.) from the sign of every pre-activation: a unit whose
    pre-activation sits within the rounding noise of the re-evaluation can land on the other side of the kink than in
    exact arithmetic.  With fp16 hi/lo operands (22 bits) that happens several times less often than with bf16 hi/lo
    (~17 bits) -- the reason the default backward re-evaluates in FP16x3.  On this seeded sample of
    19.7 M hidden units of the stress network: 29 flipped signs with the bf16 split, 3 with the fp16 split."""
    from oracle import umnn_oracle as orc
    spec, flat, inp, g = load_golden_case("cfg3_power_trained")
    rs = np.random.RandomState(2)
    B, Dx = inp["x"].shape
    reps = 256                                                                                  # abscissae per slot
    xs = (inp["x"] * rs.uniform(0, 1, size=(reps, B, Dx))).reshape(-1, Dx).astype(np.float32)
    hs = np.tile(inp["h"], (reps, 1)).astype(np.float32)
    rows = orc.slot_inputs_strided(xs, hs)
    _, pre = orc.mlp_rows(spec, flat.astype(np.float64), rows.astype(np.float64), keep=True)
    exact = [v < 0 for (_, v) in pre[1:-1]]                  # layers fed by a split product (layer 1 is plain fp32)
    flips = {}
    for fmt in ("fp16", "bf16"):
        got = []
        ops.mlp_rows_split(spec, flat, rows, fmt, signs=got)
        flips[fmt] = sum(int(np.count_nonzero(a != b)) for a, b in zip(got[1:], exact))
    n_units = sum(e.size for e in exact)
    assert flips["bf16"] > 0 and flips["fp16"] * 4 < flips["bf16"], flips


def test_hi_only_panels_average_out_unless_the_rows_are_coherent():
    """Why the weight-gradient pass keeps the lo part of the two head panels (tc_bwd_layout.cuh:
    BwdPanels).  dW = sum over rows of dz^T a with both operands rounded to bf16 ("hi-only" panels), exact accumulation:

    * incoherent rows (every row its own dz): the rounding errors are independent and the error of the sum falls like
      1/sqrt(rows);
    * the rank-1 head dz_J = dv * w_out with the SAME dv in every row (a likelihood's Jacobian rows): every row rounds
      dv * w_out[k] the same way, so the error of column k is the rounding error of that one number however many rows
      are summed -- and a second (lo) part of dz removes it."""
    rs = np.random.RandomState(0)
    K, I = 64, 48
    w_out = rs.standard_normal(K).astype(np.float32)
    errs = {}
    for n_rows in (1 << 10, 1 << 16):
        a = np.abs(rs.standard_normal((n_rows, I))).astype(np.float32) + 0.5        # activations: same sign, no cancellation
        for kind in ("incoherent", "coherent"):
            dv = (rs.uniform(0.5, 1.5, (n_rows, 1)) if kind == "incoherent" else np.full((n_rows, 1), 0.7331)).astype(np.float32)
            dz = (dv * w_out[None, :]).astype(np.float32)
            exact = dz.astype(np.float64).T @ a.astype(np.float64)
            a_hi = ops.round_bf16(a).astype(np.float64)
            dz_hi, dz_lo = ops.split(dz, "bf16")
            hi_only = dz_hi.astype(np.float64).T @ a_hi
            hi_lo = (dz_hi.astype(np.float64) + dz_lo).T @ a_hi
            nrm = np.linalg.norm(exact)
            errs[(kind, n_rows, "hi")] = np.linalg.norm(hi_only - exact) / nrm
            errs[(kind, n_rows, "hilo")] = np.linalg.norm(hi_lo - exact) / nrm
    # independent rounding errors: 64 times the rows, ~8 times less error
    assert errs[("incoherent", 1 << 16, "hi")] < 0.25 * errs[("incoherent", 1 << 10, "hi")]
    assert errs[("incoherent", 1 << 16, "hi")] < 2e-5
    # coherent rows: the error does not move with the number of rows and sits at the rounding error of one bf16 number
    assert errs[("coherent", 1 << 16, "hi")] > 0.8 * errs[("coherent", 1 << 10, "hi")] > 2e-4
    # ... until dz carries its lo part (what remains is the rounding of the activations, which does average out)
    assert errs[("coherent", 1 << 16, "hilo")] < 2e-5


def test_row_scaling_would_decorrelate_coherent_rounding():
    """The cure DESIGN.md 8 item 2 names but does not build: scale row r of dz by a pseudo-random c_r in [1, 2) and row r of the
    activations by 1 / c_r before rounding both to bf16.  Equal dz values then round differently from row to row, the products
    are unchanged up to fp32 rounding of c * (1 / c), and the error of the sum averages out again -- without a lo part."""
    rs = np.random.RandomState(1)
    K, I, n_rows = 64, 48, 1 << 16
    w_out = rs.standard_normal(K).astype(np.float32)
    a = np.abs(rs.standard_normal((n_rows, I))).astype(np.float32) + 0.5
    dz = (np.full((n_rows, 1), 0.7331, np.float32) * w_out[None, :]).astype(np.float32)
    exact = dz.astype(np.float64).T @ a.astype(np.float64)
    plain = ops.round_bf16(dz).astype(np.float64).T @ ops.round_bf16(a).astype(np.float64)
    c = (1.0 + rs.randint(0, 1 << 16, (n_rows, 1)) / 65536.0).astype(np.float32)
    inv_c = (np.float32(1.0) / c).astype(np.float32)
    scaled = ops.round_bf16((dz * c).astype(np.float32)).astype(np.float64).T @ ops.round_bf16((a * inv_c).astype(np.float32)).astype(np.float64)
    nrm = np.linalg.norm(exact)
    e_plain, e_scaled = np.linalg.norm(plain - exact) / nrm, np.linalg.norm(scaled - exact) / nrm
    assert e_plain > 2e-4 and e_scaled < 0.1 * e_plain and e_scaled < 3e-5
