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
