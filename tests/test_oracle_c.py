"""The C restatement (oracle/umnn_oracle.c) against the golden vectors and the numpy oracle."""
import time

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden_case, rel_err
from oracle import c_binding, umnn_oracle as orc


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_c_forward_matches_golden(name):
    spec, flat, inp, g = load_golden_case(name)
    z, fx, fx0 = c_binding.cc_forward(spec, flat, inp["x0"], inp["x"], inp["h"], inp["Q"], inp["layout"])
    assert rel_err(z, g["par_integral"]) < 3e-6
    assert np.all(np.abs(fx - g["f_at_x"]) <= 2e-6 * np.abs(g["f_at_x"]) + 1.2e-7)
    assert np.all(np.abs(fx0 - g["f_at_x0"]) <= 2e-6 * np.abs(g["f_at_x0"]) + 1.2e-7)


def test_c_matches_numpy_on_sigmoid_and_threads():
    spec = orc.MLPSpec((4, 24, 16, 1), orc.HIDDEN_LEAKY, orc.OUT_SIGMOID)
    flat = orc.synth_params(spec, 1, 2.0)
    x0, x, h, _ = orc.synth_inputs(9, 4, 12, 2, False)
    ref = orc.integrate_parallel(spec, flat, x0, x, h, 30)
    for nt in (1, 3):
        z, _, _ = c_binding.cc_forward(spec, flat, x0, x, h, 30, n_threads=nt)
        assert rel_err(z, ref) < 3e-6
    assert c_binding.max_threads() >= 1
