"""pytest configuration: the `gpu` marker and shared helpers for the golden fixtures."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN_CASES = ["cfg1_monotonic", "cfg1_monotonic_x0", "cfg2_toy", "cfg3_power", "cfg3_power_trained",
                "cfg4_bsds", "cfg5_mnist", "small_odd", "jit_shape"]


def load_golden_case(name):
    """Regenerate the seeded inputs of a golden case and return (spec, flat, inputs, golden dict)."""
    import hashlib
    from oracle import umnn_oracle as orc
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    B, Dx, E, Q = (int(v) for v in g["meta_shape"])
    hidden = [int(v) for v in g["meta_hidden"]]
    layout = str(g["meta_layout"])
    pseed, dseed = (int(v) for v in g["meta_seeds"])
    spec = orc.MLPSpec(tuple([1 + E] + hidden + [1]),
                       orc.HIDDEN_LEAKY if layout == "strided" else orc.HIDDEN_RELU, orc.OUT_ELU_PLUS_1)
    flat = orc.synth_params(spec, pseed, float(g["meta_gain"]))
    Hh = E * Dx if layout == "strided" else E
    x0, x, h, go = orc.synth_inputs(B, Dx, Hh, dseed, bool(g["meta_x0_zero"]))
    m = hashlib.sha256()
    for a in (flat, x0, x, h, go):
        m.update(np.ascontiguousarray(a).tobytes())
    assert m.hexdigest()[:16] == str(g["input_checksum"]), "seeded input stream drifted from the golden generator"
    return spec, flat, dict(x0=x0, x=x, h=h, grad_out=go, Q=Q, layout=layout, B=B, Dx=Dx, E=E, hidden=hidden), g


def rel_err(a, ref, floor=1e-6):
    """max |a-ref| / max(|ref|, floor)  -- the integral metric of SURVEY.md 8(d)."""
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor)))


def rel_to_max(a, ref):
    """max |a-ref| / max|ref|  -- the gradient metric of SURVEY.md 8(c)."""
    return float(np.max(np.abs(a - ref)) / max(float(np.max(np.abs(ref))), 1e-30))
