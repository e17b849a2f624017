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


# cfg3_power_adam: integrand weights of a reference flow TRAINED for 300 Adam steps and the trained conditioner's h
# (tests/golden/make_trained_golden.py); cfg3_power_trained: random weights scaled x2.5 ("trained-like")
GOLDEN_CASES = ["cfg1_monotonic", "cfg1_monotonic_x0", "cfg2_toy", "cfg3_power", "cfg3_power_trained", "cfg3_power_adam",
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
    if "stored_flat" in g:      # trained weights / real conditioner outputs travel in the file
        flat, x0, x, h, go = (g[k] for k in ("stored_flat", "stored_x0", "stored_x", "stored_h", "stored_grad_out"))
    else:
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


# ---- flip-aware gradient metric --------------------------------------------------------------------------------
# LeakyReLU / ReLU derivatives jump at 0.  Two correct evaluations of the same network that differ by rounding may
# put a unit on different sides of its kink; the gradient of THAT slot then differs by that unit's whole contribution
# (the reference's own fp32 vs fp64 d_h differ by 1.2e-4 of the largest entry for this reason).
# oracle.kink_margins() says, per slot and in float64, how close the nearest hidden pre-activation comes to its kink
# relative to the magnitude of the terms it is summed from.  A slot is "kink-ambiguous" for an implementation when
# that margin is below the implementation's rounding level (KINK_BAND); only such slots may deviate, deviating slots
# are COUNTED, and every other slot is held to a tight bound.
KINK_BAND = {"fp32": 5e-6, "auto": 5e-6, "fp16x3": 5e-6, "bf16x3": 1.5e-4}
SLOT_TIGHT_TOL = 1e-4        # per-slot bound (relative to the largest reference entry) for unambiguous slots
GRAD_NORM_TOL = 2e-3         # normwise bound over everything, flips included
GRAD_MAX_TOL = 5e-3          # max bound for sums over slots (d_params), flips included (they are diluted)


def norm_err(a, ref):
    return float(np.linalg.norm((np.asarray(a, np.float64) - ref).ravel()) / max(np.linalg.norm(np.asarray(ref).ravel()), 1e-30))


def slot_grad_report(got, ref, margins, band, per_slot_shape):
    """Per-slot comparison of a gradient with one group of entries per (sample, dim) slot.

    got/ref: arrays that reshape to per_slot_shape = [B, Dx, k] (k entries per slot; use `to_slots` for d_h).
    Returns dict(n_slots, n_ambiguous, n_flipped, max_unambiguous, max_unflipped, max_all, normwise).
    """
    got = np.asarray(got, np.float64).reshape(per_slot_shape)
    ref = np.asarray(ref, np.float64).reshape(per_slot_shape)
    scale = max(float(np.max(np.abs(ref))), 1e-30)
    err = np.max(np.abs(got - ref), axis=-1) / scale                      # [B, Dx]
    amb = np.asarray(margins).reshape(err.shape) < band
    flipped = err > SLOT_TIGHT_TOL
    return dict(n_slots=int(err.size), n_ambiguous=int(amb.sum()), n_flipped=int(flipped.sum()),
                n_flipped_unexplained=int((flipped & ~amb).sum()),
                max_unambiguous=float(err[~amb].max()) if (~amb).any() else 0.0,
                max_unflipped=float(err[~flipped].max()) if (~flipped).any() else 0.0,
                max_all=float(err.max()), normwise=norm_err(got, ref))


def dh_to_slots(d_h, B, Dx, layout):
    """d_h in the layout of h -> [B, Dx, E] (one group of E context gradients per slot)."""
    d_h = np.asarray(d_h)
    if layout == "strided":
        return d_h.reshape(B, -1, Dx).transpose(0, 2, 1)
    return d_h.reshape(B, 1, -1)


def assert_slot_grad_ok(got, ref, margins, precision, per_slot_shape, what="", max_flip_frac=0.08):
    r = slot_grad_report(got, ref, margins, KINK_BAND[precision], per_slot_shape)
    msg = f"{what} [{precision}] {r}"
    assert r["n_flipped_unexplained"] == 0, "a slot deviates although no unit is near its kink: " + msg
    assert r["max_unambiguous"] <= SLOT_TIGHT_TOL, msg
    assert r["n_flipped"] <= max(2, int(max_flip_frac * r["n_slots"])), "too many kink flips: " + msg
    assert r["normwise"] <= GRAD_NORM_TOL, msg
    return r


def assert_sum_grad_ok(got, ref, what=""):
    """Gradients that are sums over all slots (d_params): flips are diluted, so plain bounds apply."""
    n, m = norm_err(got, ref), rel_to_max(np.asarray(got, np.float64), np.asarray(ref, np.float64))
    assert n <= GRAD_NORM_TOL and m <= GRAD_MAX_TOL, f"{what}: normwise {n:.2e} (<= {GRAD_NORM_TOL}), max {m:.2e} (<= {GRAD_MAX_TOL})"
    return n, m
