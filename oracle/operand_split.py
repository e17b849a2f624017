"""CPU emulation of the tensor-core kernels' operand split (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

The CUDA kernels evaluate every hidden-to-hidden layer of the integrand MLP as
    A_hi.B_hi + A_lo.B_hi + A_hi.B_lo          (fp32 accumulation on the tensor cores)
with activations A and weights B split into a 16-bit ``hi`` and a 16-bit ``lo`` part (bf16 or fp16); layer 1, the
output layer, the activations and the quadrature sum are plain fp32.  This module restates that
arithmetic in numpy with EXACT (float64) accumulation, so that what remains is the representation error of the split
itself: it predicts the error level the GPU parity tests hold each precision to, without a GPU
(tests/test_operand_split_cpu.py).  It follows the reference's integrate(...) (ParallelNeuralIntegral.py:37-65) and
IntegrandNetwork.forward (UMNNMAF.py:263-284) through oracle/umnn_oracle.py for everything but the split layers.
"""
from __future__ import annotations

import numpy as np

from . import umnn_oracle as orc


def round_bf16(a: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32, as cvt.rn.bf16x2.f32 does."""
    b = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def round_fp16(a: np.ndarray) -> np.ndarray:
    """fp32 -> fp16 (round to nearest even, overflow to inf) -> fp32, as cvt.rn.f16x2.f32 does."""
    with np.errstate(over="ignore"):
        return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


ROUND = {"bf16": round_bf16, "fp16": round_fp16}


def split(a: np.ndarray, fmt: str):
    """hi = round(a), lo = round(a - hi) with the subtraction in fp32 (exact here: a - hi is representable)."""
    rnd = ROUND[fmt]
    a = np.asarray(a, np.float32)
    hi = rnd(a)
    with np.errstate(invalid="ignore"):
        lo = rnd((a - hi).astype(np.float32))
    return hi, lo


def mlp_rows_split(spec: orc.MLPSpec, flat: np.ndarray, rows: np.ndarray, fmt: str, signs: list = None) -> np.ndarray:
    """orc.mlp_rows with the hidden-to-hidden layers evaluated through the hi/lo split (three products, exact sums).

    ``signs`` (optional list) receives, per hidden layer, the boolean array ``pre-activation < 0`` -- what the
    backward's re-evaluation records as the derivative mask of the hidden activation."""
    layers = orc.unpack_params(spec, flat)
    a = np.asarray(rows, np.float32)
    n = len(layers)
    for li, (W, b) in enumerate(layers):
        W = W.astype(np.float32)
        b = b.astype(np.float32)
        if 0 < li < n - 1:
            a_hi, a_lo = split(a, fmt)
            w_hi, w_lo = split(W, fmt)
            # the bias rides in the GEMM as three 16-bit pieces (tc_layout.cuh): hi + lo + a third term
            b_hi, b_lo = split(b, fmt)
            b_3 = ROUND[fmt]((b - b_hi - b_lo).astype(np.float32))
            acc = (a_hi.astype(np.float64) @ w_hi.T.astype(np.float64) + a_lo.astype(np.float64) @ w_hi.T.astype(np.float64)
                   + a_hi.astype(np.float64) @ w_lo.T.astype(np.float64))
            v = (acc + (b_hi.astype(np.float64) + b_lo.astype(np.float64) + b_3.astype(np.float64))).astype(np.float32)
        else:
            v = (a.astype(np.float64) @ W.T.astype(np.float64) + b.astype(np.float64)).astype(np.float32)
        if signs is not None and li < n - 1:
            signs.append(v < 0)
        a = orc._hidden_act(v, spec.hidden_act) if li < n - 1 else orc._out_act(v, spec.out_act)
    return a.reshape(-1)


def integrate_parallel_split(spec: orc.MLPSpec, flat: np.ndarray, x0: np.ndarray, x: np.ndarray, h: np.ndarray, Q: int,
                             layout: str = "strided", fmt: str = "fp16") -> np.ndarray:
    """orc.integrate_parallel (fp32 node placement and weighting) around mlp_rows_split."""
    w, t = orc.cc_nodes_weights(Q)
    x0 = np.asarray(x0, np.float32)
    x = np.asarray(x, np.float32)
    h = np.asarray(h, np.float32)
    w = w.astype(np.float32)
    t = t.astype(np.float32)
    one, two = np.float32(1), np.float32(2)
    xT = orc._limits(x0, x, Q)
    B, Dx = x.shape
    nodes = x0[:, None, :] + (xT - x0)[:, None, :] * (t[None, :, None] + one) / two           # [B, Q+1, Dx]
    h_rep = np.broadcast_to(h[:, None, :], (B, Q + 1, h.shape[1])).reshape(B * (Q + 1), -1)
    xs = nodes.reshape(B * (Q + 1), Dx)
    rows = orc.slot_inputs_strided(xs, h_rep) if layout == "strided" else np.concatenate([xs, h_rep], axis=1)
    f = mlp_rows_split(spec, flat, rows, fmt).reshape(B, Q + 1, Dx)
    z = (f.astype(np.float64) * w[None, :, None].astype(np.float64)).sum(axis=1).astype(np.float32)
    return z * (xT - x0) / two
