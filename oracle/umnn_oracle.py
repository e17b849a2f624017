"""CPU oracle for the UMNN Clenshaw-Curtis integration hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``umnn_b200/`` or ``models/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker or
as the CPU arm that is timed beside the GPU path.

Every function restates, in plain numpy, the algorithm of a reference function
(AWehenkel/UMNN @ 59118c14) and cites the ``file:line`` it follows.  The
restatement is PINNED: ``tests/golden/make_golden.py`` imports the real reference
from ``/root/reference`` in the build container, runs it on seeded inputs and
commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this module against those vectors on every CPU test run.

Conventions
-----------
* ``Q`` = ``nb_steps``; there are ``Q + 1`` quadrature nodes.
* parameters travel as one flat fp32 vector in ``nn.Sequential`` order
  ``W1 (out x in, row-major), b1, W2, b2, ...`` -- the order of
  ``_flatten(integrand.parameters())`` (ParallelNeuralIntegral.py:6-8).
* flow layout ("strided"): ``x[N, D]``, ``h[N, E*D]`` with ``h[n, e*D + d]`` the
  e-th context value of slot (n, d)                         (UMNNMAF.py:263-284)
* contiguous layout: ``x[N, 1]``, ``h[N, E]``               (MonotonicNN.py:26-27)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

LEAKY_SLOPE = 0.01  # nn.LeakyReLU() default, UMNNMAF.py:250

HIDDEN_RELU = "relu"          # IntegrandNN, MonotonicNN.py:17-21
HIDDEN_LEAKY = "leaky_relu"   # IntegrandNetwork, UMNNMAF.py:246-251
OUT_ELU_PLUS_1 = "elu_plus_1"  # ELUPlus UMNNMAF.py:11-16, and nn.ELU()+1. MonotonicNN.py:23,27
OUT_SIGMOID = "sigmoid"        # dict_act_func["Sigmoid"], UMNNMAF.py:19


# ----------------------------------------------------------------------------
# a1: Clenshaw-Curtis nodes and weights
# ----------------------------------------------------------------------------
def cc_nodes_weights(Q: int) -> Tuple[np.ndarray, np.ndarray]:
    """Weights ``w[Q+1]`` and nodes ``t[Q+1]`` (both float32, node 0 = +1, node Q = -1).

    Follows compute_cc_weights, ParallelNeuralIntegral.py:14-34 (same as
    NeuralIntegral.py:14-34 and UMNNMAF.py:55-69): a float64 cosine matrix
    ``cos(k*i*pi/Q)`` whose first column is replaced by 1/2 and whose last column is
    halved, scaled by 2/Q, contracted against the even-moment vector
    ``2/(1-k^2)`` (k even, with the k=0 entry set to 1) and rounded to float32.
    """
    idx = np.arange(Q + 1, dtype=np.int64)
    cosmat = np.cos(np.outer(idx, idx).astype(np.float64) * math.pi / Q)  # [k, i]
    cosmat[:, 0] = 0.5
    cosmat[:, Q] = 0.5 * cosmat[:, Q]
    cosmat = cosmat * 2 / Q
    moments = np.zeros(Q + 1, dtype=np.float64)
    even = idx[idx % 2 == 0]
    moments[even] = 2.0 / (1.0 - even.astype(np.float64) ** 2)
    moments[0] = 1.0
    w = (cosmat.T @ moments.reshape(-1, 1)).reshape(-1)
    t = np.cos(idx.astype(np.float64) * math.pi / Q)
    return w.astype(np.float32), t.astype(np.float32)


# ----------------------------------------------------------------------------
# parameter handling
# ----------------------------------------------------------------------------
@dataclass
class MLPSpec:
    """Shape of an integrand MLP: widths = [n_in, H1, ..., HL, 1]."""
    widths: Tuple[int, ...]
    hidden_act: str = HIDDEN_LEAKY
    out_act: str = OUT_ELU_PLUS_1

    @property
    def n_params(self) -> int:
        return sum(i * o + o for i, o in zip(self.widths[:-1], self.widths[1:]))


def unpack_params(spec: MLPSpec, flat: np.ndarray) -> List[Tuple[np.ndarray, np.ndarray]]:
    """Split the flat vector into [(W_l [out, in], b_l [out])] views."""
    flat = np.asarray(flat).reshape(-1)
    assert flat.size == spec.n_params, (flat.size, spec.n_params)
    out, off = [], 0
    for n_in, n_out in zip(spec.widths[:-1], spec.widths[1:]):
        W = flat[off:off + n_in * n_out].reshape(n_out, n_in)
        off += n_in * n_out
        b = flat[off:off + n_out]
        off += n_out
        out.append((W, b))
    return out


def _hidden_act(v: np.ndarray, kind: str) -> np.ndarray:
    if kind == HIDDEN_LEAKY:
        return np.where(v > 0, v, v * v.dtype.type(LEAKY_SLOPE))
    if kind == HIDDEN_RELU:
        return np.maximum(v, v.dtype.type(0))
    raise ValueError(kind)


def _hidden_act_grad(v: np.ndarray, kind: str) -> np.ndarray:
    if kind == HIDDEN_LEAKY:
        return np.where(v > 0, v.dtype.type(1), v.dtype.type(LEAKY_SLOPE))
    if kind == HIDDEN_RELU:
        return (v > 0).astype(v.dtype)
    raise ValueError(kind)


def _out_act(v: np.ndarray, kind: str) -> np.ndarray:
    one = v.dtype.type(1)
    if kind == OUT_ELU_PLUS_1:
        # nn.ELU(alpha=1) then "+ 1." in the working precision: the negative branch is
        # expm1(v) + 1, which is quantised at 2^-24 in fp32 and exactly 0 for v <~ -17.3.
        return np.where(v > 0, v, np.expm1(np.minimum(v, 0))) + one
    if kind == OUT_SIGMOID:
        return one / (one + np.exp(-v))
    raise ValueError(kind)


def _out_act_grad(v: np.ndarray, kind: str) -> np.ndarray:
    one = v.dtype.type(1)
    if kind == OUT_ELU_PLUS_1:
        # torch elu_backward with is_result=False: grad * alpha * exp(x) for x <= 0
        return np.where(v > 0, one, np.exp(np.minimum(v, 0)))
    if kind == OUT_SIGMOID:
        s = one / (one + np.exp(-v))
        return s * (one - s)
    raise ValueError(kind)


def mlp_rows(spec: MLPSpec, flat: np.ndarray, rows: np.ndarray, keep: bool = False):
    """Apply the integrand MLP to ``rows[R, n_in]`` -> ``[R]``.

    The nn.Sequential of IntegrandNetwork (UMNNMAF.py:245-254) / IntegrandNN
    (MonotonicNN.py:15-24): Linear, act, ..., Linear, output activation.
    With ``keep`` the pre-activations are returned for the backward pass.
    """
    a = rows
    pre = []
    layers = unpack_params(spec, flat)
    for li, (W, b) in enumerate(layers):
        v = a @ W.T.astype(a.dtype) + b.astype(a.dtype)
        if keep:
            pre.append((a, v))
        a = _hidden_act(v, spec.hidden_act) if li < len(layers) - 1 else _out_act(v, spec.out_act)
    out = a.reshape(-1)
    return (out, pre) if keep else out


# ----------------------------------------------------------------------------
# a4 / a5: integrand networks
# ----------------------------------------------------------------------------
def slot_inputs_strided(x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """[N, D], [N, E*D] -> [N*D, 1+E]: row (n, d) = [x[n,d], h[n,0*D+d], ..., h[n,(E-1)*D+d]].

    IntegrandNetwork.forward, UMNNMAF.py:263-281: cat, view(N, 1+E, D), transpose(1,2).
    """
    N, D = x.shape
    E = h.shape[1] // D
    assert h.shape[1] == E * D
    stacked = np.concatenate([x[:, None, :], h.reshape(N, E, D)], axis=1)  # [N, 1+E, D]
    return np.ascontiguousarray(stacked.transpose(0, 2, 1)).reshape(N * D, 1 + E)


def integrand_network(spec: MLPSpec, flat: np.ndarray, x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """IntegrandNetwork.forward(x, h) -> [N, D], UMNNMAF.py:263-284."""
    N, D = x.shape
    return mlp_rows(spec, flat, slot_inputs_strided(x, h)).reshape(N, D)


def integrand_nn(spec: MLPSpec, flat: np.ndarray, x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """IntegrandNN.forward(x, h) -> [N, 1], MonotonicNN.py:26-27 (cat(x,h) -> net -> +1)."""
    return mlp_rows(spec, flat, np.concatenate([x, h], axis=1)).reshape(-1, 1)


def _integrand(spec, flat, x, h, layout):
    return integrand_network(spec, flat, x, h) if layout == "strided" else integrand_nn(spec, flat, x, h)


# ----------------------------------------------------------------------------
# a2 / a3 / a6: the integral (forward)
# ----------------------------------------------------------------------------
def _limits(x0: np.ndarray, x: np.ndarray, Q: int):
    """step = (x-x0)/Q (ParallelNeuralIntegral.py:102); xT = x0 + Q*step (:49)."""
    dt = x.dtype.type
    step = (x - x0) / dt(Q)
    xT = x0 + dt(Q) * step
    return xT


def integrate_parallel(spec: MLPSpec, flat: np.ndarray, x0: np.ndarray, x: np.ndarray, h: np.ndarray,
                       Q: int, layout: str = "strided", chunk: int = 0) -> np.ndarray:
    """ParallelNeuralIntegral.forward -> integrate(...), ParallelNeuralIntegral.py:37-65,99-108.

    All Q+1 nodes of every sample are laid out as rows (sample-major, node-minor),
    the integrand is applied once, the result is weighted and summed over the node
    axis, then scaled by (xT - x0)/2.  ``chunk`` > 0 processes that many samples at
    a time (memory only; results are per-sample independent).
    """
    w, t = cc_nodes_weights(Q)
    dt = x.dtype.type
    w = w.astype(x.dtype)
    t = t.astype(x.dtype)
    B, Dx = x.shape
    out = np.empty_like(x)
    step = chunk if chunk > 0 else B
    for s in range(0, B, step):
        sl = slice(s, min(B, s + step))
        x0c, hc = x0[sl], h[sl]
        xT = _limits(x0c, x[sl], Q)
        n = xT.shape[0]
        nodes = x0c[:, None, :] + (xT - x0c)[:, None, :] * (t[None, :, None] + dt(1)) / dt(2)  # [n, Q+1, Dx]
        h_rep = np.broadcast_to(hc[:, None, :], (n, Q + 1, hc.shape[1])).reshape(n * (Q + 1), -1)
        f = _integrand(spec, flat, nodes.reshape(n * (Q + 1), Dx), h_rep, layout).reshape(n, Q + 1, Dx)
        z = (f * w[None, :, None]).sum(axis=1)
        out[sl] = z * (xT - x0c) / dt(2)
    return out


def integrate_sequential(spec: MLPSpec, flat: np.ndarray, x0: np.ndarray, x: np.ndarray, h: np.ndarray,
                         Q: int, layout: str = "strided") -> np.ndarray:
    """NeuralIntegral.forward -> integrate(...), NeuralIntegral.py:37-66,80-88 (node loop)."""
    w, t = cc_nodes_weights(Q)
    dt = x.dtype.type
    xT = _limits(x0, x, Q)
    z = np.zeros_like(x)
    for i in range(Q + 1):
        xi = x0 + (xT - x0) * (dt(t[i]) + dt(1)) / dt(2)
        z = z + dt(w[i]) * _integrand(spec, flat, xi, h, layout)
    return z * (xT - x0) / dt(2)


# ----------------------------------------------------------------------------
# a7 / a8: backward (Leibniz rule)
# ----------------------------------------------------------------------------
def mlp_rows_vjp(spec: MLPSpec, flat: np.ndarray, rows: np.ndarray, cot: np.ndarray):
    """Vector-Jacobian product of ``mlp_rows``: returns (d_flat [P], d_rows [R, n_in]).

    Manual restatement of what torch.autograd.grad(f, params / h, cot) evaluates in
    computeIntegrand, ParallelNeuralIntegral.py:83-94.
    """
    _, pre = mlp_rows(spec, flat, rows, keep=True)
    layers = unpack_params(spec, flat)
    grads = []
    delta = cot.reshape(-1, 1).astype(rows.dtype)
    for li in range(len(layers) - 1, -1, -1):
        a_in, v = pre[li]
        if li == len(layers) - 1:
            delta = delta * _out_act_grad(v, spec.out_act)
        else:
            delta = delta * _hidden_act_grad(v, spec.hidden_act)
        gW = delta.T @ a_in
        gb = delta.sum(axis=0)
        grads.append((gW, gb))
        delta = delta @ layers[li][0].astype(rows.dtype)
    grads.reverse()
    d_flat = np.concatenate([np.concatenate([gW.reshape(-1), gb.reshape(-1)]) for gW, gb in grads])
    return d_flat, delta


def integral_backward(spec: MLPSpec, flat: np.ndarray, x0: np.ndarray, x: np.ndarray, h: np.ndarray,
                      grad_out: np.ndarray, Q: int, layout: str = "strided", chunk: int = 0):
    """ParallelNeuralIntegral.backward, ParallelNeuralIntegral.py:110-123 with
    integrate(compute_grad=True) :66-80 and computeIntegrand :83-94.

    Returns (d_x0, d_x, d_flat_params, d_h).
    """
    w, t = cc_nodes_weights(Q)
    dt = x.dtype.type
    w = w.astype(x.dtype)
    t = t.astype(x.dtype)
    B, Dx = x.shape
    d_flat = np.zeros(spec.n_params, dtype=x.dtype)
    d_h = np.zeros_like(h)
    step = chunk if chunk > 0 else B
    for s in range(0, B, step):
        sl = slice(s, min(B, s + step))
        x0c, hc, g = x0[sl], h[sl], grad_out[sl]
        xT = _limits(x0c, x[sl], Q)
        n = xT.shape[0]
        x_tot = g * (xT - x0c) / dt(2)                                   # :70
        cot = x_tot[:, None, :] * w[None, :, None]                       # :71  [n, Q+1, Dx]
        nodes = x0c[:, None, :] + (xT - x0c)[:, None, :] * (t[None, :, None] + dt(1)) / dt(2)
        if layout == "strided":
            E = hc.shape[1] // Dx
            h_rep = np.broadcast_to(hc[:, None, :], (n, Q + 1, hc.shape[1])).reshape(n * (Q + 1), -1)
            rows = slot_inputs_strided(nodes.reshape(n * (Q + 1), Dx), h_rep)     # [(n*(Q+1))*Dx, 1+E]
            gp, d_rows = mlp_rows_vjp(spec, flat, rows, cot.reshape(-1))
            # d_rows[:, 1:] is the gradient wrt the context of slot (n, i, d); sum over nodes i,
            # and map back to h[n, e*D + d]                               (:92-94)
            dctx = d_rows[:, 1:].reshape(n, Q + 1, Dx, E).sum(axis=1)     # [n, Dx, E]
            d_h[sl] = dctx.transpose(0, 2, 1).reshape(n, E * Dx)
        else:
            h_rep = np.broadcast_to(hc[:, None, :], (n, Q + 1, hc.shape[1])).reshape(n * (Q + 1), -1)
            rows = np.concatenate([nodes.reshape(n * (Q + 1), 1), h_rep], axis=1)
            gp, d_rows = mlp_rows_vjp(spec, flat, rows, cot.reshape(-1))
            d_h[sl] = d_rows[:, 1:].reshape(n, Q + 1, -1).sum(axis=1)
        d_flat += gp
    d_x = _integrand(spec, flat, x, h, layout) * grad_out                # :115,:123
    d_x0 = -_integrand(spec, flat, x0, h, layout) * grad_out             # :116,:123
    return d_x0, d_x, d_flat, d_h


def kink_margins(spec: MLPSpec, flat: np.ndarray, x0: np.ndarray, x: np.ndarray, h: np.ndarray, Q: int,
                 layout: str = "strided", chunk: int = 64) -> np.ndarray:
    """Per slot, the smallest RELATIVE distance of any hidden pre-activation from its kink over every row the
    backward evaluates (the Q+1 nodes, x and x0):   min |v| / (sum_k |W_k a_k| + |b|).

    Diagnostic of the gradient tests, not a restatement of a reference function: LeakyReLU/ReLU derivatives are
    discontinuous at 0, so an implementation whose pre-activations differ from another's by a relative rounding
    error eps may evaluate a unit on the other side of its kink wherever this margin is below eps -- the
    reference's own fp32 vs fp64 gradients differ by 1.2e-4 for that reason.  Slots whose margin
    is below an implementation's rounding level are "kink-ambiguous": their d_h / d_x may legitimately differ by
    one unit's contribution.  float64 arithmetic.  Returns [B, Dx].
    """
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    x0, x, h, flat = f64(x0), f64(x), f64(h), f64(flat)
    _, t = cc_nodes_weights(Q)
    t = t.astype(np.float64)
    B, Dx = x.shape
    layers = unpack_params(spec, flat)
    out = np.empty((B, Dx))
    for s in range(0, B, chunk):
        sl = slice(s, min(B, s + chunk))
        x0c, xc, hc = x0[sl], x[sl], h[sl]
        n = xc.shape[0]
        xT = _limits(x0c.astype(np.float32), xc.astype(np.float32), Q).astype(np.float64)
        nodes = x0c[:, None, :] + (xT - x0c)[:, None, :] * (t[None, :, None] + 1.0) / 2.0          # [n, Q+1, Dx]
        pts = np.concatenate([nodes, xc[:, None, :], x0c[:, None, :]], axis=1)                     # [n, Q+3, Dx]
        R = pts.shape[1]
        h_rep = np.broadcast_to(hc[:, None, :], (n, R, hc.shape[1])).reshape(n * R, -1)
        if layout == "strided":
            rows = slot_inputs_strided(pts.reshape(n * R, Dx), h_rep)                              # [(n*R)*Dx, 1+E]
        else:
            rows = np.concatenate([pts.reshape(n * R, 1), h_rep], axis=1)
        a = rows
        margin = np.full(a.shape[0], np.inf)
        for (W, b) in layers[:-1]:
            v = a @ W.T + b
            scale = np.abs(a) @ np.abs(W).T + np.abs(b)
            margin = np.minimum(margin, np.min(np.abs(v) / np.maximum(scale, 1e-300), axis=1))
            a = _hidden_act(v, spec.hidden_act)
        out[sl] = margin.reshape(n, R, Dx).min(axis=1)
    return out


# ----------------------------------------------------------------------------
# MADE conditioner (produces h); only what the flow path needs
# ----------------------------------------------------------------------------
def made_masks(nin: int, hidden: Sequence[int], nout: int) -> List[np.ndarray]:
    """Masks of MADE(natural_ordering=True, random=False), made.py:74-105.

    Returned in (in, out) orientation as built there; MaskedLinear stores the transpose.
    """
    L = len(hidden)
    m = {-1: np.arange(nin)}
    for l in range(L):
        m[l] = np.array([nin - 1 - (i % nin) for i in range(hidden[l])])
    masks = [m[l - 1][:, None] <= m[l][None, :] for l in range(L)]
    masks.append(m[L - 1][:, None] < m[-1][None, :])
    if nout > nin:
        masks[-1] = np.concatenate([masks[-1]] * (nout // nin), axis=1)
    return masks


def made_forward(nin: int, hidden: Sequence[int], nout: int, weights: Sequence[Tuple[np.ndarray, np.ndarray]],
                 x: np.ndarray) -> np.ndarray:
    """MADE.forward for nout != 2: masked Linear / ReLU stack, made.py:26-27,53-62,113-119."""
    masks = made_masks(nin, hidden, nout)
    a = x
    for li, ((W, b), mk) in enumerate(zip(weights, masks)):
        a = a @ (W * mk.T.astype(W.dtype)).T.astype(a.dtype) + b.astype(a.dtype)
        if li < len(weights) - 1:
            a = np.maximum(a, a.dtype.type(0))
    return a


# ----------------------------------------------------------------------------
# a9 / a10 / a11: one flow block and the stacked log-likelihood
# ----------------------------------------------------------------------------
def umnnmaf_forward(spec, flat, x, h, Q, scaling=None):
    """UMNNMAF.forward with x0 = 0: z = exp(scaling) * (integral + z0), UMNNMAF.py:76-134;
    z0 is the first E-chunk of h (:80)."""
    D = x.shape[1]
    z0 = h[:, :D]
    integ = integrate_parallel(spec, flat, np.zeros_like(x), x, h, Q, "strided")
    s = np.exp(np.zeros(D, dtype=x.dtype) if scaling is None else scaling.astype(x.dtype))
    return s[None, :] * (integ + z0)


def umnnmaf_log_jac(spec, flat, x, h, scaling=None):
    """UMNNMAF.compute_log_jac: log(f(x,h) + 1e-10) + scaling, UMNNMAF.py:136-139."""
    D = x.shape[1]
    sc = np.zeros(D, dtype=x.dtype) if scaling is None else scaling.astype(x.dtype)
    jac = integrand_network(spec, flat, x, h)
    return np.log(jac + x.dtype.type(1e-10)) + sc[None, :]


def flow_compute_ll(blocks, x, Q):
    """UMNNMAFFlow.compute_ll, UMNNMAFFlow.py:109-119.

    ``blocks`` is a list of dicts {spec, flat, made: (nin, hidden, nout, weights)}.
    Returns (ll [B], z [B, D]).
    """
    dt = x.dtype.type
    log_jac = np.zeros_like(x)
    z = x
    for blk in blocks:
        nin, hid, nout, mw = blk["made"]
        h = made_forward(nin, hid, nout, mw, x)
        z = umnnmaf_forward(blk["spec"], blk["flat"], x, h, Q)[:, ::-1]
        log_jac = log_jac + umnnmaf_log_jac(blk["spec"], blk["flat"], x, h)
        x = z
    z = z[:, ::-1]
    pi32 = np.float32(math.pi).astype(x.dtype)
    log_prob_gauss = (dt(-0.5) * (np.log(pi32 * dt(2)) + z ** 2)).sum(axis=1)
    return log_jac.sum(axis=1) + log_prob_gauss, np.ascontiguousarray(z)


# ----------------------------------------------------------------------------
# n3: sampling direction -- UMNNMAF.invert / UMNNMAFFlow.invert
# ----------------------------------------------------------------------------
def invert_grid(n_grid: int = 10) -> np.ndarray:
    """The 10 relative grid positions 0, 1/9, ..., 1 of UMNNMAF.py:183-186 (float32 like torch.arange)."""
    step = 1.0 / (n_grid - 1)
    return np.arange(0, 1 + step / 2, step).astype(np.float32)


def invert_bracket_step(integ, x_cur, grid, offset, scale, target, left=None, right=None):
    """One refinement round of UMNNMAF.invert for one dimension, UMNNMAF.py:213-231, float32.

    integ, x_cur [G, B] (integral from 0 to x_cur, grid-point major), offset/target [B], scale a scalar.
    Returns (left, right, x_next [G, B], x_mid [B]).  Faithful to the reference including the flat
    neighbour indexing of :226-227: the neighbours of grid point 0 / G-1 are read from the adjacent SAMPLE
    (index -1 wraps to the last element, the right neighbour wraps modulo B*G).
    """
    f32 = np.float32
    G, B = x_cur.shape
    z_est = f32(scale) * (offset[None, :].astype(f32) + integ.astype(f32))          # :213
    diff = np.abs(z_est - target[None, :])
    pos = np.argmin(diff, axis=0)                                                   # :218 (first minimum)
    mid = pos + np.arange(B) * G                                                    # :220
    z_val = z_est.T.reshape(-1)[mid]                                                # :221
    x_flat = x_cur.T.reshape(-1)                                                    # :222
    below = (z_val < target).astype(f32)                                            # :224
    lo = mid - 1                                                                    # :226 (negative wraps)
    hi = (mid + 1) % x_flat.shape[0]                                                # :227
    new_left = below * x_flat[mid] + (f32(1) - below) * x_flat[lo]                  # :229
    new_right = below * x_flat[hi] + (f32(1) - below) * x_flat[mid]                 # :230
    x_next = grid[:, None] * (new_right - new_left)[None, :] + new_left[None, :]    # :210 of the next round
    return new_left.astype(f32), new_right.astype(f32), x_next.astype(f32), x_flat[mid].astype(f32)


def umnnmaf_invert(blk, z, Q, n_iter=10, n_grid=10):
    """UMNNMAF.invert, UMNNMAF.py:182-232: per dimension j, `n_iter` rounds of the 10-point bracket
    refinement on [-50, 50]; every grid evaluation is a contiguous-context integral from 0."""
    f32 = np.float32
    B, D = z.shape
    spec, flat = blk["spec"], blk["flat"]
    nin, hid, nout, mw = blk["made"]
    grid = invert_grid(n_grid)
    x_inv = np.zeros((B, D), f32)
    for j in range(D):
        h_all = made_forward(nin, hid, nout, mw, x_inv)                             # :199
        offset = h_all[:, j]                                                        # :200 (first E-chunk)
        h_j = h_all[:, j::D]                                                        # :201-202
        h_rep = np.ascontiguousarray(np.broadcast_to(h_j[None], (n_grid,) + h_j.shape)).reshape(n_grid * B, -1)
        left = np.full(B, -50, f32)
        right = np.full(B, 50, f32)
        x_cur = (grid[:, None] * (right - left)[None, :] + left[None, :]).astype(f32)   # :210
        x_mid = None
        for _ in range(n_iter):
            xs = x_cur.reshape(-1, 1)
            integ = integrate_parallel(spec, flat, np.zeros_like(xs), xs, h_rep, Q, "contig").reshape(n_grid, B)
            left, right, x_cur, x_mid = invert_bracket_step(integ, x_cur, grid, offset, 1.0, z[:, j], left, right)
        x_inv[:, j] = x_mid                                                         # :231
    return x_inv


def flow_invert(blocks, z, Q, n_iter=10):
    """UMNNMAFFlow.invert, UMNNMAFFlow.py:78-90: blocks in reverse order, feature axis reversed around each."""
    z = z[:, ::-1]
    for blk in reversed(blocks):
        z = umnnmaf_invert(blk, np.ascontiguousarray(z[:, ::-1]), Q, n_iter)
    return z


# ----------------------------------------------------------------------------
# seeded synthetic data shared by the golden generator, the tests and the bench
# ----------------------------------------------------------------------------
def synth_params(spec: MLPSpec, seed: int, gain: float = 1.0) -> np.ndarray:
    """nn.Linear-like init U(-1/sqrt(in), 1/sqrt(in)) from the frozen RandomState stream.

    ``gain`` > 1 gives the "trained-like" variant (ELU goes negative, kinks exercised).
    """
    rng = np.random.RandomState(seed)
    parts = []
    for n_in, n_out in zip(spec.widths[:-1], spec.widths[1:]):
        k = 1.0 / math.sqrt(n_in)
        parts.append(rng.uniform(-k, k, size=n_in * n_out) * gain)
        parts.append(rng.uniform(-k, k, size=n_out) * gain)
    return np.concatenate(parts).astype(np.float32)


def synth_inputs(B: int, Dx: int, Hh: int, seed: int, x0_zero: bool = True):
    """x ~ 2*N(0,1), h ~ N(0,1), grad_out ~ N(0,1), x0 = 0 or 0.5*N(0,1)."""
    rng = np.random.RandomState(seed)
    x = (2.0 * rng.standard_normal((B, Dx))).astype(np.float32)
    h = rng.standard_normal((B, Hh)).astype(np.float32)
    g = rng.standard_normal((B, Dx)).astype(np.float32)
    x0 = np.zeros((B, Dx), np.float32) if x0_zero else (0.5 * rng.standard_normal((B, Dx))).astype(np.float32)
    return x0, x, h, g
