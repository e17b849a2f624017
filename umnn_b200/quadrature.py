"""Clenshaw-Curtis nodes and weights (host tables + per-device cache).

Behavioural spec: compute_cc_weights, models/UMNN/ParallelNeuralIntegral.py:14-34 -- float64
arithmetic, rounded to float32, returned as two `[Q+1, 1]` CPU tensors (weights, nodes) with node 0
at +1 (the upper limit) and node Q at -1.  Unlike the reference, which re-uploads the cached CPU
tensors on every `integrate` call (:47), the device copies are cached per (Q, device).
"""
from __future__ import annotations

import math
import threading
from typing import Dict, Tuple

import numpy as np
import torch

_host_cache: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
_device_cache: Dict[Tuple[int, str], Tuple[torch.Tensor, torch.Tensor]] = {}
_lock = threading.Lock()


def cc_tables_f64(nb_steps: int) -> Tuple[np.ndarray, np.ndarray]:
    """float64 weights w[Q+1] and nodes t[Q+1]."""
    if nb_steps < 1:
        raise ValueError("nb_steps must be >= 1")
    Q = int(nb_steps)
    i = np.arange(Q + 1, dtype=np.float64)
    # basis[k, i] = (2/Q) cos(k i pi / Q), with the i = 0 column pinned to 1/Q and the i = Q column halved
    basis = np.cos(np.outer(i, i) * math.pi / Q)
    basis[:, 0] = 0.5
    basis[:, Q] *= 0.5
    basis *= 2.0 / Q
    # integrals of the even Chebyshev polynomials over [-1, 1]: 2/(1-k^2); the k = 0 term enters with 1
    mom = np.zeros(Q + 1, dtype=np.float64)
    k_even = np.arange(0, Q + 1, 2)
    mom[k_even] = 2.0 / (1.0 - k_even.astype(np.float64) ** 2)
    mom[0] = 1.0
    w = basis.T @ mom.reshape(-1, 1)
    return w.reshape(-1), np.cos(i * math.pi / Q)


def compute_cc_weights(nb_steps: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(cc_weights [Q+1,1], steps [Q+1,1]) float32 CPU tensors, cached per nb_steps."""
    hit = _host_cache.get(nb_steps)
    if hit is not None:
        return hit
    w, t = cc_tables_f64(nb_steps)
    out = (torch.from_numpy(w.astype(np.float32)).view(-1, 1), torch.from_numpy(t.astype(np.float32)).view(-1, 1))
    with _lock:
        _host_cache[nb_steps] = out
    return out


def device_tables(nb_steps: int, device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
    """(weights [Q+1], nodes [Q+1]) float32 tensors resident on `device` (uploaded once)."""
    key = (int(nb_steps), str(device))
    hit = _device_cache.get(key)
    if hit is not None:
        return hit
    w, t = compute_cc_weights(nb_steps)
    out = (w.view(-1).to(device).contiguous(), t.view(-1).to(device).contiguous())
    with _lock:
        _device_cache[key] = out
    return out
