"""ctypes binding of oracle/libumnn_oracle.so (test infrastructure only; see umnn_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

from . import umnn_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libumnn_oracle.so")
ORC_MAX_LAYERS = 16


class _Mlp(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int), ("widths", ctypes.c_int * (ORC_MAX_LAYERS + 1)),
                ("hidden_act", ctypes.c_int), ("out_act", ctypes.c_int)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "umnn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_cc_forward.restype = ctypes.c_int
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def cc_forward(spec: orc.MLPSpec, flat, x0, x, h, Q, layout="strided", want_f=True, n_threads=0):
    """Returns (integral, f_at_x, f_at_x0) computed by the C restatement."""
    m = _Mlp()
    m.n_layers = len(spec.widths) - 1
    for i, wd in enumerate(spec.widths):
        m.widths[i] = wd
    m.hidden_act = 1 if spec.hidden_act == orc.HIDDEN_LEAKY else 0
    m.out_act = 0 if spec.out_act == orc.OUT_ELU_PLUS_1 else 1
    flat = np.ascontiguousarray(flat, np.float32)
    x0 = np.ascontiguousarray(x0, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    h = np.ascontiguousarray(h, np.float32)
    B, D = x.shape
    E = spec.widths[0] - 1
    w, t = orc.cc_nodes_weights(Q)
    out = np.empty_like(x)
    fx = np.empty_like(x) if want_f else None
    fx0 = np.empty_like(x) if want_f else None
    rc = lib().orc_cc_forward(ctypes.byref(m), _fp(flat), B, D, E, 0 if layout == "strided" else 1, Q,
                              _fp(t), _fp(w), _fp(x0), _fp(x), _fp(h), _fp(out), _fp(fx), _fp(fx0), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"orc_cc_forward failed rc={rc}")
    return out, fx, fx0


def max_threads() -> int:
    return int(lib().orc_max_threads())
