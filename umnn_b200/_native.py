"""ctypes binding of the C ABI in include/umnn_b200.h (umnn_b200/libumnn_b200.so).

The library is the product's only compute path for recognised integrands on CUDA tensors:
if it is missing or fails to load this module raises -- there is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# UMNN_B200_LIB: load an experiment variant of the library (scripts/build_variants.py) instead of the product build
LIB_PATH = os.environ.get("UMNN_B200_LIB") or os.path.join(_HERE, "libumnn_b200.so")

UMNN_ABI_VERSION = 1
UMNN_MAX_LAYERS = 8
UMNN_MAX_WIDTH = 256
UMNN_MAX_STEPS = 1024

LAYOUT_STRIDED_D, LAYOUT_CONTIG = 0, 1
ACT_RELU, ACT_LEAKY_RELU = 0, 1
OUT_ELU_PLUS_1, OUT_SIGMOID = 0, 1
PREC_FP32, PREC_BF16X3, PREC_AUTO, PREC_FP16X3 = 0, 1, 2, 3

EXPORTS = ("umnn_abi_version", "umnn_last_error", "umnn_cc_tables", "umnn_param_count",
           "umnn_packed_params_bytes", "umnn_packed_layout_id", "umnn_pack_params", "umnn_workspace_bytes", "umnn_cc_forward",
           "umnn_cc_backward", "umnn_cc_forward_host", "umnn_invert_bracket_step", "umnn_tc_forward_occupancy",
           "umnn_invert_workspace_bytes", "umnn_invert_dimension")


class Desc(ctypes.Structure):
    """Mirror of `struct umnn_desc`."""
    _fields_ = [("abi_version", ctypes.c_int32), ("layout", ctypes.c_int32), ("n_samples", ctypes.c_int64),
                ("n_dims", ctypes.c_int32), ("n_ctx", ctypes.c_int32), ("n_layers", ctypes.c_int32),
                ("widths", ctypes.c_int32 * (UMNN_MAX_LAYERS + 1)), ("hidden_act", ctypes.c_int32),
                ("out_act", ctypes.c_int32), ("nb_steps", ctypes.c_int32), ("precision", ctypes.c_int32)]


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"umnn_b200 native call failed (code {code}): {message}")
        self.code = code


_lock = threading.Lock()
_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared library once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m umnn_b200.build` (needs nvcc). "
                "umnn_b200 has no CPU/eager fallback for its CUDA path.")
        L = ctypes.CDLL(LIB_PATH)
        vp, fp, cp = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p
        dp = ctypes.POINTER(Desc)
        L.umnn_abi_version.restype = ctypes.c_int
        L.umnn_abi_version.argtypes = []
        L.umnn_last_error.restype = cp
        L.umnn_last_error.argtypes = []
        L.umnn_cc_tables.restype = ctypes.c_int
        L.umnn_cc_tables.argtypes = [ctypes.c_int32, fp, fp]
        L.umnn_param_count.restype = ctypes.c_int64
        L.umnn_param_count.argtypes = [dp]
        L.umnn_packed_params_bytes.restype = ctypes.c_size_t
        L.umnn_packed_params_bytes.argtypes = [dp]
        L.umnn_packed_layout_id.restype = ctypes.c_uint64
        L.umnn_packed_layout_id.argtypes = [dp]
        L.umnn_pack_params.restype = ctypes.c_int
        L.umnn_pack_params.argtypes = [dp, fp, vp, vp]
        L.umnn_workspace_bytes.restype = ctypes.c_size_t
        L.umnn_workspace_bytes.argtypes = [dp, ctypes.c_int32]
        L.umnn_cc_forward.restype = ctypes.c_int
        L.umnn_cc_forward.argtypes = [dp, fp, fp, fp, vp, fp, fp, fp, fp, fp, vp, ctypes.c_size_t, vp]
        L.umnn_cc_backward.restype = ctypes.c_int
        L.umnn_cc_backward.argtypes = [dp, fp, fp, fp, vp, fp, fp, fp, fp, fp, fp, fp, fp, vp, ctypes.c_size_t, vp]
        L.umnn_cc_forward_host.restype = ctypes.c_int
        L.umnn_cc_forward_host.argtypes = [dp, fp, fp, fp, fp, fp, fp, fp, ctypes.c_int32]
        L.umnn_tc_forward_occupancy.restype = ctypes.c_int
        L.umnn_tc_forward_occupancy.argtypes = [dp, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                                ctypes.POINTER(ctypes.c_int32)]
        i64 = ctypes.c_int64
        L.umnn_invert_bracket_step.restype = ctypes.c_int
        L.umnn_invert_bracket_step.argtypes = [i64, ctypes.c_int32, fp, fp, fp, fp, i64, fp, fp, i64, fp, fp, i64, fp, fp,
                                               i64, vp]
        L.umnn_invert_workspace_bytes.restype = ctypes.c_size_t
        L.umnn_invert_workspace_bytes.argtypes = [dp, i64, ctypes.c_int32]
        L.umnn_invert_dimension.restype = ctypes.c_int
        L.umnn_invert_dimension.argtypes = [dp, vp, fp, fp, i64, ctypes.c_int32, ctypes.c_int32, fp, fp, fp, fp, i64,
                                            ctypes.c_float, ctypes.c_float, fp, i64, vp, ctypes.c_size_t, vp]
        if L.umnn_abi_version() != UMNN_ABI_VERSION:
            raise RuntimeError(f"libumnn_b200.so ABI {L.umnn_abi_version()} != binding {UMNN_ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().umnn_last_error()
        raise NativeError(rc, msg.decode("utf-8", "replace") if msg else "")


def make_desc(layout: int, n_samples: int, n_dims: int, n_ctx: int, widths, hidden_act: int, out_act: int,
              nb_steps: int, precision: int = PREC_AUTO) -> Desc:
    d = Desc()
    d.abi_version = UMNN_ABI_VERSION
    d.layout = layout
    d.n_samples = n_samples
    d.n_dims = n_dims
    d.n_ctx = n_ctx
    widths = list(widths)
    if len(widths) - 1 > UMNN_MAX_LAYERS:
        raise ValueError(f"integrand has {len(widths) - 1} Linear layers; the kernel supports {UMNN_MAX_LAYERS}")
    d.n_layers = len(widths) - 1
    for i, w in enumerate(widths):
        d.widths[i] = int(w)
    d.hidden_act = hidden_act
    d.out_act = out_act
    d.nb_steps = nb_steps
    d.precision = precision
    return d
