"""Torch-tensor front end of the C ABI: packs parameters, builds descriptors, launches on the
current CUDA stream.  Tensors are passed as raw device pointers; torch is only the allocator and
the stream provider here.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import Optional

import torch

from . import _native
from .networks import KernelSpec
from .quadrature import device_tables

_PRECISION_ENV = {"fp32": _native.PREC_FP32, "bf16x3": _native.PREC_BF16X3, "fp16x3": _native.PREC_FP16X3,
                  "auto": _native.PREC_AUTO}


_ENV_DATA = getattr(os.environ, "_data", None)      # the raw bytes -> bytes dict behind os.environ (CPython, POSIX)


def _env(name: bytes) -> Optional[bytes]:
    """Current value of an environment switch.  os.environ.get() costs ~3 us per lookup (key encoding, value
    decoding), which at seven lookups per call was most of a small launch's host time; the underlying dict is read
    directly instead -- still live, so switches changed at run time (tests, A/B scripts) are seen."""
    if _ENV_DATA is not None:
        return _ENV_DATA.get(name)
    v = os.environ.get(name.decode())
    return None if v is None else v.encode()


_PRECISION_ENV_B = {k.encode(): v for k, v in _PRECISION_ENV.items()}


def default_precision() -> int:
    """UMNN_B200_PRECISION = fp32 | bf16x3 | fp16x3 | auto (default auto)."""
    v = _env(b"UMNN_B200_PRECISION")
    if v is None:
        return _native.PREC_AUTO
    return _PRECISION_ENV_B[v.lower()]


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _as_f32c(t: torch.Tensor) -> torch.Tensor:
    return t if (t.is_contiguous() and t.dtype == torch.float32) else t.contiguous().float()


_repack_always = False


@contextlib.contextmanager
def repack_every_call():
    """Inside this context the packed-parameter cache is bypassed, so the (tiny) pack launches are issued on
    every call -- used while capturing CUDA graphs, whose replays must read the live parameter values."""
    global _repack_always
    prev, _repack_always = _repack_always, True
    try:
        yield
    finally:
        _repack_always = prev


def flat_parameters(spec: KernelSpec) -> torch.Tensor:
    return torch.cat([p.detach().reshape(-1) for p in spec.param_list])


def packed_parameters(spec: KernelSpec, desc: _native.Desc, device: torch.device) -> torch.Tensor:
    """Device-side packed parameter block for this launch.

    Re-packed on every call (a few tiny launches) while the integrand is in training mode.  In eval
    mode the block is cached ON the first Linear module (so it dies with the network) and reused
    while (a) the library reports the same packed LAYOUT for the descriptor -- `umnn_packed_layout_id` covers the
    resolved precision and every offset inside the block, which depend on nb_steps through the shared-memory fit,
    so changing Q (set_steps_nb, invert with another Q) can never reuse a block in the wrong format -- and
    (b) every parameter keeps its storage address and autograd version: in-place updates (optimizer steps,
    load_state_dict, force_lipschitz) bump the version and invalidate it.  Updates made through `.data`
    (p.data.copy_/mul_/clamp_) do NOT bump the version: call `invalidate_packed(integrand)` after those, or
    keep the network in training mode.
    """
    L = _native.lib()
    owner = spec.linears[0]
    stamp = None
    layout_id = _desc_info(spec, desc, "layout")
    if layout_id == 0:
        msg = L.umnn_last_error()
        raise _native.NativeError(_native_err_unsupported, msg.decode("utf-8", "replace") if msg else "")
    if not owner.training and not _repack_always:
        stamp = (layout_id, device.index, *[v for p in spec.param_list for v in (p.data_ptr(), p._version)])
        hit = owner.__dict__.get("_umnn_packed", {}).get(layout_id)
        if hit is not None and hit[0] == stamp:
            return hit[1]
    nbytes = L.umnn_packed_params_bytes(desc)
    if nbytes == 0:
        msg = L.umnn_last_error()
        raise _native.NativeError(_native_err_unsupported, msg.decode("utf-8", "replace") if msg else "")
    for p in spec.param_list:
        if p.device != device or p.dtype != torch.float32:
            raise ValueError("umnn_b200: the integrand's parameters must be float32 on the same CUDA device as x")
    packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
    flat = _as_f32c(flat_parameters(spec))
    stream = torch.cuda.current_stream(device).cuda_stream
    with torch.cuda.device(device):
        _native.check(L.umnn_pack_params(desc, flat.data_ptr(), packed.data_ptr(), stream))
    if stamp is not None:
        cache = owner.__dict__.setdefault("_umnn_packed", {})
        if len(cache) >= 8:          # a sweep over many layouts must not pin device memory forever
            cache.clear()
        cache[layout_id] = (stamp, packed)
    else:
        owner.__dict__.pop("_umnn_packed", None)
    return packed


def invalidate_packed(integrand) -> None:
    """Drop the cached packed-parameter blocks of an integrand (IntegrandNetwork, IntegrandNN, ContiguousIntegrand or
    anything with kernel_spec()).  Needed only after parameter updates that bypass autograd's version counter
    (writes through `.data`)."""
    get_spec = getattr(integrand, "kernel_spec", None)
    spec = get_spec() if callable(get_spec) else None
    if spec is not None:
        spec.linears[0].__dict__.pop("_umnn_packed", None)


_native_err_unsupported = -4


def make_desc(spec: KernelSpec, x: torch.Tensor, nb_steps: int, precision: Optional[int] = None) -> _native.Desc:
    """Descriptor of this launch.  Descriptors (and what the library derives from them on the host: packed layout id,
    workspace sizes) are memoised on the KernelSpec, which the integrand modules cache in turn: a small-batch call
    must not spend more time describing the launch than the GPU spends running it."""
    B, Dx = x.shape
    prec = default_precision() if precision is None else precision
    key = (B, Dx, int(nb_steps), prec)
    hit = spec.desc_cache.get(key)
    if hit is not None:
        return hit
    if spec.layout == _native.LAYOUT_STRIDED_D and Dx != spec.n_dims:
        raise ValueError(f"x has {Dx} columns but the integrand network was built for {spec.n_dims}")
    if spec.layout == _native.LAYOUT_CONTIG and Dx != 1:
        raise ValueError(f"contiguous-context integrands integrate a single variable (x is [B, 1]); got [B, {Dx}]")
    d = _native.make_desc(spec.layout, B, Dx, spec.n_ctx, spec.widths, spec.hidden_act, spec.out_act, int(nb_steps), prec)
    d._umnn_key = key
    if len(spec.desc_cache) > 64:
        spec.desc_cache.clear()
    spec.desc_cache[key] = d
    return d


def _desc_info(spec: KernelSpec, desc: _native.Desc, what: str):
    """Host-side facts the library derives from a descriptor, memoised per descriptor object: 'layout' (packed layout
    id), 'ws0' / 'ws1' (forward / backward workspace bytes).  They depend on the descriptor and on process-wide
    environment switches read at call time (UMNN_B200_TC_SEGMENTS, UMNN_B200_BWD_PANELS), so the memo is keyed on
    those too."""
    L = _native.lib()
    dkey = getattr(desc, "_umnn_key", None)
    if dkey is None:        # a descriptor built elsewhere: ask the library directly
        return int(L.umnn_packed_layout_id(desc)) if what == "layout" else int(L.umnn_workspace_bytes(desc, 0 if what == "ws0" else 1))
    env = (_env(b"UMNN_B200_TC_SEGMENTS"), _env(b"UMNN_B200_BWD_PANELS"), _env(b"UMNN_B200_TC_NARROW"), _env(b"UMNN_B200_WGRAD_KBS"),
           _env(b"UMNN_B200_BWD_TILES"))
    key = (dkey, what, env)
    hit = spec.info_cache.get(key)
    if hit is not None:
        return hit
    if what == "layout":
        v = int(L.umnn_packed_layout_id(desc))
    elif what == "ws0":
        v = int(L.umnn_workspace_bytes(desc, 0))
    else:
        v = int(L.umnn_workspace_bytes(desc, 1))
    if len(spec.info_cache) > 256:
        spec.info_cache.clear()
    spec.info_cache[key] = v
    return v


_flag_ws = {}


def _forward_workspace(nbytes: int, dev: torch.device, stream: int) -> Optional[torch.Tensor]:
    """The forward's 256-byte workspace (the overflow flag of a guarded FP16X3 call), one persistent buffer per
    (device, stream): calls on one stream are ordered, so they can share it, and a small-batch call saves an
    allocation.  Buffers handed out while a CUDA graph is being captured come from the capture's own pool."""
    if nbytes == 0:
        return None
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, dtype=torch.uint8, device=dev)
    key = (dev.index, stream)
    buf = _flag_ws.get(key)
    if buf is None or buf.numel() < nbytes:
        if len(_flag_ws) > 64:
            _flag_ws.clear()
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        _flag_ws[key] = buf
    return buf


def cc_forward(spec: KernelSpec, x0: Optional[torch.Tensor], x: torch.Tensor, h: torch.Tensor, nb_steps: int,
               want_fx: bool = False, want_fx0: bool = False, precision: Optional[int] = None,
               out: Optional[torch.Tensor] = None, fx_out: Optional[torch.Tensor] = None):
    """One fused launch: (integral, f(x,h) or None, f(x0,h) or None), each [B, Dx].

    Replaces integrate(...) + IntegrandNetwork.forward of the reference (see include/umnn_b200.h).
    `out` / `fx_out`: optional preallocated contiguous float32 results on x's device (written in place).
    """
    L = _native.lib()
    x = _as_f32c(x)
    h = _as_f32c(h)
    x0 = None if x0 is None else _as_f32c(x0)
    desc = make_desc(spec, x, nb_steps, precision)
    B, Dx = x.shape
    if h.shape[0] != B or h.shape[1] != spec.n_ctx * Dx:
        raise ValueError(f"h must be [{B}, {spec.n_ctx * Dx}] for this integrand, got {tuple(h.shape)}")
    if x0 is not None and x0.shape != x.shape:
        raise ValueError("x0 and x must have the same shape")
    dev = x.device
    for t in (out, fx_out):
        if t is not None and not (t.shape == x.shape and t.dtype == torch.float32 and t.is_contiguous() and t.device == dev):
            raise ValueError("cc_forward: out / fx_out must be contiguous float32 tensors shaped like x on x's device")
    if out is None:
        out = torch.empty_like(x)
    fx = (fx_out if fx_out is not None else torch.empty_like(x)) if want_fx else None
    fx0 = torch.empty_like(x) if want_fx0 else None
    if B == 0:
        return out, fx, fx0
    packed = packed_parameters(spec, desc, dev)
    w, t = device_tables(nb_steps, dev)
    ws_bytes = _desc_info(spec, desc, "ws0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = _forward_workspace(ws_bytes, dev, stream)
    args = (desc, _ptr(x0), x.data_ptr(), h.data_ptr(), packed.data_ptr(), t.data_ptr(), w.data_ptr(), out.data_ptr(),
            _ptr(fx), _ptr(fx0), _ptr(ws), ws_bytes, stream)
    if torch.cuda.current_device() == dev.index:
        rc = L.umnn_cc_forward(*args)
    else:
        with torch.cuda.device(dev):
            rc = L.umnn_cc_forward(*args)
    if rc != 0:
        _native.check(rc)
    return out, fx, fx0


def _raw_stream(dev_index: int) -> int:
    """cudaStream_t of torch's current stream (torch.cuda.current_stream(dev).cuda_stream builds a Stream object first)."""
    get = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if get is not None:
        return get(dev_index)
    return torch.cuda.current_stream(dev_index).cuda_stream


class PreparedIntegral:
    """The fused forward for ONE call shape with everything that does not change from call to call resolved once:
    descriptor, packed parameter block, quadrature tables, workspace size, argument checks.  Calling it allocates the
    results, reads torch's current stream and makes the one native call -- for a small batch (MonotonicNN at B = 100:
    ~15 us of GPU work) the generic entry's per-call Python work was most of the time.

        prep = umnn_b200.prepare_integral(net, batch=100, nb_steps=50, want_fx=True)
        z, fx, _ = prep(x, h)              # same results, bit for bit, as cc_integrate(net, None, x, h, 50, want_fx=True)

    The parameter values are SNAPSHOTTED at preparation (like a frozen / scripted module): call `refresh()` after
    updating the integrand.  x0 = None means zeros (UMNNMAF.forward, MonotonicNN.forward)."""

    def __init__(self, spec: KernelSpec, batch: int, nb_steps: int, device: torch.device, n_dims: Optional[int] = None,
                 want_fx: bool = False, want_fx0: bool = False, precision: Optional[int] = None):
        self._lib = _native.lib()
        self.spec = spec
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("prepare_integral needs a CUDA device")
        Dx = (spec.n_dims if spec.layout == _native.LAYOUT_STRIDED_D else 1) if n_dims is None else int(n_dims)
        self.x_shape = (int(batch), Dx)
        self.h_shape = (int(batch), spec.n_ctx * Dx)
        self.nb_steps = int(nb_steps)
        self.want_fx, self.want_fx0 = bool(want_fx), bool(want_fx0)
        self._precision = precision
        self._fn = self._lib.umnn_cc_forward
        self.refresh()

    def refresh(self) -> None:
        """Re-read the integrand's parameters (and the process-wide switches) into the prepared launch."""
        probe = torch.empty(self.x_shape, device=self.device)
        why = self.spec.supported()
        if why is not None:
            raise ValueError(f"prepare_integral: the integrand is outside the fused kernel's limits ({why})")
        self._desc = make_desc(self.spec, probe, self.nb_steps, self._precision)
        global _repack_always
        prev, _repack_always = _repack_always, True          # a private block: later optimiser steps must not alias it
        try:
            self._packed = packed_parameters(self.spec, self._desc, self.device)
        finally:
            _repack_always = prev
        self._w, self._t = device_tables(self.nb_steps, self.device)
        self._ws_bytes = _desc_info(self.spec, self._desc, "ws0")
        self._packed_ptr, self._t_ptr, self._w_ptr = self._packed.data_ptr(), self._t.data_ptr(), self._w.data_ptr()
        self._desc_ref = ctypes.byref(self._desc)

    def __call__(self, x: torch.Tensor, h: torch.Tensor, x0: Optional[torch.Tensor] = None):
        dev = self.device
        if not (x.shape == self.x_shape and h.shape == self.h_shape and x.dtype is torch.float32 and h.dtype is torch.float32
                and x.device == dev and h.device == dev and x.is_contiguous() and h.is_contiguous()):
            raise ValueError(f"PreparedIntegral: expected contiguous float32 x {self.x_shape} and h {self.h_shape} on {dev}")
        if x0 is not None and not (x0.shape == self.x_shape and x0.dtype is torch.float32 and x0.device == dev and x0.is_contiguous()):
            raise ValueError("PreparedIntegral: x0 must look like x")
        out = torch.empty_like(x)
        fx = torch.empty_like(x) if self.want_fx else None
        fx0 = torch.empty_like(x) if self.want_fx0 else None
        if self.x_shape[0] == 0:
            return out, fx, fx0
        stream = _raw_stream(dev.index)
        ws = _forward_workspace(self._ws_bytes, dev, stream)
        args = (self._desc_ref, None if x0 is None else x0.data_ptr(), x.data_ptr(), h.data_ptr(), self._packed_ptr, self._t_ptr,
                self._w_ptr, out.data_ptr(), None if fx is None else fx.data_ptr(), None if fx0 is None else fx0.data_ptr(),
                None if ws is None else ws.data_ptr(), self._ws_bytes, stream)
        if torch.cuda.current_device() == dev.index:
            rc = self._fn(*args)
        else:
            with torch.cuda.device(dev):
                rc = self._fn(*args)
        if rc != 0:
            _native.check(rc)
        return out, fx, fx0


def cc_forward_host(spec_widths, layout, hidden_act, out_act, flat_params, x0, x, h, nb_steps, want_fx=False,
                    want_fx0=False, device: int = 0, precision: int = _native.PREC_AUTO):
    """Host-buffer entry (numpy arrays in, numpy arrays out) through umnn_cc_forward_host: the call a
    non-PyTorch user of the C ABI makes; H2D and D2H copies happen inside."""
    import numpy as np
    L = _native.lib()
    x = np.ascontiguousarray(x, np.float32)
    h = np.ascontiguousarray(h, np.float32)
    flat_params = np.ascontiguousarray(flat_params, np.float32)
    B, Dx = x.shape
    desc = _native.make_desc(layout, B, Dx, spec_widths[0] - 1, spec_widths, hidden_act, out_act, int(nb_steps),
                             precision)
    out = np.empty_like(x)
    fx = np.empty_like(x) if want_fx else None
    fx0 = np.empty_like(x) if want_fx0 else None
    x0p = None
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, np.float32)
        x0p = x0.ctypes.data
    _native.check(L.umnn_cc_forward_host(desc, x0p, x.ctypes.data, h.ctypes.data, flat_params.ctypes.data,
                                         out.ctypes.data, None if fx is None else fx.ctypes.data,
                                         None if fx0 is None else fx0.ctypes.data, device))
    return out, fx, fx0


def backward_precision(spec: KernelSpec, x: torch.Tensor, nb_steps: int) -> Optional[int]:
    """Which native backward serves this shape: a tensor-core precision (three tensor-core passes; PREC_AUTO lets the
    library pick its default operand split, fp16x3 with a guarded bf16 re-run), PREC_FP32 (fused FFMA kernel) or None
    (neither fits: the caller uses torch ops on the device).

    UMNN_B200_BACKWARD = auto (default: tensor cores, else FFMA) | fp16x3 | bf16x3 | fp32 | torch.
    """
    mode = (_env(b"UMNN_B200_BACKWARD") or b"auto").decode().lower()
    if mode == "torch":
        return None
    if mode == "bf16x3":
        cands = [_native.PREC_BF16X3]
    elif mode == "fp16x3":
        cands = [_native.PREC_FP16X3]
    elif mode == "fp32" or default_precision() == _native.PREC_FP32:
        cands = [_native.PREC_FP32]
    else:
        cands = [default_precision(), _native.PREC_FP32]
    if x.shape[0] == 0:
        return cands[-1]
    for prec in cands:
        if _desc_info(spec, make_desc(spec, x, nb_steps, prec), "ws1") > 0:
            return prec
    return None


def backward_supported(spec: KernelSpec, x: torch.Tensor, nb_steps: int) -> bool:
    return backward_precision(spec, x, nb_steps) is not None


def cc_backward(spec: KernelSpec, x0: Optional[torch.Tensor], x: torch.Tensor, h: torch.Tensor,
                grad_out: torch.Tensor, nb_steps: int, grad_fx: Optional[torch.Tensor] = None,
                need_x0: bool = True, need_x: bool = True, need_h: bool = True, need_params: bool = True,
                precision: Optional[int] = None):
    """Fused backward through the C ABI: (d_x0, d_x, d_flat_params, d_h); entries not requested are None.

    Replaces ParallelNeuralIntegral.backward / integrate(compute_grad=True) / computeIntegrand of the
    reference (see include/umnn_b200.h).  `grad_fx` is an optional cotangent of f(x, h) (the Jacobian point).
    """
    L = _native.lib()
    x = _as_f32c(x)
    h = _as_f32c(h)
    grad_out = _as_f32c(grad_out)
    x0 = None if x0 is None else _as_f32c(x0)
    grad_fx = None if grad_fx is None else _as_f32c(grad_fx)
    if precision is None:
        precision = backward_precision(spec, x, nb_steps)
        if precision is None:
            raise _native.NativeError(_native_err_unsupported, "no native backward serves this shape")
    desc = make_desc(spec, x, nb_steps, precision)
    dev = x.device
    d_x0 = torch.empty_like(x) if need_x0 else None
    d_x = torch.empty_like(x) if need_x else None
    d_h = torch.empty_like(h) if need_h else None
    n_params = int(L.umnn_param_count(desc))
    d_flat = torch.empty(n_params, dtype=torch.float32, device=dev) if need_params else None
    if x.shape[0] == 0:
        if d_flat is not None:
            d_flat.zero_()
        return d_x0, d_x, d_flat, d_h
    packed = packed_parameters(spec, desc, dev)
    w, t = device_tables(nb_steps, dev)
    ws_bytes = _desc_info(spec, desc, "ws1")
    if ws_bytes == 0:
        L.umnn_workspace_bytes(desc, 1)          # refresh the thread-local error string
        msg = L.umnn_last_error()
        raise _native.NativeError(_native_err_unsupported, msg.decode("utf-8", "replace") if msg else "")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(L.umnn_cc_backward(desc, _ptr(x0), x.data_ptr(), h.data_ptr(), packed.data_ptr(), t.data_ptr(),
                                         w.data_ptr(), grad_out.data_ptr(), _ptr(grad_fx), _ptr(d_x0), _ptr(d_x),
                                         _ptr(d_h), _ptr(d_flat), ws.data_ptr(), ws_bytes, stream))
    return d_x0, d_x, d_flat, d_h


def invert_bracket_step(integ: Optional[torch.Tensor], x_grid: Optional[torch.Tensor], grid: torch.Tensor,
                        offset: Optional[torch.Tensor], scale: Optional[torch.Tensor], target: Optional[torch.Tensor],
                        left: torch.Tensor, right: torch.Tensor, x_grid_next: torch.Tensor,
                        x_mid: Optional[torch.Tensor]) -> None:
    """One round of UMNNMAF.invert's bracket refinement for one dimension (UMNNMAF.py:210,213-231) as one launch.

    integ, x_grid, x_grid_next: contiguous [G, B]; offset, target, left, right, x_mid: 1-D views of length B (any
    stride, e.g. columns of [B, D] tensors); scale: 1-element tensor.  integ=None only lays the first grid.
    left/right/x_mid/x_grid_next are written in place.
    """
    L = _native.lib()
    G, B = x_grid_next.shape
    for t in (integ, x_grid, x_grid_next, grid):
        if t is not None and not (t.is_contiguous() and t.dtype == torch.float32):
            raise ValueError("invert_bracket_step: grids must be contiguous float32")
    if left.stride(0) != right.stride(0):
        raise ValueError("invert_bracket_step: left and right must share a stride")
    dev = x_grid_next.device
    stream = torch.cuda.current_stream(dev).cuda_stream

    def stride(t):
        return 0 if t is None else t.stride(0)

    with torch.cuda.device(dev):
        _native.check(L.umnn_invert_bracket_step(B, G, _ptr(integ), _ptr(x_grid), grid.data_ptr(), _ptr(offset),
                                                 stride(offset), _ptr(scale), _ptr(target), stride(target),
                                                 left.data_ptr(), right.data_ptr(), left.stride(0),
                                                 x_grid_next.data_ptr(), _ptr(x_mid), stride(x_mid), stream))
