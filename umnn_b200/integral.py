"""The neural integral: autograd Functions, the device-agnostic torch path and the kernel dispatcher.

Behavioural spec (AWehenkel/UMNN @ 59118c14), written fresh:
  * models/UMNN/ParallelNeuralIntegral.py  -- integrate :37-80, computeIntegrand :83-94,
    ParallelNeuralIntegral :97-123
  * models/UMNN/NeuralIntegral.py          -- integrate :37-66, computeIntegrand :69-75,
    NeuralIntegral :78-99

Two routes serve the same maths:
  kernel route   CUDA float32 tensors + a recognised integrand (IntegrandNetwork, IntegrandNN,
                 ContiguousIntegrand) + inv_f == False + not tracing  ->  ONE fused sm_100a launch
                 through the C ABI (umnn_b200/_native.py).  If the native library is missing, or the
                 recognised integrand is outside the kernels' limits (width > 256, > 8 Linear layers, a
                 backward neither native kernel can hold), this RAISES; it never degrades to torch ops on
                 its own.  UMNN_B200_ALLOW_TORCH_ROUTE=1 turns those raises into the torch route.
  torch route    everything else the public API allows: arbitrary callables / lambdas, CPU or MPS
                 tensors, float64, TorchScript tracing, inv_f=True.  Same results as the reference.
                 `with torch_route():` forces it for the calls inside (benchmarks of the reference's
                 algorithm as PyTorch ops on the same device; never a default).
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional, Tuple

import torch

from . import kernel
from .networks import _flatten  # noqa: F401  (re-exported: the reference modules expose it)
from .quadrature import compute_cc_weights, device_tables

# rows (= samples x nodes x dims) the torch route evaluates at once during a backward pass;
# above this the batch is processed in chunks (gradients are sums over independent samples)
_TORCH_ROUTE_MAX_ROWS = int(os.environ.get("UMNN_B200_TORCH_ROUTE_MAX_ROWS", 1 << 22))


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def _tables_for(x0: torch.Tensor, nb_steps: int, cc_weights=None, steps=None):
    """[Q+1] weight / node vectors on x0's device."""
    if cc_weights is None or steps is None:
        if x0.is_cuda:
            return device_tables(nb_steps, x0.device)
        cc_weights, steps = compute_cc_weights(nb_steps)
    return cc_weights.to(x0.device).view(-1), steps.to(x0.device).view(-1)


def _node_rows(x0, xT, h, nodes):
    """Abscissae [B*(Q+1), Dx] and replicated context [B*(Q+1), Hh], sample-major / node-minor."""
    n = nodes.shape[0]
    span = xT - x0
    X = x0.unsqueeze(1) + span.unsqueeze(1) * (nodes.view(1, n, 1) + 1) / 2
    H = h.unsqueeze(1).expand(-1, n, -1)
    return X.reshape(-1, x0.shape[1]), H.reshape(-1, h.shape[1])


def _params_of(integrand):
    get = getattr(integrand, "parameters", None)
    return list(get()) if callable(get) else []


def _call(integrand, x, h):
    fwd = getattr(integrand, "forward", None)
    return fwd(x, h) if callable(fwd) else integrand(x, h)


# --------------------------------------------------------------------------------------------------
# torch route, vectorised over the nodes ("parallel")
# --------------------------------------------------------------------------------------------------
def integrate(x0, nb_steps, step_sizes, integrand, h, compute_grad=False, x_tot=None, inv_f=False,
              cc_weights=None, steps=None):
    """Clenshaw-Curtis quadrature of `integrand` from x0 to x0 + nb_steps*step_sizes.

    compute_grad=False: returns the integral [B, Dx].
    compute_grad=True : returns (grad wrt flattened integrand parameters, grad wrt h) for the
                        cotangent x_tot [B, Dx] of the integral.
    """
    w, t = _tables_for(x0, nb_steps, cc_weights, steps)
    xT = x0 + nb_steps * step_sizes
    B, n = x0.shape[0], nb_steps + 1
    X, H = _node_rows(x0, xT, h, t)
    if not compute_grad:
        f = integrand(X, H)
        if inv_f:
            f = 1 / f
        z = (f.view(B, n, -1) * w.view(1, n, 1)).sum(1)
        return z * (xT - x0) / 2
    cot = (x_tot * (xT - x0) / 2).unsqueeze(1) * w.view(1, n, 1)
    return computeIntegrand(X, H, integrand, cot.reshape(B * n, -1), n, inv_f=inv_f)


def computeIntegrand(x, h, integrand, x_tot, nb_steps, inv_f=False):
    """VJP of the integrand rows: (flat parameter gradient, context gradient summed over nodes)."""
    h = h.detach().requires_grad_(True) if not h.requires_grad else h
    with torch.enable_grad():
        f = _call(integrand, x, h)
        if inv_f:
            f = 1 / f
        params = _params_of(integrand)
        if params:
            g_param = _flatten(torch.autograd.grad(f, params, x_tot, create_graph=True, retain_graph=True))
        else:
            g_param = None
        g_h = _flatten(torch.autograd.grad(f, h, x_tot))
    return g_param, g_h.view(int(x.shape[0] / nb_steps), nb_steps, -1).sum(1)


def _integrate_grads_chunked(x0, x, integrand, h, nb_steps, grad_output, inv_f):
    """integrate(compute_grad=True) over batch chunks so the autograd graph stays bounded."""
    B = x0.shape[0]
    rows_per_sample = (nb_steps + 1) * max(1, x0.shape[1])
    chunk = max(1, _TORCH_ROUTE_MAX_ROWS // rows_per_sample)
    if B <= chunk:
        return integrate(x0, nb_steps, (x - x0) / nb_steps, integrand, h, True, grad_output, inv_f)
    g_param_total, g_h_parts = None, []
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        gp, gh = integrate(x0[s:e], nb_steps, (x[s:e] - x0[s:e]) / nb_steps, integrand, h[s:e], True,
                           grad_output[s:e], inv_f)
        if not torch.is_grad_enabled():
            # plain backward: drop each chunk's graph as soon as it is summed (bounded memory).  Under
            # create_graph=True (double backward) the graphs are kept, as the reference keeps them.
            gp = None if gp is None else gp.detach()
            gh = gh.detach()
        if gp is not None:
            g_param_total = gp if g_param_total is None else g_param_total + gp
        g_h_parts.append(gh)
    return g_param_total, torch.cat(g_h_parts, 0)


# --------------------------------------------------------------------------------------------------
# torch route, one node at a time ("sequential", low memory)
# --------------------------------------------------------------------------------------------------
def integrate_sequential(x0, nb_steps, step_sizes, integrand, h, compute_grad=False, x_tot=None):
    w, t = _tables_for(x0, nb_steps)
    xT = x0 + nb_steps * step_sizes
    span = xT - x0
    if compute_grad:
        g_param, g_h = 0., 0.
        cot = x_tot * span / 2
        if not h.requires_grad:
            h = h.detach().requires_grad_(True)
        for i in range(nb_steps + 1):
            xi = x0 + span * (t[i] + 1) / 2
            dg_param, dg_h = computeIntegrand_sequential(xi, h, integrand, cot)
            if dg_param is not None:
                g_param = g_param + w[i] * dg_param
            g_h = g_h + w[i] * dg_h
        return (g_param if torch.is_tensor(g_param) else None), g_h
    z = 0.
    for i in range(nb_steps + 1):
        xi = x0 + span * (t[i] + 1) / 2
        z = z + w[i] * integrand(xi, h)
    return z * span / 2


def computeIntegrand_sequential(x, h, integrand, x_tot):
    with torch.enable_grad():
        f = _call(integrand, x, h)
        params = _params_of(integrand)
        g_param = (_flatten(torch.autograd.grad(f, params, x_tot, create_graph=True, retain_graph=True))
                   if params else None)
        g_h = _flatten(torch.autograd.grad(f, h, x_tot))
    return g_param, g_h


# --------------------------------------------------------------------------------------------------
# dispatcher
# --------------------------------------------------------------------------------------------------
_forced_torch_route = 0


@contextlib.contextmanager
def torch_route():
    """Inside this context every integral takes the torch route (the reference's algorithm as PyTorch ops on whatever
    device the tensors live on).  A measurement / A-B tool: nothing in the product enters it."""
    global _forced_torch_route
    _forced_torch_route += 1
    try:
        yield
    finally:
        _forced_torch_route -= 1


class UnsupportedIntegrandError(ValueError):
    """A recognised integrand on CUDA float32 tensors that the native kernels cannot serve."""


def _allow_torch_route() -> bool:
    return os.environ.get("UMNN_B200_ALLOW_TORCH_ROUTE", "0") == "1"


def kernel_route(integrand, x0, x, h, inv_f=False):
    """The KernelSpec if this call is served by the fused CUDA kernel, else None (torch route).

    A recognised integrand on CUDA float32 tensors never leaves the kernel route silently: outside the kernels'
    limits this raises UnsupportedIntegrandError unless UMNN_B200_ALLOW_TORCH_ROUTE=1."""
    if inv_f or not (torch.is_tensor(x) and x.is_cuda):
        return None
    if torch.jit.is_tracing() or torch.jit.is_scripting():
        return None
    if _forced_torch_route:
        return None
    get_spec = getattr(integrand, "kernel_spec", None)
    if not callable(get_spec):
        return None
    if not (x.dtype == torch.float32 and x0.dtype == torch.float32 and h.dtype == torch.float32):
        return None
    spec = get_spec()
    if spec is None:
        return None
    why = spec.supported()
    if why is not None:
        if _allow_torch_route():
            return None
        raise UnsupportedIntegrandError(
            f"umnn_b200: {type(integrand).__name__} is outside the fused kernel's limits ({why}). "
            "Set UMNN_B200_ALLOW_TORCH_ROUTE=1 to evaluate it with PyTorch ops instead.")
    dev = x.device
    if not (x0.is_cuda and h.is_cuda and x0.device == dev and h.device == dev):
        raise ValueError("umnn_b200: x0, x and h must live on the same CUDA device")
    # the first parameter is checked on every call, all of them whenever the parameters are (re)packed
    # (kernel.packed_parameters): a per-parameter device comparison costs more than the rest of a small launch
    p0 = spec.param_list[0]
    if p0.device != dev or p0.dtype != torch.float32:
        raise ValueError("umnn_b200: the integrand's parameters must be float32 on the same CUDA device as x")
    return spec


def integral_nograd(x0, x, integrand, h, nb_steps, parallel=True, inv_f=False, cc_weights=None, steps=None):
    """Integral value only (no autograd graph): kernel route when eligible, else the torch route."""
    spec = kernel_route(integrand, x0, x, h, inv_f)
    if spec is not None:
        return kernel.cc_forward(spec, x0, x, h, nb_steps)[0]
    if parallel:
        return integrate(x0, nb_steps, (x - x0) / nb_steps, integrand, h, False, None, inv_f, cc_weights, steps)
    return integrate_sequential(x0, nb_steps, (x - x0) / nb_steps, integrand, h, False)


def _forward_common(ctx, x0, x, integrand, h, nb_steps, inv_f, parallel):
    spec = kernel_route(integrand, x0, x, h, inv_f)
    ctx.integrand = integrand
    ctx.nb_steps = nb_steps
    ctx.inv_f = inv_f
    ctx.kernel_spec = spec
    with torch.no_grad():
        if spec is not None:
            need = ctx.needs_input_grad
            # the fused backward re-evaluates f(x), f(x0) itself; only when it cannot serve the shape are
            # the Leibniz terms taken from extra rows of the forward launch
            ctx.bwd_precision = kernel.backward_precision(spec, x, nb_steps) if any(need) else None
            ctx.native_bwd = ctx.bwd_precision is not None
            if any(need) and not ctx.native_bwd and not _allow_torch_route():
                raise UnsupportedIntegrandError(
                    "umnn_b200: neither native backward can hold this integrand shape in shared memory; set "
                    "UMNN_B200_ALLOW_TORCH_ROUTE=1 to differentiate it with PyTorch ops instead.")
            want_fx = bool(need[1]) and not ctx.native_bwd
            want_fx0 = bool(need[0]) and not ctx.native_bwd
            out, fx, fx0 = kernel.cc_forward(spec, x0, x, h, nb_steps, want_fx=want_fx, want_fx0=want_fx0)
            ctx.has_fx, ctx.has_fx0 = fx is not None, fx0 is not None
            extra = [t for t in (fx, fx0) if t is not None]
            ctx.save_for_backward(x0.clone(), x.clone(), h, *extra)
            return out
        if parallel:
            out = integrate(x0, nb_steps, (x - x0) / nb_steps, integrand, h, False, inv_f=inv_f)
        else:
            out = integrate_sequential(x0, nb_steps, (x - x0) / nb_steps, integrand, h, False)
        ctx.save_for_backward(x0.clone(), x.clone(), h)
    return out


def _backward_common(ctx, grad_output, parallel):
    saved = ctx.saved_tensors
    x0, x, h = saved[0], saved[1], saved[2]
    integrand, nb_steps, inv_f, spec = ctx.integrand, ctx.nb_steps, ctx.inv_f, ctx.kernel_spec
    need = ctx.needs_input_grad
    if spec is not None and ctx.native_bwd:
        d_x0, d_x, d_flat, d_h = kernel.cc_backward(spec, x0, x, h, grad_output, nb_steps, need_x0=bool(need[0]),
                                                    need_x=bool(need[1]), need_h=bool(need[4]),
                                                    need_params=bool(need[3]), precision=ctx.bwd_precision)
        return d_x0, d_x, d_flat, d_h
    if spec is not None:
        grad_output = grad_output.contiguous()
        extra = list(saved[3:])
        fx = extra.pop(0) if ctx.has_fx else None
        fx0 = extra.pop(0) if ctx.has_fx0 else None
        d_x = fx * grad_output if fx is not None else None
        d_x0 = -fx0 * grad_output if fx0 is not None else None
        d_flat = d_h = None
        if need[3] or need[4]:
            # UMNN_B200_ALLOW_TORCH_ROUTE=1 only (the forward raised otherwise): shapes the fused backward cannot
            # hold in shared memory go through torch ops on the same CUDA device, batch-chunked
            d_flat, d_h = _integrate_grads_chunked(x0, x, integrand, h, nb_steps, grad_output, False)
            d_h = d_h.view(h.shape)
            if not need[3]:
                d_flat = None
        return d_x0, d_x, d_flat, d_h
    # torch route: Leibniz rule for the limits, weighted VJP for parameters and context
    if parallel:
        g_param, g_h = _integrate_grads_chunked(x0, x, integrand, h, nb_steps, grad_output, inv_f)
    else:
        g_param, g_h = integrate_sequential(x0, nb_steps, (x - x0) / nb_steps, integrand, h, True, grad_output)
    d_x = integrand(x, h) * grad_output
    d_x0 = -integrand(x0, h) * grad_output
    if not need[3]:
        g_param = None
    return d_x0, d_x, g_param, g_h.view(h.shape)


class FusedIntegralAndPoint(torch.autograd.Function):
    """apply(x0, x, integrand, flat_params, h, nb_steps) -> (integral, f(x, h)) from ONE fused launch.

    Kernel route only.  Serves UMNNMAF.forward + compute_log_jac together (UMNNMAF.py:76-139): the Jacobian
    point is an extra row of the forward launch, and its cotangent enters the fused backward as
    `grad_f_at_x` instead of a second pass through the integrand network.
    """

    @staticmethod
    def forward(ctx, x0, x, integrand, flat_params, h, nb_steps):
        spec = kernel_route(integrand, x0, x, h, False)
        if spec is None:
            raise ValueError("FusedIntegralAndPoint needs the kernel route")
        ctx.kernel_spec, ctx.nb_steps = spec, nb_steps
        ctx.bwd_precision = kernel.backward_precision(spec, x, nb_steps) if any(ctx.needs_input_grad) else None
        if any(ctx.needs_input_grad) and ctx.bwd_precision is None:
            raise ValueError("FusedIntegralAndPoint: no native backward serves this shape")
        with torch.no_grad():
            out, fx, _ = kernel.cc_forward(spec, x0, x, h, nb_steps, want_fx=True)
            ctx.save_for_backward(x0.clone(), x.clone(), h)
        return out, fx

    @staticmethod
    def backward(ctx, grad_out, grad_fx):
        x0, x, h = ctx.saved_tensors
        need = ctx.needs_input_grad
        if grad_out is None:
            grad_out = torch.zeros_like(x)
        d_x0, d_x, d_flat, d_h = kernel.cc_backward(ctx.kernel_spec, x0, x, h, grad_out, ctx.nb_steps, grad_fx=grad_fx,
                                                    need_x0=bool(need[0]), need_x=bool(need[1]), need_h=bool(need[4]),
                                                    need_params=bool(need[3]), precision=ctx.bwd_precision)
        return d_x0, d_x, None, d_flat, d_h, None


def fused_point_available(integrand, x0, x, h, nb_steps, needs_grad):
    """True if FusedIntegralAndPoint can serve this call (kernel route, and a native backward when needed)."""
    spec = kernel_route(integrand, x0, x, h, False)
    if spec is None:
        return False
    return (not needs_grad) or kernel.backward_precision(spec, x, nb_steps) is not None


class ParallelNeuralIntegral(torch.autograd.Function):
    """apply(x0, x, integrand, flat_params, h, nb_steps=20, inv_f=False) -> integral [B, Dx]."""

    @staticmethod
    def forward(ctx, x0, x, integrand, flat_params, h, nb_steps=20, inv_f=False):
        return _forward_common(ctx, x0, x, integrand, h, nb_steps, inv_f, parallel=True)

    @staticmethod
    def backward(ctx, grad_output):
        d_x0, d_x, d_flat, d_h = _backward_common(ctx, grad_output, parallel=True)
        return d_x0, d_x, None, d_flat, d_h, None, None


class NeuralIntegral(torch.autograd.Function):
    """apply(x0, x, integrand, flat_params, h, nb_steps=20) -> integral [B, Dx] (node-by-node on the torch route)."""

    @staticmethod
    def forward(ctx, x0, x, integrand, flat_params, h, nb_steps=20):
        return _forward_common(ctx, x0, x, integrand, h, nb_steps, False, parallel=False)

    @staticmethod
    def backward(ctx, grad_output):
        d_x0, d_x, d_flat, d_h = _backward_common(ctx, grad_output, parallel=False)
        return d_x0, d_x, None, d_flat, d_h, None


def cc_integrate(integrand, x0, x, h, nb_steps, want_fx=False, want_fx0=False, precision=None):
    """Functional, value-only entry of the fused kernel: (integral, f(x,h) | None, f(x0,h) | None).

    One launch gives the integral of UMNNMAF.forward (UMNNMAF.py:76-134) and the Jacobian point of
    UMNNMAF.compute_log_jac (:136-139).  CUDA float32 tensors and a recognised integrand are
    required -- this entry never takes the torch route.  `precision`: None (UMNN_B200_PRECISION or auto),
    or one of umnn_b200._native.PREC_FP32 / PREC_BF16X3 / PREC_AUTO.
    """
    if x0 is None:
        x0_probe = x
    else:
        x0_probe = x0
    spec = kernel_route(integrand, x0_probe, x, h, False)
    if spec is None:
        raise ValueError("cc_integrate needs CUDA float32 tensors and a recognised integrand "
                         "(IntegrandNetwork, IntegrandNN, ContiguousIntegrand) within the kernel's limits")
    with torch.no_grad():
        return kernel.cc_forward(spec, x0, x, h, nb_steps, want_fx=want_fx, want_fx0=want_fx0, precision=precision)


def prepare_integral(integrand, batch, nb_steps, n_dims=None, want_fx=False, want_fx0=False, device=None, precision=None):
    """A `kernel.PreparedIntegral` for repeated value-only calls of ONE shape (serving loops, MonotonicNN at small
    batch): `prep(x, h, x0=None)` returns what `cc_integrate(integrand, x0, x, h, nb_steps, ...)` returns, bit for bit,
    with a fraction of the per-call host work.  Parameters are snapshotted: `prep.refresh()` after updating them."""
    get_spec = getattr(integrand, "kernel_spec", None)
    spec = get_spec() if callable(get_spec) else None
    if spec is None:
        raise ValueError("prepare_integral needs a recognised integrand (IntegrandNetwork, IntegrandNN, ContiguousIntegrand)")
    if device is None:
        device = spec.param_list[0].device
    return kernel.PreparedIntegral(spec, batch, nb_steps, device, n_dims=n_dims, want_fx=want_fx, want_fx0=want_fx0,
                                   precision=precision)


def cc_integrate_host(integrand, x_host, h_host, nb_steps, want_fx=False, out=None, fx_out=None, device=None,
                      chunks=None, precision=None):
    """Host-buffer entry of the fused kernel: (pinned) host tensors in, host tensors out -> (integral, f(x,h) | None).

    x0 = 0 as in UMNNMAF.forward (UMNNMAF.py:77).  The batch is cut into `chunks` pieces whose host->device copy,
    fused launch and device->host copy run on three streams, so only the first piece's upload and the last
    piece's download are exposed.  Results are complete once the CURRENT stream of `device` is synchronised
    (the call itself does not block).  `out` / `fx_out`: optional preallocated (pinned) result tensors.
    """
    params = _params_of(integrand)
    if device is None:
        if not params:
            raise ValueError("cc_integrate_host: pass device= for a parameter-free integrand")
        device = params[0].device
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("cc_integrate_host needs the integrand on a CUDA device")
    B = x_host.shape[0]
    if out is None:
        out = torch.empty(x_host.shape, dtype=torch.float32, pin_memory=True)
    if want_fx and fx_out is None:
        fx_out = torch.empty(x_host.shape, dtype=torch.float32, pin_memory=True)
    if B == 0:
        return out, (fx_out if want_fx else None)
    if chunks is None:
        chunks = 4 if (x_host.numel() + h_host.numel()) * 4 >= (32 << 20) else 1
    chunks = max(1, min(int(chunks), B))
    bounds = [B * c // chunks for c in range(chunks + 1)]
    if chunks == 1:
        # small batches: everything on the current stream (no side streams, events or staging bookkeeping -- at a few
        # thousand rows the host-side calls are the cost, not the copies)
        with torch.no_grad():
            xd = x_host.to(device, non_blocking=True)
            hd = h_host.to(device, non_blocking=True)
            spec = kernel_route(integrand, xd, xd, hd, False)
            if spec is None:
                raise ValueError("cc_integrate_host needs float32 inputs and a recognised integrand within the "
                                 "kernel's limits")
            od, fd, _ = kernel.cc_forward(spec, None, xd, hd, nb_steps, want_fx=want_fx, precision=precision)
            out.copy_(od, non_blocking=True)
            if want_fx:
                fx_out.copy_(fd, non_blocking=True)
        return out, (fx_out if want_fx else None)
    main = torch.cuda.current_stream(device)
    s_in, s_out = _side_streams(device)
    with torch.no_grad():
        # device staging for the whole batch, allocated on the CURRENT stream (no per-chunk allocations on the side
        # streams: their pools cannot recycle blocks while copies are in flight and fall back to cudaMalloc)
        xd = torch.empty(x_host.shape, dtype=torch.float32, device=device)
        hd = torch.empty(h_host.shape, dtype=torch.float32, device=device)
        od = torch.empty(x_host.shape, dtype=torch.float32, device=device)
        fd = torch.empty(x_host.shape, dtype=torch.float32, device=device) if want_fx else None
        spec = kernel_route(integrand, xd, xd, hd, False)
        if spec is None:
            raise ValueError("cc_integrate_host needs float32 inputs and a recognised integrand within the "
                             "kernel's limits")
        s_in.wait_stream(main)       # whatever used these blocks before is ordered on the current stream
        s_out.wait_stream(main)
        uploaded = []
        with torch.cuda.stream(s_in):
            for c in range(chunks):
                lo, hi = bounds[c], bounds[c + 1]
                xd[lo:hi].copy_(x_host[lo:hi], non_blocking=True)
                hd[lo:hi].copy_(h_host[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
                uploaded.append(ev)
        for c in range(chunks):
            lo, hi = bounds[c], bounds[c + 1]
            main.wait_event(uploaded[c])
            kernel.cc_forward(spec, None, xd[lo:hi], hd[lo:hi], nb_steps, want_fx=want_fx, precision=precision,
                              out=od[lo:hi], fx_out=None if fd is None else fd[lo:hi])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                out[lo:hi].copy_(od[lo:hi], non_blocking=True)
                if want_fx:
                    fx_out[lo:hi].copy_(fd[lo:hi], non_blocking=True)
        # the staging tensors are released when this function returns: everything that touches them is ordered
        # before whatever the current stream does next
        main.wait_stream(s_in)
        main.wait_stream(s_out)
    return out, (fx_out if want_fx else None)


_side = {}


def _side_streams(device):
    key = str(device)
    if key not in _side:
        _side[key] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _side[key]
