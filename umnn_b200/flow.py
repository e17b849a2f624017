"""UMNN-MAF flow blocks, the stacked flow and the 1-D monotone regressor (host-side mirror).

Behavioural spec (AWehenkel/UMNN @ 59118c14), written fresh:
  * EmbeddingNetwork, UMNNMAF      models/UMNN/UMNNMAF.py:37-232,304-329
  * ListModule, UMNNMAFFlow        models/UMNN/UMNNMAFFlow.py:8-151
  * MonotonicNN                    models/UMNN/MonotonicNN.py:29-54
Constructor signatures, method names (including the misspelt ones the experiment drivers call,
SURVEY.md 2.4) and state-dict keys are those of the reference, so its checkpoints load and its
drivers run unchanged.  The integral itself goes through umnn_b200.integral (fused CUDA kernel
for CUDA float32 tensors).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _native, kernel
from .integral import (FusedIntegralAndPoint, NeuralIntegral, ParallelNeuralIntegral, fused_point_available,
                       integral_nograd, kernel_route)
from .networks import (ConditionnalMADE, ContiguousIntegrand, IntegrandNN, IntegrandNetwork, MADE, _flatten, _mlp)
from .quadrature import compute_cc_weights, device_tables


class EmbeddingNetwork(nn.Module):
    """MADE conditioner + the shared integrand network of one flow block."""

    def __init__(self, in_d, hiddens_embedding=[50, 50, 50, 50], hiddens_integrand=[50, 50, 50, 50], out_made=1,
                 cond_in=0, act_func='ELU', device="cpu"):
        super().__init__()
        self.m_embeding = None
        self.device = device
        self.in_d = in_d
        if cond_in > 0:
            self.made = ConditionnalMADE(in_d, cond_in, hiddens_embedding, (in_d + cond_in) * out_made, num_masks=1,
                                         natural_ordering=True).to(device)
        else:
            self.made = MADE(in_d, hiddens_embedding, in_d * out_made, num_masks=1, natural_ordering=True).to(device)
        self.parallel_nets = IntegrandNetwork(in_d, 1 + out_made, hiddens_integrand, 1, act_func=act_func,
                                              device=device)

    def to(self, device):
        self.device = device
        self.made.to(device)
        self.parallel_nets.to(device)
        return self

    def make_embeding(self, x_made, context=None):
        self.m_embeding = self.made.forward(x_made, context)
        return self.m_embeding

    def embedding_of_dim(self, x_made, j, context=None):
        """make_embeding(x_made, context)[:, j::in_d] -- the E context values of dimension j ([B, E]; column 0 is the
        offset z0 of UMNNMAF.py:203) -- from the conditioner's hidden layers plus E rows of its last masked layer
        instead of all E*in_d of them."""
        if isinstance(self.made, ConditionnalMADE):
            hid = self.made.hidden_forward(x_made, context)
        else:
            hid = self.made.hidden_forward(x_made)
        return self.made.output_columns(hid, j, self.in_d)

    def forward(self, x_t):
        return self.parallel_nets.forward(x_t, self.m_embeding)


class UMNNMAF(nn.Module):
    """One autoregressive monotone block: z_d = exp(s_d) * (int_0^{x_d} f(t, h_d(x_<d)) dt + z0_d)."""

    def __init__(self, net, input_size, nb_steps=100, device="cpu", solver="CC"):
        super().__init__()
        self.net = net.to(device)
        self.device = device
        self.input_size = input_size
        self.nb_steps = nb_steps
        self.solver = solver
        self.register_buffer("pi", torch.tensor(math.pi))
        w, t = compute_cc_weights(nb_steps)
        self.register_buffer("cc_weights", w.clone())
        self.register_buffer("cc_steps", t.clone())
        # log-scale per dimension, frozen at 0 as in the reference (UMNNMAF.py:53)
        self.scaling = nn.Parameter(torch.zeros(input_size, device=self.device), requires_grad=False)

    def to(self, device):
        self.device = device
        super().to(device)
        return self

    # ----- the integral ---------------------------------------------------------------------------
    def _integral(self, x0, x, h):
        """int_{x0}^{x} f(t, h) dt with the block's solver; None for an unknown solver."""
        if self.solver not in ("CC", "CCParallel"):
            return None
        parallel = self.solver == "CCParallel"
        nets = self.net.parallel_nets
        tracing = torch.jit.is_tracing() or torch.jit.is_scripting()
        if tracing or ((not self.training) and (not x.requires_grad)):
            # the reference integrates directly here (UMNNMAF.py:89-106): plain torch ops, which stay differentiable
            # with respect to the integrand's parameters and h.  The fused kernel returns values only, so it serves
            # this branch only when nothing can ask for a gradient; otherwise the autograd Function below does.
            wants_grad = torch.is_grad_enabled() and (h.requires_grad or any(p.requires_grad for p in nets.parameters()))
            on_kernel = (not tracing) and kernel_route(nets, x0, x, h) is not None
            if not (on_kernel and wants_grad):
                tables_ok = parallel and self.cc_weights.shape[0] == self.nb_steps + 1
                return integral_nograd(x0, x, nets, h, self.nb_steps, parallel=parallel,
                                       cc_weights=self.cc_weights if tables_ok else None,
                                       steps=self.cc_steps if tables_ok else None)
        fn = ParallelNeuralIntegral if parallel else NeuralIntegral
        return fn.apply(x0, x, nets, _flatten(nets.parameters()), h, self.nb_steps)

    def _scale(self, batch):
        return self.scaling.unsqueeze(0).expand(batch, -1)

    def forward(self, x, method=None, x0=None, context=None):
        x0 = x0.to(x.device) if x0 is not None else torch.zeros(x.shape).to(x.device)
        h = self.net.make_embeding(x, context)
        z0 = h.view(h.shape[0], -1, x.shape[1])[:, 0, :]
        integral = self._integral(x0, x, h)
        if integral is None:
            return None
        return torch.exp(self._scale(x.shape[0])) * (integral + z0)

    def _log_jac_from_embedding(self, x):
        jac = self.net.forward(x)
        return torch.log(jac + 1e-10) + self._scale(x.shape[0])

    def compute_log_jac(self, x, context=None):
        self.net.make_embeding(x, context)
        return self._log_jac_from_embedding(x)

    def compute_log_jac_bis(self, x, context=None):
        z = self.forward(x, context=context)
        return z, self._log_jac_from_embedding(x)

    def forward_and_log_jac(self, x, context=None):
        """(z, log|dz/dx|) with ONE conditioner pass (the reference recomputes MADE, UMNNMAF.py:79,137) and, on
        the kernel route, ONE fused launch for the integral and the Jacobian point f(x, h)."""
        if self.solver in ("CC", "CCParallel") and x.is_cuda and not (torch.jit.is_tracing() or torch.jit.is_scripting()):
            x0 = torch.zeros_like(x)
            h = self.net.make_embeding(x, context)
            nets = self.net.parallel_nets
            needs_grad = torch.is_grad_enabled() and (x.requires_grad or h.requires_grad or
                                                      any(p.requires_grad for p in nets.parameters()))
            if fused_point_available(nets, x0, x, h, self.nb_steps, needs_grad):
                z0 = h.view(h.shape[0], -1, x.shape[1])[:, 0, :]
                integral, jac = FusedIntegralAndPoint.apply(x0, x, nets, _flatten(nets.parameters()), h, self.nb_steps)
                scale = self._scale(x.shape[0])
                return torch.exp(scale) * (integral + z0), torch.log(jac + 1e-10) + scale
        return self.compute_log_jac_bis(x, context=context)

    def compute_ll(self, x, context=None):
        z = self.forward(x, context=context)
        jac = self.net.forward(x)
        z.clamp_(-10., 10.)
        log_prob_gauss = -.5 * (torch.log(self.pi * 2) + z ** 2).sum(1)
        ll = log_prob_gauss + torch.log(jac + 1e-10).sum(1) + self._scale(x.shape[0]).sum(1)
        return ll, z

    computeLL = compute_ll  # name used by compute_bpp, UMNNMAF.py:166

    def compute_ll_bis(self, x, context=None):
        z = self.forward(x, context=context)
        ll = self._log_jac_from_embedding(x)
        z.clamp_(-10., 10.)
        return ll, z

    def compute_bpp(self, x, alpha=1e-6, context=None):
        d = x.shape[1]
        ll, z = self.computeLL(x, context=context)
        bpp = -ll / (d * np.log(2)) - np.log2(1 - 2 * alpha) + 8 \
            + 1 / d * (torch.log2(torch.sigmoid(x)) + torch.log2(1 - torch.sigmoid(x))).sum(1)
        z.clamp_(-10., 10.)
        return bpp, ll, z

    def set_steps_nb(self, nb_steps):
        """Change Q.  The registered node/weight buffers are rebuilt too (the reference leaves them
        stale), keeping their device."""
        self.nb_steps = nb_steps
        w, t = compute_cc_weights(nb_steps)
        self.cc_weights = w.clone().to(self.cc_weights.device)
        self.cc_steps = t.clone().to(self.cc_steps.device)

    def compute_lipschitz(self, nb_iter=10):
        return self.net.parallel_nets.computeLipshitz(nb_iter)

    def force_lipschitz(self, L=1.5):
        self.net.parallel_nets.force_lipschitz(L)

    computeLipshitz = compute_lipschitz
    forceLipshitz = force_lipschitz

    def invert(self, z, iter=10, context=None):
        """x with forward(x) ~= z: per dimension, `iter` rounds of a 10-point bracket refinement
        (UMNNMAF.py:182-232), every grid evaluation being one contiguous-context integral."""
        n_grid = 10
        B, D = z.shape
        dev = self.device
        grid = torch.arange(0, 1 + .5 / (n_grid - 1), 1 / (n_grid - 1)).to(dev)          # [10]
        derivative = ContiguousIntegrand(self.net.parallel_nets)
        # UMNN_B200_INVERT=torch keeps the reference-shaped op-by-op loop below (A/B checks of the fused bracket step)
        if iter >= 1 and B > 0 and z.dtype == torch.float32 and os.environ.get("UMNN_B200_INVERT", "") != "torch" and \
                kernel_route(derivative, z[:, :1], z[:, :1], z) is not None:
            return self._invert_native(z, iter, context, derivative, grid)
        target = z.unsqueeze(0).expand(n_grid, -1, -1)                                    # [10, B, D]
        x = target.clone()
        x_inv = torch.zeros(B, D).to(dev)
        left = -50 * torch.ones(B, D).to(dev)
        right = 50 * torch.ones(B, D).to(dev)
        s = torch.exp(self.scaling.unsqueeze(0).unsqueeze(1).expand(n_grid, B, -1))
        sample_base = torch.arange(0, B).to(dev) * n_grid
        with torch.no_grad():
            for j in range(self.input_size):
                if j % 100 == 0:
                    print(j)
                h_all = self.net.make_embeding(x_inv, context)
                offset = h_all.view(B, -1, D)[:, 0, [j]].unsqueeze(0).expand(n_grid, -1, -1)      # [10, B, 1]
                h_j = h_all[:, torch.arange(j, h_all.shape[1], D).to(dev)]
                h_j = h_j.unsqueeze(0).expand(n_grid, -1, -1).contiguous().view(n_grid * B, -1)
                x0 = torch.zeros(n_grid * B, 1).to(dev)
                for _ in range(iter):
                    x[:, :, j] = grid.view(-1, 1) * (right[:, j] - left[:, j]) + left[:, j]
                    integ = ParallelNeuralIntegral.apply(x0, x[:, :, j].contiguous().view(-1, 1), derivative, None,
                                                         h_j, self.nb_steps)
                    z_est = s[:, :, [j]] * (offset + integ.contiguous().view(n_grid, -1, 1))
                    _, z_pos = torch.abs(z_est[:, :, 0] - target[:, :, j]).min(0)
                    mid = z_pos + sample_base
                    z_val = z_est[:, :, 0].t().contiguous().view(-1)[mid]
                    x_flat = x[:, :, j].t().contiguous().view(-1)
                    below = (z_val < target[0, :, j]).float()
                    lo = mid - 1
                    hi = (mid + 1) % x_flat.shape[0]
                    left[:, j] = below * x_flat[mid] + (1 - below) * x_flat[lo]
                    right[:, j] = below * x_flat[hi] + (1 - below) * x_flat[mid]
                x_inv[:, j] = x_flat[mid]
        return x_inv


    def _invert_native(self, z, iter, context, derivative, grid):
        """invert() on the kernel route.  Per dimension: the conditioner's hidden layers plus ONLY the E rows of its last
        masked layer that dimension reads (EmbeddingNetwork.embedding_of_dim; the reference evaluates all E*D outputs,
        UMNNMAF.py:199-202), then ONE native call (umnn_invert_dimension) that enqueues the whole refinement of that
        dimension -- context replication, bracket reset, and `iter` rounds of {fused integral over the 10*B grid slots,
        fused bracket update} -- and writes the result straight into x[:, j].  Same arithmetic as the reference's
        ~15 torch ops per round (UMNNMAF.py:203-231), bit-identical bracket logic.

        UMNN_B200_INVERT=rounds keeps the previous form (two Python-level launches per round), =torch the
        reference-shaped op-by-op loop."""
        n_grid = grid.shape[0]
        B, D = z.shape
        dev = z.device
        spec = derivative.kernel_spec()
        z = z.contiguous()
        x_inv = torch.zeros(B, D, device=dev)
        s = torch.exp(self.scaling.detach()).to(dev).float().contiguous()
        E = spec.n_ctx
        if os.environ.get("UMNN_B200_INVERT", "") == "rounds":
            return self._invert_native_rounds(z, iter, context, spec, grid, x_inv, s)
        L = _native.lib()
        probe = torch.empty(n_grid * B, 1, device=dev)
        desc = kernel.make_desc(spec, probe, self.nb_steps)
        ws_bytes = int(L.umnn_invert_workspace_bytes(desc, B, n_grid))
        if ws_bytes == 0:
            _native.check(-2)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        w_tab, t_tab = device_tables(self.nb_steps, dev)
        grid = grid.float().contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.no_grad(), torch.cuda.device(dev):
            for j in range(self.input_size):
                if j % 100 == 0:
                    print(j)
                # only the E conditioner outputs dimension j reads
                h_cols = self.net.embedding_of_dim(x_inv, j, context).float().contiguous()
                packed = kernel.packed_parameters(spec, desc, dev)        # cached in eval mode, re-packed otherwise
                rc = L.umnn_invert_dimension(desc, packed.data_ptr(), t_tab.data_ptr(), w_tab.data_ptr(), B, n_grid, int(iter),
                                             h_cols.data_ptr(), grid.data_ptr(), s.data_ptr() + 4 * j, z.data_ptr() + 4 * j, D,
                                             -50.0, 50.0, x_inv.data_ptr() + 4 * j, D, ws.data_ptr(), ws_bytes, stream)
                if rc != 0:
                    _native.check(rc)
        return x_inv

    def _invert_native_rounds(self, z, iter, context, spec, grid, x_inv, s):
        """Round-by-round form of _invert_native (one fused integral launch + one bracket launch per round, issued from
        Python); kept for A/B checks of umnn_invert_dimension."""
        n_grid = grid.shape[0]
        B, D = z.shape
        dev = z.device
        E = spec.n_ctx
        h_j = torch.empty(n_grid * B, E, device=dev)
        offset = torch.empty(B, device=dev)
        target = torch.empty(B, device=dev)
        scale = torch.empty(1, device=dev)
        left = torch.empty(B, device=dev)
        right = torch.empty(B, device=dev)
        x_mid = torch.zeros(B, device=dev)
        x_a = torch.empty(n_grid, B, device=dev)
        x_b = torch.empty(n_grid, B, device=dev)
        integ = torch.empty(n_grid * B, 1, device=dev)
        with torch.no_grad():
            for j in range(self.input_size):
                h_cols = self.net.embedding_of_dim(x_inv, j, context).float()
                h_j.view(n_grid, B, E).copy_(h_cols.unsqueeze(0).expand(n_grid, -1, -1))
                offset.copy_(h_cols[:, 0])
                target.copy_(z[:, j])
                scale.copy_(s[j:j + 1])
                left.fill_(-50.)
                right.fill_(50.)
                xa, xb = x_a, x_b
                kernel.invert_bracket_step(None, None, grid, None, None, None, left, right, xa, None)
                for _ in range(iter):
                    kernel.cc_forward(spec, None, xa.view(-1, 1), h_j, self.nb_steps, out=integ)
                    kernel.invert_bracket_step(integ.view(n_grid, B), xa, grid, offset, scale, target, left, right, xb, x_mid)
                    xa, xb = xb, xa
                x_inv[:, j] = x_mid
        return x_inv


class ListModule(object):
    """Registers modules on a parent as `<prefix><i>` and indexes them like a list."""

    def __init__(self, module, prefix, *args):
        self.module = module
        self.prefix = prefix
        self.num_module = 0
        for m in args:
            self.append(m)

    def append(self, new_module):
        if not isinstance(new_module, nn.Module):
            raise ValueError('Not a Module')
        self.module.add_module(self.prefix + str(self.num_module), new_module)
        self.num_module += 1

    def __len__(self):
        return self.num_module

    def __getitem__(self, i):
        if i < 0 or i >= self.num_module:
            raise IndexError('Out of bound')
        return getattr(self.module, self.prefix + str(i))


def _reverse(t):
    """Reverse the feature axis ([:, D-1, ..., 0]) without a host-side index tensor."""
    return torch.flip(t, dims=[1])


class UMNNMAFFlow(nn.Module):
    def __init__(self, nb_flow=1, nb_in=1, hidden_derivative=[50, 50, 50, 50], hidden_embedding=[50, 50, 50, 50],
                 embedding_s=20, nb_steps=50, act_func='ELU', solver="CC", cond_in=0, device="cpu"):
        super().__init__()
        self.device = device
        self.register_buffer("pi", torch.tensor(math.pi))
        self.nets = ListModule(self, "Flow")
        for _ in range(nb_flow):
            emb = EmbeddingNetwork(nb_in, hidden_embedding, hidden_derivative, embedding_s, act_func=act_func,
                                   device=device, cond_in=cond_in).to(device)
            self.nets.append(UMNNMAF(emb, nb_in, nb_steps, device, solver=solver).to(device))

    def to(self, device):
        for net in self.nets:
            net.to(device)
        self.device = device
        super().to(device)
        return self

    def forward(self, x, context=None):
        for net in self.nets:
            x = _reverse(net.forward(x, context=context))
        return _reverse(x)

    def invert(self, z, iter=10, context=None):
        z = _reverse(z)
        for i in range(len(self.nets) - 1, -1, -1):
            z = self.nets[i].invert(_reverse(z), iter, context=context)
        return z

    def compute_log_jac(self, x, context=None):
        log_jac = 0.
        for net in self.nets:
            z, lj = net.forward_and_log_jac(x, context=context)
            log_jac += lj
            x = _reverse(z)
        return log_jac

    def compute_log_jac_bis(self, x, context=None):
        log_jac = 0.
        for net in self.nets:
            x, lj = net.compute_log_jac_bis(x, context=context)
            x = _reverse(x)
            log_jac += lj
        return _reverse(x), log_jac

    def compute_ll(self, x, context=None):
        log_jac = 0.
        z = x
        for net in self.nets:
            z, lj = net.forward_and_log_jac(x, context=context)
            z = _reverse(z)
            log_jac += lj
            x = z
        z = _reverse(z)
        log_prob_gauss = -.5 * (torch.log(self.pi * 2) + z ** 2).sum(1)
        return log_jac.sum(1) + log_prob_gauss, z

    def compute_ll_bis(self, x, context=None):
        log_jac = 0.
        for net in self.nets:
            z, lj = net.forward_and_log_jac(x, context=context)
            log_jac += lj
            x = _reverse(z)
        z = _reverse(x)
        log_prob_gauss = -.5 * (torch.log(self.pi * 2) + z ** 2)
        return log_jac + log_prob_gauss, z

    def compute_bpp(self, x, alpha=1e-6, context=None):
        d = x.shape[1]
        ll, z = self.compute_ll(x, context=context)
        bpp = -ll / (d * np.log(2)) - np.log2(1 - 2 * alpha) + 8 \
            + 1 / d * (torch.log2(torch.sigmoid(x)) + torch.log2(1 - torch.sigmoid(x))).sum(1)
        return bpp, ll, z

    def set_steps_nb(self, nb_steps):
        for net in self.nets:
            net.set_steps_nb(nb_steps)

    def compute_lipschitz(self, nb_iter=10):
        L = 1.
        for net in self.nets:
            L *= net.compute_lipschitz(nb_iter)
        return L

    def force_lipschitz(self, L=1.5):
        for net in self.nets:
            net.force_lipschitz(L)

    # spellings the reference's own drivers use (UCIExperiments.py:146,164; MNISTExperiment.py:167,225;
    # models/vae_lib/models/flows.py:325-327)
    computell = compute_ll
    forcei_lpschitz = force_lipschitz
    forceLipshitz = force_lipschitz
    computeLipshitz = compute_lipschitz


class MonotonicNN(nn.Module):
    """y = exp(s(h)) * int_0^x f(t, h) dt + o(h): monotone in x, unconstrained in h."""

    def __init__(self, in_d, hidden_layers, nb_steps=50, dev="cpu"):
        super().__init__()
        self.integrand = IntegrandNN(in_d, hidden_layers)
        self.net = nn.Sequential(*_mlp([in_d - 1] + list(hidden_layers) + [2], nn.ReLU))
        self.device = dev
        self.nb_steps = nb_steps

    def forward(self, x, h):
        x0 = torch.zeros(x.shape).to(self.device)
        out = self.net(h)
        offset = out[:, [0]]
        scaling = torch.exp(out[:, [1]])
        integral = ParallelNeuralIntegral.apply(x0, x, self.integrand, _flatten(self.integrand.parameters()), h,
                                                self.nb_steps)
        return scaling * integral + offset
