"""Integrand networks and the MADE conditioner (host-side mirror of the reference modules).

Behavioural spec (AWehenkel/UMNN @ 59118c14).  IntegrandNetwork / IntegrandNN / KernelSpec are written fresh.
MaskedLinear, MADE.update_masks / compute_ll / invert and ConditionnalMADE are NOT: they are the reference's
out-of-scope conditioner (models/UMNN/made.py:16-192, itself derived from Karpathy's public pytorch-made),
restated closely on purpose -- masks, degrees, state-dict keys and outputs must be identical for the reference's
checkpoints to load and for h to be the same tensor.  Only the mask*weight caching is ours.
  * IntegrandNetwork  models/UMNN/UMNNMAF.py:235-301 -- ONE shared MLP applied to every (sample, dim)
    slot; slot (n, d) sees [x[n,d], h[n, 0*D+d], ..., h[n, (E-1)*D+d]]; LeakyReLU(0.01) hidden,
    ELU+1 or Sigmoid output.  `forward` stays a pure-torch, TorchScript-able function
    (tests/test_jit.py:170-266); on CUDA the integral over it is served by the fused kernel, which
    recognises the module through `kernel_spec()`.
  * IntegrandNN       models/UMNN/MonotonicNN.py:12-27 -- cat(x, h) -> Linear/ReLU... -> ELU -> +1.
  * MaskedLinear / MADE / ConditionnalMADE   models/UMNN/made.py:16-192 -- the autoregressive
    conditioner that produces h.  Out of scope for the CUDA kernels (stays torch/cuBLAS); kept so that
    state-dict keys and outputs match.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native


class ELUPlus(nn.Module):
    """ELU(x) + 1: strictly positive output activation (UMNNMAF.py:11-16)."""

    def __init__(self):
        super().__init__()
        self.elu = nn.ELU()

    def forward(self, x):
        return self.elu(x) + 1.


dict_act_func = {"Sigmoid": nn.Sigmoid(), "ELU": ELUPlus()}


def _flatten(sequence):
    """Concatenate tensors into one flat vector (empty tensor if there are none)."""
    pieces = [t.contiguous().view(-1) for t in sequence]
    return torch.cat(pieces) if pieces else torch.tensor([])


def compute_lipschitz_linear(W, nb_iter=10):
    """Spectral norm of W by power iteration on W^T W (UMNNMAF.py:26-34)."""
    v = torch.randn(W.shape[1], 1).to(W.device)
    for _ in range(nb_iter):
        v = W.t() @ (W @ v)
        v = v / torch.norm(v)
    return (torch.norm(W.t() @ (W @ v)) / torch.norm(v)) ** .5


def _mlp(sizes: List[int], hidden: type, linear=nn.Linear) -> List[nn.Module]:
    layers: List[nn.Module] = []
    for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:])):
        layers.append(linear(a, b))
        if i < len(sizes) - 2:
            layers.append(hidden())
    return layers


class KernelSpec:
    """What the fused kernel needs to know about a recognised integrand."""
    __slots__ = ("layout", "widths", "hidden_act", "out_act", "linears", "n_dims", "desc_cache", "info_cache", "param_list")

    def __init__(self, layout, widths, hidden_act, out_act, linears, n_dims):
        self.layout, self.widths, self.hidden_act, self.out_act = layout, widths, hidden_act, out_act
        self.linears, self.n_dims = linears, n_dims
        self.desc_cache, self.info_cache = {}, {}     # launch descriptors and what the library derives from them
        # the Parameter objects at the time the spec was built (the spec cache is keyed on their identities)
        self.param_list = [t for lin in linears for t in (lin.weight, lin.bias)]

    @property
    def n_ctx(self):
        return self.widths[0] - 1

    def parameters(self):
        return iter(self.param_list)

    def supported(self) -> Optional[str]:
        """None if the kernel can serve this shape, else the reason (memoised: the shape is fixed per spec)."""
        hit = self.info_cache.get("supported", False)
        if hit is not False:
            return hit
        why = self._supported()
        self.info_cache["supported"] = why
        return why

    def _supported(self) -> Optional[str]:
        if len(self.widths) - 1 > _native.UMNN_MAX_LAYERS:
            return f"{len(self.widths) - 1} Linear layers > {_native.UMNN_MAX_LAYERS}"
        if len(self.widths) < 3:
            return "needs at least one hidden layer"
        if max(self.widths) > _native.UMNN_MAX_WIDTH:
            return f"width {max(self.widths)} > {_native.UMNN_MAX_WIDTH}"
        if self.widths[-1] != 1:
            return "output width must be 1"
        return None


def _cached_spec(owner: nn.Module, seq: nn.Sequential, build):
    """KernelSpec of `seq`, rebuilt only when the Sequential's module list (or an activation's slope / alpha) changes:
    parsing the stack on every call costs more than a small launch."""
    key = tuple([(id(m), m.__dict__.get("negative_slope"), m.__dict__.get("alpha"), *map(id, m._parameters.values()))
                 for m in seq._modules.values()])
    hit = owner.__dict__.get("_umnn_spec")
    if hit is not None and hit[0] == key:
        return hit[1]
    spec = build()
    owner.__dict__["_umnn_spec"] = (key, spec)
    return spec


def _sequential_spec(seq: nn.Sequential, hidden_cls, layout, n_dims, out_classes):
    """Parse Linear/act/.../Linear/out_act; returns None if the stack is not of that form.

    `out_classes`: the output-activation modules the CALLER's forward is consistent with.  IntegrandNetwork.forward
    applies the Sequential as is, so only ELUPlus (ELU + 1) and Sigmoid are positive-output forms the kernel matches;
    IntegrandNN.forward adds 1 after a plain nn.ELU.  Anything else (e.g. an IntegrandNetwork whose net was edited
    to end in nn.ELU) is not the kernel's function -> None -> torch route."""
    mods = list(seq)
    if len(mods) < 2 or len(mods) % 2 != 0:
        return None
    linears = []
    for i in range(0, len(mods), 2):
        if type(mods[i]) is not nn.Linear or mods[i].bias is None:
            return None
        linears.append(mods[i])
        if i + 1 < len(mods) - 1 and type(mods[i + 1]) is not hidden_cls:
            return None
    last = mods[-1]
    if type(last) not in out_classes:
        return None
    if isinstance(last, (ELUPlus, nn.ELU)):
        elu = last.elu if isinstance(last, ELUPlus) else last
        if elu.alpha != 1.0:
            return None
        out_act = _native.OUT_ELU_PLUS_1
    elif isinstance(last, nn.Sigmoid):
        out_act = _native.OUT_SIGMOID
    else:
        return None
    if hidden_cls is nn.LeakyReLU:
        for m in mods[1:-1:2]:
            if abs(m.negative_slope - 0.01) > 0:
                return None
        hidden_act = _native.ACT_LEAKY_RELU
    else:
        hidden_act = _native.ACT_RELU
    widths = [linears[0].in_features] + [lin.out_features for lin in linears]
    for a, b in zip(linears[:-1], linears[1:]):
        if a.out_features != b.in_features:
            return None
    return KernelSpec(layout, widths, hidden_act, out_act, linears, n_dims)


class IntegrandNetwork(nn.Module):
    def __init__(self, nnets, nin, hidden_sizes, nout, act_func='ELU', device="cpu"):
        super().__init__()
        self.nin = nin
        self.nnets = nnets
        self.nout = nout
        self.hidden_sizes = hidden_sizes
        self.device = device
        layers = _mlp([nin] + list(hidden_sizes) + [nout], nn.LeakyReLU)
        layers.append(dict_act_func[act_func])
        self.net = nn.Sequential(*layers)

    def to(self, device):
        self.device = device
        self.net.to(device)
        return self

    def forward(self, x, h):
        # [N, D] and [N, E*D] -> rows [N*D, 1+E] with the D axis fastest inside h
        n_rows = x.shape[0]
        stacked = torch.cat((x, h), 1)
        per_net = stacked.shape[1] // self.nnets
        rows = stacked.view(n_rows, per_net, self.nnets).transpose(1, 2).contiguous().view(n_rows * self.nnets, per_net)
        return self.net(rows).view(n_rows, -1)

    def independant_forward(self, x):
        return self.net(x)

    def compute_lipschitz(self, nb_iter=10):
        with torch.no_grad():
            L = 1
            for layer in self.net.modules():
                if isinstance(layer, nn.Linear):
                    L *= compute_lipschitz_linear(layer.weight, nb_iter)
        return L

    computeLipshitz = compute_lipschitz  # spelling used by UMNNMAF.compute_lipschitz, UMNNMAF.py:176

    def force_lipschitz(self, L=1.5):
        with torch.no_grad():
            for layer in self.net.modules():
                if isinstance(layer, nn.Linear):
                    layer.weight /= max(compute_lipschitz_linear(layer.weight, 10) / L, 1)

    @torch.jit.unused
    def kernel_spec(self) -> Optional[KernelSpec]:
        if self.nout != 1:
            return None
        return _cached_spec(self, self.net, lambda: _sequential_spec(self.net, nn.LeakyReLU, _native.LAYOUT_STRIDED_D,
                                                                     self.nnets, (ELUPlus, nn.Sigmoid)))


class ContiguousIntegrand(nn.Module):
    """`lambda x, h: parallel_nets.independant_forward(cat(x, h))` as a recognisable module: the
    contiguous-context view of an IntegrandNetwork used by UMNNMAF.invert (UMNNMAF.py:207)."""

    def __init__(self, parallel_nets: IntegrandNetwork):
        super().__init__()
        self.parallel_nets = parallel_nets

    def forward(self, x, h):
        return self.parallel_nets.independant_forward(torch.cat((x, h), 1))

    def kernel_spec(self) -> Optional[KernelSpec]:
        spec = self.parallel_nets.kernel_spec()
        if spec is None:
            return None
        hit = self.__dict__.get("_umnn_contig")
        if hit is None or hit[0] is not spec:
            hit = (spec, KernelSpec(_native.LAYOUT_CONTIG, spec.widths, spec.hidden_act, spec.out_act, spec.linears, 1))
            self.__dict__["_umnn_contig"] = hit
        return hit[1]


class IntegrandNN(nn.Module):
    def __init__(self, in_d, hidden_layers):
        super().__init__()
        layers = _mlp([in_d] + list(hidden_layers) + [1], nn.ReLU)
        layers.append(nn.ELU())
        self.net = nn.Sequential(*layers)

    def forward(self, x, h):
        return self.net(torch.cat((x, h), 1)) + 1.

    def kernel_spec(self) -> Optional[KernelSpec]:
        return _cached_spec(self, self.net, lambda: _sequential_spec(self.net, nn.ReLU, _native.LAYOUT_CONTIG, 1, (nn.ELU,)))


# --------------------------------------------------------------------------------------------------
# MADE conditioner
# --------------------------------------------------------------------------------------------------
class MaskedLinear(nn.Linear):
    """Linear layer whose weight is multiplied by a fixed 0/1 `mask` buffer (made.py:16-27)."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias)
        self.register_buffer('mask', torch.ones(out_features, in_features))

    def set_mask(self, mask):
        self.mask.data.copy_(torch.from_numpy(mask.astype(np.uint8).T))
        self.__dict__.pop("_masked_cache", None)

    def masked_weight(self):
        """mask * weight (made.py:26-27 recomputes it on every forward).  When no gradient can flow into the weight
        (no_grad / eval inference, sampling) the product is cached and reused until the weight or the mask changes
        (storage address + autograd version, as for the packed integrand parameters); with autograd on it is computed
        in the graph as in the reference."""
        if (torch.is_grad_enabled() and self.weight.requires_grad) or \
                (self.weight.is_cuda and torch.cuda.is_current_stream_capturing()):
            # (while a CUDA graph is being captured the product must be a node of the graph: replays read the live
            # weights, umnn_b200/graphs.py)
            return self.mask * self.weight
        stamp = (self.weight.data_ptr(), self.weight._version, self.mask.data_ptr(), self.mask._version, self.weight.device)
        hit = self.__dict__.get("_masked_cache")
        if hit is None or hit[0] != stamp:
            hit = (stamp, (self.mask * self.weight).detach())
            self.__dict__["_masked_cache"] = hit
        return hit[1]

    def forward(self, input):
        return F.linear(input, self.masked_weight(), self.bias)


class MADE(nn.Module):
    def __init__(self, nin, hidden_sizes, nout, num_masks=1, natural_ordering=False, random=False, device="cpu"):
        super().__init__()
        assert nout % nin == 0, "nout must be integer multiple of nin"
        self.random = random
        self.nin = nin
        self.nout = nout
        self.device = device
        self.pi = torch.tensor(math.pi).to(self.device)
        self.hidden_sizes = hidden_sizes
        self.net = nn.Sequential(*_mlp([nin] + list(hidden_sizes) + [nout], nn.ReLU, MaskedLinear)).to(device)
        self.natural_ordering = natural_ordering
        self.num_masks = num_masks
        self.seed = 0
        self.m = {}
        self.update_masks()

    def update_masks(self):
        if self.m and self.num_masks == 1:
            return
        depth = len(self.hidden_sizes)
        rng = np.random.RandomState(self.seed)
        self.seed = (self.seed + 1) % self.num_masks
        # degrees: inputs in (natural or permuted) order, hidden units either sampled or cycling downwards
        if self.random:
            self.m[-1] = np.arange(self.nin) if self.natural_ordering else rng.permutation(self.nin)
            for l in range(depth):
                self.m[l] = rng.randint(self.m[l - 1].min(), self.nin - 1, size=self.hidden_sizes[l])
        else:
            self.m[-1] = np.arange(self.nin)
            for l in range(depth):
                self.m[l] = np.array([self.nin - 1 - (u % self.nin) for u in range(self.hidden_sizes[l])])
        masks = [self.m[l - 1][:, None] <= self.m[l][None, :] for l in range(depth)]
        masks.append(self.m[depth - 1][:, None] < self.m[-1][None, :])
        if self.nout > self.nin:
            masks[-1] = np.concatenate([masks[-1]] * int(self.nout / self.nin), axis=1)
        for layer, mk in zip([l for l in self.net.modules() if isinstance(l, MaskedLinear)], masks):
            layer.set_mask(mk)
        self.i_map = self.m[-1].copy()
        for k in range(len(self.m[-1])):
            self.i_map[self.m[-1][k]] = k

    def _gaussian_params(self, x):
        out = self.net(x)
        return out[:, :self.nin], out[:, self.nin:]

    def forward(self, x, context=None):
        if self.nout == 2:
            mu, sigma = self._gaussian_params(x)
            return (x - mu) * torch.exp(-sigma)
        return self.net(x)

    # ----- sampling support: only the outputs one dimension needs ------------------------------------------------
    def hidden_forward(self, x):
        """Activation of the last hidden layer (everything but the final MaskedLinear)."""
        a = x
        for layer in list(self.net)[:-1]:
            a = layer(a)
        return a

    def output_columns(self, hidden, first, stride):
        """Outputs `first, first + stride, first + 2*stride, ...` of the final MaskedLinear for `hidden` =
        hidden_forward(x): out[:, first::stride] of the full forward without evaluating the other columns.
        UMNNMAF.invert (UMNNMAF.py:199-202) needs the E values h[:, e*D + j] for ONE dimension j per step, while the
        final layer has E*D outputs (23 520 at the MNIST shape)."""
        last = list(self.net)[-1]
        wm = last.masked_weight()
        return F.linear(hidden, wm[first::stride], last.bias[first::stride])

    def compute_ll(self, x):
        mu, sigma = self._gaussian_params(x)
        z = (x - mu) * torch.exp(-sigma)
        log_prob_gauss = -.5 * (torch.log(self.pi * 2) + z ** 2).sum(1)
        return -sigma.sum(1) + log_prob_gauss, z

    def invert(self, z):
        if self.nin != self.nout / 2:
            return None
        u = torch.zeros(z.shape)
        for d in range(self.nin):
            out = self.forward(u)
            mu, sigma = out[:, self.i_map[d]], out[:, self.nin + self.i_map[d]]
            u[:, self.i_map[d]] = z[:, self.i_map[d]] * torch.exp(sigma) + mu
        return u


class ConditionnalMADE(MADE):
    def __init__(self, nin, cond_in, hidden_sizes, nout, num_masks=1, natural_ordering=False, random=False,
                 device="cpu"):
        super().__init__(nin + cond_in, hidden_sizes, nout, num_masks, natural_ordering, random, device)
        self.nin_non_cond = nin
        self.cond_in = cond_in

    def _strip_context(self, out, batch):
        chunks = out.contiguous().view(batch, int(out.shape[1] / self.nin), self.nin)
        return chunks[:, :, self.cond_in:].contiguous().view(batch, -1)

    def forward(self, x, context):
        return self._strip_context(super().forward(torch.cat((context, x), 1)), x.shape[0])

    def hidden_forward(self, x, context=None):
        return super().hidden_forward(x if context is None else torch.cat((context, x), 1))

    def output_columns(self, hidden, first, stride):
        """Columns first::stride of forward()'s (context-stripped) output: chunk e of the raw output keeps its
        entries cond_in.., so stripped column e*nin_non_cond + j is raw column e*nin + cond_in + j."""
        if stride != self.nin_non_cond:
            raise ValueError("ConditionnalMADE.output_columns: stride must be the number of data dimensions")
        return super().output_columns(hidden, self.cond_in + first, self.nin)

    def invert(self, z, context=None):
        """Signature and the `None` return for non-Gaussian heads follow made.py:181-192.  The reference's loop body
        reads an undefined name (`x`) and cannot run; this one solves the trailing data dimensions one by one with
        the conditioning inputs held fixed."""
        if context is None:
            return super().invert(z)
        if self.nin != self.nout / 2:
            return None
        # conditioning variables are the leading inputs and stay fixed; only the trailing data dims are solved for
        u = torch.cat((context, torch.zeros(z.shape[0], self.nin_non_cond, dtype=z.dtype, device=z.device)), 1)
        zc = torch.cat((torch.zeros_like(context), z), 1)
        for d in range(self.cond_in, self.nin):
            out = self.net(u)
            mu, sigma = out[:, self.i_map[d]], out[:, self.nin + self.i_map[d]]
            u[:, self.i_map[d]] = zc[:, self.i_map[d]] * torch.exp(sigma) + mu
        return u[:, self.cond_in:]

    def computeLL(self, x, context):
        out = self._strip_context(self.net(torch.cat((context, x), 1)), x.shape[0])
        mu, sigma = out[:, :self.nin], out[:, self.nin:]
        z = (x - mu) * torch.exp(-sigma)
        log_prob_gauss = -.5 * (torch.log(self.pi * 2) + z ** 2).sum(1)
        return -sigma.sum(1) + log_prob_gauss, z
