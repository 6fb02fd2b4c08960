"""B-spline KAN modules with the reference's names, constructor arguments and ``state_dict`` keys
(``KANLinear`` / ``KAN`` of node_classification_clean/ekan.py:7-281), whose ``forward`` runs in the sm_100a
library instead of ATen.

Parameters / buffers (identical keys, shapes and meaning, so reference checkpoints load):
``base_weight (out,in)``, ``spline_weight (out,in,G+k)``, ``spline_scaler (out,in)``, buffer ``grid (in,G+2k+1)``.

Initialisation (cold path, runs once in torch on whatever device the module is created on) follows the reference's
recipe (ekan.py:57-77): Kaiming-uniform base weight and scaler, spline coefficients = least-squares fit of small
uniform noise sampled on the G+1 interior knots.  ``update_grid`` / ``regularization_loss`` are dead code in every
reference driver and are not provided.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence

import torch
from torch import nn

from . import _lib as L
from . import autograd, ops

Tensor = torch.Tensor


def _uniform_bspline_design(x: Tensor, t0: float, h: float, grid_size: int, order: int) -> Tensor:
    """(M, in) sample points -> (M, in, G+k) B-spline design tensor on the uniform knot vector t_j = t0 + j*h.
    Init-time helper (torch, any device): local de Boor on the fractional position inside the knot interval,
    scattered to the k+1 non-zero coefficient slots."""
    S = grid_size + order
    u = (x - t0) / h
    cell = torch.floor(u)
    frac = u - cell
    inside = (u >= 0) & (u < grid_size + 2 * order)
    local = [torch.ones_like(frac)]
    for d in range(1, order + 1):
        nxt = []
        for r in range(d + 1):
            term = torch.zeros_like(frac)
            if r > 0:
                term = term + (frac + (d - r)) / d * local[r - 1]
            if r < d:
                term = term + ((r + 1) - frac) / d * local[r]
            nxt.append(term)
        local = nxt
    out = x.new_zeros(*x.shape, S)
    base = cell.long() - order
    for r in range(order + 1):
        slot = base + r
        ok = inside & (slot >= 0) & (slot < S)
        out.scatter_add_(-1, slot.clamp(0, S - 1).unsqueeze(-1), (local[r] * ok).unsqueeze(-1))
    return out


def _module_backend_guard(x: Tensor, params: Iterable[Tensor], grad_ok: bool = False) -> bool:
    """Raises for CPU inputs (no fallback).  Returns True when autograd must record this call: modules that have a backward
    (``grad_ok``) then take their ``kagnn_b200.autograd`` path; a caller that cannot be differentiated passes False and raises."""
    if not x.is_cuda:
        raise RuntimeError("kagnn_b200 modules run on the B200 only: move the module and its inputs to a CUDA device "
                           "(there is deliberately no CPU fallback)")
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        if not grad_ok:
            raise NotImplementedError("this kagnn_b200 call has no backward: use it under "
                                      "torch.no_grad() / torch.inference_mode()")
        return True
    return False


_warned_eval_detach = False


def eval_mode_detach_notice(x: Tensor) -> None:
    """model.eval() with autograd enabled: the models take the fused inference plan and return a result that is DETACHED from
    autograd (the reference's val() / test() loops run exactly like that and never call backward on it).  Anything that does need
    gradients through an eval-mode model must not get a silently detached tensor: input gradients raise, and the first detached
    call warns once."""
    global _warned_eval_detach
    if x.is_floating_point() and x.requires_grad:
        raise NotImplementedError("kagnn_b200: gradients with respect to the input of a model in eval() mode are not implemented "
                                  "(the eval-mode forward is the fused inference plan); call model.train() -- with BatchNorm / "
                                  "Dropout modules individually in eval() if their statistics must stay frozen -- or detach the input")
    if not _warned_eval_detach:
        _warned_eval_detach = True
        import warnings
        warnings.warn("kagnn_b200: a model in eval() mode was called with autograd enabled; its output is computed by the fused "
                      "inference plan and is detached from autograd (wrap evaluation in torch.no_grad() to silence this)", stacklevel=3)


def windowed_weights(base_w: Optional[Tensor], spline_w: Tensor, scaler: Optional[Tensor], windows: int):
    """Weights of the virtual layer that evaluates a layer with more than eight slots per feature as ``windows`` copies of its
    input with eight slots each (KANLinear._windowed_spec, FastKANLayer._windowed_spec): virtual input ``w * in + i`` carries slots
    8w .. 8w+7 of input i (zero beyond the real slot count), copy 0 the base weights, every copy the scaler.
    (out, in, S) -> ((out, windows*in) | None, (out, windows*in, 8), (out, windows*in) | None)."""
    out_f, in_f, slots = spline_w.shape
    sp = torch.zeros(out_f, in_f, 8 * windows, dtype=torch.float32, device=spline_w.device)
    sp[:, :, :slots] = spline_w.detach()
    virt_spline = sp.view(out_f, in_f, windows, 8).permute(0, 2, 1, 3).reshape(out_f, windows * in_f, 8).contiguous()
    virt_base = None
    if base_w is not None:
        virt_base = torch.zeros(out_f, windows * in_f, dtype=torch.float32, device=spline_w.device)
        virt_base[:, :in_f] = base_w.detach()
    virt_scaler = None if scaler is None else scaler.detach().repeat(1, windows).contiguous()
    return virt_base, virt_spline, virt_scaler


class KANLinear(nn.Module):
    """Drop-in for ``ekan.KANLinear``: y = silu(x) @ base_weight^T + B(x) @ (spline_weight * spline_scaler)^T."""

    def __init__(self, in_features, out_features, grid_size=5, spline_order=3, scale_noise=0.1, scale_base=1.0,
                 scale_spline=1.0, enable_standalone_scale_spline=True, base_activation=nn.SiLU, grid_eps=0.02,
                 grid_range=[-1, 1]):
        super().__init__()
        if base_activation is not nn.SiLU:
            raise NotImplementedError("only the SiLU base activation (the one every reference model uses) is implemented")
        self.in_features, self.out_features = in_features, out_features
        self.grid_size, self.spline_order = grid_size, spline_order
        self.scale_noise, self.scale_base, self.scale_spline = scale_noise, scale_base, scale_spline
        self.enable_standalone_scale_spline = enable_standalone_scale_spline
        self.base_activation = base_activation()
        self.grid_eps = grid_eps
        lo, hi = grid_range
        step = (hi - lo) / grid_size
        knots = torch.arange(-spline_order, grid_size + spline_order + 1) * step + lo
        self.register_buffer("grid", knots.expand(in_features, -1).contiguous())
        self.base_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.spline_weight = nn.Parameter(torch.empty(out_features, in_features, grid_size + spline_order))
        if enable_standalone_scale_spline:
            self.spline_scaler = nn.Parameter(torch.empty(out_features, in_features))
        self._cache_key = None
        self._cache_spec: Optional[ops.KanLayerSpec] = None
        self._grid_key = None
        self._grid_t0h = None
        self.reset_parameters()

    # -- init (torch, cold) ------------------------------------------------------------------------
    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.base_weight, a=math.sqrt(5) * self.scale_base)
        with torch.no_grad():
            G, k = self.grid_size, self.spline_order
            pts = self.grid.t()[k:G + k + 1]                                   # (G+1, in) interior knots
            noise = (torch.rand(G + 1, self.in_features, self.out_features, device=pts.device) - 0.5) * self.scale_noise / G
            coeff = self.curve2coeff(pts, noise)
            self.spline_weight.copy_(coeff if self.enable_standalone_scale_spline else self.scale_spline * coeff)
            if self.enable_standalone_scale_spline:
                nn.init.kaiming_uniform_(self.spline_scaler, a=math.sqrt(5) * self.scale_spline)

    def curve2coeff(self, x: Tensor, y: Tensor) -> Tensor:
        """Least-squares spline coefficients (out,in,G+k) interpolating y (M,in,out) at x (M,in)."""
        t0 = float(self.grid[0, 0])
        h = float(self.grid[0, 1] - self.grid[0, 0])
        design = _uniform_bspline_design(x, t0, h, self.grid_size, self.spline_order)   # (M,in,S)
        sol = torch.linalg.lstsq(design.permute(1, 0, 2), y.permute(1, 0, 2)).solution  # (in,S,out)
        return sol.permute(2, 0, 1).contiguous()

    @property
    def scaled_spline_weight(self) -> Tensor:
        return self.spline_weight * (self.spline_scaler.unsqueeze(-1) if self.enable_standalone_scale_spline else 1.0)

    # -- kernel-side description -------------------------------------------------------------------
    def _knot_params(self):
        g = self.grid
        key = (g.data_ptr(), g._version, g.device)
        if key != self._grid_key:
            n = g.size(1)
            first = g[0].double()
            t0, h = float(first[0]), float((first[-1] - first[0]) / (n - 1))
            ideal = (t0 + h * torch.arange(n, device=g.device, dtype=torch.float64)).to(torch.float32)
            if not (h > 0) or float((g - ideal).abs().max()) > 1e-4 * max(abs(h), 1e-12):
                raise NotImplementedError("non-uniform knot vectors (a grid produced by update_grid) are not supported "
                                          "by the sm_100a path; no reference driver produces them")
            self._grid_key, self._grid_t0h = key, (t0, h)
        return self._grid_t0h

    def kernel_spec(self) -> ops.KanLayerSpec:
        """Pack (and cache until a parameter changes) the weights in the layout the fused kernel streams."""
        ps = [self.base_weight, self.spline_weight] + ([self.spline_scaler] if self.enable_standalone_scale_spline else [])
        key = tuple((p.data_ptr(), p._version) for p in ps) + (self.grid.data_ptr(), self.grid._version)
        if key != self._cache_key:
            t0, h = self._knot_params()
            slots = self.grid_size + self.spline_order
            if slots > 8 and 1 <= self.spline_order <= 3 and ops.tc_supported(L.BASIS_BSPLINE, 8 - self.spline_order, self.spline_order, self.out_features):
                self._cache_spec, self._cache_key = self._windowed_spec(t0, h, slots), key
                return self._cache_spec
            packed = ops.pack_kan_weights(self.base_weight, self.spline_weight,
                                          self.spline_scaler if self.enable_standalone_scale_spline else None,
                                          self.in_features, self.out_features, slots)
            scaler = self.spline_scaler if self.enable_standalone_scale_spline else None
            packed_tc = None
            if ops.tc_supported(L.BASIS_BSPLINE, self.grid_size, self.spline_order, self.out_features):
                packed_tc = ops.pack_kan_weights_tc(self.base_weight, self.spline_weight, scaler, self.in_features,
                                                    self.out_features, slots)
            self._cache_spec = ops.KanLayerSpec(L.BASIS_BSPLINE, self.in_features, self.out_features, self.grid_size,
                                                self.spline_order, t0, h, 0.0, packed, packed_w_tc=packed_tc)
            self._cache_key = key
        return self._cache_spec

    def _windowed_spec(self, t0: float, h: float, slots: int) -> ops.KanLayerSpec:
        """More than eight coefficients per (in, out) pair (the reference's search space goes to grid_size 8 + spline_order 3,
        node_classification/one_experiment.py:45-46): the tensor-core kernels hold eight slots per feature, and a uniform B-spline
        basis is shift invariant -- B_{8w+j}(x) = B_j(x - 8wh) -- so the layer is evaluated as a layer of grid_size 8 - k over
        ``windows`` copies of the input, copy w shifted by 8wh and carrying coefficients 8w .. 8w+7 (zero beyond G + k); the SiLU
        base weight rides on copy 0.  Outside a copy's knot range all of its eight bases are zero, as they should be."""
        with torch.no_grad():
            w = (slots + 7) // 8
            out_f, in_f = self.out_features, self.in_features
            virt_base, virt_spline, virt_scaler = windowed_weights(
                self.base_weight, self.spline_weight, self.spline_scaler if self.enable_standalone_scale_spline else None, w)
            g_virtual = 8 - self.spline_order
            packed = ops.pack_kan_weights(virt_base, virt_spline, virt_scaler, w * in_f, out_f, 8)
            packed_tc = ops.pack_kan_weights_tc(virt_base, virt_spline, virt_scaler, w * in_f, out_f, 8)
        return ops.KanLayerSpec(L.BASIS_BSPLINE, w * in_f, out_f, g_virtual, self.spline_order, t0, h, 0.0, packed, packed_w_tc=packed_tc,
                                windows=w, window_shift=8.0 * h, virt_spline=virt_spline, virt_scaler=virt_scaler)

    def kernel_specs(self) -> List[ops.KanLayerSpec]:
        return [self.kernel_spec()]

    def _params(self):
        return list(self.parameters(recurse=False))

    def forward(self, x: Tensor) -> Tensor:
        assert x.dim() == 2 and x.size(1) == self.in_features
        if _module_backend_guard(x, self._params(), grad_ok=True):
            return autograd.kan_linear(self, x)
        return ops.fused_layer(ops.AggSpec(L.AGG_NONE, x.to(torch.float32)), x.size(0), [self.kernel_spec()])


class KAN(nn.Module):
    """Drop-in for ``ekan.KAN``: KANLinear layers back to back; the whole chain is ONE launch."""

    def __init__(self, layers_hidden, grid_size=5, spline_order=3, scale_noise=0.1, scale_base=1.0, scale_spline=1.0,
                 base_activation=nn.SiLU, grid_eps=0.02, grid_range=[-1, 1]):
        super().__init__()
        self.grid_size, self.spline_order = grid_size, spline_order
        self.layers = nn.ModuleList(
            KANLinear(i, o, grid_size=grid_size, spline_order=spline_order, scale_noise=scale_noise, scale_base=scale_base,
                      scale_spline=scale_spline, base_activation=base_activation, grid_eps=grid_eps, grid_range=grid_range)
            for i, o in zip(layers_hidden, layers_hidden[1:]))

    def kernel_specs(self) -> List[ops.KanLayerSpec]:
        return [lay.kernel_spec() for lay in self.layers]

    def forward(self, x: Tensor, update_grid: bool = False) -> Tensor:
        if update_grid:
            raise NotImplementedError("update_grid is never used by the reference drivers and is not implemented")
        if _module_backend_guard(x, self.parameters(), grad_ok=True):
            for layer in self.layers:               # one launch per layer: every layer's input is kept for its backward
                x = layer(x)
            return x
        return chain_forward(self, x)


def chain_forward(module, x: Tensor) -> Tensor:
    """Run a KAN / FastKAN chain, ``L.MAX_LAYERS`` layers per launch."""
    specs = module.kernel_specs()
    x = x.to(torch.float32)
    for i in range(0, len(specs), L.MAX_LAYERS):
        x = ops.fused_layer(ops.AggSpec(L.AGG_NONE, x), x.size(0), specs[i:i + L.MAX_LAYERS])
    return x
