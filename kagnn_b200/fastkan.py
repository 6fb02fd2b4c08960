"""RBF KAN modules with the reference's names and ``state_dict`` keys (``SplineLinear``, ``RadialBasisFunction``,
``FastKANLayer``, ``FastKAN`` of node_classification_clean/fastkan.py:22-145); ``forward`` runs in the sm_100a library.

Keys: ``layernorm.{weight,bias}``, ``rbf.grid`` (frozen Parameter, counted by the reference's ``count_params``),
``spline_linear.weight (out, in*G)``, ``base_linear.{weight,bias}``.
``plot_curve`` and ``AttentionWithFastKANTransform`` are never referenced by any model and are not provided."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from . import autograd, ops
from .ekan import _module_backend_guard, chain_forward, windowed_weights

Tensor = torch.Tensor


class SplineLinear(nn.Linear):
    """Bias-free linear map over the flattened (in*G) RBF features, truncated-normal init."""

    def __init__(self, in_features: int, out_features: int, init_scale: float = 0.1, **kw) -> None:
        self.init_scale = init_scale
        super().__init__(in_features, out_features, bias=False, **kw)

    def reset_parameters(self) -> None:
        nn.init.trunc_normal_(self.weight, mean=0, std=self.init_scale)


class RadialBasisFunction(nn.Module):
    """Holds the G equally spaced centres and the width; evaluated inside the fused kernel."""

    def __init__(self, grid_min: float = -2., grid_max: float = 2., num_grids: int = 8, denominator: float = None):
        super().__init__()
        self.grid_min, self.grid_max, self.num_grids = grid_min, grid_max, num_grids
        self.grid = nn.Parameter(torch.linspace(grid_min, grid_max, num_grids), requires_grad=False)
        self.denominator = denominator or (grid_max - grid_min) / (num_grids - 1)

    def forward(self, x):  # pragma: no cover - the fused kernel evaluates the basis; kept for API parity
        raise RuntimeError("RadialBasisFunction is evaluated inside kagnn_fused_layer_fwd; call FastKANLayer instead")


def windowed_layernorm(weight: Tensor, bias: Tensor, windows: int, shift: float):
    """LayerNorm vectors of the virtual FastKAN layer (FastKANLayer._windowed_spec): every copy the weight, copy w the bias minus
    w * shift (the copy's centres sit w * shift further right)."""
    w = weight.detach().repeat(windows).contiguous()
    steps = torch.arange(windows, device=bias.device, dtype=torch.float32).unsqueeze(1)
    b = (bias.detach().unsqueeze(0) - shift * steps).reshape(-1).contiguous()
    return w, b


class FastKANLayer(nn.Module):
    """Drop-in for ``fastkan.FastKANLayer``: spline_linear(rbf(layernorm(x))) + base_linear(silu(x))."""

    def __init__(self, input_dim: int, output_dim: int, grid_min: float = -2., grid_max: float = 2., num_grids: int = 8,
                 use_base_update: bool = True, use_layernorm: bool = True, base_activation=F.silu,
                 spline_weight_init_scale: float = 0.1) -> None:
        super().__init__()
        if base_activation is not F.silu:
            raise NotImplementedError("only the SiLU base activation is implemented")
        self.input_dim, self.output_dim = input_dim, output_dim
        self.layernorm = None
        if use_layernorm:
            assert input_dim > 1, "Do not use layernorms on 1D inputs. Set `use_layernorm=False`."
            self.layernorm = nn.LayerNorm(input_dim)
        self.rbf = RadialBasisFunction(grid_min, grid_max, num_grids)
        self.spline_linear = SplineLinear(input_dim * num_grids, output_dim, spline_weight_init_scale)
        self.use_base_update = use_base_update
        if use_base_update:
            self.base_activation = base_activation
            self.base_linear = nn.Linear(input_dim, output_dim)
        self._cache_key = None
        self._cache_spec: Optional[ops.KanLayerSpec] = None
        self._grid_key = None
        self._grid_vals = None

    def _grid_params(self):
        """(G, first centre, spacing) of the frozen centre vector, read back (three device round trips) only when IT changes:
        the weights change every training step and must not pay for that."""
        g = self.rbf.grid.detach()
        key = (g.data_ptr(), g._version, g.device)
        if key != self._grid_key:
            G = g.numel()
            gmin = float(g[0])
            step = float((g[-1].double() - g[0].double()) / (G - 1)) if G > 1 else 1.0
            if G > 2:
                ideal = (gmin + step * torch.arange(G, device=g.device, dtype=torch.float64)).to(torch.float32)
                if float((g - ideal).abs().max()) > 1e-4 * abs(step):
                    raise NotImplementedError("non-uniform RBF centres are not supported by the sm_100a path")
            self._grid_key, self._grid_vals = key, (G, gmin, step)
        return self._grid_vals

    def kernel_spec(self) -> ops.KanLayerSpec:
        ps = [self.spline_linear.weight, self.rbf.grid]
        if self.use_base_update:
            ps += [self.base_linear.weight, self.base_linear.bias]
        if self.layernorm is not None:
            ps += [self.layernorm.weight, self.layernorm.bias]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._cache_key:
            G, gmin, step = self._grid_params()
            ln = self.layernorm
            if (G > 8 and ops.rbf_windows_enabled() and ops.tc_supported(L.BASIS_RBF, 8, 0, self.output_dim)
                    and (ln is None or ln.elementwise_affine)):
                self._cache_spec, self._cache_key = self._windowed_spec(G, gmin, step), key
                return self._cache_spec
            packed = ops.pack_kan_weights(self.base_linear.weight if self.use_base_update else None,
                                          self.spline_linear.weight, None, self.input_dim, self.output_dim, G)
            packed_tc = None
            if ops.tc_supported(L.BASIS_RBF, G, 0, self.output_dim):
                packed_tc = ops.pack_kan_weights_tc(self.base_linear.weight if self.use_base_update else None,
                                                    self.spline_linear.weight, None, self.input_dim, self.output_dim, G)
            ln = self.layernorm
            if ln is not None and abs(ln.eps - 1e-5) > 1e-12:
                raise NotImplementedError("LayerNorm eps other than 1e-5 is not supported")
            self._cache_spec = ops.KanLayerSpec(
                L.BASIS_RBF, self.input_dim, self.output_dim, G, 0, gmin, step, 1.0 / float(self.rbf.denominator), packed,
                base_bias=self.base_linear.bias.detach() if self.use_base_update else None,
                ln_weight=None if ln is None else ln.weight.detach(),
                ln_bias=None if ln is None else ln.bias.detach(), packed_w_tc=packed_tc)
            self._cache_key = key
        return self._cache_spec

    def _windowed_spec(self, G: int, gmin: float, step: float) -> ops.KanLayerSpec:
        """OPT-IN (ops.set_rbf_windows / KAGNN_RBF_WINDOWS=1; see the accuracy note there).  More than eight centres (the reference
        searches num_grids up to 32, node_classification/one_experiment.py:42): equally
        spaced Gaussians are shift invariant, phi_{8w+j}(z) = phi_j(z - 8 w step), so the layer is evaluated by the 8-centre
        tensor-core kernels over ``windows`` copies of the input.  With a LayerNorm the copies are exact duplicates (the row
        statistics of [x | x | ..] are those of x) and the shift sits in the copy's LayerNorm bias, beta - 8 w step; without one
        the input itself is shifted.  Copy w carries the spline weights of centres 8w .. 8w+7 (zero beyond G), copy 0 the SiLU
        base weights (which see the raw x)."""
        with torch.no_grad():
            w = (G + 7) // 8
            out_f, in_f = self.output_dim, self.input_dim
            dev = self.spline_linear.weight.device
            virt_base, virt_spline, _ = windowed_weights(self.base_linear.weight if self.use_base_update else None,
                                                         self.spline_linear.weight.view(out_f, in_f, G), None, w)
            ln = self.layernorm
            ln_w = ln_b = None
            shift = 8.0 * step
            if ln is not None:
                if abs(ln.eps - 1e-5) > 1e-12:
                    raise NotImplementedError("LayerNorm eps other than 1e-5 is not supported")
                ln_w, ln_b = windowed_layernorm(ln.weight, ln.bias, w, shift)
                shift = 0.0
            packed = ops.pack_kan_weights(virt_base, virt_spline.view(out_f, -1), None, w * in_f, out_f, 8)
            packed_tc = ops.pack_kan_weights_tc(virt_base, virt_spline.view(out_f, -1), None, w * in_f, out_f, 8)
        return ops.KanLayerSpec(L.BASIS_RBF, w * in_f, out_f, 8, 0, gmin, step, 1.0 / float(self.rbf.denominator), packed,
                                base_bias=self.base_linear.bias.detach() if self.use_base_update else None,
                                ln_weight=ln_w, ln_bias=ln_b, packed_w_tc=packed_tc, windows=w, window_shift=shift, virt_spline=virt_spline)

    def kernel_specs(self) -> List[ops.KanLayerSpec]:
        return [self.kernel_spec()]

    def forward(self, x: Tensor, use_layernorm: bool = True) -> Tensor:
        if not use_layernorm and self.layernorm is not None:
            raise NotImplementedError("use_layernorm=False at call time is never used by the reference models")
        lead = x.shape[:-1]
        if _module_backend_guard(x, self.parameters(), grad_ok=True):
            return autograd.fastkan_layer(self, x.reshape(-1, self.input_dim)).view(*lead, self.output_dim)
        y = ops.fused_layer(ops.AggSpec(L.AGG_NONE, x.reshape(-1, self.input_dim).to(torch.float32)),
                            x.numel() // self.input_dim, [self.kernel_spec()])
        return y.view(*lead, self.output_dim)


class FastKAN(nn.Module):
    """Drop-in for ``fastkan.FastKAN``: FastKANLayers back to back, one launch for the chain."""

    def __init__(self, layers_hidden: List[int], grid_min: float = -2., grid_max: float = 2., num_grids: int = 8,
                 use_base_update: bool = True, base_activation=F.silu, spline_weight_init_scale: float = 0.1) -> None:
        super().__init__()
        self.layers = nn.ModuleList([
            FastKANLayer(i, o, grid_min=grid_min, grid_max=grid_max, num_grids=num_grids, use_base_update=use_base_update,
                         base_activation=base_activation, spline_weight_init_scale=spline_weight_init_scale)
            for i, o in zip(layers_hidden[:-1], layers_hidden[1:])])

    def kernel_specs(self) -> List[ops.KanLayerSpec]:
        return [lay.kernel_spec() for lay in self.layers]

    def forward(self, x: Tensor) -> Tensor:
        if _module_backend_guard(x, self.parameters(), grad_ok=True):
            for layer in self.layers:               # one launch per layer: every layer's input is kept for its backward
                x = layer(x)
            return x
        return chain_forward(self, x)
