"""Message-passing layers with the PyG call surface the reference's models rely on, executed by the fused
sm_100a kernel.

torch_geometric (pinned 2.5.3 by the reference, ``requirements.txt:4``) is not available in this environment, so
``GCNConv`` / ``GINConv`` / ``GINEConv`` here are self-contained modules that keep PyG's attribute and
``state_dict`` names (``lin``, ``bias``, ``nn``, ``eps``), constructor arguments and ``forward`` signatures:

* ``GCNConv(x, edge_index, edge_weight=None)``: ``out = D^-1/2 (A'+I) D^-1/2 . lin(x) + bias``
* ``GINConv(x, edge_index, size=None)``:        ``out = nn((1+eps) x_i + sum_j x_j)``
* ``GINEConv(x, edge_index, edge_attr)``:       ``out = nn((1+eps) x_i + sum_j relu(x_j + e_ji))``

and the KAN-ised subclasses of node_classification_clean/models.py:27-92, graph_classification/models.py:153-243
and graph_regression/models.py:162-216 (``KANLayer``, ``KAGCNConv``, ``GIKANLayer``, ``FKANLayer``,
``FASTKAGCNConv``, ``GIFASTKANLayer``, ``KAGCN_Layer``, ``FASTKAGCN_Layer``).

Each layer is usable on its own (one launch for GIN/GINE, two for GCN: KAN then aggregation); the model classes
additionally fuse across layer boundaries (models_*.py).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib as L
from . import autograd, ops
from .ekan import KAN, KANLinear, _module_backend_guard
from .fastkan import FastKAN, FastKANLayer
from .graph import GraphCSR, get_graph

Tensor = torch.Tensor


def make_kan(num_features, hidden_dim, out_dim, hidden_layers, grid_size, spline_order):
    """node_classification_clean/models.py:19-21."""
    sizes = [num_features] + [hidden_dim] * (hidden_layers - 1) + [out_dim]
    return KAN(layers_hidden=sizes, grid_size=grid_size, spline_order=spline_order)


def make_fastkan(num_features, hidden_dim, out_dim, hidden_layers, grid_size):
    """node_classification_clean/models.py:23-25."""
    sizes = [num_features] + [hidden_dim] * (hidden_layers - 1) + [out_dim]
    return FastKAN(layers_hidden=sizes, num_grids=grid_size)


def _is_kan(m) -> bool:
    return hasattr(m, "kernel_specs")


class _MessagePassing(nn.Module):
    """The slice of PyG's MessagePassing surface that callers of these layers touch."""
    aggr = "add"
    flow = "source_to_target"
    node_dim = -2

    def _graph(self, x: Tensor, edge_index) -> GraphCSR:
        if isinstance(edge_index, GraphCSR):
            return edge_index
        return get_graph(edge_index, x.size(0))


class GCNConv(_MessagePassing):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, **kwargs):
        super().__init__()
        if kwargs.get("improved") or kwargs.get("cached") or kwargs.get("normalize") is False or kwargs.get("add_self_loops") is False:
            raise NotImplementedError("only GCNConv's default options are used by the reference and implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def reset_parameters(self):
        if hasattr(self.lin, "reset_parameters"):
            self.lin.reset_parameters()
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def transform(self, x: Tensor) -> Tensor:
        """``self.lin(x)``: a KAN (one launch) or whatever module the user plugged in."""
        return self.lin(x)

    def aggregate_transformed(self, h: Tensor, graph: GraphCSR, edge_weight: Optional[Tensor] = None,
                              out: Optional[Tensor] = None, extra: Optional[ops.Affine] = None) -> Tensor:
        w, sw = graph.gcn_weights(edge_weight)
        pre = extra if extra is not None else (ops.Affine(shift=self.bias.detach()) if self.bias is not None else None)
        agg = ops.AggSpec(L.AGG_WEIGHTED, h, graph.rowptr, graph.col, edge_weight=w, self_weight=sw)
        return ops.fused_layer(agg, graph.num_nodes, [], pre=pre, agg_out=out)

    def forward(self, x: Tensor, edge_index, edge_weight: Optional[Tensor] = None) -> Tensor:
        needs_grad = _module_backend_guard(x, self.parameters(), grad_ok=True)
        g = self._graph(x, edge_index)
        if needs_grad:
            if edge_weight is not None:
                raise NotImplementedError("the backward of GCNConv with user edge weights is not implemented")
            return autograd.gcn_aggregate(self.transform(x), self.bias, g)
        return self.aggregate_transformed(self.transform(x).to(torch.float32), g, edge_weight)


class GINConv(_MessagePassing):
    def __init__(self, nn: nn.Module, eps: float = 0.0, train_eps: bool = False, **kwargs):
        super().__init__()
        self.nn = nn
        self.initial_eps = float(eps)
        if train_eps:
            self.eps = torch.nn.Parameter(torch.tensor([self.initial_eps]))
        else:
            self.register_buffer("eps", torch.tensor([self.initial_eps]))
        self._eps_host = (None, None)

    def reset_parameters(self):
        for m in self.nn.modules():
            if m is not self.nn and hasattr(m, "reset_parameters"):
                m.reset_parameters()
        self.eps.data.fill_(self.initial_eps)

    def eps_value(self) -> float:
        key = (self.eps.data_ptr(), self.eps._version)
        if self._eps_host[0] != key:
            self._eps_host = (key, float(self.eps.detach().cpu()))
        return self._eps_host[1]

    def _agg_spec(self, x: Tensor, g: GraphCSR, x_halo: Optional[Tensor] = None, peer_x: Optional[Tensor] = None,
                  rows_per_rank: int = 0, halo_need: Optional[Tensor] = None, halo_flags: Optional[Tensor] = None,
                  halo_epoch: int = 0, reserve_sms: int = 0, push_y: Optional[Tensor] = None, push_ld: int = 0,
                  push_mask: Optional[Tensor] = None) -> ops.AggSpec:
        return ops.AggSpec(L.AGG_GIN, x, g.rowptr, g.col, self_scale=1.0 + self.eps_value(), x_halo=x_halo, peer_x=peer_x,
                           rows_per_rank=rows_per_rank, halo_need=halo_need, halo_flags=halo_flags, halo_epoch=halo_epoch,
                           reserve_sms=reserve_sms, push_y=push_y, push_ld=push_ld, push_mask=push_mask)

    def forward(self, x: Tensor, edge_index, size=None, out: Optional[Tensor] = None,
                post: Optional[ops.Affine] = None, **agg_kw) -> Tensor:
        needs_grad = _module_backend_guard(x, self.parameters(), grad_ok=True)
        g = self._graph(x, edge_index)
        if needs_grad:
            if post is not None or out is not None or agg_kw:
                raise NotImplementedError("fused epilogues / sharded inputs are inference-only")
            if self.eps.requires_grad:
                raise NotImplementedError("train_eps=True has no backward yet (the reference models keep eps fixed)")
            return self.nn(autograd.gin_aggregate(x, g, 1.0 + self.eps_value()))
        agg = self._agg_spec(x.to(torch.float32), g, **agg_kw)
        if _is_kan(self.nn):
            specs = self.nn.kernel_specs()
            if len(specs) <= L.MAX_LAYERS:
                return ops.fused_layer(agg, g.num_nodes, specs, post=post, out=out)
        h = self.nn(ops.fused_layer(agg, g.num_nodes, []))
        if post is not None or out is not None:
            raise NotImplementedError("fused epilogue needs a KAN/FastKAN as GINConv.nn")
        return h


class GINEConv(GINConv):
    def __init__(self, nn: nn.Module, eps: float = 0.0, train_eps: bool = False, edge_dim: Optional[int] = None, **kwargs):
        super().__init__(nn, eps, train_eps)
        if edge_dim is not None:
            raise NotImplementedError("edge_dim is never used by the reference (graph_regression/models.py:98)")
        self.lin = None

    def _agg_spec(self, x: Tensor, g: GraphCSR, edge_feat: Tensor = None, edge_row: Tensor = None,
                  x_halo: Optional[Tensor] = None) -> ops.AggSpec:
        return ops.AggSpec(L.AGG_GINE, x, g.rowptr, g.col, self_scale=1.0 + self.eps_value(), edge_feat=edge_feat,
                           edge_row=edge_row, x_halo=x_halo)

    def forward(self, x: Tensor, edge_index, edge_attr: Tensor = None, size=None, out: Optional[Tensor] = None,
                post: Optional[ops.Affine] = None, edge_row: Optional[Tensor] = None) -> Tensor:
        """``edge_attr`` is (E, F) in COO order (PyG semantics).  Internal callers may instead pass a small
        feature table plus ``edge_row`` (CSR-ordered row ids into it)."""
        if edge_attr is None:
            raise ValueError("GINEConv needs edge_attr")
        g = self._graph(x, edge_index)
        params = list(self.parameters())
        if autograd.grad_needed(x, params) or (torch.is_grad_enabled() and edge_attr.requires_grad):
            _module_backend_guard(x, params, grad_ok=True)       # backward: kagnn_gine_bwd (autograd.gine_aggregate), GPU-validated
            if edge_row is not None or out is not None or post is not None:
                raise NotImplementedError("edge-feature tables / fused epilogues are inference-only")
            if self.eps.requires_grad:
                raise NotImplementedError("train_eps=True has no backward yet (the reference models keep eps fixed)")
            if edge_attr.dim() != 2 or edge_attr.size(0) != g.csr.nnz or edge_attr.size(1) != x.size(1):
                raise ValueError("Node and edge feature dimensionalities do not match")
            return self.nn(autograd.gine_aggregate(x, edge_attr, g, 1.0 + self.eps_value()))
        if edge_row is None:
            if edge_attr.size(0) != g.csr.nnz or edge_attr.size(-1) != x.size(-1):
                raise ValueError("Node and edge feature dimensionalities do not match")
            edge_row = g.perm
        return super().forward(x, g, out=out, post=post, edge_feat=edge_attr.to(torch.float32), edge_row=edge_row)


# ---- KAN-ised layers (same names / signatures as the reference) ----------------------------------------------------
class KANLayer(KANLinear):
    def __init__(self, input_dim, output_dim, grid_size=4, spline_order=3):
        super().__init__(in_features=input_dim, out_features=output_dim, grid_size=grid_size, spline_order=spline_order)


class FKANLayer(FastKANLayer):
    def __init__(self, input_dim, output_dim, num_grids=4):
        super().__init__(input_dim=input_dim, output_dim=output_dim, num_grids=num_grids)
        self.input_dim, self.output_dim, self.num_grids = input_dim, output_dim, num_grids

    def reset_parameters(self):
        self.__init__(self.input_dim, self.output_dim, self.num_grids)


class KAGCNConv(GCNConv):
    def __init__(self, in_feat: int, out_feat: int, grid_size: int = 4, spline_order: int = 3):
        super().__init__(in_feat, out_feat)
        self.lin = KANLayer(in_feat, out_feat, grid_size, spline_order)


class FASTKAGCNConv(GCNConv):
    def __init__(self, in_feat: int, out_feat: int, grid_size: int = 4):
        super().__init__(in_channels=in_feat, out_channels=out_feat)
        self.grid_size = grid_size
        self.lin = FKANLayer(in_feat, out_feat, num_grids=grid_size)


class GIKANLayer(GINConv):
    def __init__(self, in_feat: int, out_feat: int, grid_size: int = 4, spline_order: int = 3, hidden_dim: int = 16,
                 nb_layers: int = 2):
        super().__init__(make_kan(in_feat, hidden_dim, out_feat, nb_layers, grid_size, spline_order))


class GIFASTKANLayer(GINConv):
    def __init__(self, in_feat: int, out_feat: int, grid_size: int = 4, hidden_dim: int = 16, nb_layers: int = 2):
        super().__init__(make_fastkan(in_feat, hidden_dim, out_feat, nb_layers, grid_size))


# graph_classification / graph_regression spell the GCN layers differently
KAGCN_Layer = KAGCNConv
FASTKAGCN_Layer = FASTKAGCNConv


class GATConv(_MessagePassing):
    """PyG ``GATConv`` with the options the reference uses (``GATConv(in, out, heads)``: concat=True, negative_slope=0.2,
    dropout=0, add_self_loops=True, bias=True, edge_dim=None) and PyG 2.5's parameter names: one shared projection ``lin``
    (which the KAN-ised subclasses replace), ``att_src`` / ``att_dst`` (1, heads, out_channels), ``bias`` (heads * out_channels).
    Forward = projection launch, attention launches (``ops.gat_attention``), one WEIGHTED aggregation launch per head on the
    head's column slice.  Under autograd the attention + aggregation is one ``autograd.Function`` (``autograd.gat_attend``)."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True, negative_slope: float = 0.2,
                 dropout: float = 0.0, add_self_loops: bool = True, edge_dim=None, fill_value="mean", bias: bool = True, **kwargs):
        super().__init__()
        if not concat or dropout != 0.0 or not add_self_loops or edge_dim is not None or not bias or kwargs:
            raise NotImplementedError("only GATConv's default options (as the reference uses them) are implemented")
        self.in_channels, self.out_channels, self.heads, self.negative_slope = in_channels, out_channels, heads, negative_slope
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    def reset_parameters(self):
        if hasattr(self.lin, "reset_parameters"):
            self.lin.reset_parameters()
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)
        nn.init.zeros_(self.bias)

    def attend_and_aggregate(self, h: Tensor, graph: GraphCSR, out: Optional[Tensor] = None,
                             extra: Optional[ops.Affine] = None) -> Tensor:
        """``extra``: per-column affine / activation applied after the aggregation INSTEAD of the plain bias (the models fold
        bias + eval BatchNorm or bias + SiLU into it); None = ``+ bias``."""
        n, hc = h.shape
        hd, c = self.heads, self.out_channels
        w, sw = ops.gat_attention(h, graph.csr, self.att_src, self.att_dst, hd, self.negative_slope)
        if out is None:
            out = torch.empty(n, hc, dtype=torch.float32, device=h.device)
        pre = extra if extra is not None else ops.Affine(shift=self.bias.detach())
        for k in range(hd):
            sl = slice(k * c, (k + 1) * c)
            pre_k = ops.Affine(None if pre.scale is None else pre.scale[sl], None if pre.shift is None else pre.shift[sl], pre.act)
            agg = ops.AggSpec(L.AGG_WEIGHTED, h[:, sl], graph.rowptr, graph.col, edge_weight=w[k], self_weight=sw[k])
            ops.fused_layer(agg, n, [], pre=pre_k, agg_out=out[:, sl])
        return out

    def forward(self, x: Tensor, edge_index, edge_attr=None, size=None, out: Optional[Tensor] = None,
                extra: Optional[ops.Affine] = None) -> Tensor:
        if edge_attr is not None:
            raise NotImplementedError("edge_dim is never used by the reference")
        needs_grad = _module_backend_guard(x, self.parameters(), grad_ok=True)
        g = self._graph(x, edge_index)
        if needs_grad:
            # training: projection (its own autograd Function) -> attention + aggregation with a library backward; the callers'
            # fused epilogues (``extra``) belong to the inference plan
            if extra is not None or out is not None:
                raise NotImplementedError("fused epilogues are part of the inference plan; under autograd call conv(x, edge_index)")
            return autograd.gat_attend(self.lin(x), self.att_src, self.att_dst, self.bias, g, self.heads, self.negative_slope)
        return self.attend_and_aggregate(self.lin(x).to(torch.float32), g, out=out, extra=extra)


class KAGATConv(GATConv):
    """node_classification_clean/models.py:39-46, graph_classification/models.py:165-172 (``KAGAT_Layer``)."""

    def __init__(self, in_feat: int, out_feat: int, heads: int, grid_size: int = 4, spline_order: int = 3):
        super().__init__(in_feat, out_feat, heads)
        self.lin = KANLayer(in_feat, out_feat * heads, grid_size, spline_order)


class FASTKAGATConv(GATConv):
    """node_classification_clean/models.py:76-83, graph_classification/models.py:236-243 (``FASTKAGAT_Layer``)."""

    def __init__(self, in_feat: int, out_feat: int, heads: int, grid_size: int = 4):
        super().__init__(in_feat, out_feat, heads)
        self.grid_size = grid_size
        self.lin = FKANLayer(in_feat, out_feat * heads, grid_size)


KAGAT_Layer = KAGATConv
FASTKAGAT_Layer = FASTKAGATConv
