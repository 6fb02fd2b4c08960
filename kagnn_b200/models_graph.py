"""Graph-classification models ``KAGIN`` / ``FASTKAGIN`` / ``KAGCN`` / ``FASTKAGCN`` with the positional
constructor signatures, module tree (``conv``, ``bn``, ``kan`` / ``readout``) and ``forward(data)`` of
graph_classification/models.py:95-265.  ``data`` is duck-typed: ``.x``, ``.edge_index``, ``.batch`` (sorted
graph id per node) and optionally ``.num_graphs``.

Eval-mode plan: GIN = one launch per layer (gather -> KAN chain -> BN affine); GCN = KAN_1, then per layer
[aggregate + bias -> SiLU -> KAN_{l+1}]; readout = one launch [segment pool -> KAN chain]; then one ``log_softmax`` launch over
the (graphs x classes) result."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import autograd, ops
from .conv import FASTKAGAT_Layer, FASTKAGCN_Layer, GINConv, KAGAT_Layer, KAGCN_Layer, make_fastkan, make_kan
from .ekan import _module_backend_guard, eval_mode_detach_notice
from .graph import get_graph
from .models_node import _BNFold, bn_is_foldable, bn_unfused, bn_dropout

Tensor = torch.Tensor


def _num_graphs(data) -> int:
    ng = getattr(data, "num_graphs", None)
    if ng is not None:
        return int(ng)
    return int(data.batch.max()) + 1 if data.batch.numel() else 0


def pooled_readout(x: Tensor, batch: Tensor, num_graphs: int, readout: nn.Module, mean: bool, needs_grad: bool = False) -> Tensor:
    """global_add_pool / global_mean_pool (graph_classification/models.py:117,192) fused with the readout KAN."""
    if needs_grad:
        return readout(autograd.pool(x, batch, num_graphs, mean))
    ptr = ops.segment_ptr(batch, num_graphs)
    agg = ops.AggSpec(L.AGG_SEGMENT_MEAN if mean else L.AGG_SEGMENT_SUM, x, rowptr=ptr)
    specs = readout.kernel_specs()
    out = None
    for i in range(0, len(specs), L.MAX_LAYERS):
        out = ops.fused_layer(agg, num_graphs, specs[i:i + L.MAX_LAYERS])
        agg = ops.AggSpec(L.AGG_NONE, out)
    return out


def _bn_unfused(x: Tensor, bn: nn.BatchNorm1d) -> Tensor:
    return bn_unfused(x, bn)


def _eval_without_no_grad(model: nn.Module, data) -> Tensor:
    """model.eval() with autograd enabled (graph_classification_utils.py:57-72): inference plan, result detached."""
    eval_mode_detach_notice(data.x)
    with torch.no_grad():
        return model.forward(data)


class _GINGraphModel(nn.Module):
    """conv -> bn -> dropout, xL; add-pool; KAN readout (shared by KAGIN / FASTKAGIN here and in models_regr)."""
    log_softmax = True

    def _init_common(self, gnn_layers: int, hidden_dim: int, dropout: float):
        self.n_layers = gnn_layers
        self.bn = nn.ModuleList(nn.BatchNorm1d(hidden_dim) for _ in range(gnn_layers))
        self.dropout = nn.Dropout(dropout)
        self._folds = [_BNFold() for _ in range(gnn_layers)]

    def _fusable(self) -> bool:
        return ((not self.training) or self.dropout.p == 0.0) and all(bn_is_foldable(b) for b in self.bn)

    def _conv(self, i: int, x: Tensor, g, post, extra):
        return self.conv[i](x, g, post=post)

    def _message_passing(self, x: Tensor, g, extra=None, needs_grad: bool = False) -> Tensor:
        fus = self._fusable() and not needs_grad
        for i in range(self.n_layers):
            if fus:
                x = self._conv(i, x, g, self._folds[i].get(self.bn[i]), extra)
            else:
                x = bn_dropout(self._conv(i, x, g, None, extra), self.bn[i], self.dropout, needs_grad)
        return x

    def forward(self, data) -> Tensor:
        x = data.x
        needs_grad = _module_backend_guard(x, self.parameters(), grad_ok=True)
        if needs_grad and not self.training:
            return _eval_without_no_grad(self, data)
        x = x.to(torch.float32)
        g = get_graph(data.edge_index, x.size(0))
        x = self._message_passing(x, g, needs_grad=needs_grad)
        x = pooled_readout(x, data.batch, _num_graphs(data), self.kan, mean=False, needs_grad=needs_grad)
        if not self.log_softmax:
            return x
        return autograd.log_softmax(x) if needs_grad else ops.log_softmax(x)


class KAGIN(_GINGraphModel):
    def __init__(self, gnn_layers, num_features, hidden_dim, num_classes, hidden_layers, grid_size, spline_order, dropout):
        super().__init__()
        self.conv = nn.ModuleList(
            GINConv(make_kan(num_features if i == 0 else hidden_dim, hidden_dim, hidden_dim, hidden_layers, grid_size, spline_order))
            for i in range(gnn_layers))
        self._init_common(gnn_layers, hidden_dim, dropout)
        self.kan = make_kan(hidden_dim, hidden_dim, num_classes, hidden_layers, grid_size, spline_order)


class FASTKAGIN(_GINGraphModel):
    def __init__(self, gnn_layers, num_features, hidden_dim, num_classes, hidden_layers, grid_size, dropout):
        super().__init__()
        self.conv = nn.ModuleList(
            GINConv(make_fastkan(num_features if i == 0 else hidden_dim, hidden_dim, hidden_dim, hidden_layers, grid_size))
            for i in range(gnn_layers))
        self._init_common(gnn_layers, hidden_dim, dropout)
        self.kan = make_fastkan(hidden_dim, hidden_dim, num_classes, hidden_layers, grid_size)


class _GCNGraphModel(nn.Module):
    """(conv -> silu -> dropout) xL; pool; 1-layer KAN readout (KAGCN / FASTKAGCN)."""
    log_softmax = True
    mean_pool = True

    def _message_passing(self, x: Tensor, g, needs_grad: bool = False) -> Tensor:
        n = x.size(0)
        if self.n_layers == 0:
            return x
        if needs_grad:
            for i in range(self.n_layers):
                c = self.conv[i]
                x = self.dropout(autograd.silu(autograd.gcn_aggregate(c.transform(x), c.bias, g)))
            return x
        drop_off = (not self.training) or self.dropout.p == 0.0
        if not drop_off:
            for i in range(self.n_layers):
                c = self.conv[i]
                act = ops.Affine(shift=None if c.bias is None else c.bias.detach(), act=L.ACT_SILU)
                x = self.dropout(c.aggregate_transformed(c.transform(x).to(torch.float32), g, extra=act))
            return x
        w, sw = g.gcn_weights()
        t = self.conv[0].transform(x)
        h = None
        for i in range(self.n_layers):
            c = self.conv[i]
            pre = ops.Affine(shift=None if c.bias is None else c.bias.detach(), act=L.ACT_SILU)
            agg = ops.AggSpec(L.AGG_WEIGHTED, t, g.rowptr, g.col, edge_weight=w, self_weight=sw)
            if i + 1 < self.n_layers:
                t = ops.fused_layer(agg, n, self.conv[i + 1].lin.kernel_specs(), pre=pre)
            else:
                h = ops.fused_layer(agg, n, [], pre=pre)
        return h

    def forward(self, data) -> Tensor:
        x = data.x
        needs_grad = _module_backend_guard(x, self.parameters(), grad_ok=True)
        if needs_grad and not self.training:
            return _eval_without_no_grad(self, data)
        x = self._encode(x)
        g = get_graph(data.edge_index, x.size(0))
        x = self._message_passing(x, g, needs_grad=needs_grad)
        x = pooled_readout(x, data.batch, _num_graphs(data), self.readout, mean=self.mean_pool, needs_grad=needs_grad)
        if not self.log_softmax:
            return x
        return autograd.log_softmax(x) if needs_grad else ops.log_softmax(x)

    def _encode(self, x: Tensor) -> Tensor:
        return x.to(torch.float32)


class KAGCN(_GCNGraphModel):
    def __init__(self, gnn_layers, num_features, hidden_dim, num_classes, grid_size, spline_order, dropout):
        super().__init__()
        self.n_layers = gnn_layers
        self.conv = nn.ModuleList(
            KAGCN_Layer(num_features if i == 0 else hidden_dim, hidden_dim, grid_size, spline_order) for i in range(gnn_layers))
        self.readout = make_kan(hidden_dim, hidden_dim, num_classes, 1, grid_size, spline_order)
        self.dropout = nn.Dropout(p=dropout)


class FASTKAGCN(_GCNGraphModel):
    def __init__(self, gnn_layers, num_features, hidden_dim, num_classes, grid_size, dropout):
        super().__init__()
        self.n_layers = gnn_layers
        self.conv = nn.ModuleList(
            FASTKAGCN_Layer(num_features if i == 0 else hidden_dim, hidden_dim, grid_size) for i in range(gnn_layers))
        self.readout = make_fastkan(hidden_dim, hidden_dim, num_classes, 1, grid_size)
        self.dropout = nn.Dropout(p=dropout)


class _GATGraphModel(_GCNGraphModel):
    """(GAT conv -> silu -> dropout) xL; ADD pool; 1-layer KAN read-out (graph_classification/models.py:194-216, 266-288)."""
    mean_pool = False

    def _message_passing(self, x: Tensor, g, needs_grad: bool = False) -> Tensor:
        if needs_grad:
            for i in range(self.n_layers):
                x = self.dropout(autograd.silu(self.conv[i](x, g)))        # conv = projection + attention, each with a library backward
            return x
        for i in range(self.n_layers):
            c = self.conv[i]
            x = c(x, g, extra=ops.Affine(shift=c.bias.detach(), act=L.ACT_SILU))
            if self.training and self.dropout.p > 0.0:
                x = self.dropout(x)
        return x


class KAGAT(_GATGraphModel):
    def __init__(self, gnn_layers, num_features, hidden_dim, num_classes, grid_size, spline_order, dropout, heads):
        super().__init__()
        self.n_layers = gnn_layers
        self.conv = nn.ModuleList(
            KAGAT_Layer(num_features if i == 0 else hidden_dim * heads, hidden_dim, heads, grid_size, spline_order) for i in range(gnn_layers))
        self.readout = make_kan(hidden_dim * heads, hidden_dim, num_classes, 1, grid_size, spline_order)
        self.dropout = nn.Dropout(p=dropout)


class FASTKAGAT(_GATGraphModel):
    def __init__(self, gnn_layers, num_features, hidden_dim, num_classes, grid_size, dropout, heads):
        super().__init__()
        self.n_layers = gnn_layers
        self.heads = heads
        self.conv = nn.ModuleList(
            FASTKAGAT_Layer(num_features if i == 0 else hidden_dim * heads, hidden_dim, heads=heads, grid_size=grid_size) for i in range(gnn_layers))
        self.readout = make_fastkan(hidden_dim * heads, hidden_dim, num_classes, 1, grid_size)
        self.dropout = nn.Dropout(p=dropout)
