"""Node-level models: ``GKAN_Nodes`` / ``GFASTKAN_Nodes`` with the constructor signature, module tree,
``state_dict`` keys and ``forward(x, edge_index)`` of node_classification_clean/models.py:150-257.

Execution plan in eval mode (BatchNorm folded to an affine, dropout is the identity):

* ``conv_type='gin'``: one launch per layer = CSR gather-sum -> conv's KAN chain -> BatchNorm affine -> store into
  the layer's column slice of the skip-concat buffer; then one launch for ``lay_out`` over the concat buffer.
* ``conv_type='gcn'``: ``out = A_hat . KAN(x) + b`` puts the KAN *before* the aggregation, so the fusion boundary is
  shifted half a layer: launch 0 = KAN_1(x); launch l = [aggregate(A_hat, t_l) + b_l -> BN_l -> store h_l into the
  concat buffer -> KAN_{l+1}(h_l)]; the last launch has no KAN; then ``lay_out``.

* ``conv_type='gat'``: projection (KAN launch) -> attention coefficients -> one weighted aggregation per head with bias +
  BatchNorm folded into it.

In training mode every conv is a chain of ``autograd.Function``s whose forward and backward are library launches, and
``dropout(bn(x))`` is one fused training epilogue (batch statistics + affine + Philox mask, regenerated in the backward:
``autograd.batch_norm_dropout_train``).  Eval-mode forwards of small graphs on unchanged inputs are replayed from a CUDA graph
(``_GraphReplay``)."""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import nn

from . import _lib as L
from . import autograd, ops
from .conv import FASTKAGATConv, FASTKAGCNConv, GATConv, GIFASTKANLayer, GIKANLayer, KAGATConv, KAGCNConv, GCNConv
from .ekan import KANLinear, _module_backend_guard, eval_mode_detach_notice
from .fastkan import FastKANLayer
from .graph import get_graph

Tensor = torch.Tensor


# Small graphs (BASELINE config C1: 2 708 nodes) are bound by launch latency and host work, not by the kernels: six dependent,
# mostly empty launches.  An eval-mode forward that is repeated on the SAME inputs (the reference's validation / test loops call
# the model on the one static graph every epoch) is therefore replayed from a CUDA graph: two eager calls, the third
# captured, later ones replayed.  The key holds the identity AND version of x, edge_index and every parameter / buffer, so any
# in-place change re-captures; the entry keeps x and edge_index alive, so their addresses cannot be recycled.
# KAGNN_AUTO_GRAPH_NODES=0 switches it off.
_AUTO_GRAPH_NODES = int(os.environ.get("KAGNN_AUTO_GRAPH_NODES", "20000"))


class _GraphReplay:
    def __init__(self):
        self.key = None
        self.count = 0
        self.graph = None
        self.out = None
        self.keep = None
        self.failed = False

    def run(self, model, x: Tensor, edge_index: Tensor, eager):
        key = (x.data_ptr(), tuple(x.shape), x.stride(), x._version, x.dtype, edge_index.data_ptr(), tuple(edge_index.shape),
               edge_index._version, ops.get_precision(),
               tuple((t.data_ptr(), t._version) for t in list(model.parameters()) + list(model.buffers())))
        if key != self.key:
            self.key, self.count, self.graph, self.out, self.keep = key, 0, None, None, (x, edge_index)
        self.count += 1
        if self.count < 3 or self.failed:               # two eager calls first: a validation + test pair per epoch never pays a capture
            return eager()
        if self.graph is None:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = eager()
                self.graph, self.out = g, out
            except Exception:                           # anything that cannot be captured: stay eager for good
                self.failed, self.graph, self.out = True, None, None
                torch.cuda.synchronize()
                return eager()
        self.graph.replay()
        return self.out.clone()


class _BNFold:
    """eval-mode BatchNorm1d as (scale, shift), optionally folding a preceding bias; cached per parameter version."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, bn: nn.BatchNorm1d, pre_bias: Optional[Tensor] = None):
        ts = [bn.weight, bn.bias, bn.running_mean, bn.running_var] + ([pre_bias] if pre_bias is not None else [])
        key = tuple((t.data_ptr(), t._version) for t in ts if t is not None)
        if key != self._key:
            with torch.no_grad():
                inv = torch.rsqrt(bn.running_var.double() + bn.eps)
                scale = inv * (bn.weight.double() if bn.weight is not None else 1.0)
                mean = bn.running_mean.double()
                if pre_bias is not None:
                    mean = mean - pre_bias.double()
                shift = (bn.bias.double() if bn.bias is not None else 0.0) - mean * scale
                self._val = ops.Affine(scale.float().contiguous(), shift.float().contiguous())
            self._key = key
        return self._val


def bn_is_foldable(bn: nn.BatchNorm1d) -> bool:
    return (not bn.training) and bn.track_running_stats and bn.running_mean is not None


def _bn_eval(x: Tensor, bn: nn.BatchNorm1d) -> Tensor:
    """Eval-mode BatchNorm outside a fused launch (only reached when dropout is active in training mode of a model whose
    BatchNorm layers are individually in eval mode): y = x * scale + shift through the aggregation-free fused launch."""
    fold = _BNFold().get(bn)
    return ops.fused_layer(ops.AggSpec(L.AGG_NONE, x), x.size(0), [], pre=fold)


def bn_dropout(x: Tensor, bn: nn.BatchNorm1d, dropout: nn.Dropout, needs_grad: bool = False) -> Tensor:
    """``dropout(bn(x))`` of the training loops (nc/models.py:197-198).  Batch-statistics BatchNorm followed by an active dropout
    runs as the fused epilogue (kagnn_bn_dropout_train_fwd: Philox mask, never stored); every other combination as before."""
    if dropout.training and dropout.p > 0.0 and (bn.training or bn.running_mean is None) and dropout.p < 1.0:
        return autograd.batch_norm_dropout_train(x, bn, dropout.p, needs_grad)
    return dropout(bn_unfused(x, bn, needs_grad))


def bn_unfused(x: Tensor, bn: nn.BatchNorm1d, needs_grad: bool = False) -> Tensor:
    """BatchNorm1d as its own launches: batch statistics (training) or the folded affine (eval)."""
    batch_stats = bn.training or bn.running_mean is None
    if needs_grad:
        if not batch_stats:
            raise NotImplementedError("an eval-mode BatchNorm1d inside a differentiated forward has no backward here")
        return autograd.batch_norm_train(x, bn)
    return ops.batchnorm_forward(x, bn) if batch_stats else _bn_eval(x, bn)


class _NodeModel(nn.Module):
    """Shared forward of GKAN_Nodes / GFASTKAN_Nodes."""
    convs: nn.ModuleList
    bns: nn.ModuleList

    def _init_common(self, skip: bool, dropout: float):
        self.skip = skip
        self.dropout = nn.Dropout(dropout)
        self._folds = [_BNFold() for _ in self.bns]

    def _fusable(self) -> bool:
        drop_off = (not self.training) or self.dropout.p == 0.0
        return drop_off and all(bn_is_foldable(bn) for bn in self.bns)

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        needs_grad = _module_backend_guard(x, self.parameters(), grad_ok=True)
        if needs_grad and not self.training:
            # model.eval() without torch.no_grad() (the reference's val()/test() loops of graph_classification_utils.py:57-72 do
            # that): the fused inference plan, result detached from autograd
            eval_mode_detach_notice(x)
            with torch.no_grad():
                return self.forward(x, edge_index)
        x = x.to(torch.float32)
        n, f = x.shape
        g = get_graph(edge_index, n)
        n_mp = len(self.convs)
        if n_mp == 0:
            return self.lay_out(x)
        hid = self.bns[0].num_features
        if needs_grad or not self._fusable():
            return self._forward_unfused(x, g, needs_grad)
        if (0 < n <= _AUTO_GRAPH_NODES and x.is_cuda and x.device.index == torch.cuda.current_device()
                and not torch.cuda.is_current_stream_capturing() and not getattr(self, "_in_replay", False)):
            if not hasattr(self, "_replay"):
                object.__setattr__(self, "_replay", _GraphReplay())
            object.__setattr__(self, "_in_replay", True)
            try:
                return self._replay.run(self, x, edge_index, lambda: self.forward(x, edge_index))
            finally:
                object.__setattr__(self, "_in_replay", False)
        # skip concat (models.py:196-201) without any copy: every layer writes its column slice of the hidden buffer and
        # lay_out reads two-part rows [x | h_1 .. h_L] (KagnnAggregate.x_head)
        buf = torch.empty(n, n_mp * hid, dtype=torch.float32, device=x.device) if self.skip else None
        cur = x
        is_gcn = isinstance(self.convs[0], GCNConv)
        is_gat = isinstance(self.convs[0], GATConv)
        t = self.convs[0].transform(cur) if is_gcn else None
        for l, (conv, bn) in enumerate(zip(self.convs, self.bns)):
            dst = buf[:, l * hid: (l + 1) * hid] if self.skip else torch.empty(n, hid, dtype=torch.float32, device=x.device)
            if is_gat:
                # projection (KAN launch) -> attention -> one weighted aggregation per head; bias + eval BatchNorm folded into it
                conv(cur, g, out=dst, extra=self._folds[l].get(bn, conv.bias.detach()))
            elif is_gcn:
                pre = self._folds[l].get(bn, conv.bias.detach() if conv.bias is not None else None)
                nxt = self.convs[l + 1].lin.kernel_specs() if l + 1 < n_mp else []
                w, sw = g.gcn_weights()
                agg = ops.AggSpec(L.AGG_WEIGHTED, t, g.rowptr, g.col, edge_weight=w, self_weight=sw)
                t = ops.fused_layer(agg, n, nxt, pre=pre, agg_out=dst)
            else:
                conv(cur, g, out=dst, post=self._folds[l].get(bn))
            cur = dst
        if not self.skip:
            return self.lay_out(cur)
        return ops.fused_layer(ops.AggSpec(L.AGG_NONE, buf, x_head=x), n, self.lay_out.kernel_specs())

    def _forward_unfused(self, x: Tensor, g, needs_grad: bool = False) -> Tensor:
        """Layer by layer; with ``needs_grad`` every step is an autograd.Function whose backward is a library launch."""
        feats = [x]
        for conv, bn in zip(self.convs, self.bns):
            x = conv(x, g)
            x = bn_dropout(x, bn, self.dropout, needs_grad)
            feats.append(x)
        if self.skip:
            x = torch.cat(feats, dim=1)
        return self.lay_out(x)


class GKAN_Nodes(_NodeModel):
    def __init__(self, conv_type: str, mp_layers: int, num_features: int, hidden_channels: int, num_classes: int,
                 skip: bool = True, grid_size: int = 4, spline_order: int = 3, hidden_layers: int = 2, dropout: float = 0.,
                 heads=4):
        super().__init__()
        if conv_type not in ("gcn", "gin", "gat"):
            raise ValueError("unknown conv_type")
        if conv_type != "gat":
            heads = 1
        self.convs = nn.ModuleList()
        self.bns = nn.ModuleList()
        for i in range(mp_layers):
            fin = num_features if i == 0 else hidden_channels * heads
            if conv_type == "gcn":
                self.convs.append(KAGCNConv(fin, hidden_channels, grid_size, spline_order))
            elif conv_type == "gat":
                self.convs.append(KAGATConv(fin, hidden_channels, heads, grid_size, spline_order))
            else:
                self.convs.append(GIKANLayer(fin, hidden_channels, grid_size, spline_order, hidden_channels, hidden_layers))
            self.bns.append(nn.BatchNorm1d(hidden_channels * heads))
        dim_out = num_features + mp_layers * hidden_channels * heads if skip else hidden_channels * heads
        self.lay_out = KANLinear(dim_out, num_classes, grid_size=grid_size, spline_order=spline_order)
        self._init_common(skip, dropout)


class GFASTKAN_Nodes(_NodeModel):
    def __init__(self, conv_type: str, mp_layers: int, num_features: int, hidden_channels: int, num_classes: int,
                 skip: bool = True, grid_size: int = 4, hidden_layers: int = 2, dropout: float = 0., heads=4):
        super().__init__()
        if conv_type not in ("gcn", "gin", "gat"):
            raise ValueError("unknown conv_type")
        if conv_type != "gat":
            heads = 1
        self.convs = nn.ModuleList()
        self.bns = nn.ModuleList()
        for i in range(mp_layers):
            fin = num_features if i == 0 else hidden_channels * heads
            if conv_type == "gcn":
                self.convs.append(FASTKAGCNConv(fin, hidden_channels, grid_size))
            elif conv_type == "gat":
                self.convs.append(FASTKAGATConv(fin, hidden_channels, heads, grid_size))
            else:
                self.convs.append(GIFASTKANLayer(fin, hidden_channels, grid_size, hidden_channels, hidden_layers))
            self.bns.append(nn.BatchNorm1d(hidden_channels * heads))
        dim_out = num_features + mp_layers * hidden_channels * heads if skip else hidden_channels * heads
        self.lay_out = FastKANLayer(dim_out, num_classes, num_grids=grid_size)
        self._init_common(skip, dropout)
