"""``torch.autograd.Function`` wrappers that give the modules a backward, so that the
reference's untouched training loops -- ``loss.backward()`` at node_classification_clean/utils.py:130 and
graph_classification/graph_classification_utils.py:52 -- run on them.

Every forward below is the same library launch the inference path uses; every backward is a library launch too
(``kagnn_kan_bwd_*``, ``kagnn_batchnorm_train_bwd``, ... in include/kagnn_b200.h, or the forward aggregation kernel on the
TRANSPOSED CSR).  torch contributes the autograd tape, ``torch.cat`` of the skip connection and ``nn.Dropout``'s mask.
Scope of this first backward: B-spline and FastKAN layers, GIN / GINE / GCN aggregation, BatchNorm1d, SiLU, add / mean
pooling, log_softmax.  First derivatives only (``once_differentiable``): ``create_graph=True`` raises instead of returning a
graph-less result."""
from __future__ import annotations

from typing import Optional

import dataclasses

import torch
from torch.autograd.function import once_differentiable

from . import _lib as L
from . import ops

Tensor = torch.Tensor


def grad_needed(x: Tensor, params) -> bool:
    return torch.is_grad_enabled() and ((isinstance(x, torch.Tensor) and x.requires_grad) or any(p.requires_grad for p in params))


def _rowmajor(t: Tensor) -> Tensor:
    """fp32, unit column stride, rows at least one row apart (the cotangent of ``y.sum()`` is an expanded, stride-0 tensor)."""
    t = t.to(torch.float32)
    ok = t.dim() == 2 and (t.size(1) <= 1 or t.stride(1) == 1) and (t.size(0) <= 1 or t.stride(0) >= t.size(1))
    return t if ok else t.contiguous()


class _KanLinearFn(torch.autograd.Function):
    """y = KANLinear(x) (ekan.py:154-162)."""

    @staticmethod
    def forward(ctx, x, base_w, spline_w, scaler, layer):
        spec = layer.kernel_spec()
        x = _rowmajor(x)
        y = ops.fused_layer(ops.AggSpec(L.AGG_NONE, x), x.size(0), [spec])
        ctx.spec = spec                      # keeps the packed weights of THIS forward alive
        ctx.has_scaler = scaler is not None
        ctx.save_for_backward(x, spline_w, scaler if scaler is not None else spline_w)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, spline_w, scaler = ctx.saved_tensors
        scaler = scaler if ctx.has_scaler else None
        dy = _rowmajor(dy)
        dx = ops.kan_bwd_input(ctx.spec, x, dy) if ctx.needs_input_grad[0] else None
        d_base = d_spline = d_scaler = None
        if any(ctx.needs_input_grad[1:4]):
            d_packed = ops.kan_bwd_weights(ctx.spec, x, dy)
            if ctx.spec.windows > 1:
                d_base, d_spline, d_scaler = ops.kan_unpack_windowed_grads(d_packed, ctx.spec, spline_w.size(2))
            else:
                d_base, d_spline, d_scaler = ops.kan_unpack_weight_grads(d_packed, spline_w, scaler)
        return dx, d_base, d_spline, d_scaler, None


def kan_linear(layer, x: Tensor) -> Tensor:
    scaler = layer.spline_scaler if layer.enable_standalone_scale_spline else None
    return _KanLinearFn.apply(x, layer.base_weight, layer.spline_weight, scaler, layer)


class _FastKanLayerFn(torch.autograd.Function):
    """y = FastKANLayer(x) (fastkan.py:76-85): spline_linear(rbf(layernorm(x))) + base_linear(silu(x))."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, spline_w, base_w, base_b, layer):
        spec = layer.kernel_spec()
        x = _rowmajor(x)
        y = ops.fused_layer(ops.AggSpec(L.AGG_NONE, x), x.size(0), [spec])
        ctx.spec, ctx.layer = spec, layer
        ctx.has_ln, ctx.ln_affine, ctx.has_base = layer.layernorm is not None, ln_w is not None, base_w is not None
        ctx.save_for_backward(x, spline_w, ln_w if ln_w is not None else spline_w)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, spline_w, ln_w = ctx.saved_tensors
        ln_w = ln_w if ctx.ln_affine else None
        dy = _rowmajor(dy)
        spec, lay = ctx.spec, ctx.layer
        nw, f = spec.windows, x.size(1)
        if nw > 1:
            # more than eight centres: the virtual layer over `nw` copies of x (fastkan.FastKANLayer._windowed_spec); the gradient of
            # a copied input is the sum over its copies, so are those of the LayerNorm vectors; window w holds centres 8w .. 8w+7
            x = ops.expand_windows(x, nw, spec.window_shift)
            spec = dataclasses.replace(spec, windows=1)
            if ctx.ln_affine:
                ln_w = spec.ln_weight

        def fold(t):                                                     # (.., nw * f) -> (.., f), summed over the windows
            return t if (t is None or nw == 1) else t.reshape(*t.shape[:-1], nw, f).sum(-2)

        stats = ops.layernorm_stats(x) if ctx.has_ln else None          # recomputed, not kept: (rows, 2)
        dx = d_lnw = d_lnb = d_spline = d_base = d_bb = None
        if ctx.needs_input_grad[0] or (ctx.has_ln and any(ctx.needs_input_grad[1:3])):
            dz, dxb = ops.rbf_bwd_input(spec, x, stats, dy)
            if ctx.has_ln:
                dx, d_lnw, d_lnb = ops.layernorm_backward(x, stats, ln_w, dz, dxb, ctx.ln_affine)
                dx, d_lnw, d_lnb = fold(dx), fold(d_lnw), fold(d_lnb)
            else:
                dx = fold(dz)
        if ctx.needs_input_grad[3] or (ctx.has_base and ctx.needs_input_grad[4]):
            d_packed = ops.rbf_bwd_weights(spec, x, stats, dy)
            if nw > 1:
                d_base_v, d_spline_v, _ = ops.kan_unpack_weight_grads(d_packed, spec.virt_spline, None, need_base=ctx.has_base)
                G = spline_w.size(1) // f
                d_spline = d_spline_v.view(lay.output_dim, nw, f, 8).permute(0, 2, 1, 3).reshape(lay.output_dim, f, 8 * nw)[:, :, :G]
                d_spline = d_spline.reshape(lay.output_dim, f * G).contiguous()
                d_base = None if d_base_v is None else d_base_v[:, :f].contiguous()
            else:
                sw3 = spline_w.detach().view(lay.output_dim, lay.input_dim, -1)
                d_base, d_spline3, _ = ops.kan_unpack_weight_grads(d_packed, sw3, None, need_base=ctx.has_base)
                d_spline = d_spline3.view_as(spline_w)
        if ctx.has_base and ctx.needs_input_grad[5]:
            d_bb = ops.column_sums(dy)
        return dx, d_lnw, d_lnb, d_spline, d_base, d_bb, None


def fastkan_layer(layer, x: Tensor) -> Tensor:
    ln = layer.layernorm
    base = layer.base_linear if layer.use_base_update else None
    return _FastKanLayerFn.apply(x, None if ln is None else ln.weight, None if ln is None else ln.bias, layer.spline_linear.weight,
                                 None if base is None else base.weight, None if base is None else base.bias, layer)


class _GinAggFn(torch.autograd.Function):
    """a_i = self_scale * x_i + sum_{j -> i} x_j; backward = the same sum over the reversed edges."""

    @staticmethod
    def forward(ctx, x, graph, self_scale):
        x = _rowmajor(x)
        ctx.graph, ctx.self_scale = graph, float(self_scale)
        agg = ops.AggSpec(L.AGG_GIN, x, graph.rowptr, graph.col, self_scale=ctx.self_scale)
        return ops.fused_layer(agg, graph.num_nodes, [])

    @staticmethod
    @once_differentiable
    def backward(ctx, da):
        gt = ctx.graph.transposed()
        agg = ops.AggSpec(L.AGG_GIN, _rowmajor(da), gt.rowptr, gt.col, self_scale=ctx.self_scale)
        return ops.fused_layer(agg, gt.num_nodes, []), None, None


def gin_aggregate(x: Tensor, graph, self_scale: float) -> Tensor:
    return _GinAggFn.apply(x, graph, self_scale)


class _GineAggFn(torch.autograd.Function):
    """a_i = self_scale * x_i + sum_{j -> i} relu(x_j + e_ji) (PyG GINEConv); edge features in COO order, one row per edge."""

    @staticmethod
    def forward(ctx, x, edge_feat, graph, self_scale):
        x, edge_feat = _rowmajor(x), _rowmajor(edge_feat)
        ctx.graph, ctx.self_scale = graph, float(self_scale)
        ctx.save_for_backward(x, edge_feat)
        agg = ops.AggSpec(L.AGG_GINE, x, graph.rowptr, graph.col, self_scale=ctx.self_scale, edge_feat=edge_feat, edge_row=graph.perm)
        return ops.fused_layer(agg, graph.num_nodes, [])

    @staticmethod
    @once_differentiable
    def backward(ctx, da):
        x, edge_feat = ctx.saved_tensors
        dx, de = ops.gine_backward(x, edge_feat, ctx.graph.edge_index, _rowmajor(da), ctx.self_scale, ctx.needs_input_grad[1])
        return dx, de, None, None


def gine_aggregate(x: Tensor, edge_feat: Tensor, graph, self_scale: float) -> Tensor:
    return _GineAggFn.apply(x, edge_feat, graph, self_scale)


class _GcnAggFn(torch.autograd.Function):
    """out = A_hat h + b with A_hat = D^-1/2 (A + I) D^-1/2 (PyG GCNConv.propagate + bias); dh = A_hat^T d_out."""

    @staticmethod
    def forward(ctx, h, bias, graph):
        h = _rowmajor(h)
        ctx.graph, ctx.has_bias = graph, bias is not None
        w, sw = graph.gcn_weights()
        agg = ops.AggSpec(L.AGG_WEIGHTED, h, graph.rowptr, graph.col, edge_weight=w, self_weight=sw)
        pre = ops.Affine(shift=bias.detach()) if bias is not None else None
        return ops.fused_layer(agg, graph.num_nodes, [], pre=pre)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        dout = _rowmajor(dout)
        dh = db = None
        if ctx.needs_input_grad[0]:
            gt = ctx.graph.transposed()
            w, sw = ctx.graph.gcn_weights_transposed()
            agg = ops.AggSpec(L.AGG_WEIGHTED, dout, gt.rowptr, gt.col, edge_weight=w, self_weight=sw)
            dh = ops.fused_layer(agg, gt.num_nodes, [])
        if ctx.has_bias and ctx.needs_input_grad[1]:
            db = ops.column_sums(dout)
        return dh, db, None


def gcn_aggregate(h: Tensor, bias: Optional[Tensor], graph) -> Tensor:
    return _GcnAggFn.apply(h, bias, graph)


class _GatAttendFn(torch.autograd.Function):
    """PyG GATConv after its projection: out[i] = sum_j alpha_ij h[j] + b per head, alpha = edge-softmax of
    leaky_relu(<h_j, att_src> + <h_i, att_dst>) over the incoming edges and the self loop (node_classification_clean/models.py:39-46).
    Backward: dh = sum_i alpha_ij d out[i] (WEIGHTED aggregation over the reversed edges, one launch per head) + the part through
    the scores (kagnn_gat_bwd), d att_src / d att_dst, d b = column sums."""

    @staticmethod
    def forward(ctx, h, att_src, att_dst, bias, graph, heads, slope):
        h = _rowmajor(h)
        n, hc = h.shape
        c = hc // heads
        w, sw = ops.gat_attention(h, graph.csr, att_src, att_dst, heads, slope)
        out = torch.empty(n, hc, dtype=torch.float32, device=h.device)
        for k in range(heads):
            sl = slice(k * c, (k + 1) * c)
            agg = ops.AggSpec(L.AGG_WEIGHTED, h[:, sl], graph.rowptr, graph.col, edge_weight=w[k], self_weight=sw[k])
            ops.fused_layer(agg, n, [], pre=ops.Affine(shift=bias.detach()[sl]) if bias is not None else None, agg_out=out[:, sl])
        ctx.save_for_backward(h, att_src, att_dst, w, sw)
        ctx.graph, ctx.heads, ctx.slope, ctx.has_bias = graph, heads, slope, bias is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        dout = _rowmajor(dout).contiguous()
        h, att_src, att_dst, w, sw = ctx.saved_tensors
        g, heads = ctx.graph, ctx.heads
        n, hc = h.shape
        c = hc // heads
        gt = g.transposed()
        # the coefficients in the entry order of the reversed graph (an edge keeps its coefficient)
        by_edge = torch.empty_like(w)
        by_edge[:, g.csr.perm.long()] = w
        w_t = by_edge[:, gt.csr.perm.long()].contiguous() if w.size(1) else w
        dh = torch.empty(n, hc, dtype=torch.float32, device=h.device)
        for k in range(heads):
            sl = slice(k * c, (k + 1) * c)
            agg = ops.AggSpec(L.AGG_WEIGHTED, dout[:, sl], gt.rowptr, gt.col, edge_weight=w_t[k], self_weight=sw[k])
            ops.fused_layer(agg, n, [], agg_out=dh[:, sl])
        d_as, d_ad = ops.gat_backward(h, g.csr, dout, att_src, att_dst, w, sw, heads, ctx.slope, dh)
        db = ops.column_sums(dout) if ctx.has_bias and ctx.needs_input_grad[3] else None
        return dh, d_as.view_as(att_src), d_ad.view_as(att_dst), db, None, None, None


def gat_attend(h: Tensor, att_src: Tensor, att_dst: Tensor, bias: Optional[Tensor], graph, heads: int, slope: float) -> Tensor:
    return _GatAttendFn.apply(h, att_src, att_dst, bias, graph, heads, slope)


class _BatchNormFn(torch.autograd.Function):
    """Training-mode BatchNorm1d (batch statistics; running estimates updated by the forward launch)."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn):
        x = _rowmajor(x)
        y = ops.batchnorm_forward(x, bn)
        ctx.eps = float(bn.eps)
        ctx.has_affine = weight is not None
        ctx.save_for_backward(x, weight if weight is not None else x)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx, dw, db = ops.batchnorm_backward(x, _rowmajor(dy), weight if ctx.has_affine else None, ctx.eps)
        return dx, (dw if ctx.has_affine else None), (db if ctx.has_affine else None), None


def batch_norm_train(x: Tensor, bn) -> Tensor:
    return _BatchNormFn.apply(x, bn.weight, bn.bias, bn)


class _BatchNormDropoutFn(torch.autograd.Function):
    """dropout(bn(x)) with batch statistics as one fused epilogue; the Philox mask is a function of the seed, not a saved tensor."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn, p, seed):
        x = _rowmajor(x)
        y = ops.batchnorm_dropout_forward(x, bn, p, seed)
        ctx.save_for_backward(x, weight if weight is not None else x.new_empty(0))
        ctx.has_affine, ctx.eps, ctx.p, ctx.seed = weight is not None, bn.eps, p, seed
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx, dw, db = ops.batchnorm_dropout_backward(x, _rowmajor(dy), weight if ctx.has_affine else None, ctx.eps, ctx.p, ctx.seed)
        return dx, (dw if ctx.has_affine else None), (db if ctx.has_affine else None), None, None, None


def new_dropout_seed() -> int:
    """A 63-bit seed drawn from torch's CPU generator: reproducible under torch.manual_seed, no device synchronisation."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def batch_norm_dropout_train(x: Tensor, bn, p: float, needs_grad: bool = True) -> Tensor:
    """``dropout(bn(x), p)`` in training mode (nc/models.py:197-198) through kagnn_bn_dropout_train_fwd / _bwd."""
    seed = new_dropout_seed()
    if needs_grad:
        return _BatchNormDropoutFn.apply(x, bn.weight, bn.bias, bn, float(p), seed)
    return ops.batchnorm_dropout_forward(x, bn, float(p), seed)


class _SiluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _rowmajor(x)
        ctx.save_for_backward(x)
        return ops.silu_forward(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.silu_backward(x, _rowmajor(dy))


def silu(x: Tensor) -> Tensor:
    return _SiluFn.apply(x)


class _PoolFn(torch.autograd.Function):
    """global_add_pool / global_mean_pool."""

    @staticmethod
    def forward(ctx, x, batch, num_graphs, mean):
        x = _rowmajor(x)
        ptr = ops.segment_ptr(batch, num_graphs)
        ctx.ptr, ctx.batch, ctx.n, ctx.mean = ptr, batch, x.size(0), bool(mean)
        agg = ops.AggSpec(L.AGG_SEGMENT_MEAN if mean else L.AGG_SEGMENT_SUM, x, rowptr=ptr)
        return ops.fused_layer(agg, num_graphs, [])

    @staticmethod
    @once_differentiable
    def backward(ctx, dp):
        return ops.segment_pool_backward(_rowmajor(dp), ctx.ptr, ctx.batch, ctx.n, ctx.mean), None, None, None


def pool(x: Tensor, batch: Tensor, num_graphs: int, mean: bool) -> Tensor:
    return _PoolFn.apply(x, batch, num_graphs, mean)


class _LogSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.log_softmax(_rowmajor(x))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.log_softmax_backward(y, _rowmajor(dy))


def log_softmax(x: Tensor) -> Tensor:
    return _LogSoftmaxFn.apply(x)
