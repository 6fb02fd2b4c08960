"""TEST INFRASTRUCTURE ONLY -- a dry-run harness for the Python wiring of kagnn_b200/autograd.py on a machine without a GPU.

Inside ``with cpu_double():``

* the FORWARD library calls of ``kagnn_b200.ops`` (``fused_layer``, ``csr_build``, ``gcn_norm`` ...) are replaced by small
  torch-CPU stand-ins written here;
* the BACKWARD wrappers of ``kagnn_b200.ops`` run unchanged, but ``_lib.lib()`` hands them the host-check build of
  kagnn_b200/csrc/backward.cu (tests/emul/build_emul.py) instead of libkagnn_b200.so;
* the "CUDA tensors only" guards are lifted.

Nothing here is reachable from the product: the patches are applied by the tests and undone on exit.  The GPU tests
(tests/test_gpu_train_backward.py) run the same scenarios against the real library."""
import contextlib
import ctypes as C

import torch
import torch.nn.functional as F

from kagnn_b200 import _lib as L
from kagnn_b200 import conv, ekan, fastkan, models_graph, models_node, models_regr, ops
from kagnn_b200 import graph as kgraph
from tests.emul import build_emul

_BACKWARD = ("kagnn_kan_bwd_input", "kagnn_kan_bwd_weights", "kagnn_kan_unpack_weight_grads", "kagnn_batchnorm_bwd_workspace",
             "kagnn_batchnorm_train_bwd", "kagnn_column_sums", "kagnn_log_softmax_bwd", "kagnn_silu_fwd", "kagnn_silu_bwd", "kagnn_segment_pool_bwd", "kagnn_rbf_bwd_input", "kagnn_rbf_bwd_weights", "kagnn_layernorm_bwd", "kagnn_gine_bwd")


class _HostLib:
    def __init__(self):
        h = C.CDLL(build_emul.build())
        for name in _BACKWARD:
            fn = getattr(h, name)
            fn.restype, fn.argtypes = L._SIGNATURES[name]
            setattr(self, name, fn)

    @staticmethod
    def kagnn_strerror(code):
        return b"host-check error"

    @staticmethod
    def kagnn_packed_weight_elems(i, o, s):
        return i * (s + 1) * ((o + 3) // 4 * 4)


def _csr_build(edge_index, num_nodes, num_src_nodes=None):
    src, dst = edge_index[0], edge_index[1]
    perm = torch.argsort(dst, stable=True)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.bincount(dst, minlength=num_nodes).cumsum(0)
    return ops.CSR(rowptr.to(torch.int32), src[perm].to(torch.int32), perm.to(torch.int32), num_nodes,
                   num_nodes if num_src_nodes is None else num_src_nodes, torch.zeros(1, dtype=torch.int32))


def _row_of_entry(csr):
    return torch.repeat_interleave(torch.arange(csr.num_rows), (csr.rowptr[1:] - csr.rowptr[:-1]).long())


def _gcn_norm(csr, edge_weight_csr=None):
    # PyG gcn_norm(add_self_loops=True): explicit self loops are dropped, one unit loop per node is added
    assert edge_weight_csr is None
    row, col = _row_of_entry(csr), csr.col.long()
    keep = (row != col).float()
    deg = 1.0 + torch.zeros(csr.num_rows).index_add_(0, row, keep)
    dinv = deg.rsqrt()
    return dinv[col] * dinv[row] * keep, dinv * dinv, dinv


def _gcn_degree(csr, edge_weight_csr=None):
    assert edge_weight_csr is None
    row, col = _row_of_entry(csr), csr.col.long()
    deg = 1.0 + torch.zeros(csr.num_rows).index_add_(0, row, (row != col).float())
    dinv = deg.rsqrt()
    return dinv * dinv, dinv


def _gcn_edge_weight(csr, dinv_src, dinv_dst, edge_weight_csr=None):
    assert edge_weight_csr is None
    row, col = _row_of_entry(csr), csr.col.long()
    return dinv_src[col] * dinv_dst[row] * (row != col).float()


def _gat_attention(h, csr, att_src, att_dst, heads, negative_slope=0.2):
    # PyG GATConv coefficients on the CSR: existing self loops dropped (weight 0), one appended self loop per node
    n = h.size(0)
    c = h.size(1) // heads
    hv = h.view(n, heads, c)
    a_s = (hv * att_src.detach().view(1, heads, c)).sum(-1)
    a_d = (hv * att_dst.detach().view(1, heads, c)).sum(-1)
    row, col = _row_of_entry(csr), csr.col.long()
    keep = row != col
    e = F.leaky_relu(a_s[col] + a_d[row], negative_slope)                      # (nnz, H)
    e_self = F.leaky_relu(a_s + a_d, negative_slope)                            # (n, H)
    m = e_self.clone()
    m = m.scatter_reduce(0, row[keep].unsqueeze(1).expand(-1, heads), e[keep], reduce="amax", include_self=True)
    ex = torch.where(keep.unsqueeze(1), (e - m[row]).exp(), torch.zeros_like(e))
    xs = (e_self - m).exp()
    den = xs + torch.zeros(n, heads).index_add_(0, row, ex) + 1e-16
    return (ex / den[row]).t().contiguous(), (xs / den).t().contiguous()


def _gat_backward(h, csr, dout, att_src, att_dst, w, sw, heads, negative_slope, dh):
    # stand-in of ops.gat_backward (kagnn_gat_bwd): dh holds sum_i alpha_ij d out[i] on entry and gets the part through the
    # attention scores added; returns (d att_src, d att_dst), each (heads * C,)
    n = h.size(0)
    c = h.size(1) // heads
    hv, dv = h.view(n, heads, c), dout.view(n, heads, c)
    a_src = att_src.detach().view(1, heads, c)
    a_dst = att_dst.detach().view(1, heads, c)
    a_s, a_d = (hv * a_src).sum(-1), (hv * a_dst).sum(-1)
    row, col = _row_of_entry(csr), csr.col.long()
    al, al_self = w.t(), sw.t()                                                  # (nnz, H), (n, H)
    g_e = (dv[row] * hv[col]).sum(-1)
    g_s = (dv * hv).sum(-1)
    s = al_self * g_s + torch.zeros(n, heads).index_add_(0, row, al * g_e)
    slope = lambda pre: torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, negative_slope))  # noqa: E731
    dpre = al * (g_e - s[row]) * slope(a_s[col] + a_d[row])                     # removed self loops carry alpha = 0
    dpre_self = al_self * (g_s - s) * slope(a_s + a_d)
    da_dst = dpre_self + torch.zeros(n, heads).index_add_(0, row, dpre)
    da_src = dpre_self + torch.zeros(n, heads).index_add_(0, col, dpre)
    dh.add_((da_src.unsqueeze(-1) * a_src + da_dst.unsqueeze(-1) * a_dst).reshape(n, heads * c))
    return (da_src.unsqueeze(-1) * hv).sum(0).reshape(-1), (da_dst.unsqueeze(-1) * hv).sum(0).reshape(-1)


def _gather_rows(x, index, out=None, num_rows=None):
    res = x[index.long()] if index is not None else x[: (x.size(0) if num_rows is None else num_rows)]
    if out is None:
        return res.clone()
    out.copy_(res)
    return out


def _segment_ptr(batch, num_graphs):
    ptr = torch.zeros(num_graphs + 1, dtype=torch.int64)
    ptr[1:] = torch.bincount(batch, minlength=num_graphs).cumsum(0)
    return ptr.to(torch.int32)


def _pack(base_w, spline_w, scaler, in_f, out_f, slots):
    w = torch.zeros(in_f, slots + 1, (out_f + 3) // 4 * 4)
    sw = spline_w.detach().reshape(out_f, in_f, slots)
    if scaler is not None:
        sw = sw * scaler.detach().unsqueeze(-1)
    w[:, :slots, :out_f] = sw.permute(1, 2, 0)
    if base_w is not None:
        w[:, slots, :out_f] = base_w.detach().t()
    return w.contiguous()


def _affine(a, v):
    if a is None:
        return v
    if a.scale is not None:
        v = v * a.scale
    if a.shift is not None:
        v = v + a.shift
    return F.silu(v) if a.act == L.ACT_SILU else v


def _layernorm_stats(x, x_head=None, eps=1e-5):
    assert x_head is None
    mean = x.mean(1)
    return torch.stack([mean, (x.var(1, unbiased=False) + eps).rsqrt()], dim=1).contiguous()


def _kan_layer(spec, x):
    if spec.basis == L.BASIS_RBF:
        G = spec.grid_size
        z = x
        if spec.ln_weight is not None:
            z = F.layer_norm(x, (x.size(1),), spec.ln_weight, spec.ln_bias, 1e-5)
        centres = spec.t0 + spec.h * torch.arange(G)
        phi = torch.exp(-((z.unsqueeze(-1) - centres) * spec.inv_denominator) ** 2)
        w = spec.packed_w.view(spec.in_features, G + 1, -1)[:, :, :spec.out_features]
        y = torch.einsum("nig,igo->no", phi, w[:, :G]) + F.silu(x) @ w[:, G]
        return y if spec.base_bias is None else y + spec.base_bias
    S = spec.grid_size + spec.spline_order
    bases = ekan._uniform_bspline_design(x, spec.t0, spec.h, spec.grid_size, spec.spline_order)      # (n, in, S)
    w = spec.packed_w.view(spec.in_features, S + 1, -1)[:, :, :spec.out_features]
    return torch.einsum("nis,iso->no", bases, w[:, :S]) + F.silu(x) @ w[:, S]


def _expand_windows(x, windows, shift):
    return torch.cat([x - w * shift for w in range(windows)], dim=1).contiguous()


def _fused_layer(agg, num_rows, layers, pre=None, post=None, agg_out=None, out=None):
    if any(sp.windows > 1 for sp in layers):                 # the product's own wiring of windowed layers, leaf launches emulated
        return ops._windowed_chain(agg, num_rows, layers, pre, post, agg_out, out)
    x = agg.x if agg.x_head is None else torch.cat([agg.x_head, agg.x], dim=1)
    if agg.x_halo is not None:
        x = torch.cat([x, agg.x_halo], dim=0)                # halo rows follow the owned rows in the local numbering
    if agg.src_index is not None:
        x = x[agg.src_index.long()]                          # rows of a small table addressed through an index
    if agg.mode == L.AGG_NONE:
        t = x[:num_rows]
    elif agg.mode in (L.AGG_SEGMENT_SUM, L.AGG_SEGMENT_MEAN):
        ptr = agg.rowptr.long()
        seg = torch.repeat_interleave(torch.arange(num_rows), ptr[1:] - ptr[:-1])
        t = torch.zeros(num_rows, x.size(1)).index_add_(0, seg, x[: seg.numel()])
        if agg.mode == L.AGG_SEGMENT_MEAN:
            t = t / (ptr[1:] - ptr[:-1]).clamp(min=1).unsqueeze(1)
    else:
        ptr = agg.rowptr.long()
        row = torch.repeat_interleave(torch.arange(num_rows), ptr[1:] - ptr[:-1])
        msg = x[agg.col.long()]
        if agg.mode == L.AGG_WEIGHTED:
            msg = msg * agg.edge_weight.unsqueeze(1)
            t = x[:num_rows] * agg.self_weight.unsqueeze(1)
        elif agg.mode == L.AGG_GINE:
            msg = (msg + agg.edge_feat[agg.edge_row.long()]).relu()
            t = x[:num_rows] * agg.self_scale
        else:
            assert agg.mode == L.AGG_GIN
            t = x[:num_rows] * agg.self_scale
        t = t.index_add(0, row, msg)
    t = _affine(pre, t)
    if agg_out is not None:
        agg_out.copy_(t)
    if not layers:
        return agg_out if agg_out is not None else t.contiguous()
    for spec in layers:
        t = _kan_layer(spec, t)
    t = _affine(post, t)
    if out is not None:
        out.copy_(t)
        return out
    return t.contiguous()


def _batchnorm_forward(x, bn, act=L.ACT_NONE):
    y = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight.detach() if bn.weight is not None else None,
                     bn.bias.detach() if bn.bias is not None else None, True, bn.momentum if bn.momentum is not None else 0.0, bn.eps)
    if bn.track_running_stats and bn.running_mean is not None:
        bn.num_batches_tracked += 1
    return y


@contextlib.contextmanager
def cpu_double(windows: bool = False):
    """``windows=True``: layers with more than eight slots / centres per input take the slot-window wiring (ekan._windowed_spec,
    fastkan._windowed_spec, ops._windowed_chain and the gradient folding in ops / autograd) -- which on the GPU needs the tensor-core
    kernels -- with every leaf launch emulated like all the others."""
    host = _HostLib()
    saved = []

    def patch(obj, name, val):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, val)

    def guard(x, params, grad_ok=False):
        needs = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        if needs and not grad_ok:
            raise NotImplementedError("no backward for this module")
        return needs

    patch(L, "lib", lambda: host)
    patch(ops, "_need_cuda", lambda t, name, dtype=None: None)
    patch(ops, "_stream", lambda: None)
    patch(ops, "tc_supported", (lambda *a: True) if windows else (lambda *a: False))
    if windows:
        patch(ops, "pack_kan_weights_tc", lambda *a, **k: None)
        patch(ops, "expand_windows", _expand_windows)
        patch(ops, "_rbf_windows", True)
    for name, fn in (("csr_build", _csr_build), ("gcn_norm", _gcn_norm), ("gcn_degree", _gcn_degree), ("gcn_edge_weight", _gcn_edge_weight), ("gather_rows", _gather_rows), ("segment_ptr", _segment_ptr),
                     ("pack_kan_weights", _pack), ("fused_layer", _fused_layer), ("batchnorm_forward", _batchnorm_forward),
                     ("log_softmax", lambda x: torch.log_softmax(x, dim=1)), ("layernorm_stats", _layernorm_stats),
                     ("gat_attention", _gat_attention), ("gat_backward", _gat_backward)):
        patch(ops, name, fn)
    for mod in (ekan, fastkan, conv, models_node, models_graph, models_regr):
        patch(mod, "_module_backend_guard", guard)
    patch(kgraph, "get_graph", lambda ei, n: kgraph.GraphCSR(ei, n))
    for mod in (conv, models_node, models_graph, models_regr):
        patch(mod, "get_graph", kgraph.get_graph)
    try:
        yield
    finally:
        for obj, name, val in reversed(saved):
            setattr(obj, name, val)
