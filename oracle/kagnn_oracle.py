"""CPU oracle for the KAGNN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``kagnn_b200``)
never does: it calls the sm_100a kernels through the C ABI and raises when the
library is missing.

What it is: a plain-torch, CPU, functional restatement of the reference's
arithmetic for the path named by BASELINE.json's ``north_star``:

  * the KAN primitives of ``ekan.py`` / ``fastkan.py`` (reference files live under
    ``/root/reference/{node_classification_clean,graph_classification,graph_regression}``),
  * the torch_geometric 2.5.3 layers the reference's ``models.py`` build on
    (``GCNConv`` + ``gcn_norm``, ``GINConv``, ``GINEConv``, ``global_add_pool``,
    ``global_mean_pool``).  torch_geometric is a pinned third-party dependency
    (``requirements.txt:4``) that is NOT vendored in the reference and NOT installable
    here, so its published algorithm is restated,
  * the ``forward`` glue of every KAN model class, driven by a ``state_dict`` that
    uses the reference's parameter names.

Parity pin status
-----------------
KAN half (``kan_linear``, ``kan_chain``, ``fastkan_layer``, ``fastkan_chain``): PINNED —
checked against golden vectors produced by importing the reference's own
``ekan.py`` / ``fastkan.py`` in the build container (``oracle/make_golden.py`` →
``tests/golden/*.npz``; ``tests/test_oracle_golden.py``).
Model glue (``*_forward``): PINNED to the reference's own ``models.py`` classes executed
with ``oracle/pyg_shim.py`` standing in for torch_geometric (same golden files).
PyG half (``gcn_norm``, ``gcn_conv``, ``gin_conv``, ``gine_conv``, pools): **parity unpinned** —
the reference ships no tests or vectors and torch_geometric cannot be imported; the
restatement is cross-checked only against the dense identity D^-1/2 (A'+I) D^-1/2 that the
reference itself writes out at ``node_classification_clean/time_model.py:70-80``.

All functions work in the dtype of their inputs (fp32 like the reference, or fp64 for
tighter comparisons).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# B-spline KAN (ekan.py)
# ----------------------------------------------------------------------------------------------
def uniform_knots(in_features: int, grid_size: int, spline_order: int,
                  lo: float = -1.0, hi: float = 1.0, dtype=torch.float32) -> Tensor:
    """Knot buffer ``grid`` of shape (in, G+2k+1): t_j = (j-k)*h + lo, h=(hi-lo)/G.
    Follows node_classification_clean/ekan.py:28-37 (same op order: arange*h + lo)."""
    h = (hi - lo) / grid_size
    t = torch.arange(-spline_order, grid_size + spline_order + 1) * h + lo
    return t.to(dtype).expand(in_features, -1).contiguous()


def bspline_bases(x: Tensor, grid: Tensor, spline_order: int) -> Tensor:
    """Cox-de Boor recursion, half-open level-0 indicator; zeros outside the knot range / for NaN.
    (N,in) -> (N,in,G+k).  Follows node_classification_clean/ekan.py:79-112."""
    assert x.dim() == 2 and x.size(1) == grid.size(0)
    xe = x.unsqueeze(-1)
    b = ((xe >= grid[:, :-1]) & (xe < grid[:, 1:])).to(x.dtype)
    for d in range(1, spline_order + 1):
        left = (xe - grid[:, : -(d + 1)]) / (grid[:, d:-1] - grid[:, : -(d + 1)]) * b[:, :, :-1]
        right = (grid[:, d + 1:] - xe) / (grid[:, d + 1:] - grid[:, 1:(-d)]) * b[:, :, 1:]
        b = left + right
    return b.contiguous()


def kan_linear(x: Tensor, base_weight: Tensor, spline_weight: Tensor,
               spline_scaler: Optional[Tensor], grid: Tensor, spline_order: int) -> Tensor:
    """silu(x) @ Wb^T + vec(B(x)) @ vec(Ws * scaler)^T, no bias.
    Follows node_classification_clean/ekan.py:146-162."""
    n = x.size(0)
    out_f = base_weight.size(0)
    w = spline_weight if spline_scaler is None else spline_weight * spline_scaler.unsqueeze(-1)
    base = F.linear(F.silu(x), base_weight)
    spl = F.linear(bspline_bases(x, grid, spline_order).view(n, -1), w.reshape(out_f, -1))
    return base + spl


def _kan_layer_from_sd(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    grid = sd[prefix + "grid"].to(x.dtype)
    sw = sd[prefix + "spline_weight"].to(x.dtype)
    k = (grid.size(1) - 1 - sw.size(2))            # G+2k+1 - 1 - (G+k) = k
    sc = sd.get(prefix + "spline_scaler")
    return kan_linear(x, sd[prefix + "base_weight"].to(x.dtype), sw,
                      None if sc is None else sc.to(x.dtype), grid, k)


def _count_layers(sd: Dict[str, Tensor], prefix: str, leaf: str) -> int:
    n = 0
    while f"{prefix}{n}.{leaf}" in sd:
        n += 1
    return n


def kan_chain(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """``KAN.forward``: KANLinear layers back to back, nothing in between
    (node_classification_clean/ekan.py:270-275).  ``prefix`` ends with 'layers.'."""
    for m in range(_count_layers(sd, prefix, "base_weight")):
        x = _kan_layer_from_sd(sd, f"{prefix}{m}.", x)
    return x


# ----------------------------------------------------------------------------------------------
# RBF KAN (fastkan.py)
# ----------------------------------------------------------------------------------------------
def rbf_bases(z: Tensor, grid: Tensor, denominator: float) -> Tensor:
    """exp(-((z - g)/den)^2); node_classification_clean/fastkan.py:46-47."""
    return torch.exp(-((z[..., None] - grid) / denominator) ** 2)


def fastkan_layer(x: Tensor, ln_weight: Optional[Tensor], ln_bias: Optional[Tensor], grid: Tensor,
                  spline_weight: Tensor, base_weight: Tensor, base_bias: Tensor,
                  denominator: Optional[float] = None) -> Tensor:
    """SplineLinear(rbf(LayerNorm(x))) + Linear(silu(x)); the base branch sees the RAW x.
    node_classification_clean/fastkan.py:76-85.  ``denominator`` defaults to the reference's
    (grid_max-grid_min)/(G-1) (fastkan.py:44)."""
    g = grid.numel()
    if denominator is None:
        denominator = (float(grid[-1]) - float(grid[0])) / (g - 1)
    z = x if ln_weight is None else F.layer_norm(x, (x.size(-1),), ln_weight, ln_bias, 1e-5)
    phi = rbf_bases(z, grid, denominator)
    ret = F.linear(phi.view(*phi.shape[:-2], -1), spline_weight)
    return ret + F.linear(F.silu(x), base_weight, base_bias)


def _fastkan_layer_from_sd(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    dt = x.dtype
    lw = sd.get(prefix + "layernorm.weight")
    lb = sd.get(prefix + "layernorm.bias")
    return fastkan_layer(x, None if lw is None else lw.to(dt), None if lb is None else lb.to(dt),
                         sd[prefix + "rbf.grid"].to(dt), sd[prefix + "spline_linear.weight"].to(dt),
                         sd[prefix + "base_linear.weight"].to(dt), sd[prefix + "base_linear.bias"].to(dt))


def fastkan_chain(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """``FastKAN.forward`` (node_classification_clean/fastkan.py:142-145)."""
    for m in range(_count_layers(sd, prefix, "spline_linear.weight")):
        x = _fastkan_layer_from_sd(sd, f"{prefix}{m}.", x)
    return x


# ----------------------------------------------------------------------------------------------
# torch_geometric 2.5.3 semantics, restated  (parity unpinned, see module docstring)
# ----------------------------------------------------------------------------------------------
def gcn_norm(edge_index: Tensor, num_nodes: int, dtype=torch.float32,
             edge_weight: Optional[Tensor] = None):
    """PyG ``gcn_norm(add_self_loops=True, improved=False, flow='source_to_target')``:
    existing self loops are dropped and exactly one loop of weight 1 is appended per node
    (``add_remaining_self_loops``; an existing loop's weight would be kept, but the weights are
    all ones unless the caller passed ``edge_weight``), ``deg`` = weighted in-degree at the TARGET
    (row 1 of edge_index), ``w_e = deg^-1/2[src] * w * deg^-1/2[dst]`` with inf -> 0.
    Returns (edge_index', w).  Call sites: KAGCNConv node_classification_clean/models.py:31-37."""
    row, col = edge_index[0], edge_index[1]
    if edge_weight is None:
        edge_weight = torch.ones(row.numel(), dtype=dtype)
    keep = row != col
    loop_w = torch.ones(num_nodes, dtype=dtype)
    # add_remaining_self_loops: a pre-existing loop donates its weight to the re-added loop
    inv = ~keep
    loop_w[row[inv]] = edge_weight[inv]
    loops = torch.arange(num_nodes, dtype=row.dtype)
    row2 = torch.cat([row[keep], loops])
    col2 = torch.cat([col[keep], loops])
    w = torch.cat([edge_weight[keep], loop_w])
    deg = torch.zeros(num_nodes, dtype=dtype).index_add_(0, col2, w)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return torch.stack([row2, col2]), dis[row2] * w * dis[col2]


def gcn_conv(x: Tensor, edge_index: Tensor, lin, bias: Optional[Tensor],
             edge_weight: Optional[Tensor] = None) -> Tensor:
    """PyG ``GCNConv.forward``: h = lin(x); out[i] = sum_{e: dst_e = i} w_e h[src_e]; out += bias.
    ``lin`` is a callable (the KAN that replaces ``self.lin``)."""
    n = x.size(0)
    ei, w = gcn_norm(edge_index, n, x.dtype, edge_weight)
    h = lin(x)
    out = torch.zeros(n, h.size(1), dtype=h.dtype).index_add_(0, ei[1], w.unsqueeze(1) * h.index_select(0, ei[0]))
    return out if bias is None else out + bias


def gin_conv(x: Tensor, edge_index: Tensor, nn_fn, eps: float = 0.0) -> Tensor:
    """PyG ``GINConv.forward``: nn((1+eps) x_i + sum_{j->i} x_j).
    Call sites: GIKANLayer node_classification_clean/models.py:48-56."""
    agg = torch.zeros_like(x).index_add_(0, edge_index[1], x.index_select(0, edge_index[0]))
    return nn_fn(agg + (1.0 + eps) * x)


def gine_conv(x: Tensor, edge_index: Tensor, edge_attr: Tensor, nn_fn, eps: float = 0.0) -> Tensor:
    """PyG ``GINEConv.forward`` with edge_dim=None: nn((1+eps) x_i + sum_{j->i} relu(x_j + e_ji)).
    Call site: graph_regression/models.py:98."""
    msg = (x.index_select(0, edge_index[0]) + edge_attr).relu()
    agg = torch.zeros_like(x).index_add_(0, edge_index[1], msg)
    return nn_fn(agg + (1.0 + eps) * x)


def gat_conv(x: Tensor, edge_index: Tensor, lin, att_src: Tensor, att_dst: Tensor, bias: Optional[Tensor],
             heads: int, negative_slope: float = 0.2, concat: bool = True) -> Tensor:
    """PyG 2.5 ``GATConv.forward`` for an int ``in_channels`` (one shared projection ``self.lin``, which KAGATConv replaces by a
    KAN: node_classification_clean/models.py:39-46), ``edge_dim=None``, ``dropout=0``, ``add_self_loops=True``:
        h = lin(x).view(N, H, C);  a_src = (h * att_src).sum(-1);  a_dst = (h * att_dst).sum(-1)
        edges: existing self loops removed, one loop per node appended (remove_self_loops + add_self_loops)
        e_ji = leaky_relu(a_src[j] + a_dst[i], 0.2);  alpha = softmax over the incoming edges of i (PyG softmax: exp(e - max) /
        (sum + 1e-16));  out_i = sum_j alpha_ji h_j;  heads concatenated (or averaged), + bias.
    ``lin`` is a callable.  Parity unpinned like the rest of the PyG half (torch_geometric is not installable here)."""
    n = x.size(0)
    h = lin(x)
    c = h.size(1) // heads
    h = h.view(n, heads, c)
    a_s = (h * att_src.view(1, heads, c).to(h.dtype)).sum(-1)
    a_d = (h * att_dst.view(1, heads, c).to(h.dtype)).sum(-1)
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loops = torch.arange(n, dtype=row.dtype)
    row2, col2 = torch.cat([row[keep], loops]), torch.cat([col[keep], loops])
    e = F.leaky_relu(a_s[row2] + a_d[col2], negative_slope)                                     # (E', H)
    idx = col2.unsqueeze(1).expand(-1, heads)
    m = torch.full((n, heads), float("-inf"), dtype=h.dtype).scatter_reduce(0, idx, e, reduce="amax", include_self=True)
    ex = (e - m[col2]).exp()
    den = torch.zeros(n, heads, dtype=h.dtype).index_add_(0, col2, ex) + 1e-16
    alpha = ex / den[col2]
    out = torch.zeros(n, heads, c, dtype=h.dtype).index_add_(0, col2, alpha.unsqueeze(-1) * h[row2])
    out = out.reshape(n, heads * c) if concat else out.mean(1)
    return out if bias is None else out + bias.to(out.dtype)


def global_add_pool(x: Tensor, batch: Tensor, num_graphs: Optional[int] = None) -> Tensor:
    """scatter(x, batch, dim=0, reduce='sum'), dim_size = batch.max()+1."""
    if num_graphs is None:
        num_graphs = int(batch.max()) + 1 if batch.numel() else 0
    return torch.zeros(num_graphs, x.size(1), dtype=x.dtype).index_add_(0, batch, x)


def global_mean_pool(x: Tensor, batch: Tensor, num_graphs: Optional[int] = None) -> Tensor:
    """scatter(..., reduce='mean'): sum / max(count, 1)."""
    s = global_add_pool(x, batch, num_graphs)
    cnt = torch.zeros(s.size(0), dtype=x.dtype).index_add_(0, batch, torch.ones(batch.numel(), dtype=x.dtype))
    return s / cnt.clamp(min=1).unsqueeze(1)


def batch_norm(sd: Dict[str, Tensor], prefix: str, x: Tensor, training: bool = False, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm1d forward.  eval: running stats; train: batch stats, biased variance."""
    w, b = sd[prefix + "weight"].to(x.dtype), sd[prefix + "bias"].to(x.dtype)
    if training:
        mean, var = x.mean(0), x.var(0, unbiased=False)
    else:
        mean, var = sd[prefix + "running_mean"].to(x.dtype), sd[prefix + "running_var"].to(x.dtype)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def dense_gcn_matrix(edge_index: Tensor, num_nodes: int, dtype=torch.float64) -> Tensor:
    """Independent dense statement D^-1/2 (A'+I) D^-1/2 used to cross-check ``gcn_norm``
    (the formula written out at node_classification_clean/time_model.py:70-80).  A'[i,j] counts
    edges j->i with i != j (multi-edges add up)."""
    a = torch.zeros(num_nodes, num_nodes, dtype=dtype)
    for s, d in edge_index.t().tolist():
        if s != d:
            a[d, s] += 1.0
    a += torch.eye(num_nodes, dtype=dtype)
    dis = a.sum(1).pow(-0.5)
    return dis.unsqueeze(1) * a * dis.unsqueeze(0)


# ----------------------------------------------------------------------------------------------
# Model glue (models.py forward bodies), driven by a reference-named state_dict
# ----------------------------------------------------------------------------------------------
def _is_fast(sd: Dict[str, Tensor]) -> bool:
    return any(k.endswith("rbf.grid") for k in sd)


def _kan_or_fast_chain(sd, prefix, x):
    return fastkan_chain(sd, prefix, x) if _is_fast(sd) else kan_chain(sd, prefix, x)


def _kan_or_fast_layer(sd, prefix, x):
    return _fastkan_layer_from_sd(sd, prefix, x) if _is_fast(sd) else _kan_layer_from_sd(sd, prefix, x)


def node_model_forward(sd: Dict[str, Tensor], conv_type: str, x: Tensor, edge_index: Tensor,
                       skip: bool = True, training: bool = False) -> Tensor:
    """``GKAN_Nodes.forward`` / ``GFASTKAN_Nodes.forward`` (node_classification_clean/models.py:192-203,
    :246-257) for conv_type in {'gcn','gin','gat'}; dropout p=0."""
    feats = [x]
    n_mp = _count_layers(sd, "bns.", "weight")
    for l in range(n_mp):
        p = f"convs.{l}."
        if conv_type == "gcn":
            x = gcn_conv(x, edge_index, lambda t: _kan_or_fast_layer(sd, p + "lin.", t), sd[p + "bias"].to(x.dtype))
        elif conv_type == "gin":
            x = gin_conv(x, edge_index, lambda t: _kan_or_fast_chain(sd, p + "nn.layers.", t), float(sd[p + "eps"]))
        elif conv_type == "gat":
            heads = sd[p + "att_src"].shape[1]
            x = gat_conv(x, edge_index, lambda t: _kan_or_fast_layer(sd, p + "lin.", t), sd[p + "att_src"], sd[p + "att_dst"],
                         sd[p + "bias"], heads)
        else:
            raise ValueError("unknown conv_type")
        x = batch_norm(sd, f"bns.{l}.", x, training)
        feats.append(x)
    if skip:
        x = torch.cat(feats, dim=1)
    return _kan_or_fast_layer(sd, "lay_out.", x)


class Batch:
    """Duck-typed stand-in for a PyG ``Data``/``Batch`` (x, edge_index, batch[, edge_attr])."""

    def __init__(self, x, edge_index, batch, edge_attr=None):
        self.x, self.edge_index, self.batch, self.edge_attr = x, edge_index, batch, edge_attr


def gc_kagin_forward(sd: Dict[str, Tensor], data, training: bool = False) -> Tensor:
    """graph_classification ``KAGIN.forward`` / ``FASTKAGIN.forward`` (models.py:111-119, :143-151)."""
    x = data.x
    for l in range(_count_layers(sd, "bn.", "weight")):
        p = f"conv.{l}."
        x = gin_conv(x, data.edge_index, lambda t: _kan_or_fast_chain(sd, p + "nn.layers.", t), float(sd[p + "eps"]))
        x = batch_norm(sd, f"bn.{l}.", x, training)
    x = global_add_pool(x, data.batch)
    return F.log_softmax(_kan_or_fast_chain(sd, "kan.layers.", x), dim=1)


def gc_kagcn_forward(sd: Dict[str, Tensor], data) -> Tensor:
    """graph_classification ``KAGCN.forward`` / ``FASTKAGCN.forward`` (models.py:186-194, :257-265):
    conv -> silu -> mean pool -> 1-layer KAN -> log_softmax."""
    x = data.x
    for l in range(_count_layers(sd, "conv.", "bias")):
        p = f"conv.{l}."
        x = F.silu(gcn_conv(x, data.edge_index, lambda t: _kan_or_fast_layer(sd, p + "lin.", t), sd[p + "bias"].to(x.dtype)))
    x = global_mean_pool(x, data.batch)
    return F.log_softmax(_kan_or_fast_chain(sd, "readout.layers.", x), dim=1)


def gc_kagat_forward(sd: Dict[str, Tensor], data) -> Tensor:
    """graph_classification ``KAGAT.forward`` / ``FASTKAGAT.forward`` (models.py:205-216, :277-288):
    (GAT conv -> silu) xL -> ADD pool -> 1-layer KAN -> log_softmax."""
    x = data.x
    for l in range(_count_layers(sd, "conv.", "bias")):
        p = f"conv.{l}."
        heads = sd[p + "att_src"].shape[1]
        x = F.silu(gat_conv(x, data.edge_index, lambda t: _kan_or_fast_layer(sd, p + "lin.", t), sd[p + "att_src"], sd[p + "att_dst"],
                            sd[p + "bias"], heads))
    x = global_add_pool(x, data.batch)
    return F.log_softmax(_kan_or_fast_chain(sd, "readout.layers.", x), dim=1)


def _encode(sd: Dict[str, Tensor], name: str, idx_or_x: Tensor, dtype) -> Tensor:
    """AtomEncoder/BondEncoder = sum of embedding lookups (graph_regression/models.py:244-279) when
    ``ogb_encoders`` else nn.Linear (:94-95)."""
    lst = f"{name}.{'atom' if name.startswith('atom') else 'bond'}_embedding_list."
    if f"{lst}0.weight" in sd:
        out = 0
        for c in range(idx_or_x.shape[1]):
            out = out + sd[f"{lst}{c}.weight"].to(dtype)[idx_or_x[:, c]]
        return out
    return F.linear(idx_or_x.to(dtype), sd[f"{name}.weight"].to(dtype), sd[f"{name}.bias"].to(dtype))


def gr_kagin_forward(sd: Dict[str, Tensor], data, training: bool = False, dtype=torch.float32) -> Tensor:
    """graph_regression ``KAGIN.forward`` / ``FASTKAGIN.forward`` (models.py:107-119, :146-160):
    encoders -> GINE-KAN xL -> BN -> add pool -> KAN, no log_softmax."""
    ea = data.edge_attr
    if ea.dim() == 1:
        ea = ea.unsqueeze(1)
    x = _encode(sd, "atom_encoder", data.x, dtype)
    ea = _encode(sd, "bond_encoder", ea, dtype)
    for l in range(_count_layers(sd, "bn.", "weight")):
        p = f"conv.{l}."
        x = gine_conv(x, data.edge_index, ea, lambda t: _kan_or_fast_chain(sd, p + "nn.layers.", t), float(sd[p + "eps"]))
        x = batch_norm(sd, f"bn.{l}.", x, training)
    x = global_add_pool(x, data.batch)
    return _kan_or_fast_chain(sd, "kan.layers.", x)


def gr_kagcn_forward(sd: Dict[str, Tensor], data, dtype=torch.float32) -> Tensor:
    """graph_regression ``KAGCN.forward`` / ``FASTKAGCN.forward`` (models.py:186-198, :230-242):
    encoder -> (conv -> silu) xL -> ADD pool -> 1-layer KAN."""
    x = _encode(sd, "atom_encoder", data.x, dtype)
    for l in range(_count_layers(sd, "conv.", "bias")):
        p = f"conv.{l}."
        x = F.silu(gcn_conv(x, data.edge_index, lambda t: _kan_or_fast_layer(sd, p + "lin.", t), sd[p + "bias"].to(x.dtype)))
    x = global_add_pool(x, data.batch)
    return _kan_or_fast_chain(sd, "readout.layers.", x)


def to_dtype(sd: Dict[str, Tensor], dtype) -> Dict[str, Tensor]:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def rel_err(y: Tensor, ref: Tensor) -> float:
    """The parity metric used everywhere: max|y - ref| / max|ref|."""
    ref = ref.detach().double()
    denom = float(ref.abs().max())
    return float((y.detach().double() - ref).abs().max()) / (denom if denom > 0 else 1.0)
