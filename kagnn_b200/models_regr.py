"""Graph-regression models ``KAGIN`` / ``FASTKAGIN`` / ``KAGCN`` / ``FASTKAGCN`` (+ ``AtomEncoder`` /
``BondEncoder``) with the positional constructor signatures, module tree and ``forward(data)`` of
graph_regression/models.py:86-279 (ZINC / QM9 drivers graph_regression/optuna_zinc.py:44-62).

GINE layers: ``out = KAN((1+eps) x_i + sum_j relu(x_j + e_ji))`` then BatchNorm; add-pool; KAN readout; no
log_softmax.  When the encoders are single-column embedding tables (ZINC: one atom-type column, one bond-type
column) the (N,H) / (E,H) embedded matrices are never materialised: the fused kernel reads rows of the tables
through an index (``src_index`` / ``edge_row``), which keeps the per-edge operand in L1/L2."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from .conv import FASTKAGCN_Layer, GINEConv, KAGCN_Layer, make_fastkan, make_kan
from .ekan import _module_backend_guard, eval_mode_detach_notice
from .graph import get_graph
from .models_graph import _GCNGraphModel, _GINGraphModel, _bn_unfused, _num_graphs, pooled_readout
from .models_node import bn_unfused

Tensor = torch.Tensor

# Vocabulary sizes of the OGB molecule featurisation the reference's encoders are built on
# (lengths of the categorical lists at graph_regression/models.py:281-336).
ATOM_FEATURE_DIMS = (119, 5, 12, 12, 10, 6, 6, 2, 2)
BOND_FEATURE_DIMS = (5, 6, 2)


class _SumEmbedding(nn.Module):
    """Sum over columns of per-column embedding lookups."""

    def __init__(self, attr: str, dims, emb_dim: int):
        super().__init__()
        tables = nn.ModuleList()
        for d in dims:
            emb = nn.Embedding(d, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            tables.append(emb)
        setattr(self, attr, tables)
        self._attr = attr

    def tables(self) -> nn.ModuleList:
        return getattr(self, self._attr)

    def forward(self, idx: Tensor) -> Tensor:
        out = 0
        for c in range(idx.shape[1]):
            out = out + self.tables()[c](idx[:, c])
        return out


class AtomEncoder(_SumEmbedding):
    def __init__(self, emb_dim, optional_full_atom_features_dims=None):
        super().__init__("atom_embedding_list", optional_full_atom_features_dims or ATOM_FEATURE_DIMS, emb_dim)


class BondEncoder(_SumEmbedding):
    def __init__(self, emb_dim):
        super().__init__("bond_embedding_list", BOND_FEATURE_DIMS, emb_dim)


def _encoders(model: nn.Module, num_node_features, num_edge_features, hidden_dim, ogb_encoders, with_bond=True):
    if ogb_encoders:
        model.atom_encoder = AtomEncoder(hidden_dim)
        if with_bond:
            model.bond_encoder = BondEncoder(hidden_dim)
    else:
        model.atom_encoder = nn.Linear(num_node_features, hidden_dim)
        if with_bond:
            model.bond_encoder = nn.Linear(num_edge_features, hidden_dim)


def _table_lookup(enc: nn.Module, idx: Tensor):
    """(table, int32 row ids) when the encoder is a one-column embedding, else None."""
    if isinstance(enc, _SumEmbedding) and idx.dim() == 2 and idx.size(1) == 1 and not idx.is_floating_point():
        return enc.tables()[0].weight.detach(), idx[:, 0].to(torch.int32).contiguous()
    return None


class _GINERegression(_GINGraphModel):
    log_softmax = False

    def forward(self, data) -> Tensor:
        x, edge_attr = data.x, data.edge_attr
        needs_grad = _module_backend_guard(x, self.parameters(), grad_ok=True)
        if needs_grad and not self.training:
            # eval() without no_grad (graph_regression/optuna_zinc.py:68-86): inference plan, result detached
            eval_mode_detach_notice(x)
            with torch.no_grad():
                return self.forward(data)
        if needs_grad:
            return self._forward_train(data)
        if edge_attr.dim() == 1:
            edge_attr = edge_attr.unsqueeze(1)
        n = x.size(0)
        g = get_graph(data.edge_index, n)
        fus = self._fusable()
        # edge operand: table + per-CSR-entry code, or dense (E,H) rows addressed through the CSR permutation
        bond = _table_lookup(self.bond_encoder, edge_attr)
        if bond is not None:
            edge_feat, edge_row = bond[0], bond[1].index_select(0, g.perm.long())
        else:
            edge_feat, edge_row = self.bond_encoder(edge_attr).to(torch.float32), g.perm
        atom = _table_lookup(self.atom_encoder, x)
        h = None if atom is not None else self.atom_encoder(x).to(torch.float32)
        for i in range(self.n_layers):
            conv = self.conv[i]
            post = self._folds[i].get(self.bn[i]) if fus else None
            if i == 0 and atom is not None:
                agg = ops.AggSpec(L.AGG_GINE, atom[0], g.rowptr, g.col, self_scale=1.0 + conv.eps_value(),
                                  edge_feat=edge_feat, edge_row=edge_row, src_index=atom[1])
                h = ops.fused_layer(agg, n, conv.nn.kernel_specs(), post=post)
            else:
                h = conv(h, g, edge_feat, post=post, edge_row=edge_row)
            if not fus:
                h = self.dropout(_bn_unfused(h, self.bn[i]))
        return pooled_readout(h, data.batch, _num_graphs(data), self.kan, mean=False)


    def _forward_train(self, data) -> Tensor:
        """Differentiated forward: the encoders are torch modules (their tables / weights get their gradients from torch), every
        GINE layer is aggregation -> KAN chain -> BatchNorm -> dropout as autograd.Functions over library launches."""
        x, edge_attr = data.x, data.edge_attr
        if edge_attr.dim() == 1:
            edge_attr = edge_attr.unsqueeze(1)
        g = get_graph(data.edge_index, x.size(0))
        h = self.atom_encoder(x).to(torch.float32)
        ef = self.bond_encoder(edge_attr).to(torch.float32)
        for i in range(self.n_layers):
            h = self.dropout(bn_unfused(self.conv[i](h, g, ef), self.bn[i], True))
        return pooled_readout(h, data.batch, _num_graphs(data), self.kan, mean=False, needs_grad=True)


class KAGIN(_GINERegression):
    def __init__(self, num_node_features, num_edge_features, gnn_layers, hidden_dim, hidden_layers, grid_size, spline_order,
                 num_classes, dropout, ogb_encoders):
        super().__init__()
        _encoders(self, num_node_features, num_edge_features, hidden_dim, ogb_encoders)
        self.conv = nn.ModuleList(GINEConv(make_kan(hidden_dim, hidden_dim, hidden_dim, hidden_layers, grid_size, spline_order))
                                  for _ in range(gnn_layers))
        self._init_common(gnn_layers, hidden_dim, dropout)
        self.kan = make_kan(hidden_dim, hidden_dim, num_classes, hidden_layers, grid_size, spline_order)


class FASTKAGIN(_GINERegression):
    def __init__(self, num_node_features, num_edge_features, gnn_layers, hidden_dim, hidden_layers, grid_size, num_classes,
                 dropout, ogb_encoders):
        super().__init__()
        _encoders(self, num_node_features, num_edge_features, hidden_dim, ogb_encoders)
        self.conv = nn.ModuleList(GINEConv(make_fastkan(hidden_dim, hidden_dim, hidden_dim, hidden_layers, grid_size))
                                  for _ in range(gnn_layers))
        self._init_common(gnn_layers, hidden_dim, dropout)
        self.kan = make_fastkan(hidden_dim, hidden_dim, num_classes, hidden_layers, grid_size)


class _GCNRegression(_GCNGraphModel):
    log_softmax = False
    mean_pool = False      # graph_regression/models.py:196 pools with global_add_pool

    def _encode(self, x: Tensor) -> Tensor:
        return self.atom_encoder(x).to(torch.float32)


class KAGCN(_GCNRegression):
    def __init__(self, num_node_features, gnn_layers, hidden_dim, grid_size, spline_order, num_classes, dropout, ogb_encoders):
        super().__init__()
        self.n_layers = gnn_layers
        _encoders(self, num_node_features, None, hidden_dim, ogb_encoders, with_bond=False)
        # the reference does not forward grid_size / spline_order to its layers (graph_regression/models.py:184):
        # they keep KAGCN_Layer's defaults (4, 3); reproduced so that checkpoints and results match
        self.conv = nn.ModuleList(KAGCN_Layer(hidden_dim, hidden_dim) for _ in range(gnn_layers))
        self.readout = make_kan(hidden_dim, hidden_dim, num_classes, 1, grid_size, spline_order)
        self.dropout = nn.Dropout(p=dropout)


class FASTKAGCN(_GCNRegression):
    def __init__(self, num_node_features, gnn_layers, hidden_dim, grid_size, num_classes, dropout, ogb_encoders):
        super().__init__()
        self.n_layers = gnn_layers
        _encoders(self, num_node_features, None, hidden_dim, ogb_encoders, with_bond=False)
        self.conv = nn.ModuleList(FASTKAGCN_Layer(hidden_dim, hidden_dim, grid_size) for _ in range(gnn_layers))
        self.readout = make_fastkan(hidden_dim, hidden_dim, num_classes, 1, grid_size)
        self.dropout = nn.Dropout(p=dropout)
