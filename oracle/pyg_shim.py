"""Minimal stand-in for ``torch_geometric.nn`` -- TEST INFRASTRUCTURE ONLY.

torch_geometric 2.5.3 (reference ``requirements.txt:4``) is not installed and cannot be installed
here.  ``oracle/make_golden.py`` registers this module as ``torch_geometric`` / ``torch_geometric.nn``
so that the reference's own ``models.py`` files import and run unmodified in the build container;
the golden vectors then pin the model glue (layer order, BN, pooling, skip concat, readout) to the
reference's code, while the message-passing arithmetic itself is the restatement in
``kagnn_oracle.py`` (parity unpinned for that half, see its docstring).

Attribute names (``lin``, ``bias``, ``nn``, ``eps``) follow PyG so that ``state_dict`` keys match.
"""
from __future__ import annotations

import sys
import types

import torch
from torch import nn

from . import kagnn_oracle as K


class GCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index, edge_weight=None):
        return K.gcn_conv(x, edge_index, self.lin, self.bias, edge_weight)


class GINConv(nn.Module):
    def __init__(self, nn_module, eps: float = 0.0, train_eps: bool = False, **kw):
        super().__init__()
        self.nn = nn_module
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer("eps", torch.tensor([float(eps)]))

    def forward(self, x, edge_index, size=None):
        return K.gin_conv(x, edge_index, self.nn, float(self.eps))


class GINEConv(GINConv):
    def __init__(self, nn_module, eps: float = 0.0, train_eps: bool = False, edge_dim=None, **kw):
        super().__init__(nn_module, eps, train_eps)
        self.lin = None
        if edge_dim is not None:
            raise NotImplementedError("edge_dim is never used by the reference")

    def forward(self, x, edge_index, edge_attr=None, size=None):
        return K.gine_conv(x, edge_index, edge_attr, self.nn, float(self.eps))


class GATConv(nn.Module):
    """PyG 2.5 GATConv with the defaults the reference uses (concat=True, negative_slope=0.2, dropout=0, add_self_loops=True,
    bias=True, edge_dim=None): one shared projection ``lin`` (which KAGATConv / KAGAT_Layer replace by a KAN), ``att_src`` /
    ``att_dst`` of shape (1, heads, out_channels), ``bias`` of heads * out_channels."""

    def __init__(self, in_channels, out_channels, heads=1, **kw):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    def forward(self, x, edge_index, edge_attr=None, size=None):
        return K.gat_conv(x, edge_index, self.lin, self.att_src, self.att_dst, self.bias, self.heads)


global_add_pool = K.global_add_pool
global_mean_pool = K.global_mean_pool


def install() -> None:
    """Register this shim as ``torch_geometric`` and ``torch_geometric.nn`` in ``sys.modules``."""
    pkg = types.ModuleType("torch_geometric")
    sub = types.ModuleType("torch_geometric.nn")
    for name in ("GCNConv", "GINConv", "GINEConv", "GATConv", "global_add_pool", "global_mean_pool"):
        setattr(sub, name, globals()[name])
    pkg.nn = sub
    sys.modules["torch_geometric"] = pkg
    sys.modules["torch_geometric.nn"] = sub
