"""kagnn_b200 -- B200-native forward path of KAGNN: KAN layers (B-spline ``ekan`` / RBF ``fastkan``) fused with the
GCN / GIN / GINE neighbour aggregation, behind the reference's module names and the PyG call surface.

Layout
    csrc/            hand-written sm_100a kernels + the C ABI (include/kagnn_b200.h) -> lib/libkagnn_b200.so
    _lib.py, ops.py  ctypes binding and tensor->pointer wrappers (torch = memory and streams only)
    graph.py         COO -> CSR / gcn_norm cache
    ekan.py, fastkan.py, conv.py, models_{node,graph,regr}.py
                     host-side mirror of the reference's modules (same names, signatures, state_dict keys)
    dist.py          node-range sharding + halo exchange for multi-GPU
"""
from . import _lib, ops, graph
from .ekan import KAN, KANLinear
from .fastkan import FastKAN, FastKANLayer, RadialBasisFunction, SplineLinear
from .conv import (GCNConv, GINConv, GINEConv, GATConv, KANLayer, FKANLayer, KAGCNConv, FASTKAGCNConv, GIKANLayer,
                   GIFASTKANLayer, KAGCN_Layer, FASTKAGCN_Layer, KAGATConv, FASTKAGATConv, KAGAT_Layer, FASTKAGAT_Layer,
                   make_kan, make_fastkan)
from .models_node import GKAN_Nodes, GFASTKAN_Nodes
from . import models_graph, models_regr
from .ops import set_precision, get_precision

__version__ = "0.1.0"
