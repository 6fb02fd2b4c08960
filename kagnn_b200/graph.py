"""Per-graph state the reference recomputes on every forward, built once and cached.

PyG's ``MessagePassing.propagate`` gathers ``x[edge_index[0]]`` into an (E, F) temporary and scatter-adds it at
``edge_index[1]``, and ``GCNConv`` re-runs ``gcn_norm`` on every call (``cached=False``; call sites
node_classification_clean/models.py:31-37, 48-56).  Here the COO ``edge_index`` is converted once to a
destination-sorted int32 CSR (deterministic reduction order, no atomics), the GCN normalisation is computed once
per graph, and both are kept in a small identity-keyed cache."""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import torch

from . import ops

Tensor = torch.Tensor


class GraphCSR:
    """CSR + lazily computed GCN normalisation of one ``edge_index``."""

    def __init__(self, edge_index: Tensor, num_nodes: int, num_src_nodes: Optional[int] = None):
        self.csr = ops.csr_build(edge_index, num_nodes, num_src_nodes)
        self.num_nodes = num_nodes
        self._gcn: Optional[Tuple[Tensor, Tensor, Tensor]] = None
        self._edge_index = edge_index
        self._square = num_src_nodes is None or num_src_nodes == num_nodes
        self._transposed: Optional["GraphCSR"] = None
        self._gcn_t: Optional[Tuple[Tensor, Tensor]] = None

    edge_index = property(lambda self: self._edge_index)          # the COO tensor the CSR was built from
    rowptr = property(lambda self: self.csr.rowptr)
    col = property(lambda self: self.csr.col)
    perm = property(lambda self: self.csr.perm)

    def gcn_weights(self, edge_weight: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """(edge_weight in CSR order, self_weight) of PyG ``gcn_norm(add_self_loops=True)``."""
        if edge_weight is not None:
            w_csr = ops.gather_rows(edge_weight.to(torch.float32).reshape(-1, 1), self.csr.perm).view(-1)
            w, sw, _ = ops.gcn_norm(self.csr, w_csr)
            return w, sw
        if self._gcn is None:
            self._gcn = ops.gcn_norm(self.csr)
        return self._gcn[0], self._gcn[1]


    # -- backward: the adjoint of an aggregation is the same aggregation over the reversed edges --------------------
    def transposed(self) -> "GraphCSR":
        """CSR of the reversed edges (built on first use, i.e. only when something is differentiated)."""
        if self._transposed is None:
            if not self._square:
                raise NotImplementedError("the backward of a sharded (rectangular) graph is not implemented")
            self._transposed = GraphCSR(self._edge_index.flip(0), self.num_nodes)
        return self._transposed

    def gcn_weights_transposed(self) -> Tuple[Tensor, Tensor]:
        """gcn_norm edge weights in the entry order of ``transposed()`` (an edge keeps its weight) + the self weights."""
        if self._gcn_t is None:
            w, sw = self.gcn_weights()
            gt = self.transposed()
            by_edge = torch.empty_like(w)
            by_edge[self.csr.perm.long()] = w                       # forward CSR order -> original edge order (a permutation)
            w_t = ops.gather_rows(by_edge.view(-1, 1), gt.csr.perm).view(-1) if w.numel() else w
            self._gcn_t = (w_t, sw)
        return self._gcn_t


_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CACHE_SIZE = 4


def clear_cache() -> None:
    _CACHE.clear()


def get_graph(edge_index: Tensor, num_nodes: int) -> GraphCSR:
    """Cached ``GraphCSR``.  The key is the identity of the live ``edge_index`` storage (pointer, shape, strides,
    version counter); the entry keeps a reference to the tensor so the pointer cannot be recycled while cached."""
    if not edge_index.is_cuda:
        raise RuntimeError("kagnn_b200: edge_index must be a CUDA tensor (no CPU fallback)")
    key = (edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride()), edge_index._version,
           int(num_nodes), edge_index.device.index)
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        return hit[1]
    g = GraphCSR(edge_index, num_nodes)
    _CACHE[key] = (edge_index, g)
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return g
