"""Per-graph state the reference recomputes on every forward, built once and cached.

PyG's ``MessagePassing.propagate`` gathers ``x[edge_index[0]]`` into an (E, F) temporary and scatter-adds it at
``edge_index[1]``, and ``GCNConv`` re-runs ``gcn_norm`` on every call (``cached=False``; call sites
node_classification_clean/models.py:31-37, 48-56).  Here the COO ``edge_index`` is converted once to a
destination-sorted int32 CSR (deterministic reduction order, no atomics), the GCN normalisation is computed once
per graph, and both are kept in a small identity-keyed cache."""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from . import ops

Tensor = torch.Tensor


class GraphCSR:
    """CSR + lazily computed GCN normalisation of one ``edge_index``."""

    def __init__(self, edge_index: Tensor, num_nodes: int, num_src_nodes: Optional[int] = None):
        self.csr = ops.csr_build(edge_index, num_nodes, num_src_nodes)
        _watch_index_flag(self.csr)
        self.num_nodes = num_nodes
        self._gcn: Optional[Tuple[Tensor, Tensor, Tensor]] = None
        self._edge_index = edge_index
        self._square = num_src_nodes is None or num_src_nodes == num_nodes
        self._transposed: Optional["GraphCSR"] = None
        self._gcn_t: Optional[Tuple[Tensor, Tensor]] = None

    def with_sources(self, col: Tensor, num_src_nodes: int) -> "GraphCSR":
        """The same rows and entry order over a RENUMBERED source set (node-sharded graphs: halo numbering -> replica numbering):
        no second sort, the column array is the only thing that differs."""
        g = object.__new__(GraphCSR)
        g.csr = ops.CSR(self.csr.rowptr, col, self.csr.perm, self.csr.num_rows, int(num_src_nodes), self.csr.err_flag)
        g.num_nodes, g._gcn, g._edge_index, g._square, g._transposed, g._gcn_t = self.num_nodes, None, self._edge_index, False, None, None
        return g

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


# ---- deferred range check of edge_index -------------------------------------------------------------------------------
# kagnn_csr_build never synchronises: it clamps an out-of-range node id to 0 and raises a device-side flag.  PyG / ATen would
# raise (or device-assert) on such an edge_index, so the flag must not go unread: its 4 bytes are copied to pinned host memory
# behind the build and examined -- without blocking -- whenever a later graph call finds the copy complete, which raises
# IndexError then (like any asynchronous CUDA error surfaces at a later call).  KAGNN_CHECK_INDICES=1 checks synchronously.
_PENDING: list = []


def _watch_index_flag(csr) -> None:
    if csr.err_flag is None or not csr.err_flag.is_cuda or torch.cuda.is_current_stream_capturing():
        return
    host = torch.empty(1, dtype=torch.int32).pin_memory()
    host.copy_(csr.err_flag, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _PENDING.append((ev, host, csr.num_rows))
    if os.environ.get("KAGNN_CHECK_INDICES") == "1":
        ev.synchronize()
    poll_index_flags()


def poll_index_flags(block: bool = False) -> None:
    """Raise IndexError if a finished CSR build saw node ids outside [0, num_nodes); ``block=True`` waits for every build."""
    if not _PENDING or torch.cuda.is_current_stream_capturing():
        return
    bad = None
    for item in list(_PENDING):
        ev, host, n = item
        if block:
            ev.synchronize()
        if ev.query():
            _PENDING.remove(item)
            if int(host[0]) != 0:
                bad = n
    if bad is not None:
        raise IndexError(f"kagnn_b200: edge_index contains node ids outside [0, {bad}) (detected by an earlier CSR build)")


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
    poll_index_flags()
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        return hit[1]
    g = GraphCSR(edge_index, num_nodes)
    _CACHE[key] = (edge_index, g)
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return g
