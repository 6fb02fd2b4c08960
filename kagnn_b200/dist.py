"""Node-range sharding of the node-level models across GPUs (one process per GPU, ``torch.distributed``).

The reference is single-device (node_classification_clean/utils.py:14); this is the multi-GPU form of the same
forward.  Rank ``r`` owns the contiguous node range ``[r*n_local, (r+1)*n_local)``: the rows
of ``x`` / every hidden ``h`` for those nodes and all edges whose TARGET is one of them (so every aggregation is
complete on its owner and deterministic).  The only data the layer needs from elsewhere are the feature rows of
remote SOURCE nodes ("halo rows").  Per graph, once (``build_halo_plan``):

* the distinct remote sources of the local edges, sorted by global id (hence grouped by owner), become halo rows
  ``n_local .. n_local+n_halo-1`` of a local numbering; the CSR is built over that numbering;
* one index all-to-all tells every owner which of its rows each peer needs (``send_index`` / ``send_splits``).

Per layer, one exchange (``HaloExchange.__call__``): pack the requested rows with ``kagnn_gather_rows``, ONE
``all_to_all_single`` (NCCL over NVLink on the B200 box, gloo in the CPU tests), received straight into the
contiguous halo matrix that the fused kernel reads through ``KagnnAggregate.x_halo`` -- no unpack copy.  Weights are
replicated.  GCN normalisation needs the degree of remote sources: ``dinv`` of the halo rows travels once per graph
through the same exchange.  Graph-level batches (graph_classification / graph_regression) shard by graph and need no
exchange at all: every rank simply runs the ordinary model on its slice of the batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from .graph import GraphCSR

Tensor = torch.Tensor


@dataclass
class HaloPlan:
    rank: int
    world: int
    n_local: int
    n_halo: int
    halo_global: Tensor        # (n_halo,) int64 global ids of the halo rows, ascending (grouped by owner)
    recv_splits: List[int]     # halo rows received from each peer
    send_splits: List[int]     # owned rows sent to each peer
    send_index: Tensor         # (sum(send_splits),) int32 local row ids, grouped by destination peer
    edge_index_local: Tensor   # (2, E) int64: [source in the extended local numbering; target local]
    graph: Optional[GraphCSR] = None


def relabel_edges(edge_index: Tensor, rank: int, world: int, n_local: int):
    """Pure index arithmetic of the plan (no communication): returns (edge_index_local, halo_global, recv_counts).
    ``edge_index`` holds global ids; every target must be owned by ``rank``."""
    src, dst = edge_index[0], edge_index[1]
    lo = rank * n_local
    if edge_index.numel() and (int(dst.min()) < lo or int(dst.max()) >= lo + n_local):
        raise IndexError("kagnn_b200.dist: every edge target must be owned by this rank (shard edges by target)")
    if edge_index.numel() and (int(src.min()) < 0 or int(src.max()) >= n_local * world):
        raise IndexError("kagnn_b200.dist: edge source outside [0, world*n_local)")
    remote = (src < lo) | (src >= lo + n_local)
    halo_global = torch.unique(src[remote], sorted=True)
    pos = torch.searchsorted(halo_global, src.contiguous()) if halo_global.numel() else torch.zeros_like(src)
    src_local = torch.where(remote, pos + n_local, src - lo)
    owner = torch.div(halo_global, n_local, rounding_mode="floor")
    recv_counts = torch.bincount(owner, minlength=world)[:world]
    return torch.stack([src_local, dst - lo]), halo_global, recv_counts


def build_halo_plan(edge_index: Tensor, rank: int, world: int, n_local: int, group=None, build_csr: bool = True) -> HaloPlan:
    """Collective: every rank of ``group`` must call it with its own target-sharded ``edge_index`` (global ids)."""
    ei_local, halo_global, recv_counts = relabel_edges(edge_index, rank, world, n_local)
    dev = edge_index.device
    send_counts = torch.empty_like(recv_counts)
    dist.all_to_all_single(send_counts, recv_counts, group=group)
    recv_splits = [int(v) for v in recv_counts.tolist()]
    send_splits = [int(v) for v in send_counts.tolist()]
    # tell each owner which of its rows (owner-local ids) this rank needs
    owner = torch.div(halo_global, n_local, rounding_mode="floor")
    want = (halo_global - owner * n_local).contiguous()
    send_index = torch.empty(sum(send_splits), dtype=torch.int64, device=dev)
    dist.all_to_all_single(send_index, want, send_splits, recv_splits, group=group)
    plan = HaloPlan(rank, world, n_local, int(halo_global.numel()), halo_global, recv_splits, send_splits,
                    send_index.to(torch.int32), ei_local)
    if build_csr:
        plan.graph = ShardGraph(plan, group)
    return plan


def _pack_rows(x: Tensor, index: Tensor) -> Tensor:
    return ops.gather_rows(x, index)


class HaloExchange:
    """The per-layer exchange of one plan.  ``pack`` is the row gather that fills the send buffer: the library's
    ``kagnn_gather_rows`` in the product; the CPU (gloo) tests of this host logic inject a checker instead."""

    def __init__(self, plan: HaloPlan, group=None, pack: Callable[[Tensor, Tensor], Tensor] = _pack_rows):
        self.plan, self.group, self.pack = plan, group, pack
        self.bytes_sent = 0

    def __call__(self, x_local: Tensor, out: Optional[Tensor] = None) -> Tensor:
        p = self.plan
        if x_local.size(0) != p.n_local:
            raise ValueError("halo exchange expects the rank's owned rows")
        send = self.pack(x_local, p.send_index)
        if out is None:
            out = torch.empty(p.n_halo, x_local.size(1), dtype=x_local.dtype, device=x_local.device)
        dist.all_to_all_single(out, send, p.recv_splits, p.send_splits, group=self.group)
        self.bytes_sent += send.numel() * send.element_size()
        return out


class ShardGraph(GraphCSR):
    """CSR of one shard (targets = owned rows, sources = owned + halo rows) with the sharded gcn_norm."""

    def __init__(self, plan: HaloPlan, group=None):
        super().__init__(plan.edge_index_local, plan.n_local, plan.n_local + plan.n_halo)
        self.plan, self.group = plan, group

    def gcn_weights(self, edge_weight: Optional[Tensor] = None):
        if edge_weight is not None:
            raise NotImplementedError("user edge weights are not supported on sharded graphs")
        if self._gcn is None:
            sw, dinv = ops.gcn_degree(self.csr)
            dinv_halo = HaloExchange(self.plan, self.group)(dinv.view(-1, 1)).view(-1)
            w = ops.gcn_edge_weight(self.csr, torch.cat([dinv, dinv_halo]), dinv)
            self._gcn = (w, sw, dinv)
        return self._gcn[0], self._gcn[1]


class ShardedNodeModel:
    """Runs a ``GKAN_Nodes`` / ``GFASTKAN_Nodes`` (eval mode, weights replicated on every rank) on this rank's node
    range: the fused plan of ``models_node._NodeModel.forward`` with one halo exchange in front of each aggregation."""

    def __init__(self, model, rank: int, world: int, n_local: int, group=None):
        self.model, self.rank, self.world, self.n_local, self.group = model, rank, world, n_local, group

    def prepare(self, edge_index_global: Tensor) -> HaloPlan:
        plan = build_halo_plan(edge_index_global, self.rank, self.world, self.n_local, self.group)
        plan.exchange = HaloExchange(plan, self.group)
        return plan

    @torch.no_grad()
    def forward(self, x: Tensor, plan: HaloPlan) -> Tensor:
        from .conv import GCNConv
        m = self.model
        if not m._fusable():
            raise NotImplementedError("the sharded forward implements the eval-mode plan (BatchNorm folded)")
        x = x.to(torch.float32)
        n, f = x.shape
        if n != self.n_local:
            raise ValueError("x must hold exactly the rows this rank owns")
        g, xchg = plan.graph, plan.exchange
        n_mp = len(m.convs)
        hid = m.bns[0].num_features
        if m.skip:
            buf = torch.empty(n, f + n_mp * hid, dtype=torch.float32, device=x.device)
            ops.gather_rows(x, None, out=buf[:, :f])
            cur = buf[:, :f]
        else:
            buf, cur = None, x
        is_gcn = isinstance(m.convs[0], GCNConv)
        t = m.convs[0].transform(cur) if is_gcn else None
        for l, (conv, bn) in enumerate(zip(m.convs, m.bns)):
            dst = buf[:, f + l * hid: f + (l + 1) * hid] if m.skip else torch.empty(n, hid, dtype=torch.float32, device=x.device)
            if is_gcn:
                halo = xchg(t)
                pre = m._folds[l].get(bn, conv.bias.detach() if conv.bias is not None else None)
                nxt = m.convs[l + 1].lin.kernel_specs() if l + 1 < n_mp else []
                w, sw = g.gcn_weights()
                agg = ops.AggSpec(L.AGG_WEIGHTED, t, g.rowptr, g.col, edge_weight=w, self_weight=sw, x_halo=halo)
                t = ops.fused_layer(agg, n, nxt, pre=pre, agg_out=dst)
            else:
                halo = xchg(cur)
                conv(cur, g, out=dst, post=m._folds[l].get(bn), x_halo=halo)
            cur = dst
        return m.lay_out(buf if m.skip else cur)


def shard_batch_by_graph(batch: Tensor, num_graphs: int, rank: int, world: int):
    """Graph-level batches: rank r owns graphs [r*B/P, (r+1)*B/P) -- returns (first_graph, last_graph, node mask)."""
    per = (num_graphs + world - 1) // world
    g0, g1 = rank * per, min(num_graphs, (rank + 1) * per)
    return g0, g1, (batch >= g0) & (batch < g1)
