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

import os
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


def relabel_edges_first_use(edge_index: Tensor, rank: int, world: int, n_local: int, tile: int = 128):
    """Halo numbering for the OVERLAPPED pull (mode "pull"): the distinct remote sources are numbered in the order in which the
    128-row destination tiles first use them, so that a pull which copies halo rows 0, 1, 2, ... delivers what tile t needs before
    what tile t + 1 needs.  Returns (edge_index_local, halo_global, need): ``need[t]`` = number of leading halo rows referenced by
    tiles 0..t (int32, one entry per tile).  Index arithmetic only, no communication; one host synchronisation (the halo count)."""
    src, dst = edge_index[0], edge_index[1]
    lo = rank * n_local
    dev = edge_index.device
    n_tiles = (n_local + tile - 1) // tile
    remote = (src < lo) | (src >= lo + n_local)
    rs = src[remote]
    rt = torch.div(dst[remote] - lo, tile, rounding_mode="floor")
    if rs.numel() == 0:
        return (torch.stack([src - lo, dst - lo]), torch.zeros(0, dtype=torch.int64, device=dev),
                torch.zeros(n_tiles, dtype=torch.int32, device=dev))
    key = torch.sort(rs * n_tiles + rt).values                       # by source id, then by tile: first entry of a source = first use
    ks = torch.div(key, n_tiles, rounding_mode="floor")
    uniq, counts = torch.unique_consecutive(ks, return_counts=True)
    first_tile = key[torch.cumsum(counts, 0) - counts] - uniq * n_tiles
    order = torch.argsort(first_tile * (world * n_local) + uniq)     # first-use tile, then id
    halo_global = uniq[order]
    pos_of_uniq = torch.empty_like(order)
    pos_of_uniq[order] = torch.arange(order.numel(), device=dev)
    need = torch.searchsorted(first_tile[order].contiguous(), torch.arange(n_tiles, device=dev), right=True).to(torch.int32)
    pos = pos_of_uniq[torch.searchsorted(uniq, src.contiguous()).clamp(max=uniq.numel() - 1)]
    src_local = torch.where(remote, pos + n_local, src - lo)
    return torch.stack([src_local, dst - lo]), halo_global, need


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


@dataclass
class PeerPlan:
    """Plan of the in-kernel NVLink gather: just the shard's CSR over GLOBAL source ids (no halo numbering, no send lists)."""
    graph: GraphCSR
    n_local: int
    world: int


def push_plan_arrays(edge_index_global: Tensor, rank: int, world: int, n_local: int):
    """Index arithmetic of the "push" plan (no communication, no sort, no host synchronisation).  Every layer aggregates over
    [own rows | replica of the whole matrix]: a local source j becomes j - lo, a remote one n_local + j, so ONE edge list serves
    the input features and every hidden matrix.  Returns (source ids in that numbering, local target ids, need) where
    ``need`` is a byte map over all node ids: 1 <=> this rank references the (remote) row.  An id outside the global range
    stays out of range (-1) so that the deferred check of the CSR build reports it."""
    lo, n_tot = rank * n_local, n_local * world
    src, dst = edge_index_global[0], edge_index_global[1] - lo
    ok = (src >= 0) & (src < n_tot)
    local = (src >= lo) & (src < lo + n_local)
    src_rep = torch.where(ok, torch.where(local, src - lo, src + n_local), torch.full_like(src, -1))
    need = torch.zeros(n_tot, dtype=torch.uint8, device=edge_index_global.device)
    need[src.clamp(0, n_tot - 1)] = 1
    need[lo:lo + n_local] = 0
    return src_rep, dst, need


def push_peers(rank: int, world: int) -> List[int]:
    """The ranks a rank pushes to, in the bit order of its push mask / the order of its destination table."""
    return [q for q in range(world) if q != rank]


def push_row_masks(need: Tensor, rank: int, world: int, n_local: int, group=None) -> Tensor:
    """Collective (one all-gather of the byte maps): byte r of the result has bit i set <=> rank push_peers(rank)[i] references
    row r of this rank -- the rows the producing layer's kernel copies to that peer."""
    n_tot = n_local * world
    if dist.is_available() and dist.is_initialized() and world > 1:
        if need.is_cuda:
            maps = torch.empty(world, n_tot, dtype=torch.uint8, device=need.device)
            dist.all_gather_into_tensor(maps.view(-1), need, group=group)
        else:                                              # gloo (CPU tests)
            parts = [torch.empty_like(need) for _ in range(world)]
            dist.all_gather(parts, need, group=group)
            maps = torch.stack(parts)
    else:
        maps = need.view(1, -1).expand(world, -1)
    return masks_from_maps(maps, rank, world, n_local)


def masks_from_maps(maps: Tensor, rank: int, world: int, n_local: int) -> Tensor:
    lo = rank * n_local
    peers = torch.tensor(push_peers(rank, world), dtype=torch.int64, device=maps.device)
    if peers.numel() == 0:
        return torch.zeros(n_local, dtype=torch.uint8, device=maps.device)
    shifts = torch.arange(world - 1, dtype=torch.uint8, device=maps.device).view(-1, 1)
    return torch.bitwise_left_shift(maps[peers, lo:lo + n_local], shifts).sum(0, dtype=torch.uint8)


class ShardedNodeModel:
    """Runs a ``GKAN_Nodes`` / ``GFASTKAN_Nodes`` (eval mode, weights replicated on every rank) on this rank's node
    range: the fused plan of ``models_node._NodeModel.forward`` with the remote source rows of each aggregation fetched

    * ``mode="halo"``: by one NCCL ``all_to_all_single`` of the distinct remote rows in front of the layer (works for every
      model flavour, any backend; the CPU tests run it over gloo);
    * ``mode="pull"``: like "halo", but the distinct remote rows are pulled from their owners' symmetric memory by one copy
      kernel over NVLink (``kagnn_gather_rows_peer``): no pack, no send lists, no NCCL, plan built without communication;
    * ``mode="pull_overlap"``: the same pull running CONCURRENTLY with the layer on a few reserved SMs: halo rows numbered in the
      order the destination tiles first use them, per-chunk progress counters, the gather warps of the fused kernel wait for the
      prefix their tile needs (``kagnn_gather_rows_peer_ordered``, ``KagnnAggregate.halo_flags``).  Measured on two B200s it does
      NOT pay yet: an SM sustains only ~15 GB/s of peer loads, so the 16 SMs it may take from the layer pull at 240 GB/s where the
      all-SM kernel reaches 680 GB/s (profiles/README.md); it stays available for experiments and is parity-tested;
    * ``mode="push"``: the layer that PRODUCES a hidden matrix sends it: every rank keeps a replica of each hidden matrix (all
      ranks' rows) in symmetric memory, and one warp per CTA of the fused kernel copies every finished output tile into the
      peers' replicas with bulk copies while the following tiles are computed (``KagnnAggregate.push_y``; a per-row byte mask
      restricts the copies to the ranks that reference the row).  NVLink writes are posted, so -- unlike peer loads, which
      need every SM of the GPU to fill the link -- the transfer costs no SMs and hides behind the tensor-core pipeline; the
      next layer reads the replica as an ordinary halo matrix after one device barrier.  The input features of the first layer
      are still pulled (``kagnn_gather_rows_peer``): nothing runs before them that could hide the transfer;
    * ``mode="peer"``: by the gather warps of the fused kernel themselves, straight from the owners' memory over NVLink
      (``KagnnAggregate.peer_x``): every rank keeps its skip-concat buffer in ``torch.distributed._symmetric_memory``, the
      kernel receives the table of peer-mapped base pointers, and the only cross-rank traffic besides the row loads is one
      device-side barrier between layers.  The transfer overlaps the tensor-core pipeline tile by tile.  Available for the
      GIN flavour with ``skip=True`` and B-spline chains (what the pipelined kernel runs); ``mode="auto"`` picks it when it
      applies and falls back to ``"halo"`` otherwise."""

    def __init__(self, model, rank: int, world: int, n_local: int, group=None, mode: str = "halo", resident_x_halo: bool = False):
        # resident_x_halo: keep the halo rows of the INPUT features between forwards for as long as x is the same tensor at the
        # same version (static features of a static graph: the halo is part of the partitioned input, as in partition-with-halo
        # graph stores).  Off by default: every forward then fetches the input halo again.
        self.resident_x_halo = bool(resident_x_halo)
        if mode not in ("halo", "peer", "pull", "pull_overlap", "push", "auto"):
            raise ValueError("mode must be 'halo', 'peer', 'pull', 'pull_overlap', 'push' or 'auto'")
        self.model, self.rank, self.world, self.n_local, self.group = model, rank, world, n_local, group
        self._symm = {}
        self._side = {}
        # overlapped pull: SMs left to the pull kernel and its persistent blocks (4 per reserved SM keep ~1 MB of loads in flight)
        self.pull_sms = int(os.environ.get("KAGNN_PULL_SMS", "16"))
        self.pull_ctas = self.pull_sms                 # whole-SM blocks (1024 threads): one per reserved SM
        if mode == "auto" and self.peer_supported() and not self._probe_symmetric_memory():
            mode = "halo"                                   # NVLink peer memory not available here: NCCL transport
        if mode == "auto":
            # measured on 2 x B200, arxiv-shaped bench (ms/step): push 1.00, pull 1.12, peer 1.33, NCCL halo 1.65 -- "push" moves
            # only the distinct rows a peer references AND hides the transfer behind the producing layer
            mode = "push" if self.peer_supported() else "halo"
        elif mode in ("peer", "pull", "pull_overlap", "push") and not self.peer_supported():
            raise NotImplementedError("modes 'peer' / 'pull' need a GIN-flavour GKAN_Nodes / GFASTKAN_Nodes with skip=True, "
                                      "spline_order <= 3, G + k <= 8 (FastKAN: <= 8 centres), widths <= 128 and feature widths that "
                                      "are multiples of 4; GCN flavours and skip=False use mode='halo'")
        self.mode = mode

    def _probe_symmetric_memory(self) -> bool:
        """Collective: can every rank allocate and rendezvous symmetric memory?  (False -> the NCCL halo transport.)"""
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else dist.group.WORLD
            dev = torch.device("cuda", torch.cuda.current_device())
            t = symm_mem.empty((64,), dtype=torch.float32, device=dev)
            symm_mem.rendezvous(t, grp).barrier()
        except Exception:
            ok = 0
        try:
            flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            return bool(int(flag.item()))
        except Exception:
            return False

    def peer_supported(self) -> bool:
        from .conv import GINConv, GINEConv
        from .ekan import KAN
        from .fastkan import FastKAN
        m = self.model
        if not (getattr(m, "skip", False) and len(m.convs) and all(isinstance(c, GINConv) and not isinstance(c, GINEConv) for c in m.convs)):
            return False
        for c in m.convs:
            if isinstance(c.nn, KAN):
                for lay in c.nn.layers:
                    if lay.spline_order > 3 or lay.grid_size + lay.spline_order > 8 or lay.out_features > 128 or lay.in_features % 4:
                        return False
            elif isinstance(c.nn, FastKAN):
                # what the pipelined kernel runs for RBF chains: <= 8 centres, widths <= 128 (LayerNorm statistics of the gathered
                # row are taken in the kernel, which needs the row in one 128-column unit)
                for lay in c.nn.layers:
                    if lay.rbf.num_grids > 8 or lay.output_dim > 128 or lay.input_dim > 128 or lay.input_dim % 4:
                        return False
            else:
                return False
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
        except Exception:
            return False
        return True

    def prepare(self, edge_index_global: Tensor):
        if self.mode == "pull":
            # local index arithmetic only (no communication): distinct remote sources -> halo numbering
            ei_local, halo_global, _ = relabel_edges(edge_index_global, self.rank, self.world, self.n_local)
            n_halo = int(halo_global.numel())
            plan = PeerPlan(GraphCSR(ei_local, self.n_local, self.n_local + n_halo), self.n_local, self.world)
            plan.halo_ids = halo_global.to(torch.int32)
            plan.n_halo = n_halo
            return plan
        if self.mode == "push":
            return self._prepare_push(edge_index_global)
        if self.mode == "pull_overlap":
            # the same, numbered in first-use order for the pull that runs concurrently with the layer
            ei_local, halo_global, need = relabel_edges_first_use(edge_index_global, self.rank, self.world, self.n_local)
            n_halo = int(halo_global.numel())
            plan = PeerPlan(GraphCSR(ei_local, self.n_local, self.n_local + n_halo), self.n_local, self.world)
            plan.halo_ids = halo_global.to(torch.int32)
            plan.n_halo = n_halo
            plan.need = need
            plan.flags = torch.zeros(max(1, (n_halo + ops.HALO_CHUNK - 1) // ops.HALO_CHUNK), dtype=torch.int32, device=need.device)
            plan.epoch = 0
            plan.halo_buf = {}
            return plan
        if self.mode == "peer":
            lo = self.rank * self.n_local
            dst = edge_index_global[1] - lo
            ei = torch.stack([edge_index_global[0], dst])
            g = GraphCSR(ei, self.n_local, self.n_local * self.world)       # range check of both rows happens in the build
            return PeerPlan(g, self.n_local, self.world)
        plan = build_halo_plan(edge_index_global, self.rank, self.world, self.n_local, self.group)
        plan.exchange = HaloExchange(plan, self.group)
        return plan

    def push_peers(self):
        """The ranks this one pushes to, in the bit order of the push mask / the order of the destination table."""
        return push_peers(self.rank, self.world)

    def _prepare_push(self, edge_index_global: Tensor):
        """Plan of mode "push" -- no sort besides the CSR build, no ``unique``, no host synchronisation (push_plan_arrays /
        push_row_masks: index arithmetic + one all-gather of byte maps)."""
        n, w = self.n_local, self.world
        src_rep, dst, need = push_plan_arrays(edge_index_global, self.rank, w, n)
        plan = PeerPlan(GraphCSR(torch.stack([src_rep, dst]), n, n + n * w), n, w)
        plan.push_mask = push_row_masks(need, self.rank, w, n, self.group)
        plan.need = need
        return plan

    def input_buffer(self, n_features: int, device) -> Tensor:
        """This rank's input columns INSIDE the symmetric skip-concat buffer (modes peer / pull / push): a caller that writes x
        there (H2D copy, data loader) and passes this view to ``forward`` saves the copy of x into the buffer and one barrier."""
        m = self.model
        buf, _, _, _ = self._symm_buffer(self.n_local, n_features + len(m.convs) * m.bns[0].num_features, torch.device(device))
        return buf[:, :n_features]

    def _x_halo(self, plan, x: Tensor, pull: Callable[[], Tensor]) -> Tensor:
        if not self.resident_x_halo:
            return pull()
        key = (x.data_ptr(), x._version, tuple(x.shape))
        if getattr(plan, "_x_halo_key", None) != key:
            plan._x_halo, plan._x_halo_key = pull(), key
        return plan._x_halo

    def _side_stream(self, dev):
        if dev.index not in self._side:
            self._side[dev.index] = torch.cuda.Stream(device=dev)
        return self._side[dev.index]

    def _symm_buffer(self, n: int, width: int, dev, tag: int = 0):
        """Skip-concat buffer in symmetric memory + one device table of peer base pointers per column offset (cached)."""
        key = (n, width, dev.index, tag)
        if key not in self._symm:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else dist.group.WORLD
            if hasattr(symm_mem, "enable_symm_mem_for_group"):
                try:
                    symm_mem.enable_symm_mem_for_group(grp.group_name)
                except Exception:
                    pass
            buf = symm_mem.empty((n, width), dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(buf, grp)
            self._symm[key] = (buf, hdl, [int(q) for q in hdl.buffer_ptrs], {})
        return self._symm[key]

    @torch.no_grad()
    def _forward_peer(self, x: Tensor, plan: PeerPlan) -> Tensor:
        m = self.model
        n, f = x.shape
        n_mp = len(m.convs)
        hid = m.bns[0].num_features
        buf, hdl, ptrs, tables = self._symm_buffer(n, f + n_mp * hid, x.device)

        def table(col_off: int) -> Tensor:
            if col_off not in tables:
                tables[col_off] = torch.tensor([q + 4 * col_off for q in ptrs], dtype=torch.int64, device=x.device)
            return tables[col_off]

        reps = []
        if self.mode == "push":
            # replicas of the hidden matrices that a later layer aggregates (all but the last one) + my destination tables
            for l in range(n_mp - 1):
                rbuf, _, rptrs, rtab = self._symm_buffer(self.world * n, hid, x.device, tag=1 + l)
                if "push" not in rtab:
                    rtab["push"] = torch.tensor([rptrs[q] + 4 * self.rank * n * hid for q in self.push_peers()], dtype=torch.int64,
                                                device=x.device)
                reps.append((rbuf, rtab["push"]))
        # x may already live in the buffer (input_buffer()): then it was complete before this call, and the barrier that follows
        # (every rank is done reading the previous step's buffer -- also the peers that pull x from here) is the only one needed
        own_input = x.data_ptr() == buf.data_ptr() and x.stride(0) == buf.stride(0)
        hdl.barrier()
        if not own_input:
            ops.gather_rows(x, None, out=buf[:, :f])
        col = 0
        cur = buf[:, :f]
        for l, (conv, bn) in enumerate(zip(m.convs, m.bns)):
            if l > 0 or not own_input:
                hdl.barrier()                             # the slice read below is complete on every rank
            dst = buf[:, f + l * hid: f + (l + 1) * hid]
            if self.mode == "push":
                push = dict(push_y=reps[l][1], push_ld=hid, push_mask=plan.push_mask) if l < n_mp - 1 else {}
                if l == 0:
                    # input features: the marked remote rows are pulled into this rank's replica of x (nothing runs before
                    # layer 0 that could hide a push of x)
                    key = ("xrep", f, x.device.index)
                    if key not in self._symm:
                        self._symm[key] = torch.empty(self.world * n, f, dtype=torch.float32, device=x.device)
                    xrep = self._symm[key]
                    halo = self._x_halo(plan, x, lambda: ops.gather_rows_peer_masked(table(col), buf.stride(0), self.n_local, plan.need, f, xrep))
                    conv(cur, plan.graph, out=dst, post=m._folds[l].get(bn), x_halo=halo, **push)
                else:
                    conv(cur, plan.graph, out=dst, post=m._folds[l].get(bn), x_halo=reps[l - 1][0], **push)
            elif self.mode == "pull":
                # the distinct remote rows, copied once from their owners by a pull kernel on all SMs (678 GB/s measured on two
                # B200s), then the ordinary halo layer
                pull = lambda: ops.gather_rows_peer(table(col), buf.stride(0), self.n_local, plan.halo_ids, cur.size(1))  # noqa: E731
                halo = self._x_halo(plan, x, pull) if l == 0 else pull()
                conv(cur, plan.graph, out=dst, post=m._folds[l].get(bn), x_halo=halo)
            elif self.mode == "pull_overlap":
                # the distinct remote rows, pulled from their owners WHILE the layer runs: the pull kernel (side stream, a few SMs)
                # copies them in first-use order and raises one flag per 256 rows; the gather warps of the fused kernel wait for the
                # prefix their tile needs (KagnnAggregate.halo_flags), so NVLink time hides behind the tensor-core pipeline
                width = cur.size(1)
                if width not in plan.halo_buf:
                    plan.halo_buf[width] = torch.empty(max(plan.n_halo, 1), width, dtype=torch.float32, device=x.device)
                halo = plan.halo_buf[width]
                plan.epoch += 1
                main = torch.cuda.current_stream()
                side = self._side_stream(x.device)
                side.wait_stream(main)                     # the barrier above and the previous layer
                with torch.cuda.stream(side):
                    ops.gather_rows_peer_ordered(table(col), buf.stride(0), self.n_local, plan.halo_ids, width, halo, plan.flags,
                                                 plan.epoch, self.pull_ctas)
                conv(cur, plan.graph, out=dst, post=m._folds[l].get(bn), x_halo=halo, halo_need=plan.need, halo_flags=plan.flags,
                     halo_epoch=plan.epoch, reserve_sms=self.pull_sms)
                main.wait_stream(side)
            else:
                conv(cur, plan.graph, out=dst, post=m._folds[l].get(bn), peer_x=table(col), rows_per_rank=self.n_local)
            col = f + l * hid
            cur = dst
        return m.lay_out(buf)

    @torch.no_grad()
    def forward(self, x: Tensor, plan) -> Tensor:
        from .conv import GCNConv
        m = self.model
        if len(m.convs) == 0:                               # no message passing: nothing to exchange
            return m(x, torch.zeros(2, 0, dtype=torch.int64, device=x.device))
        if self.mode in ("peer", "pull", "pull_overlap", "push"):
            if not m._fusable():
                raise NotImplementedError("the sharded forward implements the eval-mode plan (BatchNorm folded)")
            x = x.to(torch.float32)
            if x.size(0) != self.n_local:
                raise ValueError("x must hold exactly the rows this rank owns")
            return self._forward_peer(x, plan)
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
