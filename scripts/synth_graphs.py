"""Synthetic inputs of BASELINE.json's configurations: datasets cannot be downloaded, so every
benchmark and full-size parity test builds graphs of the stated shape from a seed.  Shared by bench.py, the scripts and
tests/ (test / measurement infrastructure, not product code)."""
from __future__ import annotations

import math

import torch


class Data:
    """Duck-typed stand-in of a PyG ``Batch`` (the reference's graph-level models read these attributes only)."""

    def __init__(self, x, edge_index, batch, edge_attr=None, num_graphs=None):
        self.x, self.edge_index, self.batch, self.edge_attr, self.num_graphs = x, edge_index, batch, edge_attr, num_graphs

    def to(self, device):
        mv = lambda t: t.to(device) if isinstance(t, torch.Tensor) else t
        return Data(mv(self.x), mv(self.edge_index), mv(self.batch), mv(self.edge_attr), self.num_graphs)


def batch_of_graphs(n_graphs: int, mean_nodes: float, edges_per_graph: int, gen: torch.Generator):
    """``n_graphs`` small graphs in one disjoint union: node counts Poisson(mean_nodes) clamped >= 2, ``edges_per_graph / 2``
    random intra-graph pairs emitted in both directions.  Returns (num_nodes, batch, edge_index)."""
    sizes = torch.poisson(torch.full((n_graphs,), float(mean_nodes)), generator=gen).clamp(min=2).long()
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
    batch = torch.repeat_interleave(torch.arange(n_graphs), sizes)
    und = edges_per_graph // 2
    g_of_e = torch.arange(n_graphs).repeat_interleave(und)
    a = (torch.rand(g_of_e.numel(), generator=gen) * sizes[g_of_e]).long() + ptr[g_of_e]
    b = (torch.rand(g_of_e.numel(), generator=gen) * sizes[g_of_e]).long() + ptr[g_of_e]
    ei = torch.cat([torch.stack([a, b]), torch.stack([b, a])], dim=1)
    return int(ptr[-1]), batch, ei


def zinc_batch(n_graphs: int = 1024, seed: int = 12345) -> Data:
    """BASELINE config C3: ZINC-subset-shaped batch (23.15 nodes and ~50 directed edges per graph, atom codes 0..27 as an
    int64 (N, 1) column, bond codes 1..3 as an int64 (E,) vector)."""
    gen = torch.Generator().manual_seed(seed)
    n, batch, ei = batch_of_graphs(n_graphs, 23.15, 50, gen)
    x = torch.randint(0, 28, (n, 1), generator=gen)
    ea = torch.randint(1, 4, (ei.size(1),), generator=gen)
    return Data(x, ei, batch, ea, n_graphs)


def mutag_batch(n_graphs: int = 4096, seed: int = 12345) -> Data:
    """BASELINE config C5: MUTAG-scaled batch (17.93 nodes and ~40 directed edges per graph, one-hot(7) node features)."""
    gen = torch.Generator().manual_seed(seed)
    n, batch, ei = batch_of_graphs(n_graphs, 17.93, 40, gen)
    x = torch.nn.functional.one_hot(torch.randint(0, 7, (n,), generator=gen), 7).float()
    return Data(x, ei, batch, None, n_graphs)


def rmat_edges(num_nodes: int, num_edges: int, seed: int = 12345, device="cpu", abcd=(0.57, 0.19, 0.19, 0.05),
               chunk: int = 1 << 24) -> torch.Tensor:
    """BASELINE config C4: Graph500 R-MAT edge list (a, b, c, d) = (0.57, 0.19, 0.19, 0.05) on 2^scale >= num_nodes vertices,
    vertex ids scrambled by a bijective multiplicative hash (Graph500 permutes labels so that the hubs are not the low ids)
    and reduced mod num_nodes; duplicates and self loops kept.  Returns an int64 (2, num_edges) COO tensor [source; target]."""
    scale = max(1, math.ceil(math.log2(max(num_nodes, 2))))
    a, b, c, _ = abcd
    gen = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty(2, num_edges, dtype=torch.int64, device=device)
    mask = (1 << scale) - 1
    for lo in range(0, num_edges, chunk):
        m = min(chunk, num_edges - lo)
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(m, generator=gen, device=device)
            right = ((r >= a) & (r < a + b)) | (r >= a + b + c)          # quadrants b and d: target bit set
            down = r >= a + b                                           # quadrants c and d: source bit set
            src = (src << 1) | down.long()
            dst = (dst << 1) | right.long()
        src = ((src * 0x9E3779B1 + 0x7F4A7C15) & mask) % num_nodes     # odd multiplier: a bijection on 2^scale ids
        dst = ((dst * 0x85EBCA6B + 0x165667B1) & mask) % num_nodes
        out[0, lo:lo + m] = src
        out[1, lo:lo + m] = dst
    return out
