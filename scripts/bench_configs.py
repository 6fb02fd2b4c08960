"""Forward time of the other BASELINE.json configurations on one B200 (secondary numbers; bench.py carries the headline
config 1).  Synthetic inputs per SURVEY.md section 8(d); CUDA events, L2 flushed between iterations, 5 warm-ups.

    python scripts/bench_configs.py > gpurun_out/configs.jsonl
"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb
from kagnn_b200 import models_graph, models_regr, ops

dev = torch.device("cuda")
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timeit(fn, steps=20, warmup=5):
    with torch.no_grad():
        for _ in range(warmup):
            flush.zero_()
            fn()
        ts = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


class Data:
    def __init__(self, x, edge_index, batch, edge_attr=None, num_graphs=None):
        self.x, self.edge_index, self.batch, self.edge_attr, self.num_graphs = x, edge_index, batch, edge_attr, num_graphs


def batch_of_graphs(n_graphs, mean_nodes, edges_per_graph, gen):
    sizes = torch.poisson(torch.full((n_graphs,), float(mean_nodes)), generator=gen).clamp(min=2).long()
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
    batch = torch.repeat_interleave(torch.arange(n_graphs), sizes)
    und = edges_per_graph // 2
    g_of_e = torch.arange(n_graphs).repeat_interleave(und)
    a = (torch.rand(g_of_e.numel(), generator=gen) * sizes[g_of_e]).long() + ptr[g_of_e]
    b = (torch.rand(g_of_e.numel(), generator=gen) * sizes[g_of_e]).long() + ptr[g_of_e]
    ei = torch.cat([torch.stack([a, b]), torch.stack([b, a])], dim=1)
    return int(ptr[-1]), batch, ei


def main():
    gen = torch.Generator().manual_seed(12345)
    out = []
    # config 0: Cora-shaped KAGCN
    n, f = 2708, 1433
    und = torch.randint(0, n, (2, 5278), generator=gen)
    ei = torch.cat([und, und.flip(0)], dim=1).to(dev)
    x = (torch.rand(n, f, generator=gen) < 18.17 / f).float()
    x = (x / x.sum(1, keepdim=True).clamp(min=1)).to(dev)
    m = kb.GKAN_Nodes("gcn", 2, f, 32, 7, skip=True, grid_size=5, spline_order=3).eval().to(dev)
    l0 = ops.launch_count
    ms = timeit(lambda: m(x, ei))
    out.append({"config": "0: Cora-shaped KAGCN 2 layers hidden 32 grid 5", "nodes": n, "ms": ms, "nodes_per_s": n / ms * 1e3})
    # the same forward captured in a CUDA graph (launch-latency-bound problem: replay removes the host side of 6 launches)
    try:
        with torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                m(x, ei)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                m(x, ei)
        ms = timeit(graph.replay)
        out.append({"config": "0: Cora-shaped KAGCN, CUDA-graph replay", "nodes": n, "ms": ms, "nodes_per_s": n / ms * 1e3})
    except Exception as exc:  # pragma: no cover
        out.append({"config": "0: Cora-shaped KAGCN, CUDA-graph replay", "error": repr(exc)[:200]})
    # config 2: ZINC-shaped KAGIN (GINE), batch 1024
    nn_, batch, ei = batch_of_graphs(1024, 23.15, 50, gen)
    xz = torch.randint(0, 28, (nn_, 1), generator=gen)
    ea = torch.randint(1, 4, (ei.size(1),), generator=gen)
    mz = models_regr.KAGIN(1, 1, 4, 128, 2, 5, 3, 1, 0.0, True).eval().to(dev)
    dz = Data(xz.to(dev), ei.to(dev), batch.to(dev), ea.to(dev), 1024)
    ms = timeit(lambda: mz(dz))
    out.append({"config": "2: ZINC-shaped KAGIN (GINE) 4 layers hidden 128 batch 1024", "nodes": nn_, "graphs": 1024, "ms": ms,
                "nodes_per_s": nn_ / ms * 1e3, "graphs_per_s": 1024 / ms * 1e3})
    # config 4: FastKAN KAGIN hidden 256 grid 8, MUTAG-scaled batch 4096 (fp32 here; the bf16 variant is not built)
    nn_, batch, ei = batch_of_graphs(4096, 17.93, 40, gen)
    xm = torch.nn.functional.one_hot(torch.randint(0, 7, (nn_,), generator=gen), 7).float()
    mm = models_graph.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0).eval().to(dev)
    dm = Data(xm.to(dev), ei.to(dev), batch.to(dev), None, 4096)
    ms = timeit(lambda: mm(dm))
    out.append({"config": "4: FastKAN KAGIN hidden 256 grid 8 batch 4096 (fp32)", "nodes": nn_, "graphs": 4096, "ms": ms,
                "nodes_per_s": nn_ / ms * 1e3, "graphs_per_s": 4096 / ms * 1e3})
    # FastKAN inside the Optuna ranges of the reference (hidden <= 128): arxiv-shaped GFASTKAN_Nodes, pipelined vs general kernel
    n, e = 169_343, 1_166_243
    ei = torch.randint(0, n, (2, e), generator=gen).to(dev)
    xa = (torch.randn(n, 128, generator=gen) * 0.3).to(dev)
    mf = kb.GFASTKAN_Nodes("gin", 3, 128, 64, 40, skip=True, grid_size=8, hidden_layers=2).eval().to(dev)
    for variant, tag in ((0, "pipelined kernel"), (1, "general tcgen05 kernel")):
        ops.set_tc_variant(variant)
        ms = timeit(lambda: mf(xa, ei))
        out.append({"config": f"arxiv-shaped GFASTKAN_Nodes gin 3 layers hidden 64 grid 8 ({tag})", "nodes": n, "ms": ms,
                    "nodes_per_s": n / ms * 1e3})
    ops.set_tc_variant(0)
    # config 3 (one shard's worth): RMAT-like skewed graph, KAGCN layer hidden 128 -- 1/8 of 10 M nodes / 100 M edges
    n, e = 1_250_000, 12_500_000
    src = (torch.rand(e, generator=gen) ** 3 * n).long()           # heavy-tailed source popularity
    dst = (torch.rand(e, generator=gen) ** 2 * n).long()           # and in-degree
    ei = torch.stack([src, dst]).to(dev)
    xr = torch.randn(n, 128, generator=gen).to(dev)
    conv = kb.KAGCNConv(128, 128, 5, 3).to(dev)
    ms = timeit(lambda: conv(xr, ei), steps=10)
    out.append({"config": "3 (one shard): skewed 1.25 M nodes / 12.5 M edges, KAGCNConv 128->128", "nodes": n, "edges": e, "ms": ms,
                "nodes_per_s": n / ms * 1e3})
    for o in out:
        o["counters"] = ops.launch_counters()
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
