"""Secondary measurements that ride on bench.py's JSON line (key ``extra``): the other BASELINE.json configurations on one
B200, each timed like the headline (CUDA events on the launching stream, L2 flushed by a 512 MB memset between iterations,
>= 3 warm-ups), with the parity of the timed model against the oracle where the oracle finishes in seconds.

  cora   C1  Cora-shaped KAGCN (2 layers, hidden 32, grid 5): north_star's ">= 10x the reference's PyTorch-CPU forward"
  zinc   C3  ZINC-shaped KAGIN (GINE, 4 layers, hidden 128), batch 1 024
  mutag  C5  FastKAN KAGIN hidden 256 grid 8, MUTAG-scaled batch 4 096 (fp32 and bf16)
  rmat   C4  Graph500 R-MAT 10 M nodes / 100 M edges, one KAGCN_Layer(128, 128, 5, 3): KAN launch + aggregation launch,
             HBM roofline of SURVEY.md section 8(d) (62.4 GB per layer)
"""
from __future__ import annotations

import os
import statistics
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import synth_graphs as SG  # noqa: E402


def _time(fn, flush, steps=10, warmup=3, pre_warm_s=0.2):
    with torch.no_grad():
        t_pre = time.perf_counter()                     # an idle GPU needs a moment of load to reach its boost clock (bench.py: PRE_WARM_S)
        while time.perf_counter() - t_pre < pre_warm_s:
            fn()
            torch.cuda.synchronize()
        for _ in range(warmup):
            flush.zero_()
            fn()
        ts = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda._sleep(200000)                   # keeps the GPU busy while the host enqueues (no launch gap inside the events)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def _cpu_time(fn, reps=3):
    ts = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, out


def cora(dev, flush, with_cpu=True):
    import kagnn_b200 as kb
    from oracle import kagnn_oracle as K
    n, f = 2708, 1433
    gen = torch.Generator().manual_seed(12345)
    und = torch.randint(0, n, (2, 5278), generator=gen)
    ei = torch.cat([und, und.flip(0)], dim=1)
    x = (torch.rand(n, f, generator=gen) < 18.17 / f).float()
    x = x / x.sum(1, keepdim=True).clamp(min=1)
    torch.manual_seed(12345)
    m = kb.GKAN_Nodes("gcn", 2, f, 32, 7, skip=True, grid_size=5, spline_order=3, dropout=0.0).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(dev)
    xd, eid = x.to(dev), ei.to(dev)
    from kagnn_b200 import models_node
    auto = models_node._AUTO_GRAPH_NODES
    models_node._AUTO_GRAPH_NODES = 0                      # launch by launch
    ms_eager = _time(lambda: m(xd, eid), flush, steps=20)
    models_node._AUTO_GRAPH_NODES = auto                   # the default: repeated eval forwards on the same inputs replay a CUDA graph
    ms = _time(lambda: m(xd, eid), flush, steps=20)
    out = {"workload": "Cora-shaped KAGCN: GKAN_Nodes('gcn', 2, 1433, 32, 7, skip=True, grid 5, order 3), N 2708, E 10556", "ms": ms,
           "ms_launch_by_launch": ms_eager, "nodes_per_s": n / ms * 1e3}
    # the same forward replayed from a CUDA graph (six dependent, mostly empty launches: launch latency is the cost)
    try:
        with torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                m(xd, eid)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                y_static = m(xd, eid)
        out["ms_cuda_graph"] = _time(graph.replay, flush, steps=20)
    except Exception as exc:  # pragma: no cover
        out["cuda_graph_error"] = repr(exc)[:160]
    # end to end: pinned host x / edge_index in, CSR build, forward, logits back to pinned host memory
    from kagnn_b200.graph import clear_cache
    xh, eh = x.pin_memory(), ei.pin_memory()
    yh = torch.empty(n, 7).pin_memory()

    def e2e():
        clear_cache()
        y = m(xh.to(dev, non_blocking=True), eh.to(dev, non_blocking=True))
        yh.copy_(y, non_blocking=True)
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(3):
            e2e()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            e2e()
            ts.append(time.perf_counter() - t0)
    out["ms_e2e"] = statistics.median(ts) * 1e3
    if with_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_ms, ref = _cpu_time(lambda: K.node_model_forward(sd, "gcn", x, ei, True))
        with torch.no_grad():
            y = m(xd, eid).cpu()
        out.update({"cpu_port_ms": cpu_ms, "cpu_threads": os.cpu_count(), "speedup_resident": cpu_ms / ms,
                    "speedup_e2e": cpu_ms / out["ms_e2e"], "rel_err_vs_oracle": K.rel_err(y, ref)})
    return out


def zinc(dev, flush, with_cpu=True):
    from kagnn_b200 import models_regr
    from oracle import kagnn_oracle as K
    data = SG.zinc_batch(1024, seed=12345)
    torch.manual_seed(12345)
    m = models_regr.KAGIN(1, 1, 4, 128, 2, 5, 3, 1, 0.0, True).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(dev)
    dd = data.to(dev)
    ms = _time(lambda: m(dd), flush)
    out = {"workload": "ZINC-shaped gr.KAGIN(1, 1, 4, 128, 2, 5, 3, 1, 0.0, True), batch 1024", "nodes": int(data.x.size(0)),
           "edges": int(data.edge_index.size(1)), "ms": ms, "nodes_per_s": data.x.size(0) / ms * 1e3, "graphs_per_s": 1024 / ms * 1e3}
    if with_cpu:
        cpu_ms, ref = _cpu_time(lambda: K.gr_kagin_forward(sd, K.Batch(data.x, data.edge_index, data.batch, data.edge_attr)), reps=1)
        with torch.no_grad():
            y = m(dd).cpu()
        out.update({"cpu_port_ms": cpu_ms, "rel_err_vs_oracle": K.rel_err(y, ref)})
    return out


def mutag(dev, flush, with_cpu=True):
    from kagnn_b200 import models_graph
    from oracle import kagnn_oracle as K
    data = SG.mutag_batch(4096, seed=12345)
    torch.manual_seed(12345)
    m = models_graph.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(dev)
    dd = data.to(dev)
    ms = _time(lambda: m(dd), flush)
    out = {"workload": "MUTAG-scaled gc.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0), batch 4096", "nodes": int(data.x.size(0)),
           "edges": int(data.edge_index.size(1)), "ms_fp32": ms, "nodes_per_s_fp32": data.x.size(0) / ms * 1e3}
    ref = None
    if with_cpu:
        cpu_ms, ref = _cpu_time(lambda: K.gc_kagin_forward(sd, K.Batch(data.x, data.edge_index, data.batch)), reps=1)
        with torch.no_grad():
            y = m(dd).cpu()
        out.update({"cpu_port_ms": cpu_ms, "rel_err_fp32_vs_oracle": K.rel_err(y, ref)})
    try:
        import kagnn_b200 as kb
        if hasattr(kb, "set_precision"):
            kb.set_precision("bf16")
            ms16 = _time(lambda: m(dd), flush)
            out.update({"ms_bf16": ms16, "nodes_per_s_bf16": data.x.size(0) / ms16 * 1e3})
            if ref is not None:
                with torch.no_grad():
                    out["rel_err_bf16_vs_fp32_oracle"] = K.rel_err(m(dd).cpu(), ref)
    finally:
        try:
            kb.set_precision("fp32")
        except Exception:
            pass
    return out


def rmat(dev, flush, n=10_000_000, e=100_000_000, hbm_gbs=6457.4):
    """C4 on one GPU: x (5.1 GB) and the CSR are resident, every launch streams far more than the 126 MB L2."""
    import kagnn_b200 as kb
    from kagnn_b200.graph import GraphCSR
    f = 128
    t0 = time.perf_counter()
    ei = SG.rmat_edges(n, e, seed=12345, device=dev)
    x = torch.randn(n, f, device=dev)
    torch.manual_seed(12345)
    conv = kb.KAGCN_Layer(f, 128, 5, 3).to(dev)
    g = GraphCSR(ei, n)
    g.gcn_weights()
    del ei
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    indeg_max = int((g.rowptr[1:] - g.rowptr[:-1]).max())
    h = torch.empty(n, 128, device=dev)
    y = torch.empty(n, 128, device=dev)
    from kagnn_b200 import _lib as L
    from kagnn_b200 import ops
    spec = conv.lin.kernel_specs()

    def kan():
        ops.fused_layer(ops.AggSpec(L.AGG_NONE, x), n, spec, out=h)

    def agg():
        conv.aggregate_transformed(h, g, out=y)

    ms_kan = _time(kan, flush, steps=5, warmup=3)
    ms_agg = _time(agg, flush, steps=5, warmup=3)
    ms = _time(lambda: (kan(), agg()), flush, steps=5, warmup=2)
    e_agg = e + n
    params = 128 * 128 * (8 + 2)
    alg = 4 * f * e_agg + 4 * e_agg + 4 * (n + 1) + 4 * e_agg + 4 * 128 * n + 4 * params        # SURVEY 8(d), GCN layer (no-reuse gather)
    comp = 4 * f * n + 4 * e_agg + 4 * (n + 1) + 4 * e_agg + 4 * 128 * n + 4 * params           # compulsory variant
    return {"workload": "Graph500 R-MAT (0.57, 0.19, 0.19, 0.05), 10 M nodes / 100 M directed edges, KAGCN_Layer(128, 128, 5, 3), fp32",
            "nodes": n, "edges": e, "max_in_degree": indeg_max, "setup_s": t_setup, "ms_kan": ms_kan, "ms_aggregate": ms_agg, "ms_layer": ms,
            "nodes_per_s": n / ms * 1e3,
            "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "compulsory_bytes": comp, "achieved": alg / (ms * 1e-3) / 1e9,
                         "peak": hbm_gbs, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / hbm_gbs,
                         "frac_compulsory": comp / (ms * 1e-3) / 1e9 / hbm_gbs,
                         "note": "layer = bare KAN launch (x -> h, tensor-bound: 3 bf16 products for fp32 parity) + aggregation launch (h -> y, "
                                 "HBM-bound); x (5.1 GB) >> L2 so the gather really reads DRAM"}}
