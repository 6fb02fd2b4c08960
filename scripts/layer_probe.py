"""Per-launch timing of the bench's layer shapes, split by what the launch contains (development aid).

    [KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_<variant>.so] python scripts/layer_probe.py [tag]

For the ogbn-arxiv-shaped graph of bench.py (N = 169 343, E = 1 166 243) it times, with CUDA events on the launching stream,
L2 flushed between iterations (and once more with a warm L2):
  gin128 / gin64   one fused GIN layer (gather + KAN chain 128->64->64 / 64->64->64 + BatchNorm affine)
  kan128 / kan64   the same KAN chain on plain rows (no gather): what the basis producers + tensor pipe need alone
  agg128 / agg64   the aggregation alone (n_layers = 0)
  layout           KANLinear(320 -> 40) over two-part rows [x | h]
Prints one JSON line per case."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb
from kagnn_b200 import _lib as L
from kagnn_b200 import ops
from kagnn_b200.graph import get_graph

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(L.LIB_PATH)
torch.manual_seed(0)
n, e = 169_343, 1_166_243
dev = torch.device("cuda")
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
ei = torch.randint(0, n, (2, e), device=dev)
g = get_graph(ei, n)
x128 = torch.randn(n, 128, device=dev) * 0.3
x64 = torch.randn(n, 64, device=dev) * 0.3
h192 = torch.randn(n, 192, device=dev) * 0.5
bn = ops.Affine(torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1)
conv128 = kb.GIKANLayer(128, 64, 5, 3, 64, 2).to(dev)
conv64 = kb.GIKANLayer(64, 64, 5, 3, 64, 2).to(dev)
lay_out = kb.KANLinear(320, 40, grid_size=5, spline_order=3).to(dev)
out64 = torch.empty(n, 64, device=dev)


def agg_spec(x):
    return ops.AggSpec(L.AGG_GIN, x, g.rowptr, g.col, self_scale=1.0)


cases = {
    "gin128": lambda: conv128(x128, g, out=out64, post=bn),
    "gin64": lambda: conv64(x64, g, out=out64, post=bn),
    "kan128": lambda: ops.fused_layer(ops.AggSpec(L.AGG_NONE, x128), n, conv128.nn.kernel_specs(), post=bn, out=out64),
    "kan64": lambda: ops.fused_layer(ops.AggSpec(L.AGG_NONE, x64), n, conv64.nn.kernel_specs(), post=bn, out=out64),
    "agg128": lambda: ops.fused_layer(agg_spec(x128), n, []),
    "agg64": lambda: ops.fused_layer(agg_spec(x64), n, []),
    "layout": lambda: ops.fused_layer(ops.AggSpec(L.AGG_NONE, h192, x_head=x128), n, lay_out.kernel_specs()),
}


def time_case(fn, cold, iters=12):
    ts = []
    for i in range(3 + iters):
        if cold:
            flush.zero_()
        torch.cuda._sleep(400000)          # keeps the GPU busy while the host enqueues: no launch latency inside the events
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


with torch.no_grad():
    for name, fn in cases.items():
        med_c, min_c = time_case(fn, True)
        med_w, min_w = time_case(fn, False)
        print(json.dumps({"lib": tag, "case": name, "cold_ms": round(med_c, 4), "cold_min_ms": round(min_c, 4),
                          "warm_ms": round(med_w, 4), "warm_min_ms": round(min_w, 4)}), flush=True)
