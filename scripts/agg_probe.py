"""Aggregation-only kernel under debug knobs (development aid): does the L1 carve-out / warps per SM limit the gather?"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kagnn_b200 import _lib as L, ops
from kagnn_b200.graph import get_graph
torch.manual_seed(0)
n, e = 169_343, 1_166_243
dev = torch.device("cuda")
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
ei = torch.randint(0, n, (2, e), device=dev)
g = get_graph(ei, n)
x = torch.randn(n, 128, device=dev) * 0.3
def run():
    return ops.fused_layer(ops.AggSpec(L.AGG_GIN, x, g.rowptr, g.col, self_scale=1.0), n, [])
ts = []
for i in range(15):
    flush.zero_(); torch.cuda._sleep(400000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    if i >= 3: ts.append(a.elapsed_time(b))
print(json.dumps({"grid": os.environ.get("KAGNN_DEBUG_AGG_GRID"), "smem": os.environ.get("KAGNN_DEBUG_AGG_SMEM"), "ms": round(statistics.median(ts), 4)}))
