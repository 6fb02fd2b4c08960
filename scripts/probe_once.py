"""Runs the layer-probe cases back to back (for ncu captures): gin128, kan128, layout, repeated."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb
from kagnn_b200 import _lib as L, ops
from kagnn_b200.graph import get_graph
torch.manual_seed(0)
n, e = 169_343, 1_166_243
dev = torch.device("cuda")
ei = torch.randint(0, n, (2, e), device=dev)
g = get_graph(ei, n)
x128 = torch.randn(n, 128, device=dev) * 0.3
h192 = torch.randn(n, 192, device=dev) * 0.5
bn = ops.Affine(torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1)
conv128 = kb.GIKANLayer(128, 64, 5, 3, 64, 2).to(dev)
lay_out = kb.KANLinear(320, 40, grid_size=5, spline_order=3).to(dev)
out64 = torch.empty(n, 64, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with torch.no_grad():
    for _ in range(reps):
        conv128(x128, g, out=out64, post=bn)
        ops.fused_layer(ops.AggSpec(L.AGG_NONE, x128), n, conv128.nn.kernel_specs(), post=bn, out=out64)
        ops.fused_layer(ops.AggSpec(L.AGG_NONE, h192, x_head=x128), n, lay_out.kernel_specs())
    torch.cuda.synchronize()
