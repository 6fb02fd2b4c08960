"""arxiv-shaped KAGIN with grid_size 8, spline_order 3 (eleven coefficients per pair): forward time of the windowed tensor-core
path (two virtual features of eight slots per input) against the general fp32 kernel that such layers used before.

    python scripts/windows_time.py > gpurun_out/windows_time.json
"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb
from kagnn_b200 import ops


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(12345)
    n, e, f, c = 169_343, 1_166_243, 128, 40
    ei = torch.randint(0, n, (2, e), generator=gen).to(dev)
    x = (torch.randn(n, f, generator=gen) * 0.3).to(dev)
    out = {"config": "arxiv-shaped GKAN_Nodes gin 3x64, grid_size 8, spline_order 3, KAN depth 2, eval forward", "nodes": n, "edges": e}
    res = {}
    for name, patch in (("windowed_tensor_core", None), ("general_fp32", lambda *a: False)):
        torch.manual_seed(0)
        m = kb.GKAN_Nodes("gin", 3, f, 64, c, skip=True, grid_size=8, spline_order=3, hidden_layers=2, dropout=0.0).to(dev).eval()
        saved = ops.tc_supported
        if patch is not None:
            ops.tc_supported = patch
        try:
            with torch.no_grad():
                c0 = ops.launch_counters()
                res[name] = m(x, ei)
                c1 = ops.launch_counters()
                out[name + "_ms"] = timed(lambda: m(x, ei))
                out[name + "_launches"] = {k: c1[k] - c0[k] for k in c0}
        finally:
            ops.tc_supported = saved
    d = (res["windowed_tensor_core"] - res["general_fp32"]).abs().max().item()
    out["max_abs_diff_between_paths"] = d
    out["rel_diff_between_paths"] = d / res["general_fp32"].abs().max().item()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
