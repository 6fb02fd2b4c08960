"""Is a small-batch forward bound by the host or by the device?  Eager time (CUDA events, L2 flushed), the same launches replayed
from a CUDA graph, the host time to enqueue one call, and the library launches per call -- for the ZINC- and MUTAG-shaped models.

    python scripts/host_vs_device.py > gpurun_out/host_vs_device.json
"""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import synth_graphs as SG
from kagnn_b200 import models_graph, models_regr, ops


def measure(m, dd):
    with torch.no_grad():
        for _ in range(3):
            m(dd)
        torch.cuda.synchronize()
        c0 = ops.launch_count
        m(dd)
        launches = ops.launch_count - c0
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m(dd)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t0 = time.perf_counter()
        for _ in range(50):
            m(dd)
        host = (time.perf_counter() - t0) / 50 * 1e3
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m(dd)
        g.replay()
        torch.cuda.synchronize()
        tg = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            tg.append(a.elapsed_time(b))
    return {"eager_ms": statistics.median(ts), "cuda_graph_ms": statistics.median(tg), "host_enqueue_ms": host, "library_launches": launches}


def main():
    dev = torch.device("cuda")
    out = {}
    torch.manual_seed(12345)
    m = models_regr.KAGIN(1, 1, 4, 128, 2, 5, 3, 1, 0.0, True).eval().to(dev)
    out["zinc"] = measure(m, SG.zinc_batch(1024, seed=12345).to(dev))
    torch.manual_seed(12345)
    m = models_graph.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0).eval().to(dev)
    out["mutag"] = measure(m, SG.mutag_batch(4096, seed=12345).to(dev))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
