"""BASELINE.json config 0 / north_star target: Cora-shaped KAGCN (2 layers, hidden 32, grid 5) forward on one B200 against the
reference's PyTorch-CPU forward of the same model on the box's host cores.  The CPU arm is the oracle port (torch ops in the
reference's order; torch_geometric is not installable); the GPU arm is the public module API, timed (a) with x / edge_index
resident, (b) end to end from pinned host x and edge_index to host logits, CSR build included.

    python scripts/cora_vs_cpu.py > gpurun_out/cora_vs_cpu.json
"""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb
from kagnn_b200 import graph as kgraph
from oracle import kagnn_oracle as K          # checker / CPU baseline only


def main():
    gen = torch.Generator().manual_seed(12345)
    n, f, c = 2708, 1433, 7
    und = torch.randint(0, n, (2, 5278), generator=gen)
    ei = torch.cat([und, und.flip(0)], dim=1)
    x = (torch.rand(n, f, generator=gen) < 18.17 / f).float()
    x = x / x.sum(1, keepdim=True).clamp(min=1)
    torch.manual_seed(0)
    m = kb.GKAN_Nodes("gcn", 2, f, 32, c, skip=True, grid_size=5, spline_order=3).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}

    # CPU arm: all host threads, best of 10 after 3 warm-ups
    threads = torch.get_num_threads()
    with torch.no_grad():
        for _ in range(3):
            y_cpu = K.node_model_forward(sd, "gcn", x, ei, True)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            K.node_model_forward(sd, "gcn", x, ei, True)
            ts.append((time.perf_counter() - t0) * 1e3)
    cpu_ms, cpu_med = min(ts), statistics.median(ts)

    dev = torch.device("cuda")
    m = m.to(dev)
    xh, eih = x.pin_memory(), ei.pin_memory()
    out_h = torch.empty(n, c).pin_memory()
    xd, eid = x.to(dev), ei.to(dev)

    def timed(fn, steps=50, warmup=10):
        with torch.no_grad():
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

    def e2e():
        kgraph.clear_cache()                              # a new edge_index every call: the CSR build is inside
        e = eih.to(dev, non_blocking=True)
        xx = xh.to(dev, non_blocking=True)
        out_h.copy_(m(xx, e), non_blocking=True)

    with torch.no_grad():
        y_gpu = m(xd, eid).cpu()
    resident = timed(lambda: m(xd, eid))
    end2end = timed(e2e)
    print(json.dumps({
        "config": "0: Cora-shaped KAGCN 2 layers hidden 32 grid 5 (N=2708, F=1433, E=10556)",
        "cpu_port_ms_best": cpu_ms, "cpu_port_ms_median": cpu_med, "cpu_threads": threads,
        "gpu_resident_ms": resident, "gpu_e2e_ms": end2end,
        "speedup_resident": cpu_ms / resident, "speedup_e2e": cpu_ms / end2end,
        "rel_err_vs_cpu_port": K.rel_err(y_gpu, y_cpu)}))


if __name__ == "__main__":
    main()
