"""Time of one TRAINING step (forward + backward + Adam) of the arxiv-shaped GKAN_Nodes (gin, 3 layers, hidden 64, grid 5) on
one B200 -- the first, untuned backward path (kagnn_b200/csrc/backward.cu).  CUDA events, 3 warm-ups, median of 10.

    python scripts/train_step_time.py > gpurun_out/train_step.json
"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb
from kagnn_b200 import ops


def main():
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(12345)
    n, e, f, c = 169_343, 1_166_243, 128, 40
    ei = torch.randint(0, n, (2, e), generator=gen).to(dev)
    x = (torch.randn(n, f, generator=gen) * 0.3).to(dev)
    y = torch.randint(0, c, (n,), generator=gen).to(dev)
    torch.manual_seed(0)
    model = kb.GKAN_Nodes("gin", 3, f, 64, c, skip=True, grid_size=5, spline_order=3, hidden_layers=2, dropout=0.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        model.train()
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x, ei), y)
        loss.backward()
        opt.step()
        return loss

    def fwd_only():
        model.train()
        with torch.no_grad():
            return model(x, ei)

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

    l0 = float(step())
    t_step = timed(step)
    l1 = float(step())
    t_fwd = timed(fwd_only)
    c0 = ops.launch_count
    step()
    launches = ops.launch_count - c0
    # host time to enqueue one step (no synchronisation inside): if it is close to the step time, the step is launch-bound
    import time
    torch.cuda.synchronize()
    hs = []
    for _ in range(5):
        t0 = time.perf_counter()
        step()
        hs.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
    host_ms = statistics.median(hs)
    # the same step as ONE CUDA graph (whole-network capture, Adam(capturable=True)): what the kernels alone take
    graph_ms, graph_err = None, None
    try:
        opt2 = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)

        def step2():
            opt2.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(model(x, ei), y)
            loss.backward()
            opt2.step()

        model.train()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step2()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step2()
        graph_ms = timed(g.replay)
    except Exception as exc:  # pragma: no cover
        graph_err = repr(exc)[:200]
    if len(sys.argv) > 1 and sys.argv[1] == "--kernels":
        # per-kernel device time of one step (torch.profiler / CUPTI), for the notes under profiles/
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        tot = {}
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                d = tot.setdefault(ev.name[:110], [0, 0.0])
                d[0] += 1
                d[1] += (ev.time_range.end - ev.time_range.start) / 1e3
        for name, (cnt, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
            print(f"{ms:9.3f} ms  x{cnt:<3d} {name}", file=sys.stderr)
    # steps back to back (no synchronisation between them: the host runs ahead of the device, as in a loop that reads the
    # loss only every few steps)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        step()
    b.record()
    torch.cuda.synchronize()
    b2b_ms = a.elapsed_time(b) / 20
    print(json.dumps({"config": "arxiv-shaped GKAN_Nodes gin 3x64 grid 5, training step (fwd + bwd + Adam), batch-statistics BatchNorm",
                      "train_step_ms_back_to_back": b2b_ms,
                      "nodes": n, "edges": e, "train_step_ms": t_step, "train_mode_forward_ms": t_fwd,
                      "nodes_per_s_training": n / t_step * 1e3, "library_launches_per_step": launches,
                      "host_enqueue_ms_per_step": host_ms, "train_step_ms_cuda_graph": graph_ms, "cuda_graph_error": graph_err,
                      "loss_first": l0, "loss_after_14_steps": l1,
                      "max_memory_gb": torch.cuda.max_memory_allocated() / 2**30}))


if __name__ == "__main__":
    main()
