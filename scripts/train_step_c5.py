"""Training step (forward + backward + Adam) of the C5-shaped model -- gc.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0) on the MUTAG-scaled batch of
4 096 graphs -- with the tensor-core FastKAN gradients (default) and with the fp32 CUDA-core ones (kagnn_set_backward_path(1)).
    python scripts/train_step_c5.py > gpurun_out/train_step_c5.json"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import synth_graphs as SG
from kagnn_b200 import models_graph, ops


def main():
    dev = torch.device("cuda")
    data = SG.mutag_batch(4096, seed=12345).to(dev)
    torch.manual_seed(0)
    model = models_graph.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    labels = (torch.arange(4096, device=dev) % 2)

    def step():
        model.train()
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.nll_loss(model(data), labels)
        loss.backward()
        opt.step()
        return loss

    def timed(steps=10, warmup=3):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    if len(sys.argv) > 1 and sys.argv[1] == "--host-profile":
        import cProfile
        import io
        import pstats
        import time
        from kagnn_b200 import ops as _ops
        c0 = _ops.launch_count
        step()
        print("library launches per step", _ops.launch_count - c0, file=sys.stderr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"enqueue {1e3 * (t1 - t0) / 20:.3f} ms/step, with drain {1e3 * (time.perf_counter() - t0) / 20:.3f} ms/step", file=sys.stderr)
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(20):
            step()
        pr.disable()
        torch.cuda.synchronize()
        for key in ("tottime", "cumtime"):
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(40)
            print(buf.getvalue()[:8000], file=sys.stderr)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--kernels":
        from torch.profiler import profile, ProfilerActivity
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        tot = {}
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                d = tot.setdefault(ev.name[:100], [0, 0.0])
                d[0] += 1
                d[1] += (ev.time_range.end - ev.time_range.start) / 1e3
        for name, (cnt, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:14]:
            print(f"{ms:9.3f} ms  x{cnt:<3d} {name}", file=sys.stderr)
    out = {"config": "C5-shaped gc.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0), batch 4096 graphs, training step (fwd + bwd + Adam)",
           "nodes": int(data.x.size(0)), "edges": int(data.edge_index.size(1))}
    ops.set_backward_path(0)
    out["train_step_ms_tensor_core_gradients"] = timed()
    ops.set_backward_path(1)
    try:
        out["train_step_ms_fp32_gradients"] = timed(steps=5, warmup=2)
    finally:
        ops.set_backward_path(0)
    with torch.no_grad():
        model.eval()
        for _ in range(3):
            model(data)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        model(data)
        b.record()
        torch.cuda.synchronize()
        out["eval_forward_ms"] = a.elapsed_time(b)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
