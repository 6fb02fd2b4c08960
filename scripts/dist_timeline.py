"""GPU timeline of one sharded forward step (development aid): torch.profiler (CUPTI) kernel start / duration on rank 0,
plus the host time it takes to enqueue a step.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 scripts/dist_timeline.py push"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench as B
from kagnn_b200 import dist as kd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model = B.model_state().to(dev)
n = B.N_NODES
x, ei = B.synth_graph(n, B.N_EDGES, B.N_FEAT, 12345 + rank, n_src=n * world)
x, ei = x.to(dev), ei.to(dev)
ei[1] += rank * n
mode = sys.argv[1] if len(sys.argv) > 1 else "push"
runner = kd.ShardedNodeModel(model, rank, world, n, mode=mode)
plan = runner.prepare(ei)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
with torch.no_grad():
    for _ in range(5):
        flush.zero_(); runner.forward(x, plan)
    dist.barrier(); torch.cuda.synchronize()
    host = []
    for _ in range(10):
        flush.zero_()
        t0 = time.perf_counter(); runner.forward(x, plan); host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"mode {runner.mode}: host enqueue ms/step median {sorted(host)[5]:.3f} min {min(host):.3f}", flush=True)
    dist.barrier(); torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            flush.zero_(); runner.forward(x, plan)
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        prev_end = t0
        for e in evs:
            s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
            print(f"{s:9.1f} us  +{d:8.1f} us  gap {e.time_range.start - prev_end:7.1f}  {e.name[:90]}", flush=True)
            prev_end = max(prev_end, e.time_range.end)
dist.barrier()
dist.destroy_process_group()
