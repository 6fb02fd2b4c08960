"""Where the sharded forward spends its time (development aid): CUDA-event timing of pack / all_to_all / layer launches.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 scripts/dist_breakdown.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench as B
from kagnn_b200 import dist as kd, ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model = B.model_state().to(dev)
n = B.N_NODES
x, ei = B.synth_graph(n, B.N_EDGES, B.N_FEAT, 12345 + rank, n_src=n * world)
x, ei = x.to(dev), ei.to(dev)
ei[1] += rank * n
mode = sys.argv[1] if len(sys.argv) > 1 else 'halo'
runner = kd.ShardedNodeModel(model, rank, world, n, mode=mode)
plan = runner.prepare(ei)
print(f"rank {rank}: mode {runner.mode}", flush=True)
orig_call = kd.HaloExchange.__call__
rec = []
def timed_call(self, x_local, out=None):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    send = self.pack(x_local, self.plan.send_index)
    e[1].record()
    if out is None:
        out = torch.empty(self.plan.n_halo, x_local.size(1), dtype=x_local.dtype, device=x_local.device)
    dist.all_to_all_single(out, send, self.plan.recv_splits, self.plan.send_splits, group=self.group)
    e[2].record()
    rec.append((x_local.size(1), e))
    return out
kd.HaloExchange.__call__ = timed_call
with torch.no_grad():
    for it in range(8):
        rec.clear()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); runner.forward(x, plan); e1.record()
        torch.cuda.synchronize()
        if it >= 3 and rank == 0:
            print("total %.3f ms | " % e0.elapsed_time(e1) + " ; ".join("w%d pack %.3f a2a %.3f" % (w, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])) for w, e in rec), flush=True)
dist.destroy_process_group()
