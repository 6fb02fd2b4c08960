"""NVLink pull kernels in isolation (development aid; torchrun, >= 2 GPUs): time of kagnn_gather_rows_peer (all SMs) and of
kagnn_gather_rows_peer_ordered (whole-SM blocks) for the bench's halo of one layer."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
from kagnn_b200 import ops, dist as kd
n_local, e = 169_343, 1_166_243
g = torch.Generator().manual_seed(5 + rank)
src = torch.randint(0, n_local * world, (e,), generator=g)
dst = torch.randint(0, n_local, (e,), generator=g) + rank * n_local
ei = torch.stack([src, dst]).to(dev)
_, halo_global, need = kd.relabel_edges_first_use(ei, rank, world, n_local)
ids = halo_global.to(torch.int32)
n_halo = ids.numel()
out = {}
for width in (128, 64):
    buf = symm_mem.empty((n_local, width), dtype=torch.float32, device=dev)
    buf.normal_()
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    table = torch.tensor([int(q) for q in hdl.buffer_ptrs], dtype=torch.int64, device=dev)
    halo = torch.empty(n_halo, width, device=dev)
    flags = torch.zeros((n_halo + 255) // 256, dtype=torch.int32, device=dev)
    hdl.barrier()
    def timeit(fn, reps=10):
        ts = []
        for i in range(3 + reps):
            torch.cuda._sleep(200000)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            if i >= 3: ts.append(a.elapsed_time(b))
        return statistics.median(ts)
    t_all = timeit(lambda: ops.gather_rows_peer(table, buf.stride(0), n_local, ids, width, out=halo))
    res = {"width": width, "halo_rows": n_halo, "MB": n_halo * width * 4 / 1e6, "all_sms_ms": t_all, "all_sms_GBs": n_halo * width * 4 / t_all / 1e6}
    ep = [0]
    for ctas in (8, 16, 24, 32, 64):
        def run():
            ep[0] += 1
            ops.gather_rows_peer_ordered(table, buf.stride(0), n_local, ids, width, halo, flags, ep[0], ctas)
        t = timeit(run)
        res[f"ordered_{ctas}_ms"] = t
        res[f"ordered_{ctas}_GBs"] = n_halo * width * 4 / t / 1e6
    hdl.barrier()
    if rank == 0:
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
dist.destroy_process_group()
