"""Multi-GPU parity worker (launched by torchrun from tests/test_gpu_dist.py or by hand under ``gpurun --gpus N``):
the node-sharded forward of kagnn_b200.dist on N ranks, gathered, must equal (a) the single-GPU forward of the same
model on the whole graph and (b) the CPU oracle, for GIN and GCN flavours.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_parity_worker.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import kagnn_b200 as kb
    from kagnn_b200 import dist as kd
    from oracle import kagnn_oracle as K

    n_local, e_local, f = 1500, 9000, 48
    n = n_local * world
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, f, generator=g) * 0.5
    src = torch.randint(0, n, (e_local * world,), generator=g)
    dst = torch.randint(0, n, (e_local * world,), generator=g)
    ei = torch.stack([src, dst])
    mine = (dst >= rank * n_local) & (dst < (rank + 1) * n_local)
    worst = 0.0
    for conv_type, fast in (("gin", False), ("gcn", False), ("gin", True), ("gcn", True)):
        torch.manual_seed(11)
        if fast:
            m = kb.GFASTKAN_Nodes(conv_type, 2, f, 32, 5, skip=True, grid_size=5, hidden_layers=2).eval()
        else:
            m = kb.GKAN_Nodes(conv_type, 2, f, 32, 5, skip=True, grid_size=5, spline_order=3, hidden_layers=2).eval()
        with torch.no_grad():
            for name, b in m.named_buffers():
                if name.endswith("running_var"):
                    b.uniform_(0.5, 1.5, generator=g)
                if name.endswith("running_mean"):
                    b.normal_(0, 0.1, generator=g)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        m = m.to(dev)
        runner = kd.ShardedNodeModel(m, rank, world, n_local)
        auto = kd.ShardedNodeModel(m, rank, world, n_local, mode="auto")
        assert auto.mode in (("peer", "pull", "push", "halo") if conv_type == "gin" else ("halo",)), (conv_type, fast, auto.mode)
        plan = runner.prepare(ei[:, mine].to(dev))
        y_local = runner.forward(x[rank * n_local:(rank + 1) * n_local].to(dev), plan)
        ys = [torch.empty_like(y_local) for _ in range(world)]
        dist.all_gather(ys, y_local)
        y_sharded = torch.cat(ys).cpu()
        with torch.no_grad():
            y_single = m(x.to(dev), ei.to(dev)).cpu()
        if fast:
            y_ref = K.node_model_forward(sd, conv_type, x, ei, True)
        else:
            y_ref = K.node_model_forward(sd, conv_type, x, ei, True)
        e1, e2 = K.rel_err(y_sharded, y_single), K.rel_err(y_sharded, y_ref)
        worst = max(worst, e1, e2)
        if rank == 0:
            print(f"{'fastkan' if fast else 'kan'}/{conv_type}: sharded vs single {e1:.2e}, sharded vs oracle {e2:.2e}, "
                  f"halo rows {plan.n_halo}", flush=True)
        assert e1 <= 1e-5, (conv_type, fast, e1)     # same kernels, same reduction order per row
        assert e2 <= 1e-4, (conv_type, fast, e2)
        if conv_type == "gin":
            # in-kernel NVLink gather (KagnnAggregate.peer_x) and the overlapped pull: same numbers without pack / all-to-all
            # (B-spline and FastKAN flavours; GCN flavours stay on the NCCL halo transport, which `auto` picks for them)
            for pmode in ("peer", "pull", "pull_overlap", "push"):
                peer = kd.ShardedNodeModel(m, rank, world, n_local, mode=pmode)
                pplan = peer.prepare(ei[:, mine].to(dev))
                for rep in range(3):                  # repeated steps exercise the cross-step buffer reuse barriers
                    y_peer = peer.forward(x[rank * n_local:(rank + 1) * n_local].to(dev), pplan)
                ys = [torch.empty_like(y_peer) for _ in range(world)]
                dist.all_gather(ys, y_peer)
                y_peer_all = torch.cat(ys).cpu()
                e3, e4 = K.rel_err(y_peer_all, y_single), K.rel_err(y_peer_all, y_ref)
                worst = max(worst, e3, e4)
                if rank == 0:
                    print(f"{'fastkan' if fast else 'kan'}/gin {pmode}: vs single {e3:.2e}, vs oracle {e4:.2e}", flush=True)
                assert e3 <= 1e-5 and e4 <= 1e-4, (pmode, e3, e4)
            # x written straight into the symmetric buffer (no copy, one barrier fewer) and the input halo kept between steps
            res = kd.ShardedNodeModel(m, rank, world, n_local, mode="push", resident_x_halo=True)
            rplan = res.prepare(ei[:, mine].to(dev))
            xin = res.input_buffer(x.size(1), dev)
            xin.copy_(x[rank * n_local:(rank + 1) * n_local])
            for rep in range(3):
                y_res = res.forward(xin, rplan)
            ys = [torch.empty_like(y_res) for _ in range(world)]
            dist.all_gather(ys, y_res)
            e5 = K.rel_err(torch.cat(ys).cpu(), y_single)
            assert e5 <= 1e-5, ("push, resident input", e5)
            xin.mul_(0.5)                                 # a new version of x: the kept halo must be refreshed
            y_half = res.forward(xin, rplan)
            ys = [torch.empty_like(y_half) for _ in range(world)]
            dist.all_gather(ys, y_half)
            with torch.no_grad():
                y_half_single = m(x.to(dev) * 0.5, ei.to(dev)).cpu()
            e6 = K.rel_err(torch.cat(ys).cpu(), y_half_single)
            if rank == 0:
                print(f"{'fastkan' if fast else 'kan'}/gin push resident: {e5:.2e}, after x changed {e6:.2e}", flush=True)
            assert e6 <= 1e-5, ("push, resident input, new x", e6)
    # large grids (slot windows, kagnn_b200/ekan.py: _windowed_spec): every layer is its own launch behind the halo exchange
    for conv_type, fast in (("gin", False), ("gcn", False), ("gin", True)):
        torch.manual_seed(13)
        if fast:
            m = kb.GFASTKAN_Nodes(conv_type, 2, f, 32, 5, skip=True, grid_size=12, hidden_layers=2).eval()
        else:
            m = kb.GKAN_Nodes(conv_type, 2, f, 32, 5, skip=True, grid_size=8, spline_order=3, hidden_layers=2).eval()
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        m = m.to(dev)
        big = kd.ShardedNodeModel(m, rank, world, n_local, mode="auto")
        assert big.mode == "halo", big.mode
        bplan = big.prepare(ei[:, mine].to(dev))
        y_big = big.forward(x[rank * n_local:(rank + 1) * n_local].to(dev), bplan)
        ys = [torch.empty_like(y_big) for _ in range(world)]
        dist.all_gather(ys, y_big)
        e7 = K.rel_err(torch.cat(ys).cpu(), K.node_model_forward(sd, conv_type, x, ei, True))
        worst = max(worst, e7)
        if rank == 0:
            print(f"{'fastkan grid 12' if fast else 'kan grid 8 order 3'}/{conv_type} (windows, halo): sharded vs oracle {e7:.2e}", flush=True)
        assert e7 <= 1e-4, (conv_type, fast, e7)
    if rank == 0:
        print(f"DIST_PARITY_OK world={world} worst={worst:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
