"""Host-side logic of the node-range sharding (kagnn_b200/dist.py) on CPU with the gloo backend, world_size 2 and 3:
partition book, halo lists, the index all-to-all and the per-layer row exchange.  The arithmetic of the product runs
only on the GPU, so the row gather that fills the send buffer is injected (``pack=``) and the aggregation is done by
the oracle: what is checked here is that [owned rows ; exchanged halo rows] + the relabelled edge list reproduce,
shard by shard, the aggregation over the global graph."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kagnn_oracle as K


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_problem(world, n_local, e_per_rank, f, seed=0):
    g = torch.Generator().manual_seed(seed)
    n = world * n_local
    x = torch.randn(n, f, generator=g)
    eis = []
    for r in range(world):
        src = torch.randint(0, n, (e_per_rank,), generator=g)
        dst = torch.randint(r * n_local, (r + 1) * n_local, (e_per_rank,), generator=g)
        eis.append(torch.stack([src, dst]))
    return x, eis


def _worker(rank, world, port, n_local, e_per_rank, f, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kagnn_b200 import dist as kd
        x, eis = _global_problem(world, n_local, e_per_rank, f)
        ei = eis[rank]
        plan = kd.build_halo_plan(ei, rank, world, n_local, build_csr=False)
        lo = rank * n_local
        # halo set == the distinct remote sources (CPU set computation)
        remote = sorted({int(s) for s in ei[0].tolist() if not (lo <= s < lo + n_local)})
        assert plan.halo_global.tolist() == remote
        assert plan.n_halo == len(remote) and sum(plan.recv_splits) == len(remote)
        assert plan.recv_splits[rank] == 0 and plan.send_splits[rank] == 0
        owners = [s // n_local for s in remote]
        assert plan.recv_splits == [owners.count(p) for p in range(world)]
        # relabelled edges address [owned ; halo]
        ext_ids = torch.cat([torch.arange(lo, lo + n_local), plan.halo_global])
        assert torch.equal(ext_ids[plan.edge_index_local[0]], ei[0])
        assert torch.equal(plan.edge_index_local[1] + lo, ei[1])
        # the exchange delivers exactly the halo rows (twice: the plan is reusable, widths may differ per layer)
        xchg = kd.HaloExchange(plan, pack=lambda t, idx: t.index_select(0, idx.long()))
        x_local = x[lo:lo + n_local]
        for width in (f, 1):
            halo = xchg(x_local[:, :width].contiguous())
            assert torch.equal(halo, x[plan.halo_global][:, :width])
        # aggregation over [owned ; halo] with the local edge list == the global aggregation, for the owned rows
        x_ext = torch.cat([x_local, xchg(x_local)])
        agg_local = torch.zeros(n_local, f).index_add_(0, plan.edge_index_local[1], x_ext[plan.edge_index_local[0]])
        ei_all = torch.cat(eis, dim=1)
        agg_global = torch.zeros(world * n_local, f).index_add_(0, ei_all[1], x[ei_all[0]])
        assert torch.allclose(agg_local, agg_global[lo:lo + n_local], atol=1e-5)
        # a full GIN conv through the oracle on the shard == the oracle on the global graph
        ident = lambda t: t  # noqa: E731
        full = K.gin_conv(x, ei_all, ident)[lo:lo + n_local]
        shard = K.gin_conv(x_ext, plan.edge_index_local, ident)[:n_local]
        assert torch.allclose(shard, full, atol=1e-5)
        assert xchg.bytes_sent == sum(plan.send_splits) * 4 * (2 * f + 1)
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover - reported to the parent
        import traceback
        out.put((rank, f"{type(exc).__name__}: {exc}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plan_and_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 40, 300, 5, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_relabel_rejects_foreign_targets():
    from kagnn_b200 import dist as kd
    ei = torch.tensor([[0, 5], [0, 9]])
    with pytest.raises(IndexError):
        kd.relabel_edges(ei, 0, 2, 5)          # target 9 belongs to rank 1
    with pytest.raises(IndexError):
        kd.relabel_edges(torch.tensor([[10], [0]]), 0, 2, 5)   # source outside the global range


def test_relabel_without_remote_edges():
    from kagnn_b200 import dist as kd
    ei = torch.tensor([[1, 2, 3], [0, 0, 4]])
    loc, halo, counts = kd.relabel_edges(ei, 0, 2, 5)
    assert halo.numel() == 0 and counts.tolist() == [0, 0]
    assert torch.equal(loc, ei)


def test_shard_batch_by_graph():
    from kagnn_b200 import dist as kd
    batch = torch.tensor([0, 0, 1, 2, 2, 2, 3, 4, 4])
    seen = torch.zeros_like(batch, dtype=torch.bool)
    for r in range(2):
        g0, g1, mask = kd.shard_batch_by_graph(batch, 5, r, 2)
        assert not (seen & mask).any()
        seen |= mask
        assert set(batch[mask].tolist()) == set(range(g0, g1))
    assert seen.all()


# ---- the whole sharded forward (ShardedNodeModel, mode="halo") over gloo, library launches replaced by the CPU stand-ins --------
def _model_worker(rank, world, port, conv_type, out, grid_size=5):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import kagnn_b200 as kb
        from kagnn_b200 import dist as kd
        from tests.emul.cpu_double import cpu_double
        n_local, f, c = 30, 8, 3
        x, eis = _global_problem(world, n_local, 120, f, seed=3)
        ei_all = torch.cat(eis, dim=1)
        torch.manual_seed(11)                                         # same replicated weights on every rank
        model = kb.GKAN_Nodes(conv_type, 2, f, 12, c, skip=True, grid_size=grid_size, spline_order=3, hidden_layers=2).eval()
        with torch.no_grad():
            for bn in model.bns:                                      # non-trivial eval BatchNorm
                bn.running_mean.uniform_(-0.3, 0.3)
                bn.running_var.uniform_(0.5, 1.5)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        lo = rank * n_local
        with cpu_double(windows=grid_size > 5), torch.no_grad():      # grid 8 + order 3: the slot-window wiring behind the halo exchange
            runner = kd.ShardedNodeModel(model, rank, world, n_local, mode="halo")
            if grid_size > 5:
                assert max(lay.kernel_spec().windows for conv in model.convs for lay in
                           (conv.nn.layers if hasattr(conv, "nn") else [conv.lin])) == 2
            plan = runner.prepare(eis[rank])
            y_shard = runner.forward(x[lo:lo + n_local].contiguous(), plan)
            y_shard2 = runner.forward(x[lo:lo + n_local].contiguous(), plan)          # the plan is reusable
            y_single = model(x, ei_all)                                               # the un-sharded plan, same stand-ins
        y_ref = K.node_model_forward(sd, conv_type, x, ei_all, True)
        assert torch.equal(y_shard, y_shard2)
        assert K.rel_err(y_shard, y_ref[lo:lo + n_local]) <= 1e-5
        assert K.rel_err(y_single, y_ref) <= 1e-5
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover - reported to the parent
        import traceback
        out.put((rank, f"{type(exc).__name__}: {exc}\\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("conv_type,grid_size", [("gin", 5), ("gcn", 5), ("gin", 8), ("gcn", 8)])
def test_sharded_model_forward_gloo(conv_type, grid_size):
    """world_size 2: every rank's rows of the sharded forward == the oracle on the global graph (grid 8: layers of eleven
    coefficients per pair, evaluated as two slot windows)."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_model_worker, args=(r, world, port, conv_type, out, grid_size)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


# ---- plan of the "push" transport (kagnn_b200/dist.py: push_plan_arrays / push_row_masks) --------------------------------------
def _push_worker(rank, world, port, n_local, e_per_rank, f, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kagnn_b200 import dist as kd
        x, eis = _global_problem(world, n_local, e_per_rank, f, seed=3)
        ei, lo, n = eis[rank], rank * n_local, world * n_local
        src_rep, dst, need = kd.push_plan_arrays(ei, rank, world, n_local)
        # the byte map marks exactly the distinct remote sources
        remote = sorted({int(s) for s in ei[0].tolist() if not (lo <= s < lo + n_local)})
        assert need.nonzero().view(-1).tolist() == remote
        # [own rows | replica of everything] + the renumbered edge list reproduce the aggregation over the global graph
        x_ext = torch.cat([x[lo:lo + n_local], x])
        agg_local = torch.zeros(n_local, f).index_add_(0, dst, x_ext[src_rep])
        ei_all = torch.cat(eis, dim=1)
        agg_global = torch.zeros(n, f).index_add_(0, ei_all[1], x[ei_all[0]])
        assert torch.allclose(agg_local, agg_global[lo:lo + n_local], atol=1e-5)
        # masks: bit i of byte r <=> peer i references my row r (checked against every rank's edge list)
        mask = kd.push_row_masks(need, rank, world, n_local)
        expect = torch.zeros(n_local, dtype=torch.uint8)
        for i, q in enumerate(kd.push_peers(rank, world)):
            refs = eis[q][0]
            mine = refs[(refs >= lo) & (refs < lo + n_local)] - lo
            expect[mine.unique()] |= (1 << i)
        assert torch.equal(mask, expect)
        # simulated transfer: what the peers push + what I pull of x must cover every row my edge list reads
        have = torch.zeros(n, dtype=torch.bool)
        have[need.bool()] = True                                   # masked pull of the input halo fills exactly these rows
        assert have[ei[0][(ei[0] < lo) | (ei[0] >= lo + n_local)]].all()
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover - reported to the parent
        import traceback
        out.put((rank, f"{type(exc).__name__}: {exc}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_push_plan_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_push_worker, args=(r, world, port, 40, 300, 5, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_push_plan_flags_out_of_range_sources():
    from kagnn_b200 import dist as kd
    ei = torch.tensor([[0, 7, 12, -1], [0, 1, 2, 3]])
    src_rep, dst, need = kd.push_plan_arrays(ei, 0, 2, 5)
    assert src_rep.tolist() == [0, 5 + 7, -1, -1]                  # ids outside [0, 10) stay out of range for the CSR build's check
    assert need.tolist() == [0, 0, 0, 0, 0, 0, 0, 1, 0, 1]         # (clamped ids of the bad entries may mark a row: harmless extra traffic)
