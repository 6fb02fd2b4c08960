"""The three transports of the node-sharded forward (kagnn_b200/dist.py: NCCL halo matrix, NVLink pull, in-kernel NVLink gather)
exercised on ONE GPU: the ranks are simulated inside one process -- every "rank" owns a separately allocated block of rows, the
peer table (KagnnAggregate.peer_x) holds the base pointers of all blocks, and the halo matrix is filled by the same pull kernel
(kagnn_gather_rows_peer) the multi-GPU run uses.  The kernels cannot tell a peer-mapped pointer from a local one, so this checks
exactly the device code of `mode="peer"` / `mode="pull"` / `mode="halo"` (x_halo) on the single-GPU test box; the process-group
side (symmetric memory rendezvous, barriers, all-to-all) is covered by tests/test_gpu_dist.py (>= 2 GPUs) and tests/test_dist_gloo.py.
Reference: the single-GPU forward of the same layer and the oracle (PyG GINConv / GCNConv semantics, nc/models.py:31-56)."""
import pytest
import torch

from oracle import kagnn_oracle as K

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _setup(world, n_local, f, e, seed):
    g = torch.Generator().manual_seed(seed)
    n = world * n_local
    x = torch.randn(n, f, generator=g) * 0.5
    ei = torch.randint(0, n, (2, e), generator=g)
    return n, x, ei


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("world,f", [(2, 64), (4, 128), (3, 48)])
def test_gin_layer_peer_and_pull_transports_single_gpu(world, f, fast):
    import kagnn_b200 as kb
    from kagnn_b200 import dist as kd
    from kagnn_b200 import ops
    from kagnn_b200.graph import GraphCSR
    torch.manual_seed(1)
    n_local = 1111                                                # not a multiple of the 128-row tile
    n, x, ei = _setup(world, n_local, f, 9 * world * n_local, seed=world * 100 + f)
    conv = (kb.GIFASTKANLayer(f, 32, 5, 32, 2) if fast else kb.GIKANLayer(f, 32, 5, 3, 32, 2))
    sd = {k: v.detach().clone() for k, v in conv.state_dict().items()}
    conv = conv.cuda()
    dev = torch.device("cuda")
    with torch.no_grad():
        y_single = conv(x.to(dev), ei.to(dev)).cpu()
    chain = (lambda t: K.fastkan_chain(sd, "nn.layers.", t)) if fast else (lambda t: K.kan_chain(sd, "nn.layers.", t))
    ref = K.gin_conv(x, ei, chain)
    assert K.rel_err(y_single, ref) <= TOL
    blocks = [x[r * n_local:(r + 1) * n_local].to(dev).clone() for r in range(world)]       # one allocation per "rank"
    table = torch.tensor([b.data_ptr() for b in blocks], dtype=torch.int64, device=dev)
    y_peer, y_pull = [], []
    with torch.no_grad():
        for r in range(world):
            lo = r * n_local
            mine = (ei[1] >= lo) & (ei[1] < lo + n_local)
            ei_r = ei[:, mine].to(dev)
            # in-kernel gather: CSR over GLOBAL source ids, rows read through the peer table
            g_peer = GraphCSR(torch.stack([ei_r[0], ei_r[1] - lo]), n_local, n)
            y_peer.append(conv(blocks[r], g_peer, peer_x=table, rows_per_rank=n_local).cpu())
            # pull: distinct remote rows copied by the pull kernel into a halo matrix, CSR over the local + halo numbering
            ei_local, halo_global, _ = kd.relabel_edges(ei_r, r, world, n_local)
            halo = ops.gather_rows_peer(table, blocks[r].stride(0), n_local, halo_global.to(torch.int32), f)
            assert torch.equal(halo.cpu(), x[halo_global.cpu()])
            g_pull = GraphCSR(ei_local, n_local, n_local + int(halo_global.numel()))
            y_pull.append(conv(blocks[r], g_pull, x_halo=halo).cpu())
    for name, ys in (("peer", y_peer), ("pull", y_pull)):
        y = torch.cat(ys)
        assert K.rel_err(y, y_single) <= 2e-6, name              # same kernels; only the order of equal-valued row sources may differ
        assert K.rel_err(y, ref) <= TOL, name


def test_gcn_layer_halo_matrix_single_gpu():
    """GCN flavour over a halo matrix (the transport `auto` keeps for GCN / non-skip models): h = KAN(x) of the remote sources
    arrives in x_halo, gcn_norm weights come from the degrees of the whole graph."""
    import kagnn_b200 as kb
    from kagnn_b200 import _lib as L
    from kagnn_b200 import dist as kd
    from kagnn_b200 import ops
    from kagnn_b200.graph import GraphCSR
    torch.manual_seed(2)
    world, n_local, f = 2, 1500, 64
    n, x, ei = _setup(world, n_local, f, 14_000, seed=5)
    conv = kb.KAGCNConv(f, 32, 5, 3)
    with torch.no_grad():
        conv.bias.normal_(0, 0.1)
    sd = {k: v.detach().clone() for k, v in conv.state_dict().items()}
    conv = conv.cuda()
    dev = torch.device("cuda")
    ref = K.gcn_conv(x, ei, lambda t: K._kan_layer_from_sd(sd, "lin.", t), sd["bias"])
    with torch.no_grad():
        h = conv.transform(x.to(dev))                              # KAN(x) of every node (each rank computes its own rows)
        full = GraphCSR(ei.to(dev), n)
        _, dinv_all = ops.gcn_degree(full.csr)
        ys = []
        for r in range(world):
            lo = r * n_local
            mine = (ei[1] >= lo) & (ei[1] < lo + n_local)
            ei_local, halo_global, _ = kd.relabel_edges(ei[:, mine].to(dev), r, world, n_local)
            g = GraphCSR(ei_local, n_local, n_local + int(halo_global.numel()))
            sw, dinv = ops.gcn_degree(g.csr)
            assert torch.allclose(dinv, dinv_all[lo:lo + n_local])
            w = ops.gcn_edge_weight(g.csr, torch.cat([dinv, dinv_all[halo_global]]), dinv)
            agg = ops.AggSpec(L.AGG_WEIGHTED, h[lo:lo + n_local], g.rowptr, g.col, edge_weight=w, self_weight=sw,
                              x_halo=h[halo_global].contiguous())
            ys.append(ops.fused_layer(agg, n_local, [], pre=ops.Affine(shift=conv.bias.detach())).cpu())
    assert K.rel_err(torch.cat(ys), ref) <= TOL


@pytest.mark.parametrize("fast", [False, True])
def test_overlapped_pull_with_progress_flags_single_gpu(fast):
    """mode="pull" as the multi-GPU run executes it: halo rows numbered in first-use order, pulled by the persistent copy kernel
    on a SECOND stream while the fused layer runs on a reduced grid and waits on the per-chunk flags (KagnnAggregate.halo_flags)."""
    import kagnn_b200 as kb
    from kagnn_b200 import dist as kd
    from kagnn_b200 import ops
    from kagnn_b200.graph import GraphCSR
    torch.manual_seed(3)
    world, n_local, f = 2, 20_000, 64
    n, x, ei = _setup(world, n_local, f, 10 * world * n_local, seed=77)
    conv = (kb.GIFASTKANLayer(f, 32, 5, 32, 2) if fast else kb.GIKANLayer(f, 32, 5, 3, 32, 2)).cuda()
    dev = torch.device("cuda")
    with torch.no_grad():
        y_single = conv(x.to(dev), ei.to(dev)).cpu()
    blocks = [x[r * n_local:(r + 1) * n_local].to(dev).clone() for r in range(world)]
    table = torch.tensor([b.data_ptr() for b in blocks], dtype=torch.int64, device=dev)
    side = torch.cuda.Stream()
    ys = []
    with torch.no_grad():
        for r in range(world):
            lo = r * n_local
            mine = (ei[1] >= lo) & (ei[1] < lo + n_local)
            ei_local, halo_global, need = kd.relabel_edges_first_use(ei[:, mine].to(dev), r, world, n_local)
            n_halo = int(halo_global.numel())
            assert n_halo > 10 * ops.HALO_CHUNK                       # several chunks, several tiles per chunk
            g = GraphCSR(ei_local, n_local, n_local + n_halo)
            flags = torch.zeros((n_halo + ops.HALO_CHUNK - 1) // ops.HALO_CHUNK, dtype=torch.int32, device=dev)
            halo = torch.full((n_halo, f), float("nan"), device=dev)     # a row read before it landed would poison the result
            for epoch in (1, 2):                                         # the flags are reused with a new epoch, never reset
                halo.fill_(float("nan"))
                main = torch.cuda.current_stream()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    torch.cuda._sleep(2_000_000)                         # the pull starts ~1 ms late: the layer really has to wait
                    ops.gather_rows_peer_ordered(table, blocks[r].stride(0), n_local, halo_global.to(torch.int32), f, halo, flags,
                                                 epoch, 8)
                y = conv(blocks[r], g, x_halo=halo, halo_need=need, halo_flags=flags, halo_epoch=epoch, reserve_sms=8)
                main.wait_stream(side)
            ys.append(y.cpu())
    y = torch.cat(ys)
    assert torch.isfinite(y).all()
    assert K.rel_err(y, y_single) <= 2e-6


@pytest.mark.parametrize("fast,hid,masked", [(False, 64, False), (False, 64, True), (True, 32, True), (False, 128, True)])
def test_pushed_layer_output_single_gpu(fast, hid, masked):
    """mode="push": the layer that produces a hidden matrix also copies every finished tile into the other ranks' replicas
    (KagnnAggregate.push_y, one relay warp per CTA, bulk copies), optionally restricted by the per-row byte mask; the next layer
    then aggregates over [own rows | replica].  Three "ranks" on one GPU, every replica a separate allocation."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    from kagnn_b200.graph import GraphCSR
    torch.manual_seed(4)
    world, n_local, f = 3, 5_003, 64                              # many tiles per CTA would need > 148 * 128 rows: see the second size
    if hid == 128:
        n_local = 40_001                                         # 313 tiles: every CTA relays several tiles
    n, x, ei = _setup(world, n_local, f, 8 * world * n_local, seed=11 + hid)
    mk = (lambda a, b: kb.GIFASTKANLayer(a, hid, 5, b, 2)) if fast else (lambda a, b: kb.GIKANLayer(a, hid, 5, 3, b, 2))
    conv0, conv1 = mk(f, hid).cuda(), mk(hid, 24).cuda()
    dev = torch.device("cuda")
    xd, eid = x.to(dev), ei.to(dev)
    with torch.no_grad():
        h_single = conv0(xd, eid)
        y_single = conv1(h_single, eid).cpu()
    # per "rank": replica of the hidden matrix (all ranks' rows), poisoned so that a missing push shows
    reps = [torch.full((n, hid), float("nan"), device=dev) for _ in range(world)]
    h_local = [torch.empty(n_local, hid, device=dev) for _ in range(world)]
    need = torch.zeros(world, n, dtype=torch.bool)               # need[q, j]: rank q references global row j
    for q in range(world):
        mine = (ei[1] >= q * n_local) & (ei[1] < (q + 1) * n_local)
        need[q, ei[0][mine]] = True
    with torch.no_grad():
        for r in range(world):
            lo = r * n_local
            mine = (ei[1] >= lo) & (ei[1] < lo + n_local)
            peers = [q for q in range(world) if q != r]
            table = torch.tensor([reps[q][lo:].data_ptr() for q in peers], dtype=torch.int64, device=dev)
            mask = None
            if masked:
                mask = torch.zeros(n_local, dtype=torch.uint8)
                for i, q in enumerate(peers):
                    mask |= need[q, lo:lo + n_local].to(torch.uint8) << i
                mask = mask.to(dev)
            # x of the whole graph as [own rows | everything]: source ids >= n_local address the second matrix
            conv0(xd[lo:lo + n_local], _rep_graph(ei[:, mine], lo, n_local, n, dev), out=h_local[r], x_halo=xd, push_y=table, push_ld=hid,
                  push_mask=mask)
        torch.cuda.synchronize()
        ys = []
        for r in range(world):
            lo = r * n_local
            mine = (ei[1] >= lo) & (ei[1] < lo + n_local)
            assert torch.equal(h_local[r], h_single[lo:lo + n_local]) or K.rel_err(h_local[r].cpu(), h_single[lo:lo + n_local].cpu()) <= 2e-6
            got = reps[r].clone()
            for q in range(world):                                   # what rank r received from rank q
                if q == r:
                    continue
                rows = torch.arange(q * n_local, (q + 1) * n_local)
                sent = need[r, rows] if masked else torch.ones(n_local, dtype=torch.bool)
                assert torch.equal(got[rows[sent].to(dev)], h_local[q][sent.to(dev)]), (r, q)
                assert torch.isnan(got[rows[~sent].to(dev)]).all()   # rows nobody asked for were not sent
            ys.append(conv1(h_local[r], _rep_graph(ei[:, mine], lo, n_local, n, dev), x_halo=reps[r]).cpu())
    y = torch.cat(ys)
    assert torch.isfinite(y).all()
    assert K.rel_err(y, y_single) <= 2e-6


def test_masked_pull_fills_exactly_the_marked_rows_single_gpu():
    """kagnn_gather_rows_peer_masked (the input-halo transfer of mode="push"): rows marked in the byte map are copied from their
    owner's block into the replica at their own index, everything else is left alone."""
    from kagnn_b200 import ops
    world, n_local, f = 4, 3_001, 128
    g = torch.Generator().manual_seed(21)
    n = world * n_local
    x = torch.randn(n, f + 32, generator=g)[:, 16:16 + f]                  # a column slice: leading dimension != width
    dev = torch.device("cuda")
    full = x.to(dev)
    blocks = [full[r * n_local:(r + 1) * n_local] for r in range(world)]    # views with the parent's leading dimension
    blocks = [torch.empty(n_local, f + 8, device=dev)[:, :f].copy_(b) for b in blocks]   # separate allocations, ld = f + 8
    table = torch.tensor([b.data_ptr() for b in blocks], dtype=torch.int64, device=dev)
    need = (torch.rand(n, generator=g) < 0.58).to(torch.uint8)
    need[n_local:2 * n_local] = 0                                           # "my" range is never pulled
    rep = torch.full((n, f), float("nan"), device=dev)
    ops.gather_rows_peer_masked(table, blocks[0].stride(0), n_local, need.to(dev), f, rep)
    rep = rep.cpu()
    on = need.bool()
    assert torch.equal(rep[on], x[on])
    assert torch.isnan(rep[~on]).all()


def _rep_graph(ei_mine, lo, n_local, n, dev):
    """CSR of one rank's destination rows over [own rows | replica of all rows]: local source j -> j - lo, remote -> n_local + j."""
    from kagnn_b200.graph import GraphCSR
    src, dst = ei_mine[0], ei_mine[1] - lo
    local = (src >= lo) & (src < lo + n_local)
    src = torch.where(local, src - lo, src + n_local)
    return GraphCSR(torch.stack([src, dst]).to(dev), n_local, n_local + n)
