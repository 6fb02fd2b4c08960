"""GPU parity tests (run on the B200 box with ``-m gpu``): the sm_100a path, called through the C ABI by the
kagnn_b200 modules, against (a) the golden vectors computed by the reference itself and (b) the CPU oracle on
seeded random inputs.  Tolerance: BASELINE.json's north_star asks for 1e-4 relative in fp32; the metric is
max|y - y_ref| / max|y_ref| per tensor (oracle.kagnn_oracle.rel_err)."""
import pytest
import torch

from oracle import kagnn_oracle as K
from tests.helpers import build_product_model, golden_names, load_golden, product_run

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _sd_cpu(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


@pytest.fixture(params=["auto", "tc_general", "fp32"])
def path_mode(request):
    """auto = pipelined tcgen05 kernel wherever it applies, else the general tcgen05 kernel, else fp32;
    tc_general = skip the pipelined kernel; fp32 = force the CUDA-core kernel (all are GPU paths)."""
    from kagnn_b200 import ops, _lib as L
    ops.set_path(L.PATH_FP32 if request.param == "fp32" else L.PATH_AUTO)
    ops.set_tc_variant(1 if request.param == "tc_general" else 0)
    yield request.param
    ops.set_tc_variant(0)
    ops.set_path(L.PATH_AUTO)


@pytest.mark.parametrize("name", golden_names())
def test_golden_vectors(name, path_mode):
    meta, inputs, sd, y_ref = load_golden(name)
    model = build_product_model(meta, sd)
    y = product_run(meta, inputs, model).cpu()
    assert y.shape == y_ref.shape
    assert torch.isfinite(y).all()
    assert K.rel_err(y, y_ref) <= TOL, name


@pytest.mark.parametrize("G,k,fin,fout,n", [
    (5, 3, 128, 64, 1000), (5, 3, 64, 64, 4097), (4, 3, 7, 1, 130), (5, 3, 1433, 32, 300), (5, 3, 1497, 7, 300),
    (5, 3, 320, 40, 513), (8, 1, 33, 17, 64), (3, 2, 5, 200, 77), (32, 4, 9, 256, 200), (1, 1, 2, 2, 1), (5, 3, 128, 128, 2048),
])
def test_kan_linear_random(G, k, fin, fout, n):
    import kagnn_b200 as kb
    torch.manual_seed(G * 1000 + fin)
    m = kb.KANLinear(fin, fout, grid_size=G, spline_order=k)
    x = torch.randn(n, fin) * 0.9           # ~2.5 % beyond the knot range for G=5,k=3
    y_ref = K._kan_layer_from_sd(_sd_cpu(m), "", x)
    with torch.no_grad():
        y = m.cuda()(x.cuda()).cpu()
    assert K.rel_err(y, y_ref) <= TOL


def test_kan_linear_special_values():
    import kagnn_b200 as kb
    torch.manual_seed(3)
    m = kb.KANLinear(6, 5, grid_size=5, spline_order=3)
    knots = m.grid[0].clone()
    x = torch.zeros(8, 6)
    x[0] = knots[:6]
    x[1] = knots[6:12]
    x[2] = 1e4
    x[3] = -1e4
    x[4] = torch.tensor([0.0, -0.0, 1e-30, -1e-30, 2.2, -2.2])
    x[5] = float("inf")
    x[6] = float("nan")
    x[7] = torch.nextafter(knots[-1], torch.tensor(0.0))
    y_ref = K._kan_layer_from_sd(_sd_cpu(m), "", x)
    with torch.no_grad():
        y = m.cuda()(x.cuda()).cpu()
    fin = torch.isfinite(y_ref).all(1)
    assert K.rel_err(y[fin], y_ref[fin]) <= TOL
    assert torch.isnan(y[6]).all()                      # silu(NaN) poisons the row exactly like the reference
    assert torch.equal(torch.isnan(y[5]), torch.isnan(y_ref[5]))


@pytest.mark.parametrize("sizes,G,n", [([256, 256, 256], 8, 700), ([7, 256, 256], 8, 333), ([64, 32, 16, 8, 4], 5, 129),
                                       ([1433, 16], 4, 100), ([20, 300], 3, 70)])
def test_fastkan_random(sizes, G, n):
    import kagnn_b200 as kb
    torch.manual_seed(len(sizes) * 17 + G)
    m = kb.FastKAN(sizes, num_grids=G)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1 and p.requires_grad:
                p.add_(torch.randn_like(p) * 0.1)
    x = torch.randn(n, sizes[0]) * 2.0
    y_ref = K.fastkan_chain(_sd_cpu(m), "layers.", x)
    with torch.no_grad():
        y = m.cuda()(x.cuda()).cpu()
    assert K.rel_err(y, y_ref) <= TOL


def test_kan_chain_deep_and_wide():
    import kagnn_b200 as kb
    torch.manual_seed(11)
    for sizes, G, k in [([128, 64, 64], 5, 3), ([128, 128, 128], 5, 3), ([30, 9, 9, 9, 9, 9, 9, 9, 9, 9, 3], 4, 2)]:
        m = kb.KAN(sizes, grid_size=G, spline_order=k)
        x = torch.randn(777, sizes[0])
        y_ref = K.kan_chain(_sd_cpu(m), "layers.", x)
        with torch.no_grad():
            y = m.cuda()(x.cuda()).cpu()
        assert K.rel_err(y, y_ref) <= TOL, sizes


def _rand_graph(n, e, seed, self_loops=True):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    if self_loops and e > 20:
        ei[1, :10] = ei[0, :10]
        ei[:, 10:15] = ei[:, 15:20]
    return ei


@pytest.mark.parametrize("n,e", [(1, 0), (5, 0), (64, 300), (1000, 20000), (4099, 3)])
def test_csr_build_and_gcn_norm(n, e):
    from kagnn_b200 import ops
    ei = _rand_graph(n, e, n + e)
    csr = ops.csr_build(ei.cuda(), n)
    csr.validate()
    rowptr, col, perm = csr.rowptr.cpu().long(), csr.col.cpu().long(), csr.perm.cpu().long()
    order = torch.sort(ei[1], stable=True).indices
    assert torch.equal(perm, order)                                   # bit-exact, stable
    assert torch.equal(col, ei[0][order])
    counts = torch.bincount(ei[1], minlength=n)
    assert torch.equal(rowptr, torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))
    w, sw, dinv = ops.gcn_norm(csr)
    ei2, w_ref = K.gcn_norm(ei, n)
    # scatter the oracle's weights into a dense matrix and compare with ours
    dense_ref = torch.zeros(n, n, dtype=torch.float64).index_put_((ei2[1], ei2[0]), w_ref.double(), accumulate=True) if n <= 1000 else None
    if dense_ref is not None:
        dst = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
        dense = torch.zeros(n, n, dtype=torch.float64).index_put_((dst, col), w.cpu().double(), accumulate=True)
        dense += torch.diag(sw.cpu().double())
        assert torch.allclose(dense, dense_ref, rtol=1e-5, atol=1e-7)


def test_csr_build_flags_bad_index():
    from kagnn_b200 import ops
    ei = torch.tensor([[0, 1, 7], [1, 2, 0]])
    csr = ops.csr_build(ei.cuda(), 3)
    with pytest.raises(IndexError):
        csr.validate()


@pytest.mark.parametrize("conv_type,fast", [("gin", False), ("gcn", False), ("gin", True), ("gcn", True)])
@pytest.mark.parametrize("n,e,f,h,c", [(1000, 6000, 32, 16, 7), (333, 0, 20, 8, 3), (2708, 10556, 1433, 32, 7)])
def test_node_models_random(conv_type, fast, n, e, f, h, c):
    import kagnn_b200 as kb
    torch.manual_seed(n + f)
    if fast:
        m = kb.GFASTKAN_Nodes(conv_type, 2, f, h, c, grid_size=5, hidden_layers=2)
    else:
        m = kb.GKAN_Nodes(conv_type, 2, f, h, c, grid_size=5, spline_order=3, hidden_layers=2)
    m.eval()
    with torch.no_grad():
        for name, b in m.named_buffers():
            if name.endswith("running_mean"):
                b.normal_(0, 0.3)
            if name.endswith("running_var"):
                b.uniform_(0.5, 1.5)
        for name, p in m.named_parameters():
            if name.endswith("bias") and p.dim() == 1:
                p.normal_(0, 0.2)
    ei = _rand_graph(n, e, n * 3 + e)
    x = torch.randn(n, f) * (0.1 if f > 1000 else 0.7)
    y_ref = K.node_model_forward(_sd_cpu(m), conv_type, x, ei, True)
    with torch.no_grad():
        y = m.cuda()(x.cuda(), ei.cuda()).cpu()
    assert K.rel_err(y, y_ref) <= TOL


def test_standalone_convs_match_oracle():
    import kagnn_b200 as kb
    torch.manual_seed(5)
    n, e, f, h = 500, 3000, 24, 12
    ei = _rand_graph(n, e, 99)
    x = torch.randn(n, f)
    gcn = kb.KAGCNConv(f, h, 5, 3)
    with torch.no_grad():
        gcn.bias.normal_()
    sd = _sd_cpu(gcn)
    ref = K.gcn_conv(x, ei, lambda t: K._kan_layer_from_sd(sd, "lin.", t), sd["bias"])
    with torch.no_grad():
        out = gcn.cuda()(x.cuda(), ei.cuda()).cpu()
    assert K.rel_err(out, ref) <= TOL
    # user edge weights
    ew = torch.rand(e) + 0.1
    ref = K.gcn_conv(x, ei, lambda t: K._kan_layer_from_sd(sd, "lin.", t), sd["bias"], ew)
    with torch.no_grad():
        out = gcn(x.cuda(), ei.cuda(), ew.cuda()).cpu()
    assert K.rel_err(out, ref) <= TOL
    gin = kb.GIKANLayer(f, h, 5, 3, 16, 3)
    sd = _sd_cpu(gin)
    ref = K.gin_conv(x, ei, lambda t: K.kan_chain(sd, "nn.layers.", t))
    with torch.no_grad():
        out = gin.cuda()(x.cuda(), ei.cuda()).cpu()
    assert K.rel_err(out, ref) <= TOL
    gine = kb.GINEConv(kb.make_kan(f, 16, h, 2, 4, 3))
    ea = torch.randn(e, f)
    sd = _sd_cpu(gine)
    ref = K.gine_conv(x, ei, ea, lambda t: K.kan_chain(sd, "nn.layers.", t))
    with torch.no_grad():
        out = gine.cuda()(x.cuda(), ei.cuda(), ea.cuda()).cpu()
    assert K.rel_err(out, ref) <= TOL


def test_gin_wide_input_splits_into_two_launches():
    """Aggregated tile wider than shared memory (Cora-wide GIN): aggregate to HBM, then stream the chain."""
    import kagnn_b200 as kb
    torch.manual_seed(8)
    n, e, f = 300, 2000, 1433
    ei = _rand_graph(n, e, 123)
    x = torch.randn(n, f) * 0.2
    gin = kb.GIKANLayer(f, 8, 5, 3, 8, 2)
    sd = _sd_cpu(gin)
    ref = K.gin_conv(x, ei, lambda t: K.kan_chain(sd, "nn.layers.", t))
    with torch.no_grad():
        out = gin.cuda()(x.cuda(), ei.cuda()).cpu()
    assert K.rel_err(out, ref) <= TOL


def test_pooling_matches_oracle():
    from kagnn_b200 import ops, _lib as L
    torch.manual_seed(2)
    sizes = torch.tensor([3, 1, 0, 40, 7, 0, 2])
    batch = torch.repeat_interleave(torch.arange(len(sizes)), sizes)
    x = torch.randn(int(sizes.sum()), 20)
    ptr = ops.segment_ptr(batch.cuda(), len(sizes))
    assert torch.equal(ptr.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)]))
    for mode, ref in ((L.AGG_SEGMENT_SUM, K.global_add_pool(x, batch, len(sizes))), (L.AGG_SEGMENT_MEAN, K.global_mean_pool(x, batch, len(sizes)))):
        out = ops.fused_layer(ops.AggSpec(mode, x.cuda(), rowptr=ptr), len(sizes), []).cpu()
        assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)


def test_results_are_deterministic():
    import kagnn_b200 as kb
    torch.manual_seed(1)
    m = kb.GKAN_Nodes("gin", 2, 32, 16, 4, grid_size=5, spline_order=3).eval().cuda()
    ei = _rand_graph(3000, 40000, 4).cuda()
    x = torch.randn(3000, 32).cuda()
    with torch.no_grad():
        a = m(x, ei).clone()
        kb.graph.clear_cache()
        b = m(x, ei)
    assert torch.equal(a, b)                           # CSR reduction order is fixed: bitwise reproducible


def test_requires_cuda_and_modules_without_backward_raise_under_autograd():
    import kagnn_b200 as kb
    x = torch.randn(10, 4).cuda()
    ei = torch.randint(0, 10, (2, 20)).cuda()
    gin = kb.GINConv(kb.make_kan(4, 4, 4, 1, 5, 3), train_eps=True).cuda()
    with pytest.raises(NotImplementedError):              # a trainable eps has no backward here: loud, not silent
        gin(x, ei)
    y = kb.KANLinear(4, 4).cuda()(torch.randn(3, 4).cuda())
    assert y.requires_grad                                # KAN layers record themselves for autograd
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            kb.KANLinear(4, 4).cuda()(torch.randn(3, 4))


def test_arxiv_scale_model_against_oracle():
    """BASELINE config 2 at full size: ogbn-arxiv-shaped KAGIN (3 layers, hidden 64, grid 5) vs the CPU oracle."""
    import kagnn_b200 as kb
    torch.manual_seed(12345)
    n, e, f = 169343, 1166243, 128
    m = kb.GKAN_Nodes("gin", 3, f, 64, 40, skip=True, grid_size=5, spline_order=3, hidden_layers=2).eval()
    with torch.no_grad():
        for name, b in m.named_buffers():
            if name.endswith("running_var"):
                b.uniform_(0.01, 0.05)          # keeps hidden activations O(1) like trained BN statistics
    ei = torch.randint(0, n, (2, e))
    x = torch.randn(n, f) * 0.3
    with torch.no_grad():
        y = m.cuda()(x.cuda(), ei.cuda()).cpu()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    y_ref = K.node_model_forward(_sd_cpu(m), "gin", x, ei, True)
    assert K.rel_err(y, y_ref) <= TOL


@pytest.mark.parametrize("rows,cols", [(1, 2), (37, 7), (4096, 40), (5, 300)])
def test_log_softmax_rows(rows, cols):
    from kagnn_b200 import ops
    torch.manual_seed(rows)
    x = torch.randn(rows, cols) * 5
    y = ops.log_softmax(x.cuda()).cpu()
    assert torch.allclose(y, torch.log_softmax(x, dim=1), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("rows,cols,affine", [(2, 3, True), (1000, 64, True), (50_000, 32, False), (777, 130, True)])
def test_batchnorm_training_forward_matches_torch(rows, cols, affine):
    """Training-mode BatchNorm1d (node_classification_clean/models.py:197 under model.train()): output, running statistics
    and num_batches_tracked against torch.nn.BatchNorm1d on the CPU."""
    from kagnn_b200 import ops
    torch.manual_seed(cols)
    x = torch.randn(rows, cols) * 2 + 0.5
    ref = torch.nn.BatchNorm1d(cols, affine=affine)
    if affine:
        with torch.no_grad():
            ref.weight.uniform_(0.5, 1.5)
            ref.bias.normal_()
    mine = torch.nn.BatchNorm1d(cols, affine=affine)
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    ref.train()
    mine.train()
    for _ in range(2):
        with torch.no_grad():
            y_ref = ref(x)
            y = ops.batchnorm_forward(x.cuda(), mine).cpu()
    assert K.rel_err(y, y_ref) <= TOL
    assert torch.allclose(mine.running_mean.cpu(), ref.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(mine.running_var.cpu(), ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(mine.num_batches_tracked) == int(ref.num_batches_tracked) == 2


@pytest.mark.parametrize("head,tail,out_f", [(128, 192, 40), (256, 64, 16), (100, 60, 8), (128, 3, 5)])
def test_two_part_rows_match_concatenation(head, tail, out_f):
    """KagnnAggregate.x_head: lay_out over [x | hidden] without materialising the concat (and the host-side fallback when
    the split is not on a 128-column boundary) == the oracle on torch.cat."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops, _lib as L
    torch.manual_seed(head + tail)
    n = 777
    a, b = torch.randn(n, head) * 0.7, torch.randn(n, tail) * 0.7
    m = kb.KANLinear(head + tail, out_f, grid_size=5, spline_order=3)
    y_ref = K._kan_layer_from_sd(_sd_cpu(m), "", torch.cat([a, b], 1))
    m = m.cuda()
    big = torch.zeros(n, tail + 5).cuda()                 # b as a column slice (leading dimension > width)
    big[:, :tail] = b.cuda()
    with torch.no_grad():
        y = ops.fused_layer(ops.AggSpec(L.AGG_NONE, big[:, :tail], x_head=a.cuda()), n, m.kernel_specs()).cpu()
    assert K.rel_err(y, y_ref) <= TOL


@pytest.mark.parametrize("n,e,f", [(1000, 7000, 128), (333, 900, 64), (50, 0, 32), (4097, 30000, 36), (200, 1500, 30)])
def test_aggregation_only_launches(n, e, f):
    """n_layers == 0: GIN sum, GCN-weighted sum with bias + SiLU, mean pooling -- 128-bit gather kernel for aligned shapes,
    general kernel otherwise (f = 30) -- against plain index arithmetic on the CPU."""
    from kagnn_b200 import ops, _lib as L
    from kagnn_b200.graph import get_graph
    torch.manual_seed(n + f)
    x = torch.randn(n, f)
    ei = torch.randint(0, n, (2, e))
    g = get_graph(ei.cuda(), n)
    xd = x.cuda()
    # GIN: (1 + eps) x_i + sum_j x_j
    y = ops.fused_layer(ops.AggSpec(L.AGG_GIN, xd, g.rowptr, g.col, self_scale=1.25), n, []).cpu()
    ref = 1.25 * x + torch.zeros(n, f).index_add_(0, ei[1], x[ei[0]])
    assert K.rel_err(y, ref) <= 1e-5
    # GCN: normalised sum + bias, SiLU
    w, sw = g.gcn_weights()
    bias = torch.randn(f)
    y = ops.fused_layer(ops.AggSpec(L.AGG_WEIGHTED, xd, g.rowptr, g.col, edge_weight=w, self_weight=sw), n, [],
                        pre=ops.Affine(shift=bias.cuda(), act=L.ACT_SILU)).cpu()
    ref = torch.nn.functional.silu(K.gcn_conv(x, ei, lambda t: t, bias))
    assert K.rel_err(y, ref) <= 1e-5
    # mean pooling over sorted segments (some empty)
    nb = 17
    batch = torch.sort(torch.randint(0, nb, (n,)))[0]
    ptr = ops.segment_ptr(batch.cuda(), nb)
    y = ops.fused_layer(ops.AggSpec(L.AGG_SEGMENT_MEAN, xd, rowptr=ptr), nb, []).cpu()
    ref = torch.zeros(nb, f).index_add_(0, batch, x) / torch.bincount(batch, minlength=nb).clamp(min=1).unsqueeze(1)
    assert K.rel_err(y, ref) <= 1e-5


def test_strided_inputs_are_accepted():
    """Transposed and column-sliced views (which the reference's ATen ops accept) give the same result as a fresh copy."""
    import kagnn_b200 as kb
    torch.manual_seed(9)
    lay = kb.KANLinear(48, 24).cuda()
    conv = kb.GIKANLayer(48, 24, 5, 3, 32, 2).cuda()
    ei = torch.randint(0, 500, (2, 3000)).cuda()
    base = (torch.randn(48, 500) * 0.4).cuda()
    wide = (torch.randn(500, 100) * 0.4).cuda()
    with torch.no_grad():
        xt = base.t()                                  # column stride 500
        assert torch.equal(lay(xt), lay(xt.contiguous()))
        assert torch.equal(conv(xt, ei), conv(xt.contiguous(), ei))
        xs = wide[:, 3:51]                             # row stride 100, misaligned start
        assert torch.equal(lay(xs), lay(xs.contiguous()))
        # the aligned copy takes the 128-bit gather (pairs of rows per load, the halves' partial sums meet at the row end), the
        # misaligned view the scalar one: the same neighbours summed in a different, equally deterministic order
        assert K.rel_err(conv(xs, ei).cpu(), conv(xs.contiguous(), ei).cpu()) <= 2e-5
        xe = wide[:, ::2][:, :48]                      # column stride 2
        assert torch.equal(lay(xe), lay(xe.contiguous()))


@pytest.mark.parametrize("fast", [False, True])
def test_node_model_without_message_passing_layers(fast):
    """mp_layers = 0 with skip: the model is lay_out(x) (node_classification_clean/models.py:192-203 with an empty loop)."""
    import kagnn_b200 as kb
    torch.manual_seed(3)
    cls = kb.GFASTKAN_Nodes if fast else kb.GKAN_Nodes
    m = cls("gin", 0, 12, 8, 3, skip=True).eval()
    sd = _sd_cpu(m)
    x = torch.randn(50, 12) * 0.5
    ei = torch.randint(0, 50, (2, 100))
    with torch.no_grad():
        y = m.cuda()(x.cuda(), ei.cuda()).cpu()
    assert K.rel_err(y, K.node_model_forward(sd, "gin", x, ei, True)) <= TOL
