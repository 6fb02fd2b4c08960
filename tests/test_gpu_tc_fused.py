"""GPU parity tests of the tcgen05 fused kernel (fused_tc.cu): forced onto the tensor-core path, compared with the
CPU oracle and with the general fp32 kernel on the same inputs.  Tolerance 1e-4 relative (BASELINE.json)."""
import pytest
import torch

from oracle import kagnn_oracle as K

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _sd_cpu(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


@pytest.fixture(params=["pipelined", "general"])
def tc_only(request):
    """pipelined = fused_tc2.cu (A operand in TMEM) wherever it applies, then fused_tc.cu; general = fused_tc.cu only."""
    from kagnn_b200 import ops, _lib as L
    ops.set_path(L.PATH_TC)
    ops.set_tc_variant(1 if request.param == "general" else 0)
    yield request.param
    ops.set_tc_variant(0)
    ops.set_path(L.PATH_AUTO)


@pytest.mark.parametrize("G,k,fin,fout,n", [
    (5, 3, 16, 16, 128), (5, 3, 128, 64, 1000), (5, 3, 64, 64, 4097), (4, 3, 7, 1, 130), (5, 3, 320, 40, 513),
    (2, 2, 33, 17, 64), (3, 1, 5, 200, 77), (5, 3, 128, 128, 2048), (5, 3, 1433, 32, 300), (5, 3, 256, 256, 500),
])
def test_kan_linear_tc(tc_only, G, k, fin, fout, n):
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(G * 1000 + fin)
    m = kb.KANLinear(fin, fout, grid_size=G, spline_order=k)
    x = torch.randn(n, fin) * 0.9
    y_ref = K._kan_layer_from_sd(_sd_cpu(m), "", x)
    c0 = ops.launch_counters()["tc"]
    with torch.no_grad():
        y = m.cuda()(x.cuda()).cpu()
    assert ops.launch_counters()["tc"] == c0 + 1
    assert K.rel_err(y, y_ref) <= TOL


def test_pipelined_kernel_is_the_one_that_runs():
    """The arxiv-shaped GIN layer must go through fused_tc2.cu (counted separately by the library)."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(1)
    gin = kb.GIKANLayer(128, 64, 5, 3, 64, 2).cuda()
    ei = torch.randint(0, 1000, (2, 6000)).cuda()
    c0 = ops.launch_counters()["tc2"]
    with torch.no_grad():
        gin(torch.randn(1000, 128).cuda(), ei)
    assert ops.launch_counters()["tc2"] == c0 + 1


def test_kan_special_values_tc(tc_only):
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
    assert torch.isnan(y[6]).all() and torch.isnan(y[5]).all()


@pytest.mark.parametrize("sizes,G,k,n", [([128, 64, 64], 5, 3, 3000), ([64, 64, 64], 5, 3, 777), ([128, 128, 128], 5, 3, 515),
                                         ([30, 9, 9, 9, 9, 3], 4, 2, 129), ([320, 40], 5, 3, 1000)])
def test_kan_chain_tc(tc_only, sizes, G, k, n):
    import kagnn_b200 as kb
    torch.manual_seed(len(sizes) + G)
    m = kb.KAN(sizes, grid_size=G, spline_order=k)
    x = torch.randn(n, sizes[0])
    y_ref = K.kan_chain(_sd_cpu(m), "layers.", x)
    with torch.no_grad():
        y = m.cuda()(x.cuda()).cpu()
    assert K.rel_err(y, y_ref) <= TOL


@pytest.mark.parametrize("sizes,G,n", [([256, 256, 256], 8, 700), ([7, 256, 256], 8, 333), ([64, 32, 16, 8, 4], 5, 129),
                                       ([20, 200], 3, 70), ([1433, 16], 4, 100)])
def test_fastkan_tc(tc_only, sizes, G, n):
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


def _rand_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    if e > 20:
        ei[1, :10] = ei[0, :10]
        ei[:, 10:15] = ei[:, 15:20]
    return ei


def test_fused_gin_layer_tc_matches_oracle_and_fp32(tc_only):
    import kagnn_b200 as kb
    from kagnn_b200 import ops, _lib as L
    torch.manual_seed(5)
    n, e, f, h = 5000, 40000, 128, 64
    ei = _rand_graph(n, e, 99)
    x = torch.randn(n, f) * 0.5
    gin = kb.GIKANLayer(f, h, 5, 3, h, 2)
    sd = _sd_cpu(gin)
    ref = K.gin_conv(x, ei, lambda t: K.kan_chain(sd, "nn.layers.", t))
    gin = gin.cuda()
    with torch.no_grad():
        out_tc = gin(x.cuda(), ei.cuda()).cpu()
        ops.set_path(L.PATH_FP32)
        out_fp = gin(x.cuda(), ei.cuda()).cpu()
        ops.set_path(L.PATH_TC)
    assert K.rel_err(out_tc, ref) <= TOL
    assert K.rel_err(out_fp, ref) <= TOL
    assert K.rel_err(out_tc, out_fp) <= 5e-5


def test_gine_and_gcn_tc(tc_only):
    import kagnn_b200 as kb
    torch.manual_seed(6)
    n, e, f, h = 700, 5000, 32, 16
    ei = _rand_graph(n, e, 7)
    x = torch.randn(n, f)
    ea = torch.randn(e, f)
    gine = kb.GINEConv(kb.make_kan(f, 16, h, 2, 4, 3))
    sd = _sd_cpu(gine)
    ref = K.gine_conv(x, ei, ea, lambda t: K.kan_chain(sd, "nn.layers.", t))
    with torch.no_grad():
        out = gine.cuda()(x.cuda(), ei.cuda(), ea.cuda()).cpu()
    assert K.rel_err(out, ref) <= TOL


def test_node_model_auto_path_uses_tc_and_matches():
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(12)
    n, e, f = 20000, 150000, 128
    for conv in ("gin", "gcn"):
        m = kb.GKAN_Nodes(conv, 3, f, 64, 40, skip=True, grid_size=5, spline_order=3, hidden_layers=2).eval()
        with torch.no_grad():
            for name, b in m.named_buffers():
                if name.endswith("running_var"):
                    b.uniform_(0.01, 0.05)
            for name, p_ in m.named_parameters():
                if name.endswith("bias") and p_.dim() == 1:
                    p_.normal_(0, 0.2)
        ei = _rand_graph(n, e, 3)
        x = torch.randn(n, f) * 0.3
        y_ref = K.node_model_forward(_sd_cpu(m), conv, x, ei, True)
        c0 = ops.launch_counters()
        with torch.no_grad():
            y = m.cuda()(x.cuda(), ei.cuda()).cpu()
        c1 = ops.launch_counters()
        assert c1["tc"] - c0["tc"] >= 4, (c0, c1)
        assert K.rel_err(y, y_ref) <= TOL, conv


@pytest.mark.parametrize("sizes,G,n,ln", [([128, 128, 128], 8, 1000, True), ([7, 64, 64], 8, 300, True), ([64, 32, 16, 8, 4], 5, 129, True),
                                          ([100, 128], 6, 515, False), ([16, 16], 2, 128, True), ([128, 40], 8, 4097, True)])
def test_fastkan_pipelined_kernel(sizes, G, n, ln):
    """FastKAN chains up to 128 wide run in the pipelined kernel (RBF basis + LayerNorm statistics in the producers, base bias
    folded into the TMEM read-back / epilogue); result against the oracle."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(sum(sizes) + G)
    m = kb.FastKAN(sizes, num_grids=G)
    if not ln:
        for lay in m.layers:
            lay.layernorm = None
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1 and p.requires_grad:
                p.add_(torch.randn_like(p) * 0.1)
    x = torch.randn(n, sizes[0]) * 1.5 + 0.3
    y_ref = K.fastkan_chain(_sd_cpu(m), "layers.", x)
    c0 = ops.launch_counters()["tc2"]
    with torch.no_grad():
        y = m.cuda()(x.cuda()).cpu()
    assert ops.launch_counters()["tc2"] == c0 + 1
    assert K.rel_err(y, y_ref) <= TOL


def test_fastkan_gin_layer_pipelined():
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(9)
    n, e, f = 3000, 20000, 128
    ei = _rand_graph(n, e, 5)
    x = torch.randn(n, f) * 0.4
    conv = kb.GIFASTKANLayer(f, 64, 8, 64, 2)
    sd = _sd_cpu(conv)
    c0 = ops.launch_counters()["tc2"]
    with torch.no_grad():
        y = conv.cuda()(x.cuda(), ei.cuda()).cpu()
    assert ops.launch_counters()["tc2"] == c0 + 1
    ref = K.gin_conv(x, ei, lambda t: K.fastkan_chain(sd, "nn.layers.", t))
    assert K.rel_err(y, ref) <= TOL


@pytest.mark.parametrize("fin,fout,n,split", [(320, 40, 1500, 128), (300, 64, 500, 0), (1433, 16, 200, 0), (256, 128, 777, 128)])
def test_fastkan_wide_input_uses_layernorm_prepass(fin, fout, n, split):
    """FastKAN layers whose input is wider than one 128-column tile unit (the skip-concat read-out): LayerNorm statistics from
    the kagnn_layernorm_stats pre-pass, the layer itself in the pipelined kernel (optionally over two-part rows)."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops, _lib as L
    torch.manual_seed(fin + fout)
    m = kb.FastKANLayer(fin, fout, num_grids=8)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1 and p.requires_grad:
                p.add_(torch.randn_like(p) * 0.1)
    x = torch.randn(n, fin) * 1.3 - 0.2
    y_ref = K._fastkan_layer_from_sd(_sd_cpu(m), "", x)
    m = m.cuda()
    c0 = ops.launch_counters()["tc2"]
    with torch.no_grad():
        if split:
            xd = x.cuda()
            y = ops.fused_layer(ops.AggSpec(L.AGG_NONE, xd[:, split:], x_head=xd[:, :split].contiguous()), n, m.kernel_specs()).cpu()
        else:
            y = m(x.cuda()).cpu()
    assert ops.launch_counters()["tc2"] == c0 + 1
    assert K.rel_err(y, y_ref) <= TOL
