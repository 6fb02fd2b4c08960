"""GAT flavour: KAGATConv / FASTKAGATConv and the models built on them against the oracle's
restatement of PyG 2.5 GATConv (parity unpinned for the attention arithmetic, like the rest of the PyG half; the model glue is
pinned by the nc_*_gat / gc_*gat fixtures generated from the reference's own models.py -- tests/test_gpu_parity.py runs those)."""
import pytest
import torch

from oracle import kagnn_oracle as K

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _sd_cpu(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def _graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n - 5, (2, e), generator=g)            # the last 5 nodes are isolated: attention = the self loop alone
    ei[1, :40] = ei[0, :40]                                       # existing self loops (PyG removes them, then adds one per node)
    ei[:, 40:80] = ei[:, 80:120]                                  # duplicate edges
    return ei, g


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("f,c,heads", [(33, 8, 4), (64, 16, 1), (16, 6, 3), (128, 32, 2)])
def test_gat_conv_against_oracle(f, c, heads, fast):
    import kagnn_b200 as kb
    torch.manual_seed(f + c + heads)
    n, e = 3000, 14_000
    ei, g = _graph(n, e, seed=f)
    x = torch.randn(n, f, generator=g) * 0.6
    conv = kb.FASTKAGATConv(f, c, heads, 6) if fast else kb.KAGATConv(f, c, heads, 5, 3)
    with torch.no_grad():
        conv.bias.normal_(0, 0.2)
        conv.att_src.mul_(3.0)                                    # spread the attention logits: softmax far from uniform
        conv.att_dst.mul_(3.0)
    sd = _sd_cpu(conv)
    with torch.no_grad():
        y = conv.cuda()(x.cuda(), ei.cuda()).cpu()
    lin = (lambda t: K._fastkan_layer_from_sd(sd, "lin.", t)) if fast else (lambda t: K._kan_layer_from_sd(sd, "lin.", t))
    ref = K.gat_conv(x, ei, lin, sd["att_src"], sd["att_dst"], sd["bias"], heads)
    assert y.shape == ref.shape == (n, heads * c)
    assert K.rel_err(y, ref) <= TOL


def test_gat_attention_rows_sum_to_one_and_ignore_existing_self_loops():
    from kagnn_b200 import ops
    from kagnn_b200.graph import GraphCSR
    torch.manual_seed(0)
    n, heads, c = 500, 3, 8
    ei, g = _graph(n, 3000, seed=9)
    h = torch.randn(n, heads * c, generator=g).cuda()
    gr = GraphCSR(ei.cuda(), n)
    att = torch.randn(2, 1, heads, c, generator=g).cuda()
    w, sw = ops.gat_attention(h, gr.csr, att[0], att[1], heads)
    assert w.shape == (heads, ei.size(1)) and sw.shape == (heads, n)
    row = torch.repeat_interleave(torch.arange(n), (gr.rowptr[1:] - gr.rowptr[:-1]).long().cpu())
    tot = sw.cpu() + torch.zeros(heads, n).index_add_(1, row, w.cpu())
    assert torch.allclose(tot, torch.ones(heads, n), atol=1e-5)
    loops = gr.col.cpu().long() == row
    assert loops.any() and float(w.cpu()[:, loops].abs().max()) == 0.0
    assert torch.allclose(sw.cpu()[:, -5:], torch.ones(heads, 5))              # isolated nodes attend to themselves only


GRAD_TOL = 1e-3


def _kink_margin(h, att_src, att_dst, ei, heads):
    """Smallest |a_src[j] + a_dst[i]| over the attended pairs.  leaky_relu has a kink at 0: an edge whose pre-activation changes
    sign between the GPU forward and the CPU oracle (their projections differ by ~3e-6) gets a gradient 5x different, which a
    max-norm comparison of gradients cannot absorb -- the gradient tests therefore run on graphs that keep every pair away from 0."""
    n = h.size(0)
    c = h.size(1) // heads
    hv = h.double().view(n, heads, c)
    a_s = (hv * att_src.double().view(1, heads, c)).sum(-1)
    a_d = (hv * att_dst.double().view(1, heads, c)).sum(-1)
    keep = ei[0] != ei[1]
    loops = torch.arange(n)
    r, cc = torch.cat([ei[0][keep], loops]), torch.cat([ei[1][keep], loops])
    return float((a_s[r] + a_d[cc]).abs().min())


def _safe_graph(n, e, seed, project, att_src, att_dst, heads, margin=5e-5):
    """A graph (seed, seed + 1000, ...) on which no attended pair sits within `margin` of the leaky_relu kink."""
    for k in range(200):
        ei, g = _graph(n, e, seed + 1000 * k)
        if _kink_margin(project, att_src, att_dst, ei, heads) > margin:
            return ei
    raise AssertionError("no kink-free graph found")


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("f,c,heads", [(33, 8, 4), (64, 16, 1), (16, 6, 3)])
def test_gat_conv_gradients_against_autograd_through_the_oracle(f, c, heads, fast):
    """loss.backward() through KAGATConv / FASTKAGATConv (projection Function + autograd.gat_attend -> kagnn_gat_bwd) against torch
    autograd through the oracle's restatement of GATConv: d x, d att_src, d att_dst, d bias and the projection's weights."""
    import kagnn_b200 as kb
    torch.manual_seed(f * 3 + c + heads)
    n, e = 300, 600                                 # few attended pairs, so that a kink-free graph exists (see _kink_margin)
    g = torch.Generator().manual_seed(f + 1)
    x = torch.randn(n, f, generator=g) * 0.6
    dy = torch.randn(n, heads * c, generator=g)
    conv = kb.FASTKAGATConv(f, c, heads, 6) if fast else kb.KAGATConv(f, c, heads, 5, 3)
    with torch.no_grad():
        conv.bias.normal_(0, 0.2)
        conv.att_src.mul_(3.0)
        conv.att_dst.mul_(3.0)
    sd = _sd_cpu(conv)
    proj = (K._fastkan_layer_from_sd(sd, "lin.", x) if fast else K._kan_layer_from_sd(sd, "lin.", x))
    ei = _safe_graph(n, e, f + 1, proj, sd["att_src"], sd["att_dst"], heads)
    conv = conv.cuda()
    xd = x.cuda().requires_grad_(True)
    y = conv(xd, ei.cuda())
    y.backward(dy.cuda())
    # reference: autograd through the oracle
    ps = {k: v.clone().requires_grad_(v.is_floating_point() and "grid" not in k) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    lin = (lambda t: K._fastkan_layer_from_sd(ps, "lin.", t)) if fast else (lambda t: K._kan_layer_from_sd(ps, "lin.", t))
    ref = K.gat_conv(xr, ei, lin, ps["att_src"], ps["att_dst"], ps["bias"], heads)
    assert K.rel_err(y.detach().cpu(), ref.detach()) <= TOL
    ref.backward(dy)
    assert K.rel_err(xd.grad.cpu(), xr.grad) <= GRAD_TOL
    got = dict(conv.named_parameters())
    checked = 0
    for k, v in ps.items():
        if v.grad is None or k not in got:
            continue
        assert got[k].grad is not None, k
        assert K.rel_err(got[k].grad.cpu(), v.grad) <= GRAD_TOL, k
        checked += 1
    assert checked >= 4                                         # att_src, att_dst, bias and the projection's weights


@pytest.mark.parametrize("fast", [False, True])
def test_gat_node_model_training_step_matches_the_oracle(fast):
    """GKAN_Nodes / GFASTKAN_Nodes with conv_type='gat' in training mode (batch-statistics BatchNorm): loss and gradients of a
    cross-entropy step against autograd through the oracle's model forward."""
    import kagnn_b200 as kb
    torch.manual_seed(3)
    n, f, classes = 400, 24, 5                     # few attended pairs: none of them near the leaky_relu kink at these seeds (see _kink_margin)
    ei, g = _graph(n, 900, seed=4)
    x = torch.randn(n, f, generator=g) * 0.5
    labels = torch.randint(0, classes, (n,), generator=g)
    if fast:
        m = kb.GFASTKAN_Nodes("gat", 2, f, 8, classes, skip=True, grid_size=6, hidden_layers=2, dropout=0.0, heads=2).train()
    else:
        m = kb.GKAN_Nodes("gat", 2, f, 8, classes, skip=True, grid_size=5, spline_order=3, hidden_layers=2, dropout=0.0, heads=2).train()
    sd = _sd_cpu(m)
    m = m.cuda()
    loss = torch.nn.functional.cross_entropy(m(x.cuda(), ei.cuda()), labels.cuda())
    loss.backward()
    ps = {k: v.clone().requires_grad_(v.is_floating_point() and "grid" not in k and "running" not in k) for k, v in sd.items()}
    ref_loss = torch.nn.functional.cross_entropy(K.node_model_forward(ps, "gat", x, ei, True, training=True), labels)
    ref_loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss.detach())) <= 1e-4 * max(1.0, abs(float(ref_loss.detach())))
    got = dict(m.named_parameters())
    worst, checked = 0.0, 0
    for k, v in ps.items():
        if v.grad is None or k not in got or got[k].grad is None:
            continue
        if float(v.grad.abs().max()) < 1e-6:          # a conv bias in front of a batch-statistics BatchNorm: its gradient is rounding noise
            continue
        err = K.rel_err(got[k].grad.cpu(), v.grad)
        assert err <= GRAD_TOL, (k, err)
        worst = max(worst, err)
        checked += 1
    assert checked >= 10, checked


@pytest.mark.parametrize("n,heads,c,e", [(300, 4, 8, 1500), (2000, 1, 16, 9000), (2000, 3, 6, 9000), (5000, 2, 64, 40_000)])
def test_gat_attention_backward_on_identical_inputs(n, heads, c, e):
    """autograd.gat_attend (kagnn_gat_bwd + the aggregation over the reversed edges) with h as a leaf: both sides see exactly the
    same projected features, so every attended pair is on the same side of the leaky_relu kink and the comparison is tight."""
    from kagnn_b200 import autograd
    from kagnn_b200.graph import GraphCSR
    ei, g = _graph(n, e, seed=n + heads)
    h0 = torch.randn(n, heads * c, generator=g)
    a_s, a_d = torch.randn(1, heads, c, generator=g) * 1.5, torch.randn(1, heads, c, generator=g) * 1.5
    bias, dy = torch.randn(heads * c, generator=g), torch.randn(n, heads * c, generator=g)
    h, p1, p2, pb = (t.clone().double().requires_grad_(True) for t in (h0, a_s, a_d, bias))
    K.gat_conv(h, ei, lambda t: t, p1, p2, pb, heads).backward(dy.double())
    hd, q1, q2, qb = (t.cuda().requires_grad_(True) for t in (h0, a_s, a_d, bias))
    out = autograd.gat_attend(hd, q1, q2, qb, GraphCSR(ei.cuda(), n), heads, 0.2)
    out.backward(dy.cuda())
    for got, ref in ((hd, h), (q1, p1), (q2, p2), (qb, pb)):
        assert K.rel_err(got.grad.cpu().double(), ref.grad) <= 2e-5


def test_gat_graph_classification_model_trains():
    """gc.KAGAT (graph_classification/models.py:194-216) under autograd: loss decreases over a few Adam steps and every parameter
    receives a finite gradient."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import synth_graphs as SG
    from kagnn_b200 import models_graph
    torch.manual_seed(0)
    data = SG.mutag_batch(64, seed=3).to("cuda")
    m = models_graph.KAGAT(2, 7, 8, 2, 5, 3, 0.0, 2).cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    labels = (torch.arange(64, device="cuda") % 2)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = torch.nn.functional.nll_loss(m(data), labels)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0]
