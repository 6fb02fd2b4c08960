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


def test_gat_has_no_backward_and_says_so():
    import kagnn_b200 as kb
    conv = kb.KAGATConv(8, 4, 2, 5, 3).cuda()
    x = torch.randn(20, 8).cuda()
    ei = torch.randint(0, 20, (2, 50)).cuda()
    with pytest.raises(NotImplementedError):
        conv(x, ei)
    with torch.no_grad():
        assert conv(x, ei).shape == (20, 8)
