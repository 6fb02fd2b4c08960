"""GPU parity at the STATED shapes of BASELINE.json's configurations C3, C4 (one shard's worth) and C5 -- the models and the
synthetic inputs of SURVEY.md section 8(d), built by scripts/synth_graphs.py exactly as the benchmarks build them.

  C3  gr.KAGIN(1, 1, 4, 128, 2, 5, 3, 1, 0.0, True) on the ZINC-shaped batch of 1 024 graphs (GINE, bond-table lookup):
      the whole model against the oracle (graph_regression/models.py:86-119).
  C5  gc.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0) on the MUTAG-scaled batch of 4 096 graphs, fp32: the whole model against the
      oracle (graph_classification/models.py:125-151).
  C4  KAGCN_Layer(128, 128, 5, 3) on a Graph500 R-MAT graph (0.57, 0.19, 0.19, 0.05) with 2^20 nodes and 10.5 M edges:
      the oracle on a sample of destination rows that contains the heaviest hubs (graph_classification/models.py:157-163).

Tolerance 1e-4 relative (north_star), fp32."""
import os
import sys

import pytest
import torch

from oracle import kagnn_oracle as K

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import synth_graphs as SG  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _sd_cpu(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


def test_c3_zinc_kagin_batch_1024():
    from kagnn_b200 import models_regr
    torch.manual_seed(3)
    data = SG.zinc_batch(1024, seed=12345)
    m = models_regr.KAGIN(1, 1, 4, 128, 2, 5, 3, 1, 0.0, True).eval()
    sd = _sd_cpu(m)
    with torch.no_grad():
        y = m.cuda()(data.to("cuda")).cpu()
    ref = K.gr_kagin_forward(sd, K.Batch(data.x, data.edge_index, data.batch, data.edge_attr))
    assert y.shape == ref.shape == (1024, 1) and torch.isfinite(y).all()
    assert K.rel_err(y, ref) <= TOL


def test_c5_mutag_fastkagin_hidden_256_batch_4096():
    from kagnn_b200 import models_graph
    torch.manual_seed(5)
    data = SG.mutag_batch(4096, seed=12345)
    m = models_graph.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0).eval()
    sd = _sd_cpu(m)
    with torch.no_grad():
        y = m.cuda()(data.to("cuda")).cpu()
    ref = K.gc_kagin_forward(sd, K.Batch(data.x, data.edge_index, data.batch))
    assert y.shape == ref.shape == (4096, 2) and torch.isfinite(y).all()
    assert K.rel_err(y, ref) <= TOL


def _gcn_rows_oracle(x, ei, rows, sd, prefix=""):
    """GCNConv(lin = KANLinear) for the selected destination rows: h = KAN(x) on the rows' 1-hop neighbourhood only, PyG
    gcn_norm weights from the degrees of the whole graph (oracle.gcn_norm semantics: self loops replaced by one unit loop)."""
    n = x.size(0)
    src, dst = ei
    keep = src != dst
    src, dst = src[keep], dst[keep]
    deg = torch.ones(n, dtype=torch.float64)                       # the added self loop
    deg.index_add_(0, dst, torch.ones(dst.numel(), dtype=torch.float64))
    dis = deg.pow(-0.5)
    pos = torch.full((n,), -1, dtype=torch.long)
    pos[rows] = torch.arange(rows.numel())
    sel = pos[dst] >= 0
    s_src, s_dst = src[sel], dst[sel]
    need = torch.unique(torch.cat([s_src, rows]))
    loc = torch.full((n,), -1, dtype=torch.long)
    loc[need] = torch.arange(need.numel())
    h = K._kan_layer_from_sd(sd, prefix + "lin.", x[need]).double()
    out = (dis[rows] * dis[rows]).unsqueeze(1) * h[loc[rows]]
    w = (dis[s_src] * dis[s_dst]).unsqueeze(1)
    out.index_add_(0, pos[s_dst], w * h[loc[s_src]])
    return (out + sd[prefix + "bias"].double()).float()


def test_c4_rmat_kagcn_layer_sampled_rows():
    import kagnn_b200 as kb
    torch.manual_seed(4)
    n, e, f = 1 << 20, 10_500_000, 128
    ei = SG.rmat_edges(n, e, seed=12345, device="cuda").cpu()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(n, f, generator=g)
    conv = kb.KAGCN_Layer(f, 128, 5, 3)
    with torch.no_grad():
        conv.bias.copy_(torch.randn(128, generator=g) * 0.1)
    sd = _sd_cpu(conv)
    with torch.no_grad():
        y = conv.cuda()(x.cuda(), ei.cuda()).cpu()
    assert y.shape == (n, 128) and torch.isfinite(y).all()
    indeg = torch.bincount(ei[1], minlength=n)
    assert int(indeg.max()) > 2000                                 # the skew this test is about: mega-hub rows exist
    hubs = torch.topk(indeg, 6).indices
    rows = torch.unique(torch.cat([torch.randperm(n, generator=g)[:1500], hubs, torch.tensor([0, 127, 128, n - 1])]))
    ref = _gcn_rows_oracle(x, ei, rows, sd)
    assert K.rel_err(y[rows], ref) <= TOL
