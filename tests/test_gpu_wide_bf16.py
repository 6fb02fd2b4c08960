"""Layers wider than 128 outputs (BASELINE config C5: hidden 256) and the bf16 single-product mode of the pipelined kernel.

fp32 mode (default): every configuration must match the oracle within 1e-4 relative (north_star) -- also the wide ones, which
run one launch per layer (kagnn_b200.ops._wide_chain) with a 256-column accumulator.
bf16 mode (kagnn_set_precision(KAGNN_PREC_BF16), config C5): operands rounded to bf16 once, fp32 accumulate; north_star states no
tolerance for it, SURVEY.md section 8(d) sets 2e-2 relative against the fp32 oracle -- written here as BF16_TOL."""
import os
import sys

import pytest
import torch

from oracle import kagnn_oracle as K

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import synth_graphs as SG  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-4
BF16_TOL = 2e-2


def _sd_cpu(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


@pytest.fixture
def bf16_mode():
    import kagnn_b200 as kb
    kb.set_precision("bf16")
    assert kb.get_precision() == "bf16"
    yield
    kb.set_precision("fp32")


@pytest.mark.parametrize("sizes", [[256, 256, 2], [7, 256, 256], [200, 256, 40], [64, 136], [300, 256, 256, 16]])
@pytest.mark.parametrize("fast", [False, True])
def test_wide_chains_fp32(sizes, fast):
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(sum(sizes))
    n = 1000
    x = torch.randn(n, sizes[0]) * 0.7
    m = kb.FastKAN(sizes, num_grids=8) if fast else kb.KAN(sizes, grid_size=5, spline_order=3)
    sd = _sd_cpu(m)
    before = ops.launch_counters()
    with torch.no_grad():
        y = m.cuda()(x.cuda()).cpu()
    after = ops.launch_counters()
    ref = K.fastkan_chain(sd, "layers.", x) if fast else K.kan_chain(sd, "layers.", x)
    assert K.rel_err(y, ref) <= TOL
    assert after["tc2"] - before["tc2"] >= len(sizes) - 1 or max(sizes[1:]) <= 128      # the pipelined kernel ran every wide layer


def test_wide_gin_layer_with_gather_and_batchnorm_fp32():
    """GIN aggregation in front of a 256-wide FastKAN chain (config C5's second layer): aggregation launch, LayerNorm statistics
    pre-pass, one launch per layer, eval BatchNorm folded into the last one."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(0)
    n, e, f = 3000, 9000, 256
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, f, generator=g) * 0.4
    ei = torch.randint(0, n, (2, e), generator=g)
    conv = kb.GIFASTKANLayer(f, 256, 8, 256, 2)
    sd = _sd_cpu(conv)
    post = ops.Affine((torch.rand(256, generator=g) + 0.5).cuda(), (torch.randn(256, generator=g) * 0.1).cuda())
    with torch.no_grad():
        y = conv.cuda()(x.cuda(), ei.cuda(), post=post).cpu()
    ref = K.gin_conv(x, ei, lambda t: K.fastkan_chain(sd, "nn.layers.", t)) * post.scale.cpu() + post.shift.cpu()
    assert K.rel_err(y, ref) <= TOL


@pytest.mark.parametrize("sizes,fast", [([128, 64, 64], False), ([256, 256, 2], True), ([33, 7], False), ([64, 64, 64], True),
                                        ([320, 40], False), ([200, 256, 40], False)])
def test_bf16_mode_chains(sizes, fast, bf16_mode):
    import kagnn_b200 as kb
    torch.manual_seed(sum(sizes) + 1)
    n = 1500
    x = torch.randn(n, sizes[0]) * 0.6
    m = kb.FastKAN(sizes, num_grids=8) if fast else kb.KAN(sizes, grid_size=5, spline_order=3)
    sd = _sd_cpu(m)
    with torch.no_grad():
        y = m.cuda()(x.cuda()).cpu()
    ref = K.fastkan_chain(sd, "layers.", x) if fast else K.kan_chain(sd, "layers.", x)
    err = K.rel_err(y, ref)
    assert err <= BF16_TOL
    assert err > 1e-6                      # it really ran the single-product path (fp32 mode lands near 1e-5 or below)


def test_c5_mutag_fastkagin_hidden_256_bf16(bf16_mode):
    """BASELINE config C5 as stated: fastkan KAGIN hidden 256 grid 8 in bf16 on the MUTAG-scaled batch of 4 096 graphs."""
    from kagnn_b200 import models_graph
    torch.manual_seed(5)
    data = SG.mutag_batch(4096, seed=12345)
    m = models_graph.FASTKAGIN(2, 7, 256, 2, 2, 8, 0.0).eval()
    sd = _sd_cpu(m)
    with torch.no_grad():
        y = m.cuda()(data.to("cuda")).cpu()
    ref = K.gc_kagin_forward(sd, K.Batch(data.x, data.edge_index, data.batch))
    assert y.shape == ref.shape == (4096, 2) and torch.isfinite(y).all()
    assert K.rel_err(y, ref) <= BF16_TOL


def test_precision_switch_restores_fp32():
    import kagnn_b200 as kb
    torch.manual_seed(9)
    x = torch.randn(700, 64) * 0.5
    m = kb.KAN([64, 64, 16], grid_size=5, spline_order=3)
    sd = _sd_cpu(m)
    ref = K.kan_chain(sd, "layers.", x)
    mc = m.cuda()
    with torch.no_grad():
        kb.set_precision("bf16")
        try:
            e16 = K.rel_err(mc(x.cuda()).cpu(), ref)
        finally:
            kb.set_precision("fp32")
        e32 = K.rel_err(mc(x.cuda()).cpu(), ref)
    assert e32 <= TOL < e16 <= BF16_TOL
