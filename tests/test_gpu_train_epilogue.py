"""Fused training-mode epilogue dropout(bn(x)) (nc/models.py:197-198): batch statistics + affine +
running-estimate update + Philox dropout in two launches, mask regenerated in the backward.  Reference = torch on the GPU with the
mask read back from the output (a dropout mask is random by definition; what must agree is everything else)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,c,p", [(5000, 64, 0.5), (777, 33, 0.1), (4096, 128, 0.9)])
def test_bn_dropout_forward_and_backward_match_torch_under_the_same_mask(n, c, p):
    from kagnn_b200 import autograd as KA
    torch.manual_seed(n + c)
    bn = torch.nn.BatchNorm1d(c).cuda().train()
    ref_bn = torch.nn.BatchNorm1d(c).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
        ref_bn.load_state_dict(bn.state_dict())
    x = (torch.randn(n, c, device="cuda") * 2.0 + 0.7).requires_grad_(True)
    y = KA.batch_norm_dropout_train(x, bn, p)
    keep = (y != 0)
    frac = 1.0 - float(keep.float().mean())
    assert abs(frac - p) < 4.0 * (p * (1 - p) / (n * c)) ** 0.5 + 1e-3            # the mask drops a fraction p
    xr = x.detach().clone().requires_grad_(True)
    yr = ref_bn(xr) * keep / (1.0 - p)
    assert float((y - yr).abs().max()) <= 2e-5 * max(1.0, float(yr.abs().max()))
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, atol=1e-6) and torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-5)
    assert int(bn.num_batches_tracked) == 1
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy)
    scale = float(xr.grad.abs().max())
    assert float((x.grad - xr.grad).abs().max()) <= 2e-4 * scale
    assert torch.allclose(bn.weight.grad, ref_bn.weight.grad, rtol=2e-4, atol=2e-4 * float(ref_bn.weight.grad.abs().max()))
    assert torch.allclose(bn.bias.grad, ref_bn.bias.grad, rtol=2e-4, atol=2e-4 * float(ref_bn.bias.grad.abs().max()))


def test_mask_is_a_function_of_the_seed():
    from kagnn_b200 import ops
    bn = torch.nn.BatchNorm1d(16).cuda().train()
    x = torch.randn(300, 16, device="cuda")
    a = ops.batchnorm_dropout_forward(x, bn, 0.4, 1234)
    b = ops.batchnorm_dropout_forward(x, bn, 0.4, 1234)
    c = ops.batchnorm_dropout_forward(x, bn, 0.4, 1235)
    assert torch.equal(a != 0, b != 0) and not torch.equal(a != 0, c != 0)
    rows = (a != 0).float().mean(1)
    assert float(rows.min()) > 0.1 and float(rows.max()) < 0.98               # no structure along rows


def test_training_loop_with_dropout_learns():
    """The reference's train() (nc/utils.py:125-132) on a model with dropout > 0: the fused epilogue is on the path, the loss goes down."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(0)
    n, f, c = 600, 24, 4
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, f, generator=g).cuda()
    labels = (x[:, :c].argmax(1)).cuda()
    ei = torch.randint(0, n, (2, 2400), generator=g).cuda()
    model = kb.GKAN_Nodes("gin", 2, f, 16, c, skip=True, grid_size=5, spline_order=3, hidden_layers=2, dropout=0.3).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    losses = []
    for _ in range(40):
        model.train()
        opt.zero_grad()
        loss = F.cross_entropy(model(x, ei), labels)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.7 * losses[0]
    model.eval()
    with torch.no_grad():
        acc = float((model(x, ei).argmax(1) == labels).float().mean())
    assert acc > 0.5
