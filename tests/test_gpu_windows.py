"""B-spline layers with more than eight coefficients per (in, out) pair (grid_size + spline_order > 8, the upper part of the
reference's search space: node_classification/one_experiment.py:45-46) run on the tensor-core kernels as `windows` virtual
features of eight slots each (kagnn_b200/ekan.py: KANLinear._windowed_spec, kagnn_expand_windows).  Forward against the oracle's
restatement of KANLinear.forward / the node model, and the launch counters say which kernel ran."""
import pytest
import torch

from oracle import kagnn_oracle as K

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("rows,in_f,out_f,G,k", [(1000, 24, 32, 8, 3), (513, 40, 64, 6, 3), (300, 7, 5, 13, 3), (260, 16, 128, 20, 2),
                                                  (150, 9, 12, 30, 1)])
def test_windowed_layer_matches_oracle_on_the_tensor_core_kernel(rows, in_f, out_f, G, k):
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(rows + G)
    lay = kb.KANLinear(in_f, out_f, grid_size=G, spline_order=k)
    with torch.no_grad():
        lay.spline_weight.normal_(0, 0.3)
        lay.spline_scaler.uniform_(0.5, 1.5)
    sd = {kk: v.detach().clone() for kk, v in lay.state_dict().items()}
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, in_f, generator=g) * 0.8
    x[0, 0], x[1, 0], x[2, 0] = 1.0, -1.0, 5.0                      # the ends of the grid range, and outside the knots
    lay = lay.cuda()
    spec = lay.kernel_specs()[0]
    assert spec.windows == (G + k + 7) // 8 and spec.in_features == spec.windows * in_f and spec.grid_size + spec.spline_order == 8
    c0 = ops.launch_counters()
    with torch.no_grad():
        y = lay(x.cuda())
    c1 = ops.launch_counters()
    assert c1["fp32"] == c0["fp32"] and (c1["tc"] + c1["tc2"]) > (c0["tc"] + c0["tc2"])     # not the general fp32 kernel
    ref = K.kan_linear(x, sd["base_weight"], sd["spline_weight"], sd.get("spline_scaler"), sd["grid"], k)
    assert K.rel_err(y.cpu(), ref) <= TOL


@pytest.mark.parametrize("conv", ["gcn", "gin"])
def test_node_model_with_a_large_grid(conv):
    """GKAN_Nodes with grid_size 8, spline_order 3 (S = 11): aggregation, windowed KAN chains, eval BatchNorm, skip read-out."""
    import kagnn_b200 as kb
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(11)
    n, f = 1500, 20
    x = torch.randn(n, f, generator=g) * 0.6
    ei = torch.randint(0, n, (2, 6000), generator=g)
    m = kb.GKAN_Nodes(conv, 2, f, 16, 5, skip=True, grid_size=8, spline_order=3, hidden_layers=2, dropout=0.0).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        y = m.cuda()(x.cuda(), ei.cuda())
    ref = K.node_model_forward(sd, conv, x, ei, True)
    assert K.rel_err(y.cpu(), ref) <= TOL


def test_training_step_through_a_windowed_layer():
    """loss.backward() through KANLinear(G = 8, k = 3): gradients against autograd through the oracle."""
    import kagnn_b200 as kb
    torch.manual_seed(5)
    lay = kb.KANLinear(12, 10, grid_size=8, spline_order=3)
    sd = {kk: v.detach().clone() for kk, v in lay.state_dict().items()}
    x = torch.randn(400, 12) * 0.7
    dy = torch.randn(400, 10)
    lay = lay.cuda().train()
    xd = x.cuda().requires_grad_(True)
    lay(xd).backward(dy.cuda())
    xr = x.clone().requires_grad_(True)
    params = {kk: v.clone().requires_grad_(kk != "grid") for kk, v in sd.items()}
    K.kan_linear(xr, params["base_weight"], params["spline_weight"], params.get("spline_scaler"), params["grid"], 3).backward(dy)
    assert K.rel_err(xd.grad.cpu(), xr.grad) <= 2e-4
    assert K.rel_err(lay.base_weight.grad.cpu(), params["base_weight"].grad) <= 2e-4
    assert K.rel_err(lay.spline_weight.grad.cpu(), params["spline_weight"].grad) <= 2e-4
    assert K.rel_err(lay.spline_scaler.grad.cpu(), params["spline_scaler"].grad) <= 2e-4


@pytest.mark.parametrize("rows,in_f,out_f,G,ln", [(600, 24, 32, 12, True), (500, 40, 64, 16, True), (300, 10, 12, 32, True),
                                                   (400, 16, 20, 20, False), (700, 96, 256, 10, True)])
def test_windowed_fastkan_layer_forward_and_gradients(rows, in_f, out_f, G, ln):
    """FastKANLayer with more than eight centres (the reference searches num_grids up to 32): 8-centre tensor-core kernels over
    copies of the input, the shift in the copies' LayerNorm bias (or in the input when there is no LayerNorm); forward and every
    gradient against autograd through the oracle's restatement of FastKANLayer.forward."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    ops.set_rbf_windows(True)                                        # opt-in: see the accuracy note at ops.set_rbf_windows
    try:
        _windowed_fastkan_case(rows, in_f, out_f, G, ln)
    finally:
        ops.set_rbf_windows(False)


def _windowed_fastkan_case(rows, in_f, out_f, G, ln):
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(rows + G)
    lay = kb.FastKANLayer(in_f, out_f, num_grids=G, use_layernorm=ln)
    with torch.no_grad():
        lay.spline_linear.weight.normal_(0, 0.3)
        if ln:
            lay.layernorm.weight.uniform_(0.5, 1.5)
            lay.layernorm.bias.normal_(0, 0.2)
    sd = {kk: v.detach().clone() for kk, v in lay.state_dict().items()}
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, in_f, generator=g) * 0.9
    dy = torch.randn(rows, out_f, generator=g)
    lay = lay.cuda().train()
    spec = lay.kernel_specs()[0]
    assert spec.windows == (G + 7) // 8 and spec.grid_size == 8 and spec.in_features == spec.windows * in_f
    c0 = ops.launch_counters()
    xd = x.cuda().requires_grad_(True)
    y = lay(xd)
    c1 = ops.launch_counters()
    assert c1["fp32"] == c0["fp32"] and (c1["tc"] + c1["tc2"]) > (c0["tc"] + c0["tc2"])
    y.backward(dy.cuda())
    xr = x.clone().requires_grad_(True)
    p = {kk: v.clone().requires_grad_(kk != "rbf.grid") for kk, v in sd.items()}
    ref = K.fastkan_layer(xr, p.get("layernorm.weight"), p.get("layernorm.bias"), p["rbf.grid"], p["spline_linear.weight"],
                          p["base_linear.weight"], p["base_linear.bias"])
    ref.backward(dy)
    assert K.rel_err(y.detach().cpu(), ref.detach()) <= TOL
    assert K.rel_err(xd.grad.cpu(), xr.grad) <= 1e-3
    assert K.rel_err(lay.spline_linear.weight.grad.cpu(), p["spline_linear.weight"].grad) <= 1e-3
    assert K.rel_err(lay.base_linear.weight.grad.cpu(), p["base_linear.weight"].grad) <= 1e-3
    assert K.rel_err(lay.base_linear.bias.grad.cpu(), p["base_linear.bias"].grad) <= 1e-3
    if ln:
        assert K.rel_err(lay.layernorm.weight.grad.cpu(), p["layernorm.weight"].grad) <= 1e-3
        assert K.rel_err(lay.layernorm.bias.grad.cpu(), p["layernorm.bias"].grad) <= 1e-3


def test_fastkan_with_many_centres_defaults_to_the_fp32_kernel():
    """Without the opt-in a FastKANLayer with more than eight centres keeps the general fp32 kernel (parity bound of 1e-4 on deep
    models; ops.set_rbf_windows)."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    assert not ops.rbf_windows_enabled()
    torch.manual_seed(2)
    lay = kb.FastKANLayer(16, 12, num_grids=12)
    sd = {kk: v.detach().clone() for kk, v in lay.state_dict().items()}
    x = torch.randn(300, 16) * 0.8
    lay = lay.cuda()
    assert lay.kernel_specs()[0].windows == 1
    c0 = ops.launch_counters()
    with torch.no_grad():
        y = lay(x.cuda())
    c1 = ops.launch_counters()
    assert c1["fp32"] > c0["fp32"]
    ref = K.fastkan_layer(x, sd["layernorm.weight"], sd["layernorm.bias"], sd["rbf.grid"], sd["spline_linear.weight"],
                          sd["base_linear.weight"], sd["base_linear.bias"])
    assert K.rel_err(y.cpu(), ref) <= TOL
