"""The tcgen05 gradient kernels of the B-spline KAN layer (kagnn_b200/csrc/backward_tc.cu) against the fp32 CUDA-core kernels of
the same library (kagnn_set_backward_path(1)) and against torch autograd through the oracle's restatement of KANLinear.forward
(ekan.py:154-162).  The tensor-core path forms every product from bf16 hi/lo pairs (three products, fp32 accumulate), so it must
agree with the fp32 kernels to ~1e-5 relative; the tolerance against autograd is the gradient tolerance of the suite (1e-3)."""
import pytest
import torch

from oracle import kagnn_oracle as K

pytestmark = pytest.mark.gpu
TOL_PATHS = 5e-5
TOL_AUTOGRAD = 1e-3

SHAPES = [  # (rows, in, out, G, k)
    (1000, 128, 64, 5, 3),
    (4096, 64, 64, 5, 3),
    (777, 320, 40, 5, 3),          # the read-out: out not a multiple of 16, rows not a multiple of 128
    (300, 33, 7, 5, 3),
    (512, 20, 8, 3, 2),            # S = 5: fewer slots than the 8 of a unit
    (260, 16, 16, 4, 1),
    (2000, 64, 128, 5, 3),
    (640, 48, 256, 5, 3),          # wide output (BASELINE C5 width)
    (1500, 24, 32, 8, 3),          # S = 11: two windows of eight slots (ekan.KANLinear._windowed_spec)
    (900, 20, 16, 13, 3),          # S = 16: two full windows
    (700, 12, 24, 20, 2),          # S = 22: three windows
]


def _layer(in_f, out_f, G, k, seed):
    import kagnn_b200 as kb
    torch.manual_seed(seed)
    lay = kb.KANLinear(in_f, out_f, grid_size=G, spline_order=k)
    with torch.no_grad():
        lay.spline_scaler.uniform_(0.5, 1.5)
    return lay


@pytest.mark.parametrize("rows,in_f,out_f,G,k", SHAPES)
def test_tc_gradients_match_fp32_kernels_and_autograd(rows, in_f, out_f, G, k):
    from kagnn_b200 import ops
    lay = _layer(in_f, out_f, G, k, rows + in_f)
    sd = {kk: v.detach().clone() for kk, v in lay.state_dict().items()}
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, in_f, generator=g) * 0.7
    x[0, 0], x[1 % rows, 0] = 3.5, -3.5                             # outside the knot range: only the base branch has a gradient
    dy = torch.randn(rows, out_f, generator=g)
    lay = lay.cuda()
    spec = lay.kernel_specs()[0]
    xd, dyd = x.cuda(), dy.cuda()
    ops.set_backward_path(0)
    dx_tc, dp_tc = ops.kan_bwd_input(spec, xd, dyd), ops.kan_bwd_weights(spec, xd, dyd)
    ops.set_backward_path(2)                                         # d input without the packed-weight operand
    try:
        dx_tc2, dp_tc2 = ops.kan_bwd_input(spec, xd, dyd), ops.kan_bwd_weights(spec, xd, dyd)
        ops.set_backward_path(3)                                     # d weights: one feature block per CTA; d input: no look-ahead
        dp_tc3, dx_tc3 = ops.kan_bwd_weights(spec, xd, dyd), ops.kan_bwd_input(spec, xd, dyd)
        ops.set_backward_path(1)
        dx_32, dp_32 = ops.kan_bwd_input(spec, xd, dyd), ops.kan_bwd_weights(spec, xd, dyd)
    finally:
        ops.set_backward_path(0)
    assert K.rel_err(dx_tc.cpu(), dx_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dx_tc2.cpu(), dx_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dx_tc3.cpu(), dx_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dp_tc.cpu(), dp_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dp_tc2.cpu(), dp_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dp_tc3.cpu(), dp_32.cpu()) <= TOL_PATHS
    # autograd through the oracle
    xr = x.clone().requires_grad_(True)
    params = {kk: v.clone().requires_grad_(kk != "grid") for kk, v in sd.items()}
    y = K.kan_linear(xr, params["base_weight"], params["spline_weight"], params.get("spline_scaler"), params["grid"], k)
    y.backward(dy)
    assert K.rel_err(dx_tc.cpu(), xr.grad) <= TOL_AUTOGRAD
    if spec.windows > 1:                                             # G + k > 8: gradient of the virtual (windowed) packing
        d_base, d_spline, d_scaler = ops.kan_unpack_windowed_grads(dp_tc, spec, G + k)
    else:
        d_base, d_spline, d_scaler = ops.kan_unpack_weight_grads(dp_tc, lay.spline_weight, lay.spline_scaler)
    assert K.rel_err(d_base.cpu(), params["base_weight"].grad) <= TOL_AUTOGRAD
    assert K.rel_err(d_spline.cpu(), params["spline_weight"].grad) <= TOL_AUTOGRAD
    assert K.rel_err(d_scaler.cpu(), params["spline_scaler"].grad) <= TOL_AUTOGRAD


def test_tc_gradients_full_size_strided():
    """arxiv-sized rows, operands that are column slices of wider matrices (leading dimensions != widths)."""
    from kagnn_b200 import ops
    lay = _layer(64, 64, 5, 3, 7).cuda()
    spec = lay.kernel_specs()[0]
    n = 169_343
    g = torch.Generator(device="cuda").manual_seed(3)
    xb = torch.randn(n, 320, generator=g, device="cuda") * 0.5
    dyb = torch.randn(n, 96, generator=g, device="cuda")
    x, dy = xb[:, 128:192], dyb[:, 16:80]
    ops.set_backward_path(0)
    dx_tc, dp_tc = ops.kan_bwd_input(spec, x, dy), ops.kan_bwd_weights(spec, x, dy)
    ops.set_backward_path(1)
    try:
        dx_32, dp_32 = ops.kan_bwd_input(spec, x, dy), ops.kan_bwd_weights(spec, x, dy)
    finally:
        ops.set_backward_path(0)
    assert K.rel_err(dx_tc.cpu(), dx_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dp_tc.cpu(), dp_32.cpu()) <= TOL_PATHS


RBF_SHAPES = [  # (rows, in, out, num_grids, layernorm)
    (1000, 64, 64, 8, True),
    (777, 256, 256, 8, True),      # BASELINE C5 width
    (300, 33, 7, 4, True),
    (2048, 7, 256, 8, True),
    (640, 48, 40, 5, False),       # use_layernorm=False: z = x, one output matrix
]


@pytest.mark.parametrize("rows,in_f,out_f,G,ln", RBF_SHAPES)
def test_tc_gradients_of_the_fastkan_layer_match_the_fp32_kernels(rows, in_f, out_f, G, ln):
    """FastKANLayer (fastkan.py:76-85): dz / dx_base / dP from the tensor-core kernels against the fp32 ones."""
    import kagnn_b200 as kb
    from kagnn_b200 import ops
    torch.manual_seed(rows + in_f)
    lay = kb.FastKANLayer(in_f, out_f, num_grids=G, use_layernorm=ln).cuda()
    if ln:
        with torch.no_grad():
            lay.layernorm.weight.uniform_(0.5, 1.5)
            lay.layernorm.bias.normal_(0, 0.2)
    spec = lay.kernel_specs()[0]
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, in_f, generator=g, device="cuda") * 0.9
    dy = torch.randn(rows, out_f, generator=g, device="cuda")
    stats = ops.layernorm_stats(x) if ln else None
    ops.set_backward_path(0)
    dz_tc, dxb_tc = ops.rbf_bwd_input(spec, x, stats, dy)
    dp_tc = ops.rbf_bwd_weights(spec, x, stats, dy)
    ops.set_backward_path(2)
    try:
        dz_tc2, _ = ops.rbf_bwd_input(spec, x, stats, dy)
        dp_tc2 = ops.rbf_bwd_weights(spec, x, stats, dy)
        ops.set_backward_path(3)
        dp_tc3 = ops.rbf_bwd_weights(spec, x, stats, dy)
        dz_tc3, _ = ops.rbf_bwd_input(spec, x, stats, dy)
        ops.set_backward_path(1)
        dz_32, dxb_32 = ops.rbf_bwd_input(spec, x, stats, dy)
        dp_32 = ops.rbf_bwd_weights(spec, x, stats, dy)
    finally:
        ops.set_backward_path(0)
    assert K.rel_err(dz_tc.cpu(), dz_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dz_tc2.cpu(), dz_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dz_tc3.cpu(), dz_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dp_tc.cpu(), dp_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dp_tc2.cpu(), dp_32.cpu()) <= TOL_PATHS
    assert K.rel_err(dp_tc3.cpu(), dp_32.cpu()) <= TOL_PATHS
    if ln:
        assert K.rel_err(dxb_tc.cpu(), dxb_32.cpu()) <= TOL_PATHS
    else:
        assert dxb_tc is None and dxb_32 is None
