"""CPU check of the slot-window construction (kagnn_b200/ekan.py: windowed_weights, kagnn_b200/fastkan.py: windowed_layernorm):
a layer with more than eight coefficients / centres per input, evaluated by the ORACLE as the virtual eight-slot layer over shifted
(or duplicated) copies of the input, equals the oracle's evaluation of the layer itself.  This is the identity the GPU path relies on
(B_{8w+j}(x) = B_j(x - 8wh), phi_{8w+j}(z) = phi_j(z - 8w step)); tests/test_gpu_windows.py checks the kernels."""
import pytest
import torch

from oracle import kagnn_oracle as K


@pytest.mark.parametrize("G,k", [(8, 3), (6, 3), (13, 3), (20, 2), (30, 1), (9, 1)])
def test_bspline_layer_equals_its_windowed_form(G, k):
    from kagnn_b200.ekan import windowed_weights
    torch.manual_seed(G * 10 + k)
    in_f, out_f, S = 7, 5, G + k
    h = 2.0 / G
    knots = (torch.arange(-k, G + k + 1, dtype=torch.float64) * h - 1.0)
    grid = knots.expand(in_f, -1).contiguous()
    base = torch.randn(out_f, in_f, dtype=torch.float64)
    spline = torch.randn(out_f, in_f, S, dtype=torch.float64)
    scaler = torch.rand(out_f, in_f, dtype=torch.float64) + 0.5
    x = torch.randn(400, in_f, dtype=torch.float64) * 0.8
    x[0, 0], x[1, 0], x[2, 0], x[3, 0] = 1.0, -1.0, 4.0, -4.0       # ends of the grid range, outside the knots
    ref = K.kan_linear(x, base, spline, scaler, grid, k)
    w = (S + 7) // 8
    vb, vs, vsc = windowed_weights(base.float(), spline.float(), scaler.float(), w)
    assert vs.shape == (out_f, w * in_f, 8) and vb.shape == (out_f, w * in_f) and vsc.shape == (out_f, w * in_f)
    # values survive the float32 staging of the helper up to rounding: redo the rearrangement in fp64 for an exact comparison
    sp = torch.zeros(out_f, in_f, 8 * w, dtype=torch.float64)
    sp[:, :, :S] = spline
    vs64 = sp.view(out_f, in_f, w, 8).permute(0, 2, 1, 3).reshape(out_f, w * in_f, 8)
    assert torch.allclose(vs.double(), vs64, atol=1e-6)
    vb64 = torch.zeros(out_f, w * in_f, dtype=torch.float64)
    vb64[:, :in_f] = base
    xv = torch.cat([x - 8.0 * h * c for c in range(w)], dim=1)
    gv = knots[: (8 - k) + 2 * k + 1].expand(w * in_f, -1).contiguous()                # grid_size 8 - k, same origin and spacing
    got = K.kan_linear(xv, vb64, vs64, scaler.repeat(1, w), gv, k)
    assert K.rel_err(got, ref) <= 1e-12


@pytest.mark.parametrize("G,ln", [(12, True), (16, True), (32, True), (20, False), (9, True)])
def test_fastkan_layer_equals_its_windowed_form(G, ln):
    from kagnn_b200.ekan import windowed_weights
    from kagnn_b200.fastkan import windowed_layernorm
    torch.manual_seed(G)
    in_f, out_f = 6, 4
    grid = torch.linspace(-2.0, 2.0, G, dtype=torch.float64)
    step = 4.0 / (G - 1)
    spline = torch.randn(out_f, in_f * G, dtype=torch.float64)
    base_w, base_b = torch.randn(out_f, in_f, dtype=torch.float64), torch.randn(out_f, dtype=torch.float64)
    ln_w = torch.rand(in_f, dtype=torch.float64) + 0.5 if ln else None
    ln_b = torch.randn(in_f, dtype=torch.float64) * 0.2 if ln else None
    x = torch.randn(300, in_f, dtype=torch.float64)
    ref = K.fastkan_layer(x, ln_w, ln_b, grid, spline, base_w, base_b, denominator=step)
    w = (G + 7) // 8
    sp = torch.zeros(out_f, in_f, 8 * w, dtype=torch.float64)
    sp[:, :, :G] = spline.view(out_f, in_f, G)
    vs = sp.view(out_f, in_f, w, 8).permute(0, 2, 1, 3).reshape(out_f, w * in_f * 8)
    vb = torch.zeros(out_f, w * in_f, dtype=torch.float64)
    vb[:, :in_f] = base_w
    vb32, vs32, _ = windowed_weights(base_w.float(), spline.view(out_f, in_f, G).float(), None, w)
    assert torch.allclose(vs32.double().reshape(out_f, -1), vs, atol=1e-6) and torch.allclose(vb32.double(), vb, atol=1e-6)
    if ln:
        vw32, vbias32 = windowed_layernorm(ln_w.float(), ln_b.float(), w, 8.0 * step)
        vw = ln_w.repeat(w)
        vbias = torch.cat([ln_b - 8.0 * step * c for c in range(w)])
        assert torch.allclose(vw32.double(), vw, atol=1e-6) and torch.allclose(vbias32.double(), vbias, atol=1e-5)
        xv = x.repeat(1, w)                                                            # exact duplicates: same row statistics
        got = K.fastkan_layer(xv, vw, vbias, grid[:8], vs, vb, base_b, denominator=step)
    else:
        xv = torch.cat([x - 8.0 * step * c for c in range(w)], dim=1)
        # the base branch sees the shifted copies too, but only copy 0 (unshifted) has base weights
        got = K.fastkan_layer(xv, None, None, grid[:8], vs, vb, base_b, denominator=step)
    assert K.rel_err(got, ref) <= 1e-12
