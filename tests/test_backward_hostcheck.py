"""CPU check of the backward kernels' index arithmetic: kagnn_b200/csrc/backward.cu compiled as serial host code
(tests/emul/build_emul.py -- test infrastructure, never loaded by the product) against torch autograd through the oracle
(the reference's own forward restated, node_classification_clean/ekan.py:79-162).  The real CUDA build of the same source is
checked on the B200 by tests/test_gpu_train_backward.py."""
import ctypes as C

import pytest
import torch

from oracle import kagnn_oracle as K
from tests.emul import build_emul

TOL = 1e-4


class Layer(C.Structure):                      # struct KagnnKanLayer
    _fields_ = [("basis", C.c_int32), ("in_features", C.c_int32), ("out_features", C.c_int32), ("grid_size", C.c_int32),
                ("spline_order", C.c_int32), ("t0", C.c_float), ("h", C.c_float), ("inv_denominator", C.c_float),
                ("packed_w", C.c_void_p), ("base_bias", C.c_void_p), ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p),
                ("packed_w_tc", C.c_void_p), ("ln_stats", C.c_void_p)]


@pytest.fixture(scope="module")
def lib():
    return C.CDLL(build_emul.build())


def p(t):
    return C.c_void_p(t.data_ptr())


def pack(base_w, spline_w, scaler):
    """[in][S+1][out_pad4] (the layout of kagnn_pack_kan_weights, kagnn_b200/csrc/api.cu)."""
    out_f, in_f, S = spline_w.shape
    out_pad = (out_f + 3) // 4 * 4
    w = torch.zeros(in_f, S + 1, out_pad)
    w[:, :S, :out_f] = (spline_w * scaler.unsqueeze(-1)).permute(1, 2, 0)
    w[:, S, :out_f] = base_w.t()
    return w.contiguous()


@pytest.mark.parametrize("G,k,fin,fout,n", [(5, 3, 16, 8, 300), (4, 3, 7, 5, 129), (8, 1, 3, 2, 64), (2, 2, 5, 9, 77),
                                            (16, 4, 4, 3, 50), (5, 3, 1, 1, 1), (5, 3, 6, 300, 40)])
def test_kan_layer_gradients(lib, G, k, fin, fout, n):
    torch.manual_seed(G * 100 + k)
    S = G + k
    grid = K.uniform_knots(fin, G, k)
    x = (torch.randn(n, fin) * 0.9).requires_grad_(True)
    with torch.no_grad():
        x[:: 7, 0] = grid[0, k + 1]                          # exactly on a knot
        if n > 3:
            x[1, :] = 50.0                                   # outside the knot range
            x[2, :] = -50.0
    bw = (torch.randn(fout, fin) * 0.3).requires_grad_(True)
    sw = (torch.randn(fout, fin, S) * 0.3).requires_grad_(True)
    sc = (torch.randn(fout, fin) * 0.5 + 1.0).requires_grad_(True)
    dy = torch.randn(n, fout)
    K.kan_linear(x, bw, sw, sc, grid, k).backward(dy)

    packed = pack(bw.detach(), sw.detach(), sc.detach())
    lay = Layer(0, fin, fout, G, k, float(grid[0, 0]), float(grid[0, 1] - grid[0, 0]), 0.0, packed.data_ptr(), None, None, None,
                None, None)
    # padded leading dimensions on purpose
    xs = torch.zeros(n, fin + 3)
    xs[:, :fin] = x.detach()
    dys = torch.zeros(n, fout + 2)
    dys[:, :fout] = dy
    dx = torch.full((n, fin + 1), 7.0)
    assert lib.kagnn_kan_bwd_input(C.byref(lay), p(xs), C.c_int64(fin + 3), p(dys), C.c_int64(fout + 2), C.c_int64(n), p(dx),
                                   C.c_int64(fin + 1), None) == 0
    assert K.rel_err(dx[:, :fin], x.grad) <= TOL
    assert torch.all(dx[:, fin] == 7.0)

    dP = torch.full_like(packed, 3.0)
    assert lib.kagnn_kan_bwd_weights(C.byref(lay), p(xs), C.c_int64(fin + 3), p(dys), C.c_int64(fout + 2), C.c_int64(n), p(dP),
                                     None) == 0
    d_base, d_spline, d_sc = torch.empty(fout, fin), torch.empty(fout, fin, S), torch.empty(fout, fin)
    assert lib.kagnn_kan_unpack_weight_grads(p(dP), p(sw.detach().contiguous()), p(sc.detach().contiguous()), fin, fout, S,
                                             p(d_base), p(d_spline), p(d_sc), None) == 0
    assert K.rel_err(d_base, bw.grad) <= TOL
    assert K.rel_err(d_spline, sw.grad) <= TOL
    assert K.rel_err(d_sc, sc.grad) <= TOL


def test_kan_weight_gradient_many_row_slabs(lib):
    """More rows than one slab: the slabs' partial sums meet through the (emulated) atomics."""
    torch.manual_seed(1)
    G, k, fin, fout, n = 5, 3, 3, 4, 5000
    grid = K.uniform_knots(fin, G, k)
    x = torch.randn(n, fin) * 0.7
    bw = torch.randn(fout, fin, requires_grad=True)
    sw = torch.randn(fout, fin, G + k, requires_grad=True)
    sc = torch.ones(fout, fin, requires_grad=True)
    dy = torch.randn(n, fout)
    K.kan_linear(x, bw, sw, sc, grid, k).backward(dy)
    packed = pack(bw.detach(), sw.detach(), sc.detach())
    lay = Layer(0, fin, fout, G, k, float(grid[0, 0]), float(grid[0, 1] - grid[0, 0]), 0.0, packed.data_ptr(), None, None, None,
                None, None)
    dP = torch.empty_like(packed)
    assert lib.kagnn_kan_bwd_weights(C.byref(lay), p(x), C.c_int64(fin), p(dy), C.c_int64(fout), C.c_int64(n), p(dP), None) == 0
    d_base, d_spline = torch.empty(fout, fin), torch.empty(fout, fin, G + k)
    assert lib.kagnn_kan_unpack_weight_grads(p(dP), p(sw.detach()), None, fin, fout, G + k, p(d_base), p(d_spline), None, None) == 0
    assert K.rel_err(d_base, bw.grad) <= TOL and K.rel_err(d_spline, sw.grad) <= TOL


@pytest.mark.parametrize("rows,cols,affine", [(2, 3, True), (1000, 20, True), (333, 130, False)])
def test_batchnorm_backward(lib, rows, cols, affine):
    torch.manual_seed(rows)
    x = (torch.randn(rows, cols) * 2 + 0.5).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(cols, affine=affine).train()
    if affine:
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-1, 1)
    dy = torch.randn(rows, cols)
    # reference in fp64 (with two rows dx is a pure cancellation residue: torch's own fp32 backward is noise there)
    bn64 = torch.nn.BatchNorm1d(cols, affine=affine).double().train()
    if affine:
        bn64.load_state_dict({k_: v.double() if v.is_floating_point() else v for k_, v in bn.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    bn64(x64).backward(dy.double())
    x.grad = x64.grad.float()
    if affine:
        bn.weight.grad, bn.bias.grad = bn64.weight.grad.float(), bn64.bias.grad.float()
    dx = torch.empty(rows, cols)
    dw, db = torch.empty(cols), torch.empty(cols)
    nbytes = lib.kagnn_batchnorm_bwd_workspace(cols)
    assert nbytes == cols * 32
    ws = torch.empty(cols * 4, dtype=torch.float64)
    lib.kagnn_batchnorm_train_bwd.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_float,
                                              C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    rc = lib.kagnn_batchnorm_train_bwd(p(x.detach()), cols, p(dy), cols, rows, cols, p(bn.weight.detach()) if affine else None,
                                       1e-5, p(dx), cols, p(dw), p(db), p(ws), cols * 32, None)
    assert rc == 0
    assert K.rel_err(dx, x.grad) <= TOL
    if affine:
        assert K.rel_err(dw, bn.weight.grad) <= TOL and K.rel_err(db, bn.bias.grad) <= TOL


def test_small_epilogue_backwards(lib):
    torch.manual_seed(5)
    # column sums
    x = torch.randn(700, 9)
    out = torch.full((9,), 5.0)
    assert lib.kagnn_column_sums(p(x), C.c_int64(9), C.c_int64(700), 9, p(out), None) == 0
    assert K.rel_err(out, x.sum(0)) <= 1e-5
    # log_softmax
    z = torch.randn(37, 6, requires_grad=True)
    y = torch.log_softmax(z, dim=1)
    dy = torch.randn(37, 6)
    y.backward(dy)
    dz = torch.empty(37, 6)
    assert lib.kagnn_log_softmax_bwd(p(y.detach()), C.c_int64(6), p(dy), C.c_int64(6), C.c_int64(37), 6, p(dz), C.c_int64(6), None) == 0
    assert K.rel_err(dz, z.grad) <= 1e-5
    # pooling
    sizes = torch.tensor([3, 1, 5, 2])
    batch = torch.repeat_interleave(torch.arange(4), sizes)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)]).to(torch.int32)
    for mean in (0, 1):
        h = torch.randn(11, 5, requires_grad=True)
        pooled = (K.global_mean_pool if mean else K.global_add_pool)(h, batch, 4)
        dp = torch.randn(4, 5)
        pooled.backward(dp)
        dh = torch.empty(11, 5)
        assert lib.kagnn_segment_pool_bwd(p(dp), C.c_int64(5), p(ptr), p(batch), C.c_int64(11), 5, mean, p(dh), C.c_int64(5), None) == 0
        assert K.rel_err(dh, h.grad) <= 1e-6


@pytest.mark.parametrize("G,fin,fout,n,ln", [(8, 7, 5, 120, True), (4, 16, 3, 77, False), (32, 3, 9, 40, True), (1, 2, 2, 10, False),
                                            (6, 5, 260, 30, True)])
def test_fastkan_layer_gradients(lib, G, fin, fout, n, ln):
    """rbf_bwd_input / rbf_bwd_weights / layernorm_bwd against autograd through the oracle's fastkan_layer (fastkan.py:76-85)."""
    torch.manual_seed(G * 10 + fin)
    grid = torch.linspace(-2.0, 2.0, G) if G > 1 else torch.tensor([-2.0])
    den = 4.0 / (G - 1) if G > 1 else 1.0
    x = (torch.randn(n, fin) * 1.3).requires_grad_(True)
    lw = (1.0 + 0.3 * torch.randn(fin)).requires_grad_(True) if ln else None
    lb = (0.2 * torch.randn(fin)).requires_grad_(True) if ln else None
    sw = (torch.randn(fout, fin * G) * 0.3).requires_grad_(True)
    bw = (torch.randn(fout, fin) * 0.3).requires_grad_(True)
    bb = torch.randn(fout).requires_grad_(True)
    dy = torch.randn(n, fout)
    K.fastkan_layer(x, lw, lb, grid, sw, bw, bb, den).backward(dy)

    packed = pack(bw.detach(), sw.detach().view(fout, fin, G), torch.ones(fout, fin))
    lay = Layer(1, fin, fout, G, 0, -2.0, (4.0 / (G - 1)) if G > 1 else 1.0, 1.0 / den, packed.data_ptr(), None,
                lw.data_ptr() if ln else None, lb.data_ptr() if ln else None, None, None)
    xd = x.detach()
    stats = None
    if ln:
        stats = torch.stack([xd.mean(1), (xd.var(1, unbiased=False) + 1e-5).rsqrt()], dim=1).contiguous()
    dz = torch.empty(n, fin + 2)
    dxb = torch.empty(n, fin + 1) if ln else None
    lib.kagnn_rbf_bwd_input.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    assert lib.kagnn_rbf_bwd_input(C.byref(lay), p(xd), fin, p(stats) if ln else None, p(dy), fout, n, p(dz), fin + 2,
                                   p(dxb) if ln else None, fin + 1 if ln else 0, None) == 0
    if ln:
        dx = torch.empty(n, fin)
        d_lw, d_lb = torch.empty(fin), torch.empty(fin)
        lib.kagnn_layernorm_bwd.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                            C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        assert lib.kagnn_layernorm_bwd(p(xd), fin, p(stats), p(lw.detach()), p(dz), fin + 2, p(dxb), fin + 1, n, fin, p(dx), fin,
                                       p(d_lw), p(d_lb), None) == 0
        assert K.rel_err(d_lw, lw.grad) <= TOL and K.rel_err(d_lb, lb.grad) <= TOL
    else:
        dx = dz[:, :fin]
    assert K.rel_err(dx, x.grad) <= TOL

    dP = torch.full_like(packed, 9.0)
    lib.kagnn_rbf_bwd_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                          C.c_void_p]
    assert lib.kagnn_rbf_bwd_weights(C.byref(lay), p(xd), fin, p(stats) if ln else None, p(dy), fout, n, p(dP), None) == 0
    d_base, d_spline = torch.empty(fout, fin), torch.empty(fout, fin, G)
    assert lib.kagnn_kan_unpack_weight_grads(p(dP), p(sw.detach()), None, fin, fout, G, p(d_base), p(d_spline), None, None) == 0
    assert K.rel_err(d_base, bw.grad) <= TOL
    assert K.rel_err(d_spline.view(fout, fin * G), sw.grad) <= TOL


@pytest.mark.parametrize("n,e,h", [(20, 90, 6), (5, 0, 3), (64, 400, 17)])
def test_gine_aggregation_gradients(lib, n, e, h):
    torch.manual_seed(n + e)
    x = torch.randn(n, h, requires_grad=True)
    ef = torch.randn(e, h, requires_grad=True)
    ei = torch.randint(0, n, (2, e))
    da = torch.randn(n, h)
    K.gine_conv(x, ei, ef, lambda t: t, eps=0.25).backward(da)
    dx = torch.full((n, h + 2), 4.0)
    de = torch.full((max(e, 1), h + 1), 4.0)
    lib.kagnn_gine_bwd.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                   C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    rc = lib.kagnn_gine_bwd(p(x.detach()), h, p(ef.detach()) if e else None, h, p(ei.contiguous()) if e else None, e, n, h, p(da), h,
                            1.25, p(dx), h + 2, p(de) if e else None, h + 1, None)
    assert rc == 0
    assert K.rel_err(dx[:, :h], x.grad) <= 1e-6 and torch.all(dx[:, h:] == 4.0)
    if e:
        assert K.rel_err(de[:, :h], ef.grad) <= 1e-6 and torch.all(de[:, h:] == 4.0)
