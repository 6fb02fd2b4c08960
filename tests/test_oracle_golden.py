"""Pins the oracle (oracle/kagnn_oracle.py) to numbers computed by the reference itself
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference)."""
import pytest
import torch

from oracle import kagnn_oracle as K
from tests.helpers import golden_names, load_golden, oracle_run


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_output(name):
    meta, inputs, sd, y_ref = load_golden(name)
    y = oracle_run(meta, inputs, sd)
    assert y.shape == y_ref.shape
    # same ops in the same order as the reference on the same CPU -> essentially bitwise
    assert K.rel_err(y, y_ref) <= 2e-6, name


@pytest.mark.parametrize("name", golden_names("kanlinear_"))
def test_oracle_bases_match_reference(name):
    meta, inputs, sd, _ = load_golden(name)
    b = K.bspline_bases(inputs["x"], sd["grid"], meta["k"])
    assert torch.equal(b, sd["__bases"])
    # at most k+1 non-zeros per (row, feature); rows 1 and 2 are out of range -> all zero
    assert int((b != 0).sum(-1).max()) <= meta["k"] + 1
    assert float(b[1].abs().max()) == 0.0 and float(b[2].abs().max()) == 0.0


@pytest.mark.parametrize("name", golden_names())
def test_fp64_oracle_agrees_with_fp32_reference(name):
    meta, inputs, sd, y_ref = load_golden(name)
    y64 = oracle_run(meta, inputs, sd, torch.float64)
    assert K.rel_err(y64, y_ref) <= 5e-5, name


def test_gcn_norm_matches_dense_identity():
    """The only in-tree statement of the GCN normalisation is the dense algebra at
    node_classification_clean/time_model.py:70-80: D^-1/2 (A'+I) D^-1/2."""
    g = torch.Generator().manual_seed(7)
    n = 23
    ei = torch.randint(0, n - 2, (2, 90), generator=g)
    ei[1, :5] = ei[0, :5]            # self loops
    ei[:, 5:9] = ei[:, 9:13]         # duplicates
    h = torch.randn(n, 6, generator=g, dtype=torch.float64)
    bias = torch.randn(6, generator=g, dtype=torch.float64)
    out = K.gcn_conv(h, ei, lambda t: t, bias)
    dense = K.dense_gcn_matrix(ei, n) @ h + bias
    assert torch.allclose(out, dense, rtol=1e-12, atol=1e-12)


def test_gin_gine_pool_small_known_answers():
    x = torch.tensor([[1.0, -2.0], [3.0, 0.5], [-1.0, 4.0]])
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 2]])       # 0->1, 1->0, 1->2, 2->2
    ident = lambda t: t
    out = K.gin_conv(x, ei, ident, eps=0.5)
    exp = 1.5 * x + torch.stack([x[1], x[0], x[1] + x[2]])
    assert torch.allclose(out, exp)
    ea = torch.tensor([[0.0, 0.0], [10.0, 10.0], [-10.0, -10.0], [0.0, 0.0]])
    out = K.gine_conv(x, ei, ea, ident)
    exp = x + torch.stack([(x[1] + 10).relu(), (x[0]).relu(), (x[1] - 10).relu() + x[2].relu()])
    assert torch.allclose(out, exp)
    batch = torch.tensor([0, 0, 2])
    assert torch.allclose(K.global_add_pool(x, batch), torch.stack([x[0] + x[1], torch.zeros(2), x[2]]))
    assert torch.allclose(K.global_mean_pool(x, batch), torch.stack([(x[0] + x[1]) / 2, torch.zeros(2), x[2]]))


# ---- gradients: the oracle differentiated by autograd == the reference's own modules differentiated by autograd ----------
from tests.helpers import grad_err, grad_golden_names, grad_scale, load_grad_golden, oracle_grads


@pytest.mark.parametrize("name", grad_golden_names())
def test_oracle_gradients_match_reference(name):
    meta, inputs, sd, y_ref, g_ref = load_grad_golden(name)
    y, g = oracle_grads(meta, inputs, sd)
    assert K.rel_err(y.detach(), y_ref) <= 2e-6
    assert set(g) == set(g_ref), (sorted(set(g) ^ set(g_ref)))
    scale = grad_scale(g_ref)
    for k in g_ref:
        assert grad_err(g[k].detach(), g_ref[k], scale) <= 1e-5, (name, k)
