"""CPU dry run of the autograd wiring (kagnn_b200/autograd.py and the modules' training paths): forward launches replaced
by torch-CPU stand-ins, backward launches served by the host-check build of backward.cu (tests/emul/cpu_double.py -- test
infrastructure).  Checked against gradients computed by the reference's own modules (tests/golden/grad/).  The same
scenarios run against the real library on the B200 in tests/test_gpu_train_backward.py."""
import pytest
import torch

from oracle import kagnn_oracle as K
from tests.emul.cpu_double import cpu_double
from tests.helpers import build_product_model, grad_err, grad_golden_names, grad_scale, load_grad_golden

TOL = 1e-4


def run_product_grads(meta, inputs, sd, device):
    model = build_product_model(meta, sd, device=device)
    model.train(bool(meta.get("training", True)))
    inp = {k: v.to(device) for k, v in inputs.items()}
    x = inp["x"].clone()
    if x.is_floating_point():
        x.requires_grad_(True)
    if meta["kind"] in ("kan_linear", "kan_chain", "fastkan_chain"):
        y = model(x)
    elif meta["kind"] == "node":
        y = model(x, inp["edge_index"])
    else:
        y = model(K.Batch(x, inp["edge_index"], inp["batch"], inp.get("edge_attr")))
    y.backward(inp["dy"])
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    grads["__x"] = x.grad if x.is_floating_point() else torch.zeros(1, device=device)     # integer atom codes: no input gradient
    return y.detach(), grads, model


def check_against_fixture(name, device, grad_tol=TOL):
    meta, inputs, sd, y_ref, g_ref = load_grad_golden(name)
    y, g, model = run_product_grads(meta, inputs, sd, device)
    assert K.rel_err(y.cpu(), y_ref) <= TOL
    assert set(g) == set(g_ref), sorted(set(g) ^ set(g_ref))
    scale = grad_scale(g_ref)
    for k in g_ref:
        assert g[k].shape == g_ref[k].shape, k
        assert grad_err(g[k].cpu(), g_ref[k], scale) <= grad_tol, (name, k)
    return model


@pytest.mark.parametrize("name", grad_golden_names())
def test_module_gradients_match_reference_dry_run(name):
    with cpu_double():
        check_against_fixture(name, "cpu")


@pytest.mark.parametrize("name", ["grad_kanlinear_g8_k1_5x9", "grad_fastkan_5_6_7_g32", "grad_kanlinear_g5_k3_33x7", "grad_nc_gkan_gin_skip1",
                                  "grad_nc_gfastkan_gcn"])
def test_slot_window_wiring_matches_reference_dry_run(name):
    """Layers with more than eight coefficients (KANLinear grid 8, order 1) or centres (FastKAN 5-6-7 with 32 grids) through the
    slot-window wiring -- virtual layer over copies of the input, dx / LayerNorm gradients summed over the copies, weight gradients
    folded back -- against the gradients the reference's own modules produced; the other fixtures must be untouched by it."""
    from kagnn_b200 import ops
    with cpu_double(windows=True):
        model = check_against_fixture(name, "cpu")
        specs = [m.kernel_spec() for m in model.modules() if hasattr(m, "kernel_spec")]
        if "g8_k1" in name:
            assert [sp.windows for sp in specs] == [2]
        if "g32" in name:
            assert [sp.windows for sp in specs] == [4, 4]
        if "g5_k3" in name or "nc_" in name:
            assert all(sp.windows == 1 for sp in specs)
    assert not ops.rbf_windows_enabled()


def test_eval_mode_with_autograd_enabled_takes_the_inference_plan():
    """graph_classification_utils.py:57-72 evaluates with model.eval() but without torch.no_grad()."""
    import kagnn_b200 as kb
    with cpu_double():
        m = kb.GKAN_Nodes("gin", 2, 6, 8, 3).eval()
        x = torch.randn(20, 6)
        ei = torch.randint(0, 20, (2, 50))
        y = m(x, ei)
        assert not y.requires_grad
        with torch.no_grad():
            assert torch.equal(y, m(x, ei))


def test_modules_without_backward_raise_under_autograd():
    import kagnn_b200 as kb
    with cpu_double():
        x = torch.randn(10, 4)
        gin = kb.GINConv(kb.make_kan(4, 4, 4, 1, 5, 3), train_eps=True)
        with pytest.raises(NotImplementedError):
            gin(x, torch.randint(0, 10, (2, 20)))


@pytest.mark.parametrize("family,args", [("KAGCN", (1, 2, 16, 5, 3, 1, 0.0, True)), ("FASTKAGCN", (1, 2, 16, 4, 1, 0.0, True))])
def test_graph_regression_gcn_models_train(family, args):
    """graph_regression KAGCN / FASTKAGCN (OGB atom encoder -> GCN layers -> add pool -> read-out, one output column) with the
    loss ``y.sum()``, whose cotangent is an expanded stride-0 tensor."""
    from kagnn_b200 import models_regr
    with cpu_double():
        g = torch.Generator().manual_seed(0)
        n = 30
        x = torch.randint(0, 28, (n, 1), generator=g)
        ei = torch.randint(0, n, (2, 80), generator=g)
        batch = torch.sort(torch.randint(0, 4, (n,), generator=g))[0]
        torch.manual_seed(1)
        m = getattr(models_regr, family)(*args).train()
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        y = m(K.Batch(x, ei, batch))
        y.sum().backward()
        ref_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith("grid") else v) for k, v in sd.items()}
        yr = K.gr_kagcn_forward(ref_sd, K.Batch(x, ei, batch))
        yr.sum().backward()
        assert K.rel_err(y, yr) <= 1e-5
        g_ref = {k: v.grad for k, v in ref_sd.items() if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
        scale = grad_scale(g_ref)
        got = {nm: p.grad for nm, p in m.named_parameters() if p.grad is not None}
        assert set(got) == set(g_ref)
        for k in g_ref:
            assert grad_err(got[k], g_ref[k], scale) <= TOL, k
