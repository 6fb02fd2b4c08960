"""Generate the GRADIENT golden vectors under tests/golden/grad/ -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_grad          (build container only: needs /root/reference)

Same recipe as oracle/make_golden.py, but every case runs the reference's own module in training mode, back-propagates a
seeded cotangent ``dy`` through torch autograd (``y.backward(dy)``) and stores ``x.grad`` and every parameter's ``.grad``
next to the inputs.  These pin the backward path the same way the forward fixtures pin the
forward."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import kagnn_oracle as K
from .make_golden import GC, NC, _load, _randomise, batched_graphs, small_graph

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "grad")


def _save(name, meta, inputs, sd, y, grads):
    arrs = {"meta": np.array(json.dumps(meta))}
    for k, v in inputs.items():
        arrs["in/" + k] = v.detach().cpu().numpy()
    for k, v in sd.items():
        arrs["sd/" + k] = v.detach().cpu().numpy()
    arrs["out/y"] = y.detach().cpu().numpy()
    for k, v in grads.items():
        arrs["grad/" + k] = v.detach().cpu().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(f"{name}: y {tuple(y.shape)}, {len(grads)} gradients, |dx|max={float(grads['__x'].abs().max()):.4f}")


def _backprop(model, fwd, x, dy_gen):
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    xg = x.clone().requires_grad_(True)
    y = fwd(xg)
    dy = torch.randn(y.shape, generator=dy_gen)
    y.backward(dy)
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    grads["__x"] = xg.grad
    return sd0, y, dy, grads


def main() -> None:
    from . import pyg_shim
    pyg_shim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(777)
    gen = torch.Generator().manual_seed(777)
    ekan = _load(os.path.join(NC, "ekan.py"), "ref_nc_ekan")

    for (g, k, fin, fout, n) in [(5, 3, 33, 7, 64), (4, 3, 16, 16, 40), (8, 1, 5, 9, 33), (3, 2, 6, 4, 50), (16, 4, 12, 4, 50)]:
        lay = ekan.KANLinear(fin, fout, grid_size=g, spline_order=k)
        x = torch.randn(n, fin, generator=gen) * 0.8
        knots = lay.grid[0]
        x[1, :] = knots[0] - 1e-3                       # outside the knot range on both sides
        x[2, :] = knots[-1] + 0.5
        x[3, :] = 0.0                                   # on a knot when G is even
        sd0, y, dy, grads = _backprop(lay, lay, x, gen)
        _save(f"grad_kanlinear_g{g}_k{k}_{fin}x{fout}", dict(kind="kan_linear", G=g, k=k), dict(x=x, dy=dy), sd0, y, grads)

    for sizes, g, k in [([7, 16, 5], 5, 3), ([20, 8, 8, 3], 3, 2)]:
        net = ekan.KAN(sizes, grid_size=g, spline_order=k)
        x = torch.randn(48, sizes[0], generator=gen)
        sd0, y, dy, grads = _backprop(net, net, x, gen)
        _save("grad_kan_" + "_".join(map(str, sizes)) + f"_g{g}k{k}", dict(kind="kan_chain", G=g, k=k), dict(x=x, dy=dy), sd0, y, grads)

    ncm = _load(os.path.join(NC, "models.py"), "ref_nc_models")
    n, e, f, c = 70, 260, 19, 5
    ei = small_graph(n, e, gen)
    x = torch.randn(n, f, generator=gen) * 0.7
    for conv in ("gcn", "gin"):
        for skip in (True, False):
            m = ncm.GKAN_Nodes(conv, 2, f, 12, c, skip=skip, grid_size=5, spline_order=3, hidden_layers=2, dropout=0.0).train()
            _randomise(m, gen)
            sd0, y, dy, grads = _backprop(m, lambda t: m(t, ei), x, gen)
            _save(f"grad_nc_gkan_{conv}_skip{int(skip)}", dict(kind="node", conv_type=conv, skip=skip, fast=False, training=True,
                  mp_layers=2, num_features=f, hidden=12, classes=c, G=5, k=3, hidden_layers=2), dict(x=x, edge_index=ei, dy=dy),
                  sd0, y, grads)

    gcm = _load(os.path.join(GC, "models.py"), "ref_gc_models")
    ei, batch, n = batched_graphs(9, gen)
    x = torch.randn(n, 7, generator=gen)
    for name, mk, meta in [
        ("grad_gc_kagin", lambda: gcm.KAGIN(2, 7, 16, 3, 2, 5, 3, 0.0), dict(family="KAGIN", args=[2, 7, 16, 3, 2, 5, 3, 0.0])),
        ("grad_gc_kagcn", lambda: gcm.KAGCN(3, 7, 12, 4, 4, 2, 0.0), dict(family="KAGCN", args=[3, 7, 12, 4, 4, 2, 0.0])),
    ]:
        m = mk().train()
        _randomise(m, gen)
        sd0, y, dy, grads = _backprop(m, lambda t: m(K.Batch(t, ei, batch)), x, gen)
        _save(name, dict(kind="gc", training=True, **meta), dict(x=x, edge_index=ei, batch=batch, dy=dy), sd0, y, grads)


def main_fastkan() -> None:
    """FastKAN cases, own seed (added after the B-spline fixtures were validated on the GPU: those files stay byte-identical)."""
    from . import pyg_shim
    pyg_shim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(4242)
    gen = torch.Generator().manual_seed(4242)
    fastkan = _load(os.path.join(NC, "fastkan.py"), "ref_nc_fastkan")
    for sizes, g in [([7, 32], 8), ([33, 5], 4), ([16, 16, 4], 8), ([5, 6, 7], 32)]:
        net = fastkan.FastKAN(sizes, num_grids=g)
        _randomise(net, gen)
        x = torch.randn(40, sizes[0], generator=gen) * 1.5
        sd0, y, dy, grads = _backprop(net, net, x, gen)
        _save("grad_fastkan_" + "_".join(map(str, sizes)) + f"_g{g}", dict(kind="fastkan_chain", G=g), dict(x=x, dy=dy), sd0, y, grads)

    ncm = _load(os.path.join(NC, "models.py"), "ref_nc_models")
    n, e, f, c = 70, 260, 19, 5
    ei = small_graph(n, e, gen)
    x = torch.randn(n, f, generator=gen) * 0.7
    for conv in ("gcn", "gin"):
        m = ncm.GFASTKAN_Nodes(conv, 2, f, 10, c, skip=True, grid_size=6, hidden_layers=2, dropout=0.0).train()
        _randomise(m, gen)
        sd0, y, dy, grads = _backprop(m, lambda t: m(t, ei), x, gen)
        _save(f"grad_nc_gfastkan_{conv}", dict(kind="node", conv_type=conv, skip=True, fast=True, training=True, mp_layers=2,
              num_features=f, hidden=10, classes=c, G=6, hidden_layers=2), dict(x=x, edge_index=ei, dy=dy), sd0, y, grads)

    gcm = _load(os.path.join(GC, "models.py"), "ref_gc_models")
    ei, batch, n = batched_graphs(9, gen)
    x = torch.randn(n, 7, generator=gen)
    for name, mk, meta in [
        ("grad_gc_fastkagin", lambda: gcm.FASTKAGIN(2, 7, 16, 2, 2, 8, 0.0), dict(family="FASTKAGIN", args=[2, 7, 16, 2, 2, 8, 0.0])),
        ("grad_gc_fastkagcn", lambda: gcm.FASTKAGCN(2, 7, 12, 3, 5, 0.0), dict(family="FASTKAGCN", args=[2, 7, 12, 3, 5, 0.0])),
    ]:
        m = mk().train()
        _randomise(m, gen)
        sd0, y, dy, grads = _backprop(m, lambda t: m(K.Batch(t, ei, batch)), x, gen)
        _save(name, dict(kind="gc", training=True, **meta), dict(x=x, edge_index=ei, batch=batch, dy=dy), sd0, y, grads)


def main_gine() -> None:
    """graph_regression GINE models (ZINC-style OGB encoders and QM9-style linear encoders), own seed."""
    from . import pyg_shim
    from .make_golden import GR
    pyg_shim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(999)
    gen = torch.Generator().manual_seed(999)
    grm = _load(os.path.join(GR, "models.py"), "ref_gr_models", extra_path=(NC,))
    ei, batch, n = batched_graphs(8, gen)
    xz = torch.randint(0, 28, (n, 1), generator=gen)
    ea = torch.randint(1, 4, (ei.size(1),), generator=gen)
    xq = torch.randn(n, 11, generator=gen)
    eq = torch.randn(ei.size(1), 4, generator=gen)
    for name, mk, args, x, e_attr in [
        ("grad_gr_kagin", grm.KAGIN, [1, 1, 3, 16, 2, 5, 3, 1, 0.0, True], xz, ea),
        ("grad_gr_fastkagin", grm.FASTKAGIN, [1, 1, 2, 16, 2, 6, 1, 0.0, True], xz, ea),
        ("grad_gr_kagin_linear_enc", grm.KAGIN, [11, 4, 2, 16, 2, 4, 3, 3, 0.0, False], xq, eq),
    ]:
        m = mk(*args).train()
        _randomise(m, gen)
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        xin = x.clone().requires_grad_(True) if x.is_floating_point() else x
        y = m(K.Batch(xin, ei, batch, e_attr))
        dy = torch.randn(y.shape, generator=gen)
        y.backward(dy)
        grads = {nm: p.grad for nm, p in m.named_parameters() if p.grad is not None}
        grads["__x"] = xin.grad if x.is_floating_point() else torch.zeros(1)
        fam = "KAGIN" if mk is grm.KAGIN else "FASTKAGIN"
        _save(name, dict(kind="gr", family=fam, args=args, training=True), dict(x=x, edge_index=ei, batch=batch, edge_attr=e_attr, dy=dy),
              sd0, y, grads)


def main_gat() -> None:
    """GAT flavour (node and graph-classification models of the reference, GATConv from oracle/pyg_shim.py), own seed."""
    from . import pyg_shim
    pyg_shim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(2024)
    gen = torch.Generator().manual_seed(2024)
    ncm = _load(os.path.join(NC, "models.py"), "ref_nc_models_gat_grad")
    n, e, f, c = 60, 200, 13, 4
    ei = small_graph(n, e, gen)
    x = torch.randn(n, f, generator=gen) * 0.7
    m = ncm.GKAN_Nodes("gat", 2, f, 8, c, skip=True, grid_size=5, spline_order=3, dropout=0.0, heads=3).train()
    _randomise(m, gen)
    sd0, y, dy, grads = _backprop(m, lambda t: m(t, ei), x, gen)
    _save("grad_nc_gkan_gat", dict(kind="node", conv_type="gat", skip=True, fast=False, training=True, mp_layers=2, num_features=f, hidden=8,
          classes=c, G=5, k=3, hidden_layers=2, heads=3), dict(x=x, edge_index=ei, dy=dy), sd0, y, grads)
    m = ncm.GFASTKAN_Nodes("gat", 2, f, 6, c, skip=True, grid_size=6, dropout=0.0, heads=2).train()
    _randomise(m, gen)
    sd0, y, dy, grads = _backprop(m, lambda t: m(t, ei), x, gen)
    _save("grad_nc_gfastkan_gat", dict(kind="node", conv_type="gat", skip=True, fast=True, training=True, mp_layers=2, num_features=f,
          hidden=6, classes=c, G=6, hidden_layers=2, heads=2), dict(x=x, edge_index=ei, dy=dy), sd0, y, grads)
    gcm = _load(os.path.join(GC, "models.py"), "ref_gc_models_gat_grad")
    ei, batch, n = batched_graphs(8, gen)
    x = torch.randn(n, 7, generator=gen)
    for name, mk, meta in [
        ("grad_gc_kagat", lambda: gcm.KAGAT(2, 7, 8, 3, 4, 3, 0.0, 2), dict(family="KAGAT", args=[2, 7, 8, 3, 4, 3, 0.0, 2])),
        ("grad_gc_fastkagat", lambda: gcm.FASTKAGAT(2, 7, 6, 2, 5, 0.0, 2), dict(family="FASTKAGAT", args=[2, 7, 6, 2, 5, 0.0, 2])),
    ]:
        m = mk().train()
        _randomise(m, gen)
        sd0, y, dy, grads = _backprop(m, lambda t: m(K.Batch(t, ei, batch)), x, gen)
        _save(name, dict(kind="gc", training=True, **meta), dict(x=x, edge_index=ei, batch=batch, dy=dy), sd0, y, grads)


if __name__ == "__main__":
    import sys
    if "--gine-only" in sys.argv:
        main_gine()
    elif "--fastkan-only" in sys.argv:
        main_fastkan()
    elif "--gat-only" in sys.argv:
        main_gat()
    else:
        main()
        main_fastkan()
        main_gine()
        main_gat()
