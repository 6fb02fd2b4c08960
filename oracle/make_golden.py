"""Generate the golden vectors under tests/golden/ -- TEST INFRASTRUCTURE ONLY.

Run in the BUILD container (needs /root/reference, which does not exist on the GPU box):

    python -m oracle.make_golden

It imports the reference's own ``ekan.py`` / ``fastkan.py`` / ``models.py`` (unmodified, from
/root/reference; models.py with ``oracle/pyg_shim.py`` standing in for torch_geometric), runs them
on small seeded inputs with randomised parameters, and stores inputs, ``state_dict`` and outputs as
``.npz``.  Nothing from the reference is copied into the repo: only numbers it computed.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
NC, GC, GR = (os.path.join(REF, d) for d in ("node_classification_clean", "graph_classification", "graph_regression"))
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(path: str, name: str, extra_path=()):
    """Import a reference file under a private module name (three packages share module names)."""
    for m in ("ekan", "fastkan", "models"):
        sys.modules.pop(m, None)
    old = list(sys.path)
    sys.path[:0] = [os.path.dirname(path), *extra_path]
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = old
    return mod


def _randomise(model: torch.nn.Module, gen: torch.Generator) -> None:
    """Make every tensor that the default init leaves trivial non-trivial: GCN bias (zeros), BN affine
    and running stats, LayerNorm affine.  Spline/base weights keep the reference's init."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "bias" and p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.2)
            if (".bn" in "." + name or "bns." in name or "layernorm" in name) and leaf == "weight":
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=gen))
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=gen) * 0.3)
            if name.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=gen))


def small_graph(n: int, e: int, gen: torch.Generator) -> torch.Tensor:
    """Directed multigraph with self loops, duplicate edges and isolated nodes (the last 3 nodes)."""
    ei = torch.randint(0, n - 3, (2, e), generator=gen)
    ei[:, :4] = ei[:, 4:8]                      # duplicates
    ei[1, 8:12] = ei[0, 8:12]                   # self loops
    ei[:, 12] = ei[:, 8]                        # a duplicated self loop
    return ei


def batched_graphs(n_graphs: int, gen: torch.Generator, lo=2, hi=9):
    """PyG-style batch: node ids grouped by graph, ``batch`` sorted; one graph has no edges."""
    sizes = torch.randint(lo, hi, (n_graphs,), generator=gen)
    off = torch.cumsum(sizes, 0) - sizes
    srcs, dsts = [], []
    for g in range(n_graphs):
        if g == 1:
            continue
        m = int(sizes[g]) * 2
        s = torch.randint(0, int(sizes[g]), (m,), generator=gen) + off[g]
        d = torch.randint(0, int(sizes[g]), (m,), generator=gen) + off[g]
        srcs += [s, d]
        dsts += [d, s]
    ei = torch.stack([torch.cat(srcs), torch.cat(dsts)])
    batch = torch.repeat_interleave(torch.arange(n_graphs), sizes)
    return ei, batch, int(sizes.sum())


def _save(name: str, meta: dict, inputs: dict, sd: dict, out: torch.Tensor) -> None:
    arrs = {"meta": np.array(json.dumps(meta))}
    for k, v in inputs.items():
        arrs["in/" + k] = v.detach().cpu().numpy()
    for k, v in sd.items():
        arrs["sd/" + k] = v.detach().cpu().numpy()
    arrs["out/y"] = out.detach().cpu().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(f"{name}: out {tuple(out.shape)}  |y|max={float(out.abs().max()):.4f}")


def main() -> None:
    from . import pyg_shim
    pyg_shim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(12345)
    gen = torch.Generator().manual_seed(12345)
    ekan = _load(os.path.join(NC, "ekan.py"), "ref_nc_ekan")
    fastkan = _load(os.path.join(NC, "fastkan.py"), "ref_nc_fastkan")

    # ---- a4: KANLinear, incl. knots hit exactly, out-of-range inputs, odd widths, every k ------
    for (g, k, fin, fout, n) in [(5, 3, 33, 7, 64), (4, 3, 16, 16, 40), (1, 1, 2, 3, 17), (8, 2, 5, 9, 33),
                                 (16, 4, 12, 4, 50), (32, 4, 3, 2, 25), (2, 2, 128, 64, 9)]:
        lay = ekan.KANLinear(fin, fout, grid_size=g, spline_order=k)
        x = torch.randn(n, fin, generator=gen) * 0.8
        knots = lay.grid[0]
        x[0, :] = knots[torch.arange(fin) % knots.numel()]            # exactly on knots (incl. t_0, t_last)
        x[1, :] = knots[0] - 1e-3
        x[2, :] = knots[-1] + 0.5
        x[3, :] = 0.0
        x[4, :] = torch.nextafter(knots[-1], torch.tensor(-10.0))
        with torch.no_grad():
            y = lay(x)
            bases = lay.b_splines(x)
        sd = dict(lay.state_dict())
        sd["__bases"] = bases
        _save(f"kanlinear_g{g}_k{k}_{fin}x{fout}", dict(kind="kan_linear", G=g, k=k), dict(x=x), sd, y)

    # ---- a5: KAN chains -------------------------------------------------------------------------
    for sizes, g, k in [([7, 16, 5], 5, 3), ([20, 8, 8, 8, 3], 3, 2), ([64, 64, 64], 5, 3)]:
        net = ekan.KAN(sizes, grid_size=g, spline_order=k)
        x = torch.randn(48, sizes[0], generator=gen)
        with torch.no_grad():
            y = net(x)
        _save("kan_" + "_".join(map(str, sizes)) + f"_g{g}k{k}", dict(kind="kan_chain", G=g, k=k), dict(x=x), net.state_dict(), y)

    # ---- a7/a8: FastKANLayer / FastKAN ------------------------------------------------------------
    for sizes, g in [([7, 32], 8), ([33, 5], 4), ([2, 3], 2), ([16, 16, 4], 8), ([256, 64, 2], 8), ([5, 6, 7, 8], 32)]:
        net = fastkan.FastKAN(sizes, num_grids=g)
        _randomise(net, gen)
        x = torch.randn(40, sizes[0], generator=gen) * 1.5
        x[0] = 7.0                                      # constant row: LayerNorm variance 0
        with torch.no_grad():
            y = net(x)
        _save("fastkan_" + "_".join(map(str, sizes)) + f"_g{g}", dict(kind="fastkan_chain", G=g), dict(x=x), net.state_dict(), y)

    # ---- a9/a10/a12: node models ---------------------------------------------------------------------
    ncm = _load(os.path.join(NC, "models.py"), "ref_nc_models")
    n, e, f, c = 70, 260, 19, 5
    ei = small_graph(n, e, gen)
    x = torch.randn(n, f, generator=gen) * 0.7
    for conv in ("gcn", "gin"):
        for skip in (True, False):
            m = ncm.GKAN_Nodes(conv, 2, f, 12, c, skip=skip, grid_size=5, spline_order=3, hidden_layers=2, dropout=0.0).eval()
            _randomise(m, gen)
            with torch.no_grad():
                y = m(x, ei)
            _save(f"nc_gkan_{conv}_skip{int(skip)}", dict(kind="node", conv_type=conv, skip=skip, fast=False, mp_layers=2,
                  num_features=f, hidden=12, classes=c, G=5, k=3, hidden_layers=2), dict(x=x, edge_index=ei), m.state_dict(), y)
        m = ncm.GFASTKAN_Nodes(conv, 3, f, 10, c, skip=True, grid_size=6, hidden_layers=2, dropout=0.0).eval()
        _randomise(m, gen)
        with torch.no_grad():
            y = m(x, ei)
        _save(f"nc_gfastkan_{conv}", dict(kind="node", conv_type=conv, skip=True, fast=True, mp_layers=3, num_features=f,
              hidden=10, classes=c, G=6, hidden_layers=2), dict(x=x, edge_index=ei), m.state_dict(), y)
    # train-mode BN (batch statistics) for one model
    m = ncm.GKAN_Nodes("gin", 2, f, 12, c, grid_size=4, spline_order=3, hidden_layers=1).train()
    _randomise(m, gen)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        y = m(x, ei)
    _save("nc_gkan_gin_trainbn", dict(kind="node", conv_type="gin", skip=True, fast=False, training=True, mp_layers=2,
          num_features=f, hidden=12, classes=c, G=4, k=3, hidden_layers=1), dict(x=x, edge_index=ei), sd0, y)

    # ---- a13: graph classification -------------------------------------------------------------------
    gcm = _load(os.path.join(GC, "models.py"), "ref_gc_models")
    ei, batch, n = batched_graphs(9, gen)
    x = torch.nn.functional.one_hot(torch.randint(0, 7, (n,), generator=gen), 7).float()
    data = K.Batch(x, ei, batch)
    for name, mk, meta in [
        ("gc_kagin", lambda: gcm.KAGIN(2, 7, 16, 3, 2, 5, 3, 0.0), dict(family="KAGIN", args=[2, 7, 16, 3, 2, 5, 3, 0.0])),
        ("gc_fastkagin", lambda: gcm.FASTKAGIN(2, 7, 16, 2, 2, 8, 0.0), dict(family="FASTKAGIN", args=[2, 7, 16, 2, 2, 8, 0.0])),
        ("gc_kagcn", lambda: gcm.KAGCN(3, 7, 12, 4, 4, 2, 0.0), dict(family="KAGCN", args=[3, 7, 12, 4, 4, 2, 0.0])),
        ("gc_fastkagcn", lambda: gcm.FASTKAGCN(2, 7, 12, 3, 5, 0.0), dict(family="FASTKAGCN", args=[2, 7, 12, 3, 5, 0.0])),
    ]:
        m = mk().eval()
        _randomise(m, gen)
        with torch.no_grad():
            y = m(data)
        _save(name, dict(kind="gc", **meta), dict(x=x, edge_index=ei, batch=batch), m.state_dict(), y)

    # ---- a11/a14: graph regression (GINE, OGB encoders); gr/ lacks fastkan.py -> borrow nc's ---------
    grm = _load(os.path.join(GR, "models.py"), "ref_gr_models", extra_path=(NC,))
    ei, batch, n = batched_graphs(8, gen)
    xz = torch.randint(0, 28, (n, 1), generator=gen)
    ea = torch.randint(1, 4, (ei.size(1),), generator=gen)
    data = K.Batch(xz, ei, batch, ea)
    for name, mk, meta in [
        ("gr_kagin", lambda: grm.KAGIN(1, 1, 3, 16, 2, 5, 3, 1, 0.0, True), dict(family="KAGIN", args=[1, 1, 3, 16, 2, 5, 3, 1, 0.0, True])),
        ("gr_fastkagin", lambda: grm.FASTKAGIN(1, 1, 2, 16, 2, 6, 1, 0.0, True), dict(family="FASTKAGIN", args=[1, 1, 2, 16, 2, 6, 1, 0.0, True])),
        ("gr_kagcn", lambda: grm.KAGCN(1, 2, 16, 5, 3, 1, 0.0, True), dict(family="KAGCN", args=[1, 2, 16, 5, 3, 1, 0.0, True])),
        ("gr_fastkagcn", lambda: grm.FASTKAGCN(1, 2, 16, 4, 1, 0.0, True), dict(family="FASTKAGCN", args=[1, 2, 16, 4, 1, 0.0, True])),
    ]:
        m = mk().eval()
        _randomise(m, gen)
        with torch.no_grad():
            y = m(data)
        _save(name, dict(kind="gr", **meta), dict(x=xz, edge_index=ei, batch=batch, edge_attr=ea), m.state_dict(), y)
    # QM9-style linear encoders (ogb_encoders=False): float node / edge features
    xq = torch.randn(n, 11, generator=gen)
    eq = torch.randn(ei.size(1), 4, generator=gen)
    m = grm.KAGIN(11, 4, 2, 16, 2, 4, 3, 3, 0.0, False).eval()
    _randomise(m, gen)
    with torch.no_grad():
        y = m(K.Batch(xq, ei, batch, eq))
    _save("gr_kagin_linear_enc", dict(kind="gr", family="KAGIN", args=[11, 4, 2, 16, 2, 4, 3, 3, 0.0, False]),
          dict(x=xq, edge_index=ei, batch=batch, edge_attr=eq), m.state_dict(), y)


def main_large_grids() -> None:
    """Layers with more than eight coefficients per (in, out) pair -- the upper part of the reference's search space
    (node_classification_clean/one_experiment.py:45-46) -- from the reference's own KANLinear / GKAN_Nodes: the shapes the product
    evaluates as slot windows.  Own generator, written next to the other fixtures without touching them
    (``python -m oracle.make_golden --large-grids-only``)."""
    from . import pyg_shim
    pyg_shim.install()
    torch.manual_seed(2468)
    gen = torch.Generator().manual_seed(2468)
    ekan = _load(os.path.join(NC, "ekan.py"), "ref_nc_ekan")
    for (g, k, fin, fout, n) in [(8, 3, 24, 32, 300), (13, 3, 20, 16, 200), (20, 2, 12, 24, 150), (6, 3, 9, 5, 140)]:
        lay = ekan.KANLinear(fin, fout, grid_size=g, spline_order=k)
        x = torch.randn(n, fin, generator=gen) * 0.8
        knots = lay.grid[0]
        x[0, :] = knots[torch.arange(fin) % knots.numel()]            # exactly on knots (incl. t_0, t_last)
        x[1, :] = knots[0] - 1e-3
        x[2, :] = knots[-1] + 0.5
        x[3, :] = 0.0
        x[4, :] = torch.nextafter(knots[-1], torch.tensor(-10.0))
        x[5, :] = knots[8] if knots.numel() > 8 else 0.0               # the boundary between the first two windows of eight slots
        with torch.no_grad():
            y = lay(x)
            bases = lay.b_splines(x)
        sd = dict(lay.state_dict())
        sd["__bases"] = bases
        _save(f"kanlinear_g{g}_k{k}_{fin}x{fout}", dict(kind="kan_linear", G=g, k=k), dict(x=x), sd, y)
    ncm = _load(os.path.join(NC, "models.py"), "ref_nc_models")
    n, e, f, c = 400, 1500, 19, 5
    ei = small_graph(n, e, gen)
    x = torch.randn(n, f, generator=gen) * 0.7
    for conv in ("gcn", "gin"):
        m = ncm.GKAN_Nodes(conv, 2, f, 12, c, skip=True, grid_size=8, spline_order=3, hidden_layers=2, dropout=0.0).eval()
        _randomise(m, gen)
        with torch.no_grad():
            y = m(x, ei)
        _save(f"nc_gkan_{conv}_g8k3", dict(kind="node", conv_type=conv, skip=True, fast=False, mp_layers=2, num_features=f, hidden=12,
              classes=c, G=8, k=3, hidden_layers=2), dict(x=x, edge_index=ei), m.state_dict(), y)


if __name__ == "__main__":
    from . import kagnn_oracle as K
    if "--large-grids-only" in sys.argv:
        main_large_grids()
    else:
        main()
        main_large_grids()
