"""Golden vectors of the GAT flavour -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_gat          (build container: needs /root/reference)

Same recipe as oracle/make_golden.py, kept separate so that the fixtures that file generates keep their numbers: the reference's
own ``models.py`` (node_classification_clean and graph_classification) run unmodified with ``oracle/pyg_shim.py`` standing in
for torch_geometric -- so the model glue (``KAGATConv`` replacing ``GATConv.lin`` by a KAN, BatchNorm width hidden * heads, skip
concat, read-out) is the reference's code, while the attention arithmetic is the restatement ``kagnn_oracle.gat_conv``
(parity unpinned: torch_geometric 2.5.3 cannot be installed here)."""
from __future__ import annotations

import os

import torch

from . import kagnn_oracle as K
from .make_golden import GC, NC, _load, _randomise, _save, batched_graphs, small_graph


def main() -> None:
    from . import pyg_shim
    pyg_shim.install()
    gen = torch.Generator().manual_seed(424242)
    torch.manual_seed(424242)
    ncm = _load(os.path.join(NC, "models.py"), "ref_nc_models_gat")
    n, e, f, c = 70, 260, 19, 5
    ei = small_graph(n, e, gen)
    x = torch.randn(n, f, generator=gen) * 0.7
    m = ncm.GKAN_Nodes("gat", 2, f, 8, c, skip=True, grid_size=5, spline_order=3, dropout=0.0, heads=3).eval()
    _randomise(m, gen)
    with torch.no_grad():
        y = m(x, ei)
    _save("nc_gkan_gat", dict(kind="node", conv_type="gat", skip=True, fast=False, mp_layers=2, num_features=f, hidden=8, classes=c,
          G=5, k=3, hidden_layers=2, heads=3), dict(x=x, edge_index=ei), m.state_dict(), y)
    m = ncm.GFASTKAN_Nodes("gat", 2, f, 6, c, skip=False, grid_size=6, dropout=0.0, heads=2).eval()
    _randomise(m, gen)
    with torch.no_grad():
        y = m(x, ei)
    _save("nc_gfastkan_gat", dict(kind="node", conv_type="gat", skip=False, fast=True, mp_layers=2, num_features=f, hidden=6,
          classes=c, G=6, hidden_layers=2, heads=2), dict(x=x, edge_index=ei), m.state_dict(), y)

    gcm = _load(os.path.join(GC, "models.py"), "ref_gc_models_gat")
    ei, batch, n = batched_graphs(9, gen)
    x = torch.nn.functional.one_hot(torch.randint(0, 7, (n,), generator=gen), 7).float()
    data = K.Batch(x, ei, batch)
    for name, mk, meta in [
        ("gc_kagat", lambda: gcm.KAGAT(2, 7, 8, 3, 4, 3, 0.0, 4), dict(family="KAGAT", args=[2, 7, 8, 3, 4, 3, 0.0, 4])),
        ("gc_fastkagat", lambda: gcm.FASTKAGAT(2, 7, 6, 2, 5, 0.0, 2), dict(family="FASTKAGAT", args=[2, 7, 6, 2, 5, 0.0, 2])),
    ]:
        m = mk().eval()
        _randomise(m, gen)
        with torch.no_grad():
            y = m(data)
        _save(name, dict(kind="gc", **meta), dict(x=x, edge_index=ei, batch=batch), m.state_dict(), y)


if __name__ == "__main__":
    main()
