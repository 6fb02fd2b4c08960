"""Accuracy of the slot-window paths on a whole node model: against the oracle in fp64 -- the windowed tensor-core path, the general
fp32 kernel, and the oracle itself in fp32 (what the reference's own fp32 arithmetic is worth on this model)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb
from kagnn_b200 import ops
from oracle import kagnn_oracle as K


def main():
    n, f = 3000, 48
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, f, generator=g) * 0.5
    ei = torch.randint(0, n, (2, 18000), generator=g)
    out = {}
    for name, make in (("fastkan_grid12", lambda: kb.GFASTKAN_Nodes("gin", 2, f, 32, 5, skip=True, grid_size=12, hidden_layers=2)),
                       ("fastkan_grid8", lambda: kb.GFASTKAN_Nodes("gin", 2, f, 32, 5, skip=True, grid_size=8, hidden_layers=2)),
                       ("fastkan_grid32", lambda: kb.GFASTKAN_Nodes("gin", 2, f, 32, 5, skip=True, grid_size=32, hidden_layers=2)),
                       ("kan_grid8_k3", lambda: kb.GKAN_Nodes("gin", 2, f, 32, 5, skip=True, grid_size=8, spline_order=3, hidden_layers=2))):
        torch.manual_seed(13)
        m = make().eval()
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        ref64 = K.node_model_forward(sd64, "gin", x.double(), ei, True)
        ref32 = K.node_model_forward(sd, "gin", x, ei, True)
        m = m.cuda()
        with torch.no_grad():
            y_win = m(x.cuda(), ei.cuda()).cpu()
            saved = ops.tc_supported
            ops.tc_supported = lambda *a: False
            try:
                for mod in m.modules():
                    if hasattr(mod, "_cache_key"):
                        mod._cache_key = None
                y_gen = m(x.cuda(), ei.cuda()).cpu()
            finally:
                ops.tc_supported = saved
        out[name] = {"tensor_core_vs_fp64": K.rel_err(y_win.double(), ref64), "general_fp32_vs_fp64": K.rel_err(y_gen.double(), ref64),
                     "oracle_fp32_vs_fp64": K.rel_err(ref32.double(), ref64), "tensor_core_vs_oracle_fp32": K.rel_err(y_win, ref32)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
