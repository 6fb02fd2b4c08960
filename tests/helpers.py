"""Shared test helpers: golden-vector loading and the oracle dispatch per golden ``kind``."""
import glob
import json
import os

import numpy as np
import torch

from oracle import kagnn_oracle as K

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    inputs = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in/")}
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    return meta, inputs, sd, torch.from_numpy(z["out/y"])


def oracle_run(meta, inputs, sd, dtype=torch.float32):
    """Run the oracle for one golden case (any kind)."""
    sd = K.to_dtype({k: v for k, v in sd.items() if not k.startswith("__")}, dtype)
    x = inputs["x"].to(dtype) if inputs["x"].is_floating_point() else inputs["x"]
    kind = meta["kind"]
    if kind == "kan_linear":
        return K._kan_layer_from_sd(sd, "", x)
    if kind == "kan_chain":
        return K.kan_chain(sd, "layers.", x)
    if kind == "fastkan_chain":
        return K.fastkan_chain(sd, "layers.", x)
    if kind == "node":
        return K.node_model_forward(sd, meta["conv_type"], x, inputs["edge_index"], meta["skip"], meta.get("training", False))
    ea = inputs.get("edge_attr")
    if ea is not None and ea.is_floating_point():
        ea = ea.to(dtype)
    data = K.Batch(x, inputs["edge_index"], inputs["batch"], ea)
    fam = meta["family"]
    if kind == "gc":
        if fam.endswith("GAT"):
            return K.gc_kagat_forward(sd, data)
        return (K.gc_kagin_forward(sd, data, meta.get("training", False)) if fam.endswith("GIN") else K.gc_kagcn_forward(sd, data))
    if kind == "gr":
        return (K.gr_kagin_forward(sd, data, meta.get("training", False), dtype=dtype) if fam.endswith("GIN")
                else K.gr_kagcn_forward(sd, data, dtype=dtype))
    raise ValueError(kind)


# ---- product-side construction for the golden cases (used by the GPU parity tests) ------------------------------
def build_product_model(meta, sd, device="cuda"):
    """Instantiate the kagnn_b200 model that corresponds to a golden case and load the reference state_dict."""
    import kagnn_b200 as kb
    kind = meta["kind"]
    sd = {k: v for k, v in sd.items() if not k.startswith("__")}
    if kind == "kan_linear":
        out_f, in_f, s = sd["spline_weight"].shape
        m = kb.KANLinear(in_f, out_f, grid_size=meta["G"], spline_order=meta["k"])
    elif kind == "kan_chain":
        n = len([k for k in sd if k.endswith("base_weight")])
        sizes = [sd["layers.0.base_weight"].shape[1]] + [sd[f"layers.{i}.base_weight"].shape[0] for i in range(n)]
        m = kb.KAN(sizes, grid_size=meta["G"], spline_order=meta["k"])
    elif kind == "fastkan_chain":
        n = len([k for k in sd if k.endswith("base_linear.weight")])
        sizes = [sd["layers.0.base_linear.weight"].shape[1]] + [sd[f"layers.{i}.base_linear.weight"].shape[0] for i in range(n)]
        m = kb.FastKAN(sizes, num_grids=meta["G"])
    elif kind == "node":
        if meta["fast"]:
            m = kb.GFASTKAN_Nodes(meta["conv_type"], meta["mp_layers"], meta["num_features"], meta["hidden"], meta["classes"],
                                  skip=meta["skip"], grid_size=meta["G"], hidden_layers=meta["hidden_layers"], heads=meta.get("heads", 4))
        else:
            m = kb.GKAN_Nodes(meta["conv_type"], meta["mp_layers"], meta["num_features"], meta["hidden"], meta["classes"],
                              skip=meta["skip"], grid_size=meta["G"], spline_order=meta["k"], hidden_layers=meta["hidden_layers"],
                              heads=meta.get("heads", 4))
    elif kind == "gc":
        m = getattr(kb.models_graph, meta["family"])(*meta["args"])
    elif kind == "gr":
        m = getattr(kb.models_regr, meta["family"])(*meta["args"])
    else:
        raise ValueError(kind)
    m.load_state_dict(sd, strict=True)
    m = m.to(device)
    m.train(bool(meta.get("training", False)))
    return m


def product_run(meta, inputs, model, device="cuda"):
    inp = {k: v.to(device) for k, v in inputs.items()}
    with torch.no_grad():
        if meta["kind"] in ("kan_linear", "kan_chain", "fastkan_chain"):
            return model(inp["x"])
        if meta["kind"] == "node":
            return model(inp["x"], inp["edge_index"])
        return model(K.Batch(inp["x"], inp["edge_index"], inp["batch"], inp.get("edge_attr")))


# ---- gradient fixtures (tests/golden/grad/, produced by oracle/make_golden_grad.py) -----------------------------------
def grad_golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "grad", prefix + "*.npz")))


def load_grad_golden(name):
    """-> meta, inputs (x, dy[, edge_index, batch]), state_dict, y, grads ('__x' = d loss / d x, else parameter names)."""
    z = np.load(os.path.join(GOLDEN, "grad", name + ".npz"))
    meta = json.loads(str(z["meta"]))
    pick = lambda pre: {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}
    return meta, pick("in/"), pick("sd/"), torch.from_numpy(z["out/y"]), pick("grad/")


def oracle_grads(meta, inputs, sd):
    """Autograd through the oracle: the same quantities the fixture holds."""
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("grid", "running_mean", "running_var", "eps"))
              else v.clone()) for k, v in sd.items()}
    x = inputs["x"].clone()
    if x.is_floating_point():
        x.requires_grad_(True)
    ins = dict(inputs, x=x)
    y = oracle_run(meta, ins, sd)
    y.backward(inputs["dy"])
    grads = {k: v.grad for k, v in sd.items() if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
    grads["__x"] = x.grad if x.is_floating_point() else torch.zeros(1)          # integer atom codes: no input gradient
    return y, grads


def grad_err(g, g_ref, scale):
    """max|g - g_ref| relative to max|g_ref|, floored at 5 % of the largest parameter gradient of the case: a gradient that is
    analytically zero (a bias in front of a train-mode BatchNorm) is rounding noise in the reference itself."""
    denom = max(float(g_ref.abs().max()), 0.05 * scale)
    return float((g.double() - g_ref.double()).abs().max()) / denom


def grad_scale(g_ref):
    return max(float(v.abs().max()) for k, v in g_ref.items() if k != "__x")
