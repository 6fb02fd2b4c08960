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
        return K.gc_kagin_forward(sd, data) if fam.endswith("GIN") else K.gc_kagcn_forward(sd, data)
    if kind == "gr":
        return K.gr_kagin_forward(sd, data, dtype=dtype) if fam.endswith("GIN") else K.gr_kagcn_forward(sd, data, dtype=dtype)
    raise ValueError(kind)
