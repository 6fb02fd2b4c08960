"""CPU check of the host-side execution plans (kagnn_b200/models_*.py, conv.py, ekan.py, fastkan.py): which launches a model
forward issues, with which column slices, folded BatchNorm affines, half-layer-shifted GCN fusion, pooling and read-out --
with every library launch replaced by a torch-CPU stand-in of the same call (tests/emul/cpu_double.py, test infrastructure).
The numbers are compared with the outputs the reference itself produced (tests/golden/*.npz); the real kernels behind the same
plans are checked on the B200 by tests/test_gpu_parity.py."""
import pytest
import torch

from oracle import kagnn_oracle as K
from tests.emul.cpu_double import cpu_double
from tests.helpers import build_product_model, golden_names, load_golden, product_run


@pytest.mark.parametrize("name", golden_names())
def test_inference_plan_reproduces_reference_output(name):
    meta, inputs, sd, y_ref = load_golden(name)
    if meta["kind"] == "kan_linear" and not torch.isfinite(inputs["x"]).all():
        pytest.skip("special values are a kernel matter")
    with cpu_double():
        model = build_product_model(meta, sd, device="cpu")
        y = product_run(meta, inputs, model, device="cpu")
    assert y.shape == y_ref.shape
    assert K.rel_err(y, y_ref) <= 1e-4, name          # the stand-ins are fp32 torch: north_star tolerance


@pytest.mark.parametrize("name", ["kanlinear_g8_k3_24x32", "kanlinear_g13_k3_20x16", "kanlinear_g20_k2_12x24", "kanlinear_g8_k2_5x9",
                                  "nc_gkan_gcn_g8k3", "nc_gkan_gin_g8k3", "fastkan_5_6_7_8_g32", "nc_gkan_gin_skip1", "gc_kagin"])
def test_slot_window_plan_reproduces_reference_output(name):
    """The same, with the slot-window wiring switched on (tests/emul/cpu_double.py: windows=True): layers with more than eight
    coefficients or centres per input run as ops._windowed_chain -- the aggregation materialised, every layer over copies of its
    input -- and must still reproduce what the reference computed; models within eight slots must be untouched."""
    meta, inputs, sd, y_ref = load_golden(name)
    with cpu_double(windows=True):
        model = build_product_model(meta, sd, device="cpu")
        y = product_run(meta, inputs, model, device="cpu")
        windows = [m.kernel_spec().windows for m in model.modules() if hasattr(m, "kernel_spec")]
    assert (max(windows) > 1) == any(t in name for t in ("g8_k3", "g13", "g20", "g8_k2", "g8k3", "g32")), (name, windows)
    assert y.shape == y_ref.shape
    assert K.rel_err(y, y_ref) <= 1e-4, name
