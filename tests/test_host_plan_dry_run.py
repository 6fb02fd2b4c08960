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
