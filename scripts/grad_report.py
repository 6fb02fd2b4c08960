"""Worst forward / gradient error per gradient fixture on the GPU (numbers behind the tolerances of
tests/test_gpu_train_backward.py).  python scripts/grad_report.py > gpurun_out/grad_report.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kagnn_oracle as K
from tests.helpers import grad_err, grad_golden_names, grad_scale, load_grad_golden
from tests.test_backward_wiring import run_product_grads

only = sys.argv[1] if len(sys.argv) > 1 else ""
for name in [n for n in grad_golden_names() if only in n]:
    meta, inputs, sd, y_ref, g_ref = load_grad_golden(name)
    try:
        y, g, _ = run_product_grads(meta, inputs, sd, "cuda")
        scale = grad_scale(g_ref)
        errs = {k: grad_err(g[k].cpu(), g_ref[k], scale) for k in g_ref}
        worst = max(errs, key=errs.get)
        print(f"{name:32s} fwd {K.rel_err(y.cpu(), y_ref):.2e}  dx {errs['__x']:.2e}  worst {errs[worst]:.2e} ({worst})", flush=True)
    except Exception as exc:
        print(f"{name:32s} ERROR {exc!r}"[:300], flush=True)
