"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/kagnn_b200.h declares, and the host-side mirror refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "kagnn_b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kagnn_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from kagnn_b200 import build, _lib
    path = build.build()
    h = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 11
    for sym in declared:
        assert hasattr(h, sym), f"{sym} declared in include/kagnn_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared          # the ctypes binding covers the whole header
    assert _lib.lib().kagnn_version() == 100
    assert _lib.lib().kagnn_strerror(-2).decode().startswith("configuration not supported")


def test_host_only_queries_need_no_gpu():
    from kagnn_b200 import _lib
    n = _lib.lib().kagnn_packed_weight_elems(128, 64, 8)
    assert n == 128 * 9 * 64
    assert _lib.lib().kagnn_packed_weight_elems(5, 7, 3) == 5 * 4 * 8        # out padded to 4


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct as gcc lays the header out == the ctypes mirror in kagnn_b200/_lib.py."""
    import subprocess
    from kagnn_b200 import _lib
    structs = {"KagnnAffine": _lib.KagnnAffine, "KagnnKanLayer": _lib.KagnnKanLayer, "KagnnAggregate": _lib.KagnnAggregate}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "kagnn_b200.h")}"', "int main(void) {"]
    for name, cls in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == ctypes.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got[f"{name}.{field}"]) == getattr(cls, field).offset, (name, field)
    assert ctypes.sizeof(_lib.KagnnAggregate) == 216


def test_no_cpu_fallback():
    import kagnn_b200 as kb
    m = kb.KANLinear(3, 2)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(4, 3))
    with torch.no_grad(), pytest.raises(RuntimeError):
        kb.GKAN_Nodes("gin", 1, 3, 4, 2)(torch.randn(4, 3), torch.zeros(2, 0, dtype=torch.long))


def test_unknown_and_out_of_scope_configurations_raise():
    import kagnn_b200 as kb
    with pytest.raises(ValueError, match="unknown conv_type"):
        kb.GKAN_Nodes("sage", 1, 3, 4, 2)
    m = kb.GKAN_Nodes("gat", 1, 3, 4, 2, heads=2)               # the GAT flavour exists (inference); its non-default options do not
    assert m.bns[0].num_features == 8 and m.lay_out.in_features == 3 + 8
    with pytest.raises(NotImplementedError):
        kb.GATConv(3, 4, heads=2, concat=False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kagnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_init_matches_reference_recipe_statistics():
    """reset_parameters is cold-path torch code; check that the fitted spline coefficients reproduce the noise
    they were fitted to (curve2coeff is an exact interpolation when G+1 <= G+k)."""
    import kagnn_b200 as kb
    from kagnn_b200.ekan import _uniform_bspline_design
    from oracle import kagnn_oracle as K
    torch.manual_seed(0)
    lay = kb.KANLinear(6, 4, grid_size=5, spline_order=3)
    x = torch.linspace(-2.3, 2.3, 97).unsqueeze(1).expand(-1, 6).contiguous()
    mine = _uniform_bspline_design(x, float(lay.grid[0, 0]), float(lay.grid[0, 1] - lay.grid[0, 0]), 5, 3)
    ref = K.bspline_bases(x, lay.grid, 3)
    assert torch.allclose(mine, ref, atol=2e-6)
    assert lay.spline_weight.abs().max() < 1.0 and lay.spline_weight.abs().max() > 0


def test_aggspec_makes_column_strided_views_row_major():
    from kagnn_b200 import ops, _lib as L
    base = torch.randn(6, 10)
    spec = ops.AggSpec(L.AGG_NONE, base.t())
    assert spec.x.stride(1) == 1 and torch.equal(spec.x, base.t())
    sliced = base[:, 2:7]                              # unit column stride: kept as a view (the kernels take a leading dimension)
    assert ops.AggSpec(L.AGG_NONE, sliced).x.data_ptr() == sliced.data_ptr()


def test_host_check_seam_is_not_part_of_the_product():
    """kagnn_b200/csrc/launch.cuh lets the CPU suite compile backward.cu as serial host code (tests/emul/).  That build is test
    infrastructure: the product's Python never refers to it, the product build never defines the macro, and the only source
    that mentions the macro is the seam header itself."""
    from kagnn_b200 import build
    pkg = os.path.join(ROOT, "kagnn_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            with open(os.path.join(pkg, f)) as fh:
                src = fh.read()
            assert "emul" not in src and "host_check" not in src.lower() and "HOST_CHECK" not in src, f
    assert not any("HOST_CHECK" in flag for flag in build.NVCC_FLAGS)
    users = []
    for f in os.listdir(os.path.join(pkg, "csrc")):
        with open(os.path.join(pkg, "csrc", f)) as fh:
            if "KAGNN_HOST_CHECK" in fh.read():
                users.append(f)
    assert users == ["launch.cuh"]
