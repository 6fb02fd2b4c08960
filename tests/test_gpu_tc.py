"""GPU tests of the tcgen05 building blocks (descriptor encodings, TMEM addressing, bulk TMA, bf16 hi/lo split)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(16, 16), (64, 64), (64, 128), (128, 128), (256, 64), (48, 32), (64, 256)])
def test_tc_selftest_matches_fp64(n, k):
    from kagnn_b200 import ops
    torch.manual_seed(n * 7 + k)
    a = torch.randn(128, k)
    b = torch.randn(n, k)
    ref = a.double() @ b.double().t()
    scale = float(ref.abs().max())
    d1 = ops.tc_selftest(a.cuda(), b.cuda(), 1).cpu().double()
    d3 = ops.tc_selftest(a.cuda(), b.cuda(), 3).cpu().double()
    e1 = float((d1 - ref).abs().max()) / scale
    e3 = float((d3 - ref).abs().max()) / scale
    assert e1 < 2e-2, f"single bf16 pass is structurally wrong: {e1}"
    assert e3 < 2e-5, f"hi/lo compensated product too inaccurate: {e3}"


def test_tc_selftest_row_and_column_identity():
    """A = one-hot rows, B = distinct integers: every (row, column) of D must land where it belongs."""
    from kagnn_b200 import ops
    k, n = 128, 64
    a = torch.zeros(128, k)
    a[torch.arange(128), torch.arange(128) % k] = 1.0
    b = (torch.arange(n * k, dtype=torch.float32).view(n, k) % 251) - 125.0      # exactly representable in bf16
    d = ops.tc_selftest(a.cuda(), b.cuda(), 1).cpu()
    assert torch.equal(d, a @ b.t())


@pytest.mark.parametrize("n,k", [(16, 16), (64, 64), (64, 128), (128, 128), (256, 64), (48, 32), (64, 256)])
def test_tc_selftest_a_operand_in_tmem(n, k):
    """Same product with the A operand written to tensor memory by tcgen05.st and consumed by the TS-form MMA
    (the layout the fused kernel's basis producers use): modes 11 (one bf16 pass) and 13 (hi/lo compensated)."""
    from kagnn_b200 import ops
    torch.manual_seed(n * 11 + k)
    a = torch.randn(128, k)
    b = torch.randn(n, k)
    ref = a.double() @ b.double().t()
    scale = float(ref.abs().max())
    d1 = ops.tc_selftest(a.cuda(), b.cuda(), 11).cpu().double()
    d3 = ops.tc_selftest(a.cuda(), b.cuda(), 13).cpu().double()
    assert float((d1 - ref).abs().max()) / scale < 2e-2
    assert float((d3 - ref).abs().max()) / scale < 2e-5


def test_tc_selftest_tmem_a_identity():
    from kagnn_b200 import ops
    k, n = 128, 64
    a = torch.zeros(128, k)
    a[torch.arange(128), (torch.arange(128) * 5) % k] = 1.0
    b = (torch.arange(n * k, dtype=torch.float32).view(n, k) % 251) - 125.0
    d = ops.tc_selftest(a.cuda(), b.cuda(), 11).cpu()
    assert torch.equal(d, a @ b.t())
