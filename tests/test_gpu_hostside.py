"""Host-side behaviour around the launches that a drop-in must get right (found by the round-1 code review): version
counters of tensors the kernels update through raw pointers, the deferred range check of edge_index, the device guard."""
import pytest
import torch

from oracle import kagnn_oracle as K

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _sd_cpu(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def test_eval_after_train_mode_forward_sees_the_new_running_statistics():
    """eval -> train-mode forward under no_grad (updates running_mean / running_var in the kernel, no optimizer step touches
    the BatchNorm affine) -> eval: the second eval must use the NEW statistics (nc/models.py:197 with nn.BatchNorm1d
    semantics), i.e. the cached eval-mode fold must notice the update."""
    import kagnn_b200 as kb
    torch.manual_seed(0)
    n, f = 700, 32
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, f, generator=g) * 0.5
    ei = torch.randint(0, n, (2, 4000), generator=g)
    m = kb.GKAN_Nodes("gin", 2, f, 32, 5, skip=True, grid_size=5, spline_order=3, hidden_layers=2).cuda()
    xd, eid = x.cuda(), ei.cuda()
    with torch.no_grad():
        m.eval()
        y0 = m(xd, eid).cpu()
        ref0 = K.node_model_forward(_sd_cpu(m), "gin", x, ei, True)
        assert K.rel_err(y0, ref0) <= TOL
        m.train()
        for _ in range(3):
            m(xd, eid)                                      # batch statistics; running estimates move
        m.eval()
        sd1 = _sd_cpu(m)
        assert not torch.allclose(sd1["bns.0.running_mean"], torch.zeros(32))
        y1 = m(xd, eid).cpu()
    ref1 = K.node_model_forward(sd1, "gin", x, ei, True)
    assert K.rel_err(y1, ref1) <= TOL
    assert K.rel_err(y1, ref0) > 1e-3                       # and it really is a different function now


def test_out_of_range_edge_index_raises():
    """PyG / ATen raise (or device-assert) on node ids outside [0, N); the CSR build flags them on the device and the flag is
    examined at a later graph call (immediately with poll_index_flags(block=True))."""
    import kagnn_b200 as kb
    from kagnn_b200 import graph
    graph.clear_cache()
    graph.poll_index_flags(block=True)
    n = 100
    x = torch.randn(n, 16).cuda()
    conv = kb.GIKANLayer(16, 16, 5, 3, 16, 2).cuda()
    good = torch.randint(0, n, (2, 300)).cuda()
    with torch.no_grad():
        conv(x, good)
        graph.poll_index_flags(block=True)                  # nothing wrong: no raise
        for bad_row, bad_val in ((0, n), (1, n + 7), (0, -1)):
            bad = good.clone()
            bad[bad_row, 17] = bad_val
            with pytest.raises(IndexError):                 # at the build if its flag is already back, else at the next check
                conv(x, bad)
                graph.poll_index_flags(block=True)
    graph.clear_cache()


def test_tensors_on_another_device_run_there():
    """The library launches on the CURRENT device; the wrappers must make the tensors' device current (ATen's device guard)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import kagnn_b200 as kb
    torch.manual_seed(0)
    m = kb.KAN([24, 16, 8], grid_size=5, spline_order=3)
    sd = _sd_cpu(m)
    x = torch.randn(300, 24)
    m1 = m.to("cuda:1")
    with torch.no_grad():
        assert torch.cuda.current_device() == 0
        y = m1(x.to("cuda:1"))
    assert y.device == torch.device("cuda:1")
    assert K.rel_err(y.cpu(), K.kan_chain(sd, "layers.", x)) <= TOL
    from kagnn_b200 import _lib as L
    from kagnn_b200 import ops
    with pytest.raises(RuntimeError):
        ops.fused_layer(ops.AggSpec(L.AGG_NONE, x.to("cuda:0")), 300, m1.kernel_specs())


def test_small_graph_forward_is_replayed_from_a_cuda_graph_and_tracks_every_input():
    """Eval-mode forwards of a small node model on the same inputs are replayed from a CUDA graph from the third call on
    (kagnn_b200/models_node.py: _GraphReplay); an in-place change of x, of edge_index or of a parameter must show in the result."""
    import kagnn_b200 as kb
    from kagnn_b200 import models_node
    from oracle import kagnn_oracle as K
    assert models_node._AUTO_GRAPH_NODES > 0
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(5)
    n, f = 900, 70
    x = torch.randn(n, f, generator=g) * 0.5
    ei = torch.randint(0, n, (2, 4000), generator=g)
    m = kb.GKAN_Nodes("gcn", 2, f, 16, 5, skip=True, grid_size=5, spline_order=3, dropout=0.0).eval()
    md = m.cuda()
    xd, eid = x.cuda(), ei.cuda()

    def ref():
        sd = {k: v.detach().cpu().clone() for k, v in md.state_dict().items()}
        return K.node_model_forward(sd, "gcn", xd.cpu(), eid.cpu(), True)

    with torch.no_grad():
        ys = [md(xd, eid) for _ in range(5)]
        assert md._replay.graph is not None                          # calls 4 and 5 were replays
        assert ys[3].data_ptr() != ys[4].data_ptr()                  # every call returns its own tensor
        for y in ys:
            assert K.rel_err(y.cpu(), ref()) <= 1e-4
        assert torch.equal(ys[0], ys[4])
        xd.mul_(0.5)                                                 # new version of x
        y = md(xd, eid)
        assert K.rel_err(y.cpu(), ref()) <= 1e-4 and not torch.equal(y, ys[4])
        for _ in range(3):
            y = md(xd, eid)
        assert K.rel_err(y.cpu(), ref()) <= 1e-4
        eid[0, :100] = (eid[0, :100] + 1) % n                        # new version of edge_index
        for _ in range(4):
            y = md(xd, eid)
        assert K.rel_err(y.cpu(), ref()) <= 1e-4
        md.lay_out.base_weight.mul_(1.5)                             # new version of a parameter
        for _ in range(4):
            y = md(xd, eid)
        assert K.rel_err(y.cpu(), ref()) <= 1e-4
