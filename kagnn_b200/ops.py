"""Thin tensor -> (pointer, sizes, stream) wrappers over the C ABI (include/kagnn_b200.h).

torch is used for device memory and streams only; every arithmetic step below runs in libkagnn_b200.so.
Nothing here falls back to torch math: a CPU tensor or a missing library raises."""
from __future__ import annotations

import ctypes as C
import dataclasses
import functools
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib as L

Tensor = torch.Tensor

# launch counter: bench.py reports how many of OUR kernels ran inside the timed region
launch_count = 0
# optional per-launch timing: set to a list and every fused launch appends (label, start_event, end_event)
fused_timing = None


# torch.cuda.current_stream() builds a Stream object through three Python layers (17 us a call, a seventh of the host time of a
# training step); the raw handle is one C call
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda dev: torch.cuda.current_stream(dev).cuda_stream)
_current_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def _stream() -> C.c_void_p:
    """Current stream of the current device; every public op below runs under ``_on_device``, which makes the device of
    its tensor arguments the current one first (the library launches on cudaGetDevice and never switches devices itself)."""
    return C.c_void_p(_raw_stream(_current_device()))


def _devices_of(obj, found: set) -> None:
    """Devices of the tensors of a launch argument.  Records (AggSpec, KanLayerSpec, Affine, CSR) name the tensors that stand
    for them in ``_anchors`` -- enough to catch inputs / weights / outputs on different devices without walking every field on
    every launch."""
    if obj is None:
        return
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            found.add(obj.device.index)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _devices_of(o, found)
    elif hasattr(obj, "_anchors"):
        d = obj.__dict__
        for n in obj._anchors:
            v = d[n]
            if v is not None and v.is_cuda:
                found.add(v.device.index)
    elif isinstance(obj, torch.nn.Module):
        for t in obj._parameters.values():
            _devices_of(t, found)
        for t in obj._buffers.values():
            _devices_of(t, found)


_n_devices = None


def _single_device() -> bool:
    global _n_devices
    if _n_devices is None:
        if not torch.cuda.is_initialized():
            return False                           # not known yet (and nothing launched so far): take the full check
        _n_devices = torch.cuda.device_count()
    return _n_devices == 1


def _on_device(fn):
    """Device guard of a launch wrapper (what ATen ops do implicitly): all CUDA tensor arguments -- also inside the AggSpec /
    KanLayerSpec / Affine / CSR records and BatchNorm modules -- must live on ONE device, and that device is made current for
    the call, so the kernels, workspaces and the stream handed to the library belong to the tensors and not to whatever device
    happened to be current."""
    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        if _single_device():                      # one visible GPU: every CUDA tensor lives on it and it is current
            return fn(*args, **kwargs)
        found: set = set()
        _devices_of(args, found)
        _devices_of(tuple(kwargs.values()), found)
        if len(found) > 1:
            raise RuntimeError(f"kagnn_b200.{fn.__name__}: tensor arguments live on different CUDA devices {sorted(found)}")
        if not found or next(iter(found)) == _current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(next(iter(found))):
            return fn(*args, **kwargs)
    return guarded


def _p(t: Optional[Tensor]) -> Optional[C.c_void_p]:
    return None if t is None else C.c_void_p(t.data_ptr())


_dummies = {}


def _addr(t: Tensor) -> int:
    """Device address; an empty tensor (edgeless graph) still yields a valid non-null pointer."""
    if t.numel() > 0:
        return t.data_ptr()
    key = (t.device, t.dtype)
    if key not in _dummies:
        _dummies[key] = torch.zeros(4, dtype=t.dtype, device=t.device)
    return _dummies[key].data_ptr()


def _need_cuda(t: Tensor, name: str, dtype=None) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"kagnn_b200: {name} must be a CUDA tensor (the sm_100a path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"kagnn_b200: {name} must be {dtype}, got {t.dtype}")


def _rows(t: Tensor, name: str) -> int:
    """Leading dimension of a 2-D fp32 row-major (possibly column-sliced) tensor."""
    _need_cuda(t, name, torch.float32)
    if t.dim() != 2 or (t.size(1) > 1 and t.stride(1) != 1):
        raise ValueError(f"kagnn_b200: {name} must be 2-D with unit column stride")
    return t.stride(0) if t.size(0) > 1 else max(t.stride(0), t.size(1))


# ---------------------------------------------------------------------------------------------------
# graph ingestion
# ---------------------------------------------------------------------------------------------------
@dataclass
class CSR:
    """Destination-sorted CSR of a COO ``edge_index`` (PyG flow source_to_target)."""
    rowptr: Tensor        # (num_rows+1,) int32
    col: Tensor           # (nnz,) int32 source row of each entry
    perm: Tensor          # (nnz,) int32 original edge id of each entry
    num_rows: int
    num_src: int
    err_flag: Tensor      # int32 scalar: non-zero if an index was out of range
    _anchors = ("rowptr",)

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def validate(self) -> None:
        """Synchronising range check (the build itself never syncs)."""
        if int(self.err_flag.item()) != 0:
            raise IndexError("kagnn_b200: edge_index contains node ids outside [0, num_nodes)")


@_on_device
def csr_build(edge_index: Tensor, num_nodes: int, num_src_nodes: Optional[int] = None) -> CSR:
    global launch_count
    _need_cuda(edge_index, "edge_index", torch.int64)
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must have shape (2, E)")
    ei = edge_index.contiguous()
    E = ei.size(1)
    nsrc = num_nodes if num_src_nodes is None else num_src_nodes
    dev = ei.device
    rowptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    perm = torch.empty(E, dtype=torch.int32, device=dev)
    wbytes = L.lib().kagnn_csr_build_workspace(E, num_nodes)
    ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    L.check(L.lib().kagnn_csr_build(_p(ei), E, num_nodes, nsrc, _p(rowptr), _p(col), _p(perm), _p(ws), wbytes, _stream()),
            "csr_build")
    launch_count += 5 if E else 0
    # the range-check flag is the first word of the workspace: keep a 4-byte copy so the workspace (16 B per edge + sort
    # scratch) is released as soon as the build has run instead of living as long as the cached graph
    return CSR(rowptr, col, perm, num_nodes, nsrc, ws[:4].view(torch.int32).clone())


@_on_device
def segment_ptr(batch: Tensor, num_graphs: int) -> Tensor:
    global launch_count
    _need_cuda(batch, "batch", torch.int64)
    b = batch.contiguous()
    ptr = torch.empty(num_graphs + 1, dtype=torch.int32, device=b.device)
    L.check(L.lib().kagnn_segment_ptr(_p(b), b.numel(), num_graphs, _p(ptr), _stream()), "segment_ptr")
    launch_count += 1
    return ptr


@_on_device
def gcn_norm(csr: CSR, edge_weight_csr: Optional[Tensor] = None):
    """PyG gcn_norm on the CSR -> (edge_weight (nnz), self_weight (N), dinv (N))."""
    global launch_count
    dev = csr.rowptr.device
    n = csr.num_rows
    w = torch.empty(csr.nnz, dtype=torch.float32, device=dev)
    sw = torch.empty(n, dtype=torch.float32, device=dev)
    dinv = torch.empty(n, dtype=torch.float32, device=dev)
    if edge_weight_csr is not None:
        _need_cuda(edge_weight_csr, "edge_weight", torch.float32)
        edge_weight_csr = edge_weight_csr.contiguous()
    L.check(L.lib().kagnn_gcn_norm(_p(csr.rowptr), _p(csr.col), n, _p(edge_weight_csr), _p(w), _p(sw), _p(dinv), _stream()),
            "gcn_norm")
    launch_count += 2
    return w, sw, dinv


@_on_device
def gcn_degree(csr: CSR, edge_weight_csr: Optional[Tensor] = None):
    """First half of gcn_norm: (self_weight (N), dinv (N)) of the destination rows."""
    global launch_count
    dev = csr.rowptr.device
    sw = torch.empty(csr.num_rows, dtype=torch.float32, device=dev)
    dinv = torch.empty(csr.num_rows, dtype=torch.float32, device=dev)
    L.check(L.lib().kagnn_gcn_degree(_p(csr.rowptr), _p(csr.col), csr.num_rows, _p(edge_weight_csr), _p(sw), _p(dinv),
                                     _stream()), "gcn_degree")
    launch_count += 1
    return sw, dinv


@_on_device
def gcn_edge_weight(csr: CSR, dinv_src: Tensor, dinv_dst: Tensor, edge_weight_csr: Optional[Tensor] = None) -> Tensor:
    """Second half of gcn_norm: w_e = dinv_src[col_e] * w_e * dinv_dst[row]; dinv_src covers halo rows too."""
    global launch_count
    _need_cuda(dinv_src, "dinv_src", torch.float32)
    _need_cuda(dinv_dst, "dinv_dst", torch.float32)
    if dinv_src.numel() < csr.num_src or dinv_dst.numel() < csr.num_rows:
        raise ValueError("dinv vectors shorter than the graph")
    w = torch.empty(csr.nnz, dtype=torch.float32, device=csr.rowptr.device)
    L.check(L.lib().kagnn_gcn_edge_weight(_p(csr.rowptr), _p(csr.col), csr.num_rows, _p(edge_weight_csr),
                                          _p(dinv_src.contiguous()), _p(dinv_dst.contiguous()), _p(w), _stream()),
            "gcn_edge_weight")
    launch_count += 1
    return w


@_on_device
def gather_rows(x: Tensor, index: Optional[Tensor], out: Optional[Tensor] = None, num_rows: Optional[int] = None) -> Tensor:
    """out[r] = x[index[r]] (halo send-buffer packing); ``index=None`` copies the first ``num_rows`` rows."""
    global launch_count
    ldx = _rows(x, "x")
    if index is not None:
        _need_cuda(index, "index", torch.int32)
        num_rows = index.numel()
    elif num_rows is None:
        num_rows = x.size(0)
    if out is None:
        out = torch.empty(num_rows, x.size(1), dtype=torch.float32, device=x.device)
    ldo = _rows(out, "out")
    if num_rows == 0:
        return out
    L.check(L.lib().kagnn_gather_rows(_p(x), ldx, _p(index), num_rows, x.size(1), _p(out), ldo, _stream()), "gather_rows")
    launch_count += 1
    return out


@_on_device
def expand_windows(x: Tensor, windows: int, shift: float) -> Tensor:
    """(rows, windows * cols): copy w of the columns is x - w * shift (kagnn_expand_windows)."""
    global launch_count
    _need_cuda(x, "x", torch.float32)
    ldx = _rows(x, "x")
    out = torch.empty(x.size(0), windows * x.size(1), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    L.check(L.lib().kagnn_expand_windows(_p(x), ldx, x.size(0), x.size(1), int(windows), float(shift), _p(out), out.size(1), _stream()),
            "expand_windows")
    launch_count += 1
    return out


@_on_device
def layernorm_stats(x: Tensor, x_head: Optional[Tensor] = None, eps: float = 1e-5) -> Tensor:
    """(rows, 2) per-row (mean, rstd) of LayerNorm over the two-part rows [x_head | x] (kagnn_layernorm_stats)."""
    global launch_count
    ldx = _rows(x, "x")
    stats = torch.empty(x.size(0), 2, dtype=torch.float32, device=x.device)
    if x.size(0):
        L.check(L.lib().kagnn_layernorm_stats(_p(x), ldx, x.size(1), _p(x_head) if x_head is not None else None,
                                              _rows(x_head, "x_head") if x_head is not None else 0,
                                              x_head.size(1) if x_head is not None else 0, x.size(0), eps, _p(stats), _stream()),
                "layernorm_stats")
        launch_count += 1
    return stats


@_on_device
def log_softmax(x: Tensor) -> Tensor:
    """Row-wise log_softmax of (rows, classes) logits (kagnn_log_softmax_rows)."""
    global launch_count
    ldx = _rows(x, "x")
    y = torch.empty(x.size(0), x.size(1), dtype=torch.float32, device=x.device)
    if x.numel():
        L.check(L.lib().kagnn_log_softmax_rows(_p(x), ldx, x.size(0), x.size(1), _p(y), _rows(y, "y"), _stream()), "log_softmax")
        launch_count += 1
    return y


@_on_device
def batchnorm_forward(x: Tensor, bn, act: int = L.ACT_NONE) -> Tensor:
    """``bn(x)`` of a ``torch.nn.BatchNorm1d`` in training mode (or without running statistics): batch statistics, running
    estimates updated like torch does (kagnn_batchnorm_train_fwd).  Eval mode with running statistics is folded into the
    fused kernels' affine epilogue by the models and never comes here."""
    global launch_count
    ldx = _rows(x, "x")
    n, c = x.shape
    if n == 0:
        return x.clone()
    use_batch = bn.training or bn.running_mean is None
    if not use_batch:
        raise RuntimeError("eval-mode BatchNorm is folded into the fused layer (models_node._BNFold)")
    if bn.momentum is None and bn.track_running_stats and bn.running_mean is not None:
        momentum = 1.0 / float(int(bn.num_batches_tracked) + 1)          # cumulative moving average
    else:
        momentum = 0.0 if bn.momentum is None else float(bn.momentum)
    track = bn.training and bn.track_running_stats and bn.running_mean is not None
    y = torch.empty(n, c, dtype=torch.float32, device=x.device)
    wbytes = L.lib().kagnn_batchnorm_train_workspace(c)
    ws = torch.empty(wbytes, dtype=torch.uint8, device=x.device)
    L.check(L.lib().kagnn_batchnorm_train_fwd(
        _p(x), ldx, n, c, _p(bn.weight.detach()) if bn.weight is not None else None, _p(bn.bias.detach()) if bn.bias is not None else None,
        float(bn.eps), momentum, _p(bn.running_mean) if track else None, _p(bn.running_var) if track else None, act,
        _p(y), _rows(y, "y"), _p(ws), wbytes, _stream()), "batchnorm_train_fwd")
    if track:
        bn.num_batches_tracked += 1
        # the kernel updated the running statistics through raw pointers: bump their version counters like an in-place ATen
        # op would, so that caches keyed on them (models_node._BNFold) see the change
        torch.autograd.graph.increment_version(bn.running_mean)
        torch.autograd.graph.increment_version(bn.running_var)
    launch_count += 3
    return y


def _bn_momentum(bn) -> float:
    if bn.momentum is None and bn.track_running_stats and bn.running_mean is not None:
        return 1.0 / float(int(bn.num_batches_tracked) + 1)                  # cumulative moving average
    return 0.0 if bn.momentum is None else float(bn.momentum)


@_on_device
def batchnorm_dropout_forward(x: Tensor, bn, p: float, seed: int) -> Tensor:
    """``dropout(bn(x), p)`` in training mode as two launches (kagnn_bn_dropout_train_fwd): batch statistics, running estimates
    updated like torch does, Philox mask keyed by ``seed`` (regenerated by ``batchnorm_dropout_backward``)."""
    global launch_count
    ldx = _rows(x, "x")
    n, c = x.shape
    if n == 0:
        return x.clone()
    track = bn.training and bn.track_running_stats and bn.running_mean is not None
    y = torch.empty(n, c, dtype=torch.float32, device=x.device)
    wbytes = L.lib().kagnn_bn_dropout_train_workspace(c)
    ws = torch.empty(wbytes, dtype=torch.uint8, device=x.device)
    L.check(L.lib().kagnn_bn_dropout_train_fwd(
        _p(x), ldx, n, c, _p(bn.weight.detach()) if bn.weight is not None else None, _p(bn.bias.detach()) if bn.bias is not None else None,
        float(bn.eps), _bn_momentum(bn), _p(bn.running_mean) if track else None, _p(bn.running_var) if track else None, float(p),
        int(seed) & 0xFFFFFFFFFFFFFFFF, _p(y), _rows(y, "y"), _p(ws), wbytes, _stream()), "bn_dropout_train_fwd")
    if track:
        bn.num_batches_tracked += 1
        torch.autograd.graph.increment_version(bn.running_mean)
        torch.autograd.graph.increment_version(bn.running_var)
    launch_count += 3
    return y


@_on_device
def batchnorm_dropout_backward(x: Tensor, dy: Tensor, weight: Optional[Tensor], eps: float, p: float, seed: int):
    """Backward of ``batchnorm_dropout_forward`` -> (dx, d weight, d bias) (kagnn_bn_dropout_train_bwd)."""
    global launch_count
    n, c = x.shape
    dx = torch.empty(n, c, dtype=torch.float32, device=x.device)
    dw = torch.empty(c, dtype=torch.float32, device=x.device)
    db = torch.empty(c, dtype=torch.float32, device=x.device)
    wbytes = L.lib().kagnn_bn_dropout_train_workspace(c)
    ws = torch.empty(wbytes, dtype=torch.uint8, device=x.device)
    L.check(L.lib().kagnn_bn_dropout_train_bwd(_p(x), _rows(x, "x"), _p(dy), _rows(dy, "dy"), n, c, _p(weight.detach()) if weight is not None else None,
                                               float(eps), float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(dx), _rows(dx, "dx"), _p(dw), _p(db),
                                               _p(ws), wbytes, _stream()), "bn_dropout_train_bwd")
    launch_count += 3
    return dx, dw, db


@_on_device
def gather_rows_peer(peer_x: Tensor, ldx: int, rows_per_rank: int, ids: Tensor, num_cols: int, out: Optional[Tensor] = None) -> Tensor:
    """out[r] = row ids[r] % rows_per_rank of rank ids[r] // rows_per_rank, pulled over NVLink (kagnn_gather_rows_peer)."""
    global launch_count
    _need_cuda(peer_x, "peer_x", torch.int64)
    _need_cuda(ids, "ids", torch.int32)
    if out is None:
        out = torch.empty(ids.numel(), num_cols, dtype=torch.float32, device=ids.device)
    if ids.numel() == 0:
        return out
    L.check(L.lib().kagnn_gather_rows_peer(_p(peer_x), ldx, rows_per_rank, _p(ids), ids.numel(), num_cols, _p(out), _rows(out, "out"),
                                           _stream()), "gather_rows_peer")
    launch_count += 1
    return out


@_on_device
def gather_rows_peer_masked(peer_x: Tensor, ldx: int, rows_per_rank: int, need: Tensor, num_cols: int, out: Tensor) -> Tensor:
    """out[id] = row id % rows_per_rank of rank id // rows_per_rank for every id with need[id] != 0, pulled over NVLink
    (kagnn_gather_rows_peer_masked); ``out`` has one row per id (a replica of the whole matrix)."""
    global launch_count
    _need_cuda(peer_x, "peer_x", torch.int64)
    _need_cuda(need, "need", torch.uint8)
    if out.size(0) < need.numel():
        raise ValueError("out needs one row per candidate id")
    if need.numel() == 0:
        return out
    L.check(L.lib().kagnn_gather_rows_peer_masked(_p(peer_x), ldx, rows_per_rank, _p(need), need.numel(), num_cols, _p(out),
                                                  _rows(out, "out"), _stream()), "gather_rows_peer_masked")
    launch_count += 1
    return out


@_on_device
def gat_attention(h: Tensor, csr: CSR, att_src: Tensor, att_dst: Tensor, heads: int, negative_slope: float = 0.2):
    """PyG GATConv's attention coefficients for h = lin(x) of shape (N, heads * C) on a destination-sorted CSR:
    returns (edge_weight (heads, nnz), self_weight (heads, N)) -- per head the operands of the WEIGHTED aggregation
    (kagnn_gat_scores + kagnn_gat_edge_softmax)."""
    global launch_count
    ldh = _rows(h, "h")
    n, hc = h.shape
    c = hc // heads
    dev = h.device
    a_s = torch.empty(n, heads, dtype=torch.float32, device=dev)
    a_d = torch.empty(n, heads, dtype=torch.float32, device=dev)
    ats, atd = att_src.detach().reshape(-1).contiguous(), att_dst.detach().reshape(-1).contiguous()
    _need_cuda(ats, "att_src", torch.float32)
    _need_cuda(atd, "att_dst", torch.float32)
    L.check(L.lib().kagnn_gat_scores(_p(h), ldh, n, heads, c, _p(ats), _p(atd), _p(a_s), _p(a_d), _stream()), "gat_scores")
    w = torch.empty(heads, max(csr.nnz, 1), dtype=torch.float32, device=dev)
    sw = torch.empty(heads, n, dtype=torch.float32, device=dev)
    L.check(L.lib().kagnn_gat_edge_softmax(_p(csr.rowptr), C.c_void_p(_addr(csr.col)), n, csr.nnz, heads, _p(a_s), _p(a_d),
                                           float(negative_slope), _p(w), _p(sw), _stream()), "gat_edge_softmax")
    launch_count += 2
    return w[:, :csr.nnz] if csr.nnz else w[:, :0], sw


@_on_device
def gat_backward(h: Tensor, csr: CSR, dout: Tensor, att_src: Tensor, att_dst: Tensor, w: Tensor, sw: Tensor, heads: int,
                 negative_slope: float, dh: Tensor):
    """Attention part of PyG GATConv's backward (kagnn_gat_bwd): ``dh`` holds the aggregation part on entry and gets the part
    that flows through the scores added; returns (d att_src, d att_dst), each (heads * C,)."""
    global launch_count
    n, hc = h.shape
    c = hc // heads
    dev = h.device
    ats, atd = att_src.detach().reshape(-1).contiguous(), att_dst.detach().reshape(-1).contiguous()
    d_as = torch.empty(hc, dtype=torch.float32, device=dev)
    d_ad = torch.empty(hc, dtype=torch.float32, device=dev)
    nbytes = int(L.lib().kagnn_gat_bwd_workspace(n, csr.nnz, heads))
    ws = torch.empty(max(nbytes, 4) // 4 + 1, dtype=torch.float32, device=dev)
    w = w.contiguous()
    L.check(L.lib().kagnn_gat_bwd(_p(csr.rowptr), C.c_void_p(_addr(csr.col)), n, csr.nnz, heads, c, _p(h), _rows(h, "h"), _p(dout),
                                  _rows(dout, "dout"), _p(ats), _p(atd), C.c_void_p(_addr(w)) if csr.nnz else None, _p(sw),
                                  float(negative_slope), _p(ws), ws.numel() * 4, _p(dh), _rows(dh, "dh"), _p(d_as), _p(d_ad), _stream()),
            "gat_bwd")
    launch_count += 3
    return d_as, d_ad


HALO_CHUNK = 256          # rows per progress flag of gather_rows_peer_ordered (kHaloChunk in csrc/graph.cu)


@_on_device
def gather_rows_peer_ordered(peer_x: Tensor, ldx: int, rows_per_rank: int, ids: Tensor, num_cols: int, out: Tensor, flags: Tensor,
                             epoch: int, num_ctas: int) -> Tensor:
    """The pull of gather_rows_peer as a persistent kernel on the CURRENT stream that publishes per-chunk progress flags
    (kagnn_gather_rows_peer_ordered); meant to run on a side stream next to the fused layer that reads ``out`` as x_halo."""
    global launch_count
    _need_cuda(peer_x, "peer_x", torch.int64)
    _need_cuda(ids, "ids", torch.int32)
    _need_cuda(flags, "flags", torch.int32)
    if flags.numel() * HALO_CHUNK < ids.numel():
        raise ValueError("one flag per 256 halo rows")
    if ids.numel() == 0:
        return out
    L.check(L.lib().kagnn_gather_rows_peer_ordered(_p(peer_x), ldx, rows_per_rank, _p(ids), ids.numel(), num_cols, _p(out), _rows(out, "out"),
                                                   _p(flags), int(epoch), int(num_ctas), _stream()), "gather_rows_peer_ordered")
    launch_count += 1
    return out


# ---------------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------------
@_on_device
def pack_kan_weights(base_w: Optional[Tensor], spline_w: Tensor, scaler: Optional[Tensor], in_f: int, out_f: int,
                     slots: int) -> Tensor:
    global launch_count
    _need_cuda(spline_w, "spline_weight", torch.float32)
    n = L.lib().kagnn_packed_weight_elems(in_f, out_f, slots)
    packed = torch.empty(n, dtype=torch.float32, device=spline_w.device)
    bw = None if base_w is None else base_w.detach().contiguous()
    sc = None if scaler is None else scaler.detach().contiguous()
    L.check(L.lib().kagnn_pack_kan_weights(_p(bw), _p(spline_w.detach().contiguous()), _p(sc), in_f, out_f, slots, _p(packed),
                                           _stream()), "pack_kan_weights")
    launch_count += 1
    return packed


# FastKAN layers with more than eight centres as slot windows on the tensor-core kernels (fastkan.FastKANLayer._windowed_spec): OFF
# unless asked for.  The bf16 hi/lo split carries ~17 bits per operand; narrow Gaussians (denominator = range / (G - 1)) amplify
# that through a deep model: against an fp64 evaluation, a 2 x 2-layer GIN model with 12 centres came out at 1.7e-4 on the tensor
# cores, 1.0e-5 on the general fp32 kernel and 0.8e-5 in the reference's own fp32 (32 centres: 9e-3 / 1.3e-3 / 0.9e-3;
# profiles/r2_windows_accuracy.json) -- outside the 1e-4 parity bound, so the default stays the fp32 kernel.  B-spline windows
# (8.6e-6 on the same model) are always on.
_rbf_windows = os.environ.get("KAGNN_RBF_WINDOWS", "0") == "1"


def set_rbf_windows(on: bool) -> None:
    global _rbf_windows
    _rbf_windows = bool(on)


def rbf_windows_enabled() -> bool:
    return _rbf_windows


def tc_supported(basis: int, grid_size: int, spline_order: int, out_features: int) -> bool:
    """Shapes the tcgen05 kernel handles (8 slots per feature, <= 4 non-zero bases, N <= 256)."""
    if out_features > 256:
        return False
    if basis == L.BASIS_BSPLINE:
        return 1 <= spline_order <= 3 and grid_size + spline_order <= 8
    return 1 <= grid_size <= 8


@_on_device
def pack_kan_weights_tc(base_w: Optional[Tensor], spline_w: Tensor, scaler: Optional[Tensor], in_f: int, out_f: int,
                        slots: int) -> Tensor:
    """bf16 hi/lo weights in the UMMA canonical layout, chunked for the tcgen05 kernel."""
    global launch_count
    _need_cuda(spline_w, "spline_weight", torch.float32)
    nbytes = L.lib().kagnn_packed_weight_tc_bytes(in_f, out_f)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=spline_w.device)
    bw = None if base_w is None else base_w.detach().contiguous()
    sc = None if scaler is None else scaler.detach().contiguous()
    L.check(L.lib().kagnn_pack_kan_weights_tc(_p(bw), _p(spline_w.detach().contiguous()), _p(sc), in_f, out_f, slots, _p(packed),
                                              _stream()), "pack_kan_weights_tc")
    launch_count += 2
    return packed


def set_path(mode: int) -> None:
    """PATH_AUTO / PATH_FP32 / PATH_TC (see kagnn_set_path)."""
    L.check(L.lib().kagnn_set_path(mode), "set_path")


def set_backward_path(mode: int) -> None:
    """0 = tcgen05 gradient kernels when the shape fits (default), 1 = fp32 CUDA-core kernels only, 2 / 3 = the tcgen05 kernels'
    alternative variants (kagnn_set_backward_path; tests compare them all)."""
    L.check(L.lib().kagnn_set_backward_path(int(mode)), "set_backward_path")


def set_precision(mode) -> None:
    """``"fp32"`` (default: bf16 hi/lo split, three products per K step, matches the reference's fp32 forward within 1e-4) or
    ``"bf16"`` (one bf16 product per K step, fp32 accumulate: BASELINE config C5) -- see kagnn_set_precision."""
    code = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16}.get(mode, mode)
    L.check(L.lib().kagnn_set_precision(int(code)), "set_precision")


def get_precision() -> str:
    return "bf16" if L.lib().kagnn_get_precision() == L.PREC_BF16 else "fp32"


def set_tc_variant(variant: int) -> None:
    """0 = pipelined tcgen05 kernel first (default), 1 = only the shared-memory-A tcgen05 kernel."""
    global _tc_variant
    L.check(L.lib().kagnn_set_tc_variant(variant), "set_tc_variant")
    _tc_variant = variant


def launch_counters():
    a, b = C.c_int64(0), C.c_int64(0)
    L.lib().kagnn_get_launch_counters(C.byref(a), C.byref(b))
    return {"tc": a.value, "fp32": b.value, "tc2": L.lib().kagnn_get_tc2_launches()}


# ---------------------------------------------------------------------------------------------------
# the fused layer
# ---------------------------------------------------------------------------------------------------
@dataclass
class KanLayerSpec:
    basis: int
    in_features: int
    out_features: int
    grid_size: int
    spline_order: int
    t0: float
    h: float
    inv_denominator: float
    packed_w: Tensor
    base_bias: Optional[Tensor] = None
    ln_weight: Optional[Tensor] = None
    ln_bias: Optional[Tensor] = None
    packed_w_tc: Optional[Tensor] = None
    ln_stats: Optional[Tensor] = None     # per-row (mean, rstd) of the first layer's LayerNorm (layernorm_stats); per call
    # B-spline layers with more than eight slots per feature (G + k > 8) as `windows` virtual features of eight slots each:
    # in_features = windows * real inputs, the input is expand_windows(x, windows, window_shift = 8 h) (ekan.KANLinear.kernel_spec);
    # virt_* = the repacked (out, windows * in, 8) / (out, windows * in) weights, for the chain rule of the backward
    windows: int = 1
    window_shift: float = 0.0
    virt_spline: Optional[Tensor] = None
    virt_scaler: Optional[Tensor] = None
    _anchors = ("packed_w",)

    def fill(self, s: L.KagnnKanLayer) -> None:
        s.basis, s.in_features, s.out_features = self.basis, self.in_features, self.out_features
        s.grid_size, s.spline_order = self.grid_size, self.spline_order
        s.t0, s.h, s.inv_denominator = self.t0, self.h, self.inv_denominator
        s.packed_w = self.packed_w.data_ptr()
        s.base_bias = None if self.base_bias is None else self.base_bias.data_ptr()
        s.ln_weight = None if self.ln_weight is None else self.ln_weight.data_ptr()
        s.ln_bias = None if self.ln_bias is None else self.ln_bias.data_ptr()
        s.packed_w_tc = None if self.packed_w_tc is None else self.packed_w_tc.data_ptr()
        s.ln_stats = None if self.ln_stats is None else self.ln_stats.data_ptr()


@dataclass
class Affine:
    scale: Optional[Tensor] = None
    shift: Optional[Tensor] = None
    act: int = L.ACT_NONE
    _anchors = ("scale", "shift")

    def struct(self) -> L.KagnnAffine:
        for t, n in ((self.scale, "scale"), (self.shift, "shift")):
            if t is not None:
                _need_cuda(t, n, torch.float32)
                if not t.is_contiguous():
                    raise ValueError("affine vectors must be contiguous")
        return L.KagnnAffine(None if self.scale is None else self.scale.data_ptr(),
                             None if self.shift is None else self.shift.data_ptr(), self.act, 0)


@dataclass
class AggSpec:
    mode: int
    x: Tensor
    rowptr: Optional[Tensor] = None
    col: Optional[Tensor] = None
    edge_weight: Optional[Tensor] = None
    self_weight: Optional[Tensor] = None
    self_scale: float = 1.0
    edge_feat: Optional[Tensor] = None
    edge_row: Optional[Tensor] = None
    src_index: Optional[Tensor] = None
    x_halo: Optional[Tensor] = None       # node-sharded graphs: source rows >= x.size(0) live here (dist.py)
    peer_x: Optional[Tensor] = None       # node-sharded graphs, in-kernel NVLink gather: (world,) int64 peer base pointers
    rows_per_rank: int = 0
    x_head: Optional[Tensor] = None       # AGG_NONE two-part rows: logical row = [x_head[r] | x[r]] (skip concat without the copy)
    halo_need: Optional[Tensor] = None    # x_halo filled while the layer runs (gather_rows_peer_ordered): per tile, halo rows needed so far
    halo_flags: Optional[Tensor] = None   # ... per chunk of 256 halo rows: == halo_epoch once the chunk has landed
    halo_epoch: int = 0
    reserve_sms: int = 0
    push_y: Optional[Tensor] = None       # output rows pushed to the peers while the layer runs: (num_push,) int64 destination pointers
    push_ld: int = 0                      # ... leading dimension of the destinations (floats)
    push_mask: Optional[Tensor] = None    # ... optional (num_rows,) uint8: bit i set <=> destination i needs the row
    _anchors = ("x", "rowptr", "x_head", "x_halo", "edge_feat")

    def __post_init__(self) -> None:
        # the C ABI takes row-major matrices with a leading dimension: a transposed / column-strided view (which the
        # reference's ATen ops accept) is made row-major here -- a copy, no arithmetic
        for name in ("x", "x_head", "x_halo", "edge_feat"):
            t = getattr(self, name)
            if isinstance(t, torch.Tensor) and t.dim() == 2 and t.size(1) > 1 and t.stride(1) != 1:
                setattr(self, name, t.contiguous())

    def struct(self) -> L.KagnnAggregate:
        ldx = _rows(self.x, "x")
        s = L.KagnnAggregate()
        s.mode, s.num_cols = self.mode, self.x.size(1)
        if self.x_head is not None:
            if self.x_head.size(0) != self.x.size(0):
                raise ValueError("x_head must have as many rows as x")
            s.ld_head = _rows(self.x_head, "x_head")
            s.x_head = _addr(self.x_head)
            s.num_head_cols = self.x_head.size(1)
            s.num_cols = self.x_head.size(1) + self.x.size(1)
        s.x, s.ldx = _addr(self.x), ldx
        for name in ("src_index", "rowptr", "col", "edge_row"):
            t = getattr(self, name)
            if t is not None:
                _need_cuda(t, name, torch.int32)
                setattr(s, name, _addr(t))
        for name in ("edge_weight", "self_weight"):
            t = getattr(self, name)
            if t is not None:
                _need_cuda(t, name, torch.float32)
                setattr(s, name, _addr(t))
        s.self_scale = float(self.self_scale)
        if self.edge_feat is not None:
            s.ld_edge = _rows(self.edge_feat, "edge_feat")
            s.edge_feat = _addr(self.edge_feat)
        if self.x_halo is not None:
            if self.x_halo.size(1) != self.x.size(1):
                raise ValueError("x_halo must have the same number of columns as x")
            s.ld_halo = _rows(self.x_halo, "x_halo")
            s.x_halo = _addr(self.x_halo)
            s.num_local_src = self.x.size(0)
        if self.halo_flags is not None:
            if self.x_halo is None or self.halo_need is None:
                raise ValueError("halo_flags needs x_halo and halo_need")
            _need_cuda(self.halo_flags, "halo_flags", torch.int32)
            _need_cuda(self.halo_need, "halo_need", torch.int32)
            s.halo_flags, s.halo_need = _addr(self.halo_flags), _addr(self.halo_need)
            s.halo_epoch, s.reserve_sms = int(self.halo_epoch), int(self.reserve_sms)
        if self.push_y is not None and self.push_y.numel():
            _need_cuda(self.push_y, "push_y", torch.int64)
            if self.push_ld <= 0 or self.push_y.numel() > 8:
                raise ValueError("push_y needs push_ld and at most 8 destinations")
            s.push_y, s.ld_push, s.num_push = _addr(self.push_y), int(self.push_ld), int(self.push_y.numel())
            if self.push_mask is not None:
                _need_cuda(self.push_mask, "push_mask", torch.uint8)
                s.push_mask = _addr(self.push_mask)
        if self.peer_x is not None:
            _need_cuda(self.peer_x, "peer_x", torch.int64)
            if self.rows_per_rank <= 0:
                raise ValueError("peer_x needs rows_per_rank")
            s.peer_x = _addr(self.peer_x)
            s.rows_per_rank = int(self.rows_per_rank)
            s.num_ranks = int(self.peer_x.numel())
        return s


def _launch_fused(agg: AggSpec, num_rows: int, layers: Sequence[KanLayerSpec], pre, post, agg_out, y) -> int:
    if fused_timing is not None:
        label = "agg%d[%d]%s" % (agg.mode, agg.x.size(1) + (agg.x_head.size(1) if agg.x_head is not None else 0), "".join("->%d" % sp.out_features for sp in layers))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        code = _launch_fused_raw(agg, num_rows, layers, pre, post, agg_out, y)
        ev1.record()
        fused_timing.append((label, ev0, ev1))
        return code
    return _launch_fused_raw(agg, num_rows, layers, pre, post, agg_out, y)


def _launch_fused_raw(agg: AggSpec, num_rows: int, layers: Sequence[KanLayerSpec], pre, post, agg_out, y) -> int:
    arr = (L.KagnnKanLayer * max(len(layers), 1))()
    for i, sp in enumerate(layers):
        sp.fill(arr[i])
    a = agg.struct()
    pre_s = pre.struct() if pre is not None else None
    post_s = post.struct() if post is not None else None
    return L.lib().kagnn_fused_layer_fwd(
        C.byref(a), num_rows, C.byref(pre_s) if pre_s is not None else None,
        _p(agg_out), _rows(agg_out, "agg_out") if agg_out is not None else 0,
        len(layers), arr, C.byref(post_s) if post_s is not None else None,
        _p(y), _rows(y, "y") if y is not None else 0, _stream())


@_on_device
def fused_layer(agg: AggSpec, num_rows: int, layers: Sequence[KanLayerSpec], pre: Optional[Affine] = None,
                post: Optional[Affine] = None, agg_out: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Optional[Tensor]:
    """tile = aggregate(x) -> pre -> [agg_out] -> KAN chain -> post -> out, in one launch
    (kagnn_fused_layer_fwd).  Returns ``out`` (allocated if None and there is at least one layer)."""
    global launch_count
    if len(layers) > L.MAX_LAYERS:
        raise NotImplementedError(f"KAN chains deeper than {L.MAX_LAYERS} are not supported")
    dev = agg.x.device
    if layers and out is None:
        out = torch.empty(num_rows, layers[-1].out_features, dtype=torch.float32, device=dev)
    if not layers and agg_out is None:
        agg_out = torch.empty(num_rows, agg.x.size(1), dtype=torch.float32, device=dev)
    if num_rows == 0:
        return out if layers else agg_out
    if any(sp.windows > 1 for sp in layers):
        return _windowed_chain(agg, num_rows, layers, pre, post, agg_out, out)
    if len(layers) > 1 and any(sp.out_features > 128 for sp in layers) and all(sp.packed_w_tc is not None for sp in layers) \
            and _tc_variant_is_pipelined():
        return _wide_chain(agg, num_rows, layers, pre, post, agg_out, out)
    wide_single = len(layers) == 1 and layers[0].out_features > 128
    if (layers and layers[0].basis == L.BASIS_RBF and layers[0].ln_weight is not None and layers[0].ln_stats is None
            and agg.mode == L.AGG_NONE and pre is None and agg.src_index is None and layers[0].packed_w_tc is not None
            and layers[0].in_features > (64 if wide_single else 128) and (wide_single or all(sp.out_features <= 128 for sp in layers))):
        # FastKAN over rows wider than one tile unit (the skip-concat read-out; every input of a wide layer, whose units are 64
        # columns): LayerNorm statistics by a small pre-pass so that the pipelined kernel can stream the row unit by unit
        stats = layernorm_stats(agg.x, agg.x_head)
        layers = [dataclasses.replace(layers[0], ln_stats=stats)] + list(layers[1:])
    code = _launch_fused(agg, num_rows, layers, pre, post, agg_out, out)
    launch_count += 1
    if code == L.E_UNSUPPORTED and agg.x_head is not None:
        # two-part rows not available for this shape / kernel: materialise the concatenation (two strided row copies)
        cat = torch.empty(agg.x.size(0), agg.x_head.size(1) + agg.x.size(1), dtype=torch.float32, device=dev)
        gather_rows(agg.x_head, None, out=cat[:, :agg.x_head.size(1)])
        gather_rows(agg.x, None, out=cat[:, agg.x_head.size(1):])
        agg = AggSpec(agg.mode, cat)
        code = _launch_fused(agg, num_rows, layers, pre, post, agg_out, out)
        launch_count += 1
    if code == L.E_UNSUPPORTED and layers and (agg.mode != L.AGG_NONE or pre is not None or agg.src_index is not None):
        # the aggregated tile is too wide for shared memory: aggregate to HBM, then stream the chain
        tmp = agg_out if agg_out is not None else torch.empty(num_rows, agg.x.size(1), dtype=torch.float32, device=dev)
        L.check(_launch_fused(agg, num_rows, [], pre, None, tmp, None), "fused_layer(aggregate)")
        code = _launch_fused(AggSpec(L.AGG_NONE, tmp), num_rows, layers, None, post, None, out)
        launch_count += 1
    L.check(code, "fused_layer")
    return out if layers else agg_out


_tc_variant = 0


def _tc_variant_is_pipelined() -> bool:
    return _tc_variant == 0


def _windowed_chain(agg: AggSpec, num_rows: int, layers, pre, post, agg_out, out):
    """Chains with a B-spline layer of more than eight slots per feature (KanLayerSpec.windows > 1): the aggregation (if any)
    and every layer are their own launches, because the input of a windowed layer is expanded in HBM first."""
    dev = agg.x.device
    plain = agg.mode == L.AGG_NONE and pre is None and agg.src_index is None and agg.x_head is None
    if plain and agg_out is None:
        cur = agg.x
    else:
        cur = agg_out if agg_out is not None else torch.empty(num_rows, layers[0].in_features // layers[0].windows, dtype=torch.float32, device=dev)
        fused_layer(agg, num_rows, [], pre=pre, agg_out=cur)
    for i, sp in enumerate(layers):
        last = i == len(layers) - 1
        xin = expand_windows(cur, sp.windows, sp.window_shift) if sp.windows > 1 else cur
        y = out if last else torch.empty(num_rows, sp.out_features, dtype=torch.float32, device=dev)
        cur = fused_layer(AggSpec(L.AGG_NONE, xin), num_rows, [dataclasses.replace(sp, windows=1)], post=post if last else None, out=y)
    return cur


def _wide_chain(agg: AggSpec, num_rows: int, layers, pre, post, agg_out, out):
    """KAN / FastKAN chains with a layer wider than 128 outputs (BASELINE config C5: hidden 256).  The pipelined kernel keeps
    chained activations in tensor memory, which holds two 128-column accumulators or ONE of 256: such a chain runs as one launch
    per layer with the (rows x width) activations passing through HBM in between -- a few tens of MB against hundreds of GFLOP.
    An aggregation in front of a layer that needs whole-row LayerNorm statistics (FastKAN, inputs wider than one 64-column tile
    unit) becomes its own launch too, so that the statistics pre-pass sees the aggregated rows."""
    dev = agg.x.device
    first = layers[0]
    needs_stats = first.basis == L.BASIS_RBF and first.ln_weight is not None and first.in_features > 64
    cur_agg, cur_pre = agg, pre
    if needs_stats and (agg.mode != L.AGG_NONE or pre is not None or agg.src_index is not None or agg.x_head is not None):
        tmp = agg_out if agg_out is not None else torch.empty(num_rows, first.in_features, dtype=torch.float32, device=dev)
        fused_layer(agg, num_rows, [], pre=pre, agg_out=tmp)
        cur_agg, cur_pre, agg_out = AggSpec(L.AGG_NONE, tmp), None, None
    x = None
    for i, sp in enumerate(layers):
        last = i == len(layers) - 1
        a = cur_agg if i == 0 else AggSpec(L.AGG_NONE, x)
        y = out if last else torch.empty(num_rows, sp.out_features, dtype=torch.float32, device=dev)
        x = fused_layer(a, num_rows, [sp], pre=cur_pre if i == 0 else None, post=post if last else None,
                        agg_out=agg_out if i == 0 else None, out=y)
    return x


@_on_device
def tc_selftest(a: Tensor, b: Tensor, nprod: int = 3) -> Tensor:
    """D = A @ B^T for one 128-row tile through the tcgen05 path (kagnn_tc_selftest); tests only."""
    global launch_count
    _need_cuda(a, "A", torch.float32)
    _need_cuda(b, "B", torch.float32)
    a, b = a.contiguous(), b.contiguous()
    if a.size(0) != 128 or a.size(1) != b.size(1):
        raise ValueError("A must be (128, K) and B (N, K)")
    n, k = b.size(0), b.size(1)
    d = torch.empty(128, n, dtype=torch.float32, device=a.device)
    wb = L.lib().kagnn_tc_selftest_workspace(n, k)
    ws = torch.empty(max(wb, 16), dtype=torch.uint8, device=a.device)
    L.check(L.lib().kagnn_tc_selftest(_p(a), _p(b), n, k, _p(d), nprod, _p(ws), wb, _stream()), "tc_selftest")
    launch_count += 2
    return d


# ---------------------------------------------------------------------------------------------------
# backward (include/kagnn_b200.h "Backward"; used by kagnn_b200/autograd.py)
# ---------------------------------------------------------------------------------------------------
def _layer_struct(spec: KanLayerSpec) -> L.KagnnKanLayer:
    s = L.KagnnKanLayer()
    spec.fill(s)
    return s


@_on_device
def kan_bwd_input(spec: KanLayerSpec, x: Tensor, dy: Tensor) -> Tensor:
    """d loss / d x of one B-spline KAN layer (kagnn_kan_bwd_input)."""
    global launch_count
    if spec.windows > 1:                     # d x = sum over the windows of d (x - w shift)
        dxw = kan_bwd_input(dataclasses.replace(spec, windows=1), expand_windows(x, spec.windows, spec.window_shift), dy)
        return dxw.view(x.size(0), spec.windows, x.size(1)).sum(1)
    ldx, ld_dy = _rows(x, "x"), _rows(dy, "dy")
    dx = torch.empty(x.size(0), spec.in_features, dtype=torch.float32, device=x.device)
    s = _layer_struct(spec)
    L.check(L.lib().kagnn_kan_bwd_input(C.byref(s), _p(x), ldx, _p(dy), ld_dy, x.size(0), _p(dx), _rows(dx, "dx"), _stream()),
            "kan_bwd_input")
    launch_count += 1
    return dx


@_on_device
def kan_bwd_weights(spec: KanLayerSpec, x: Tensor, dy: Tensor) -> Tensor:
    """Gradient of the packed fp32 weights [in][slots+1][out_pad4] (kagnn_kan_bwd_weights); of the VIRTUAL layer's packing when the
    layer is windowed (kan_unpack_windowed_grads folds it back)."""
    global launch_count
    if spec.windows > 1:
        return kan_bwd_weights(dataclasses.replace(spec, windows=1), expand_windows(x, spec.windows, spec.window_shift), dy)
    ldx, ld_dy = _rows(x, "x"), _rows(dy, "dy")
    d_packed = torch.empty_like(spec.packed_w)
    s = _layer_struct(spec)
    L.check(L.lib().kagnn_kan_bwd_weights(C.byref(s), _p(x), ldx, _p(dy), ld_dy, x.size(0), _p(d_packed), _stream()),
            "kan_bwd_weights")
    launch_count += 1
    return d_packed


@_on_device
def kan_unpack_weight_grads(d_packed: Tensor, spline_w: Tensor, scaler: Optional[Tensor], need_base: bool = True):
    """d_packed -> (d base_weight | None, d spline_weight, d spline_scaler | None) (kagnn_kan_unpack_weight_grads)."""
    global launch_count
    out_f, in_f, slots = spline_w.shape
    sw = spline_w.detach().contiguous()
    sc = None if scaler is None else scaler.detach().contiguous()
    dev = spline_w.device
    d_base = torch.empty(out_f, in_f, dtype=torch.float32, device=dev) if need_base else None
    d_spline = torch.empty(out_f, in_f, slots, dtype=torch.float32, device=dev)
    d_scaler = torch.empty(out_f, in_f, dtype=torch.float32, device=dev) if sc is not None else None
    L.check(L.lib().kagnn_kan_unpack_weight_grads(_p(d_packed), _p(sw), _p(sc), in_f, out_f, slots, _p(d_base), _p(d_spline),
                                                  _p(d_scaler), _stream()), "kan_unpack_weight_grads")
    launch_count += 1
    return d_base, d_spline, d_scaler


def kan_unpack_windowed_grads(d_packed: Tensor, spec: KanLayerSpec, slots: int):
    """Gradients of a windowed layer's REAL parameters from the gradient of its virtual packing: the chain rule through
    scaled_spline_weight on the virtual (out, windows * in, 8) weights, then window w's slots go back to slots 8 w .. 8 w + 7,
    the base weight is window 0's, the scaler's gradient the sum over the windows."""
    d_base_v, d_spline_v, d_scaler_v = kan_unpack_weight_grads(d_packed, spec.virt_spline, spec.virt_scaler)
    out_f, vin, _ = spec.virt_spline.shape
    w, f = spec.windows, vin // spec.windows
    d_spline = d_spline_v.view(out_f, w, f, 8).permute(0, 2, 1, 3).reshape(out_f, f, 8 * w)[:, :, :slots].contiguous()
    d_base = d_base_v[:, :f].contiguous()
    d_scaler = None if d_scaler_v is None else d_scaler_v.view(out_f, w, f).sum(1)
    return d_base, d_spline, d_scaler


@_on_device
def batchnorm_backward(x: Tensor, dy: Tensor, weight: Optional[Tensor], eps: float):
    """Backward of training-mode BatchNorm1d -> (dx, d weight, d bias) (kagnn_batchnorm_train_bwd)."""
    global launch_count
    ldx, ld_dy = _rows(x, "x"), _rows(dy, "dy")
    n, c = x.shape
    dx = torch.empty(n, c, dtype=torch.float32, device=x.device)
    dw = torch.empty(c, dtype=torch.float32, device=x.device)
    db = torch.empty(c, dtype=torch.float32, device=x.device)
    wbytes = L.lib().kagnn_batchnorm_bwd_workspace(c)
    ws = torch.empty(wbytes, dtype=torch.uint8, device=x.device)
    L.check(L.lib().kagnn_batchnorm_train_bwd(_p(x), ldx, _p(dy), ld_dy, n, c, _p(weight.detach()) if weight is not None else None,
                                              float(eps), _p(dx), _rows(dx, "dx"), _p(dw), _p(db), _p(ws), wbytes, _stream()),
            "batchnorm_train_bwd")
    launch_count += 2
    return dx, dw, db


@_on_device
def column_sums(x: Tensor) -> Tensor:
    global launch_count
    ldx = _rows(x, "x")
    out = torch.empty(x.size(1), dtype=torch.float32, device=x.device)
    L.check(L.lib().kagnn_column_sums(_p(x), ldx, x.size(0), x.size(1), _p(out), _stream()), "column_sums")
    launch_count += 1
    return out


@_on_device
def log_softmax_backward(y: Tensor, dy: Tensor) -> Tensor:
    global launch_count
    dx = torch.empty(y.size(0), y.size(1), dtype=torch.float32, device=y.device)
    if y.numel():
        L.check(L.lib().kagnn_log_softmax_bwd(_p(y), _rows(y, "y"), _p(dy), _rows(dy, "dy"), y.size(0), y.size(1), _p(dx),
                                              _rows(dx, "dx"), _stream()), "log_softmax_bwd")
        launch_count += 1
    return dx


@_on_device
def silu_forward(x: Tensor) -> Tensor:
    global launch_count
    y = torch.empty(x.size(0), x.size(1), dtype=torch.float32, device=x.device)
    if x.numel():
        L.check(L.lib().kagnn_silu_fwd(_p(x), _rows(x, "x"), x.size(0), x.size(1), _p(y), _rows(y, "y"), _stream()), "silu_fwd")
        launch_count += 1
    return y


@_on_device
def silu_backward(x: Tensor, dy: Tensor) -> Tensor:
    global launch_count
    dx = torch.empty(x.size(0), x.size(1), dtype=torch.float32, device=x.device)
    if x.numel():
        L.check(L.lib().kagnn_silu_bwd(_p(x), _rows(x, "x"), _p(dy), _rows(dy, "dy"), x.size(0), x.size(1), _p(dx), _rows(dx, "dx"),
                                       _stream()), "silu_bwd")
        launch_count += 1
    return dx


@_on_device
def segment_pool_backward(d_pooled: Tensor, ptr: Tensor, batch: Tensor, num_rows: int, mean: bool) -> Tensor:
    global launch_count
    _need_cuda(ptr, "segment_ptr", torch.int32)
    _need_cuda(batch, "batch", torch.int64)
    cols = d_pooled.size(1)
    dx = torch.empty(num_rows, cols, dtype=torch.float32, device=d_pooled.device)
    if num_rows:
        L.check(L.lib().kagnn_segment_pool_bwd(_p(d_pooled), _rows(d_pooled, "d_pooled"), _p(ptr), _p(batch.contiguous()), num_rows, cols,
                                               1 if mean else 0, _p(dx), _rows(dx, "dx"), _stream()), "segment_pool_bwd")
        launch_count += 1
    return dx


@_on_device
def rbf_bwd_input(spec: KanLayerSpec, x: Tensor, stats: Optional[Tensor], dy: Tensor):
    """FastKAN layer: (dz, dx_base) with a LayerNorm (stats given), else (complete dx, None) (kagnn_rbf_bwd_input)."""
    global launch_count
    n = x.size(0)
    dz = torch.empty(n, spec.in_features, dtype=torch.float32, device=x.device)
    dxb = torch.empty(n, spec.in_features, dtype=torch.float32, device=x.device) if stats is not None else None
    s = _layer_struct(spec)
    L.check(L.lib().kagnn_rbf_bwd_input(C.byref(s), _p(x), _rows(x, "x"), _p(stats), _p(dy), _rows(dy, "dy"), n, _p(dz), _rows(dz, "dz"),
                                        _p(dxb), _rows(dxb, "dxb") if dxb is not None else 0, _stream()), "rbf_bwd_input")
    launch_count += 1
    return dz, dxb


@_on_device
def rbf_bwd_weights(spec: KanLayerSpec, x: Tensor, stats: Optional[Tensor], dy: Tensor) -> Tensor:
    """Gradient of the packed FastKAN weights [in][G+1][out_pad4] (kagnn_rbf_bwd_weights)."""
    global launch_count
    d_packed = torch.empty_like(spec.packed_w)
    s = _layer_struct(spec)
    L.check(L.lib().kagnn_rbf_bwd_weights(C.byref(s), _p(x), _rows(x, "x"), _p(stats), _p(dy), _rows(dy, "dy"), x.size(0), _p(d_packed),
                                          _stream()), "rbf_bwd_weights")
    launch_count += 1
    return d_packed


@_on_device
def layernorm_backward(x: Tensor, stats: Tensor, ln_weight: Optional[Tensor], dz: Tensor, dx_base: Optional[Tensor], affine: bool):
    """LayerNorm backward -> (dx, d weight | None, d bias | None) (kagnn_layernorm_bwd)."""
    global launch_count
    n, c = x.shape
    dx = torch.empty(n, c, dtype=torch.float32, device=x.device)
    dw = torch.empty(c, dtype=torch.float32, device=x.device) if affine else None
    db = torch.empty(c, dtype=torch.float32, device=x.device) if affine else None
    L.check(L.lib().kagnn_layernorm_bwd(_p(x), _rows(x, "x"), _p(stats), _p(ln_weight.detach()) if ln_weight is not None else None,
                                        _p(dz), _rows(dz, "dz"), _p(dx_base), _rows(dx_base, "dx_base") if dx_base is not None else 0,
                                        n, c, _p(dx), _rows(dx, "dx"), _p(dw), _p(db), _stream()), "layernorm_bwd")
    launch_count += 2 if affine else 1
    return dx, dw, db


@_on_device
def gine_backward(x: Tensor, edge_feat: Tensor, edge_index: Tensor, da: Tensor, self_scale: float, need_edge_grad: bool = True):
    """Backward of the GINE aggregation on the COO edge list -> (dx, d edge_feat | None) (kagnn_gine_bwd)."""
    global launch_count
    _need_cuda(edge_index, "edge_index", torch.int64)
    ei = edge_index.contiguous()
    n, c = x.shape
    e = ei.size(1)
    dx = torch.empty(n, c, dtype=torch.float32, device=x.device)
    de = torch.empty(e, c, dtype=torch.float32, device=x.device) if need_edge_grad else None
    L.check(L.lib().kagnn_gine_bwd(_p(x), _rows(x, "x"), _p(edge_feat) if e else None, _rows(edge_feat, "edge_feat") if e else c,
                                   _p(ei) if e else None, e, n, c, _p(da), _rows(da, "da"), float(self_scale), _p(dx), _rows(dx, "dx"),
                                   _p(de) if (de is not None and e) else None, _rows(de, "d_edge") if (de is not None and e) else c,
                                   _stream()), "gine_bwd")
    launch_count += 2 if e else 1
    return dx, de
