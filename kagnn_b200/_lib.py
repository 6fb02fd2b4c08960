"""ctypes binding of libkagnn_b200.so (the C ABI declared in include/kagnn_b200.h).

There is no fallback of any kind: if the library is missing or fails to load, every op raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KAGNN_LIB") or os.path.join(_HERE, "lib", "libkagnn_b200.so")

# enums (mirror include/kagnn_b200.h)
BASIS_BSPLINE, BASIS_RBF = 0, 1
AGG_NONE, AGG_GIN, AGG_GINE, AGG_WEIGHTED, AGG_SEGMENT_SUM, AGG_SEGMENT_MEAN = range(6)
ACT_NONE, ACT_SILU = 0, 1
PATH_AUTO, PATH_FP32, PATH_TC = 0, 1, 2
PREC_FP32, PREC_BF16 = 0, 1
MAX_LAYERS = 8

_ERRORS = {-1: ValueError, -2: NotImplementedError, -3: ValueError, -4: RuntimeError, -5: RuntimeError, -6: IndexError}
E_UNSUPPORTED = -2


class KagnnAffine(C.Structure):
    _fields_ = [("scale", C.c_void_p), ("shift", C.c_void_p), ("act", C.c_int32), ("_pad", C.c_int32)]


class KagnnKanLayer(C.Structure):
    _fields_ = [
        ("basis", C.c_int32), ("in_features", C.c_int32), ("out_features", C.c_int32),
        ("grid_size", C.c_int32), ("spline_order", C.c_int32),
        ("t0", C.c_float), ("h", C.c_float), ("inv_denominator", C.c_float),
        ("packed_w", C.c_void_p), ("base_bias", C.c_void_p), ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p),
        ("packed_w_tc", C.c_void_p), ("ln_stats", C.c_void_p),
    ]


class KagnnAggregate(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("num_cols", C.c_int32),
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("src_index", C.c_void_p), ("rowptr", C.c_void_p), ("col", C.c_void_p),
        ("edge_weight", C.c_void_p), ("self_weight", C.c_void_p),
        ("self_scale", C.c_float), ("_pad", C.c_int32),
        ("edge_feat", C.c_void_p), ("ld_edge", C.c_int64), ("edge_row", C.c_void_p),
        ("x_halo", C.c_void_p), ("ld_halo", C.c_int64), ("num_local_src", C.c_int64),
        ("peer_x", C.c_void_p), ("rows_per_rank", C.c_int64), ("num_ranks", C.c_int32), ("num_head_cols", C.c_int32),
        ("x_head", C.c_void_p), ("ld_head", C.c_int64),
        ("halo_need", C.c_void_p), ("halo_flags", C.c_void_p), ("halo_epoch", C.c_int32), ("reserve_sms", C.c_int32),
        ("push_y", C.c_void_p), ("push_mask", C.c_void_p), ("ld_push", C.c_int64), ("num_push", C.c_int32), ("_pad2", C.c_int32),
    ]


class KagnnError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None

_SIGNATURES = {
    "kagnn_version": (C.c_int, []),
    "kagnn_strerror": (C.c_char_p, [C.c_int]),
    "kagnn_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 4),
    "kagnn_csr_build_workspace": (C.c_size_t, [C.c_int64, C.c_int64]),
    "kagnn_csr_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_size_t, C.c_void_p]),
    "kagnn_segment_ptr": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "kagnn_gcn_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kagnn_gcn_degree": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kagnn_gcn_edge_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "kagnn_packed_weight_elems": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "kagnn_pack_kan_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "kagnn_fused_layer_fwd": (C.c_int, [C.POINTER(KagnnAggregate), C.c_int64, C.POINTER(KagnnAffine), C.c_void_p, C.c_int64,
                                        C.c_int32, C.POINTER(KagnnKanLayer), C.POINTER(KagnnAffine), C.c_void_p, C.c_int64,
                                        C.c_void_p]),
    "kagnn_packed_weight_tc_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "kagnn_pack_kan_weights_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "kagnn_set_path": (C.c_int, [C.c_int]),
    "kagnn_set_backward_path": (C.c_int, [C.c_int32]),
    "kagnn_gat_bwd_workspace": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int32]),
    "kagnn_gat_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p,
                                C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kagnn_get_launch_counters": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kagnn_set_tc_variant": (C.c_int, [C.c_int]),
    "kagnn_set_precision": (C.c_int, [C.c_int]),
    "kagnn_get_precision": (C.c_int, []),
    "kagnn_get_tc2_launches": (C.c_int64, []),
    "kagnn_tc_selftest_workspace": (C.c_size_t, [C.c_int32, C.c_int32]),
    "kagnn_tc_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "kagnn_layernorm_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_float,
                                        C.c_void_p, C.c_void_p]),
    "kagnn_log_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "kagnn_batchnorm_train_workspace": (C.c_size_t, [C.c_int32]),
    "kagnn_batchnorm_train_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                            C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kagnn_gather_rows_peer": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_void_p]),
    "kagnn_gather_rows_peer_masked": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                                C.c_void_p]),
    "kagnn_gather_rows_peer_ordered": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                                 C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "kagnn_bn_dropout_train_workspace": (C.c_size_t, [C.c_int32]),
    "kagnn_bn_dropout_train_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                             C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t,
                                             C.c_void_p]),
    "kagnn_bn_dropout_train_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_float,
                                             C.c_float, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                             C.c_void_p]),
    "kagnn_gat_scores": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "kagnn_gat_edge_softmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "kagnn_kan_bwd_input": (C.c_int, [C.POINTER(KagnnKanLayer), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_int64, C.c_void_p]),
    "kagnn_kan_bwd_weights": (C.c_int, [C.POINTER(KagnnKanLayer), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "kagnn_kan_unpack_weight_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "kagnn_batchnorm_bwd_workspace": (C.c_size_t, [C.c_int32]),
    "kagnn_batchnorm_train_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_float,
                                            C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kagnn_column_sums": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "kagnn_log_softmax_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                        C.c_void_p]),
    "kagnn_silu_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "kagnn_silu_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "kagnn_segment_pool_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_int64, C.c_void_p]),
    "kagnn_rbf_bwd_input": (C.c_int, [C.POINTER(KagnnKanLayer), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "kagnn_rbf_bwd_weights": (C.c_int, [C.POINTER(KagnnKanLayer), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_void_p]),
    "kagnn_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kagnn_gine_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                 C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "kagnn_gather_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "kagnn_expand_windows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_int64,
                                       C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load (once) and return the ctypes handle.  Raises if the shared object is absent: build it with
    ``python -m kagnn_b200.build`` (or ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise KagnnError(f"{LIB_PATH} not found: the sm_100a library has not been built "
                                 "(run `python -m kagnn_b200.build`); there is no fallback path")
            h = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(h, name)
                fn.restype = res
                fn.argtypes = args
            _lib = h
    return _lib


def check(code: int, what: str = "") -> None:
    if code == 0:
        return
    msg = lib().kagnn_strerror(code).decode()
    raise _ERRORS.get(code, KagnnError)(f"kagnn_b200 {what}: {msg} (code {code})")
