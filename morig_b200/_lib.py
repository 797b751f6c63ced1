"""ctypes binding of `libmorig_b200.so` (declared in include/morig_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MORIG_LIB") or os.path.join(_HERE, "libmorig_b200.so")   # MORIG_LIB: debug builds (role tracer)

c_f32p = C.c_void_p   # device pointers travel as integers
c_i32p = C.c_void_p
c_i64p = C.c_void_p


class DenseDesc(C.Structure):
    """mirror of `morig_dense_desc`"""
    _fields_ = [
        ("A", c_f32p), ("lda", C.c_int32),
        ("W", c_f32p), ("ldw", C.c_int32),
        ("bias", c_f32p), ("scale", c_f32p), ("shift", c_f32p),
        ("rowbias", c_f32p), ("ldrb", C.c_int32),
        ("batch", c_i32p),
        ("n_vtx", C.c_int32), ("n_graphs", C.c_int32),
        ("C", c_f32p), ("ldc", C.c_int32),
        ("pool", c_f32p), ("ldpool", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("relu", C.c_int32),
        ("Wtc", c_f32p), ("tc_bn", C.c_int32),
        ("tc_kind", C.c_int32), ("tc_w_inv", C.c_float),
        ("a_amax", c_f32p), ("c_amax", c_f32p),
        ("tc_w_inv_dev", c_f32p),
    ]


class EdgeDesc(C.Structure):
    """mirror of `morig_edge_desc`"""
    _fields_ = [
        ("PQ", c_f32p), ("ldpq", C.c_int32), ("p_off", C.c_int32), ("q_off", C.c_int32),
        ("rowptr", c_i32p), ("col", c_i32p), ("tgt", c_i32p),
        ("N", C.c_int32), ("E_max", C.c_int32), ("n_frames", C.c_int32), ("out_repeat", C.c_int32),
        ("W1", c_f32p), ("ldw", C.c_int32),
        ("b1", c_f32p), ("scale", c_f32p), ("shift", c_f32p),
        ("out", c_f32p), ("ldo", C.c_int32), ("out_off", C.c_int32),
        ("H", C.c_int32),
        ("W1tc", c_f32p),
        ("tc_kind", C.c_int32), ("tc_w_inv", C.c_float),
        ("pq_amax", c_f32p), ("out_amax", c_f32p),
    ]


_P, _I = C.c_void_p, C.c_int32


class EdgeLayerDesc(C.Structure):
    """mirror of `morig_edge_layer_desc`"""
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int32), ("P", C.c_void_p), ("Q", C.c_void_p), ("ldpq", C.c_int32),
                ("rowptr", C.c_void_p), ("col", C.c_void_p), ("tgt", C.c_void_p), ("n_targets", C.c_int32), ("E", C.c_int32),
                ("W", C.c_void_p), ("ldw", C.c_int32), ("bias", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("C", C.c_void_p), ("ldc", C.c_int32), ("out", C.c_void_p), ("ldo", C.c_int32), ("N", C.c_int32),
                ("K", C.c_int32)]


_SIGNATURES = {
    "morig_version": (C.c_int, []),
    "morig_last_error": (C.c_char_p, []),
    "morig_sm_count": (C.c_int, []),
    "morig_graph_prep_workspace": (C.c_size_t, [C.c_int64, C.c_int32]),
    "morig_graph_prep": (C.c_int, [c_i64p, C.c_int64, C.c_int32, c_i32p, c_i32p, c_i32p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "morig_knn_graph": (C.c_int, [c_f32p, c_i32p, C.c_int32, C.c_int32, C.c_int32, c_i64p, C.c_void_p]),
    "morig_dense_fwd": (C.c_int, [C.POINTER(DenseDesc), C.c_void_p]),
    "morig_pack_tc_f16_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "morig_pack_tc_f16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "morig_edgeconv_fwd": (C.c_int, [C.POINTER(EdgeDesc), C.c_void_p]),
    "morig_edgeconv_fwd_batch": (C.c_int, [C.POINTER(EdgeDesc), C.c_int32, C.c_void_p]),
    "morig_surface_geodesic_workspace": (C.c_size_t, [C.c_int32, C.c_int32]),
    "morig_surface_geodesic": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.c_void_p]),
    "morig_geo_ball_edges": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "morig_tpl_edges_workspace": (C.c_size_t, [C.c_int64]),
    "morig_tpl_edges": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "morig_meanshift_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "morig_temporal_attn_fwd": (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, c_f32p, C.c_void_p]),
    "morig_row_normalize": (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p, C.c_int32, C.c_int32,
                                      C.c_void_p]),
    "morig_frame_reduce": (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_f32p, C.c_int32,
                                     C.c_void_p]),
    "morig_gather_cols": (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, c_i32p, C.c_int32, C.c_int32,
                                    C.c_int32, c_f32p, C.c_int32, C.c_int32, c_f32p, C.c_void_p]),
    "morig_fill_f32": (C.c_int, [c_f32p, C.c_int64, C.c_float, C.c_void_p]),
    "morig_fill_many_f32": (C.c_int, [_P, _P, _I, C.c_float, _P]),
    "morig_fill_cut_f32": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, C.c_float, _P]),
    "morig_absmax_f32": (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p, C.c_void_p]),
    "morig_nms_meanshift_workspace": (C.c_size_t, [_I]),
    "morig_nms_meanshift": (C.c_int, [_P, _P, _I, C.c_double, C.c_double, C.c_double, _P, _P, C.c_size_t, _P]),
    "morig_nn_dist_f32": (C.c_int, [_P, _I, _P, _I, _I, _P, _P, _P]),
    "morig_nn_dist_f64": (C.c_int, [_P, _I, _P, _I, _I, _P, _P, _P]),
    "morig_chamfer_bwd_f32": (C.c_int, [_P, _I, _P, _I, _I, _P, _P, _P, _P, C.c_float, C.c_float, _P, _P]),
    "morig_info_nce_fwd": (C.c_int, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _I, C.c_float, _P, _P, _P]),
    "morig_info_nce_bwd": (C.c_int, [_P, _I, _P, _I, _P, _P, _I, _P, _P, _I, _I, _I, C.c_float, _P, _I, _P, _I, _P]),
    "morig_fps": (C.c_int, [_P, _P, _P, _P, _I, _I, _P, _P]),
    "morig_ball_query": (C.c_int, [_P, _P, _P, _P, _I, C.c_float, _I, _P, _P, _P]),
    "morig_knn_topk": (C.c_int, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P]),
    "morig_knn_interpolate": (C.c_int, [_P, _I, _P, _P, _P, _I, _I, _I, _P, _I, _P]),
    "morig_sample_surface": (C.c_int, [_P, _P, _I, _I, C.c_uint64, _P, _P, _P, _P]),
    "morig_edge_mlp_layer": (C.c_int, [C.POINTER(EdgeLayerDesc), _P]),
    # ---- training path ----
    "morig_transpose_pad_f32": (C.c_int, [_P, _I, _I, _I, _P, _I, _P]),
    "morig_wgrad_workspace": (C.c_size_t, [_I, _I, _I]),
    "morig_wgrad_f32": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _I, _P, C.c_size_t, _P]),
    "morig_colstats_workspace": (C.c_size_t, [_I, _I]),
    "morig_bn_train_fwd": (C.c_int, [_P, _I, _I, _I, _P, _P, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P,
                                     C.c_size_t, _P]),
    "morig_bn_relu_bwd": (C.c_int, [_P, _I, _P, _I, _I, _I, _P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "morig_col_affine": (C.c_int, [_P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "morig_relu_bwd": (C.c_int, [_P, _I, _P, _I, _I, _I, _P, _I, _P]),
    "morig_edge_gather_relu": (C.c_int, [_P, _I, _P, _I, _P, _P, _I, _I, _P, _I, _P, _P]),
    "morig_edge_gather_relu_bwd": (C.c_int, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P, _I, _P, _I, _P]),
    "morig_segmax_fwd": (C.c_int, [_P, _I, _P, _I, _I, _P, _I, _P, _I, _P]),
    "morig_segmax_bwd": (C.c_int, [_P, _I, _P, _I, _I, _I, _P, _I, _I, _P]),
    "morig_seg_ptr": (C.c_int, [_P, _I, _I, _P, _P]),
    "morig_row_gather": (C.c_int, [_P, _I, _P, _I, _I, _P, _I, _P]),
    "morig_seg_sum": (C.c_int, [_P, _I, _P, _I, _I, _P, _I, _P]),
    "morig_normalize_fwd": (C.c_int, [_P, _I, _I, _I, _P, _I, _P]),
    "morig_normalize_bwd": (C.c_int, [_P, _I, _P, _I, _I, _I, _P, _I, _P]),
    "morig_attn_cls_fwd": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "morig_attn_cls_bwd_workspace": (C.c_size_t, [_I, _I]),
    "morig_attn_cls_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
}

EXPORTS = tuple(_SIGNATURES)
ABI_VERSION = 7
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m morig_b200.build` "
                "(morig_b200 has no CPU / PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.morig_version() != ABI_VERSION:
            raise ImportError(f"{LIB_PATH}: ABI version {lib.morig_version()} != {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().morig_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {code}): {msg}")


def stream_ptr(device=None) -> int:
    """raw handle of torch's current stream on `device` (default: the current device -- every top-level forward of
    this package runs under `torch.cuda.device(<device of its inputs>)`, see FusedModule.__call__)"""
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"morig_b200: `{name}` must live on a CUDA device (got {t.device}); "
                           "there is no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"morig_b200: `{name}` must be {dtype} (got {t.dtype})")
    return t.contiguous()
