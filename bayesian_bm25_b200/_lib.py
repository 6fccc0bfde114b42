"""ctypes binding of libbb25.so (include/bb25.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is usable, the first call that needs it raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BB25_LIB") or os.path.join(_HERE, "libbb25.so")  # BB25_LIB: tuning builds only

GATING = {"none": 0, "relu": 1, "swish": 2, "gelu": 3, "softplus": 4}


class Params(C.Structure):
    """struct bb25_params"""

    _fields_ = [
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("has_base_rate", C.c_int),
        ("base_rate", C.c_double),
        ("prior_mode", C.c_int),
    ]


def make_params(alpha, beta, base_rate=None, prior_free=False) -> Params:
    return Params(float(alpha), float(beta), int(base_rate is not None),
                  float(base_rate) if base_rate is not None else 0.0, 1 if prior_free else 0)


class FusedField(C.Structure):
    """struct bb25_fused_field"""

    _fields_ = [
        ("index", C.c_void_p),
        ("params", Params),
        ("weight", C.c_double),
        ("q_terms", C.c_void_p),
        ("q_off", C.c_void_p),
        ("term_base", C.c_int64),
        ("n_terms_total", C.c_int64),
    ]


_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
_PP = C.POINTER(Params)

# name -> (restype, argtypes); every symbol include/bb25.h declares
SIGNATURES = {
    "bb25_last_error": (C.c_char_p, []),
    "bb25_version": (_i32, []),
    "bb25_launch_count": (C.c_ulonglong, []),
    "bb25_device_count": (_i32, []),
    "bb25_measure_read_bandwidth": (_i32, [_i32, _i64, _i32, _i32, C.POINTER(_dbl), C.POINTER(_dbl)]),
    "bb25_index_create": (_i32, [_i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _dbl, _i64, C.POINTER(_vp)]),
    "bb25_index_destroy": (None, [_vp]),
    "bb25_index_info": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32),
                               C.POINTER(_i32), C.POINTER(_i64)]),
    "bb25_index_table_info": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "bb25_get_scores": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "bb25_get_probabilities": (_i32, [_vp, _PP, _vp, _i32, _vp, _i64, _vp]),
    "bb25_retrieve_batch": (_i32, [_vp, _PP, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "bb25_retrieve_batch_ex": (_i32, [_vp, _PP, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp]),
    "bb25_retrieve_batch_host": (_i32, [_vp, _PP, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "bb25_retrieve_one_dense": (_i32, [_vp, _PP, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "bb25_retrieve_sync_stats": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "bb25_index_set_threshold_exchange": (_i32, [_vp, _vp, _vp, _i32]),
    "bb25_index_kth_values": (_i32, [_vp, _i32, C.POINTER(_vp), _vp]),
    "bb25_index_set_kth_values": (_i32, [_vp, _i32, _vp, _vp]),
    "bb25_quantile_ranks": (None, [_i32, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "bb25_apply_quantiles": (_i32, [_i32, _vp, _i32, _i64, _i32, _vp, _vp]),
    "bb25_merge_topk_peers": (_i32, [_i32, _vp, _vp, _i32, _i64, _i64, _i32, _vp]),
    "bb25_unpack_topk": (_i32, [_i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "bb25_memcpy_device": (_i32, [_i32, _vp, _vp, _i64, _vp]),
    "bb25_retrieve_stats": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "bb25_index_set_pruning": (_i32, [_vp, _i32]),
    "bb25_retrieve_prune_stats": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "bb25_retrieve_route_stats": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "bb25_retrieve_sparse_units": (_i32, [_vp, C.POINTER(_i64)]),
    "bb25_retrieve_timing": (_i32, [_vp, C.POINTER(_dbl), C.POINTER(_i64)]),
    "bb25_merge_topk": (_i32, [_i32, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp]),
    "bb25_pack_topk": (_i32, [_i32, _vp, _vp, _vp, _i64, _vp, _vp]),
    "bb25_merge_topk_packed": (_i32, [_i32, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp]),
    "bb25_topk_f64": (_i32, [_i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "bb25_sigmoid": (_i32, [_i32, _vp, _i64, _vp, _vp]),
    "bb25_logit": (_i32, [_i32, _vp, _i64, _vp, _vp]),
    "bb25_likelihood": (_i32, [_i32, _PP, _vp, _i64, _vp, _vp]),
    "bb25_tf_prior": (_i32, [_i32, _vp, _i64, _vp, _vp]),
    "bb25_norm_prior": (_i32, [_i32, _vp, _i64, _vp, _vp]),
    "bb25_composite_prior": (_i32, [_i32, _vp, _vp, _i64, _vp, _vp]),
    "bb25_posterior": (_i32, [_i32, _vp, _vp, _i32, _dbl, _i64, _vp, _vp]),
    "bb25_score_to_probability": (_i32, [_i32, _PP, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "bb25_wand_upper_bound": (_i32, [_i32, _PP, _vp, _dbl, _i64, _vp, _vp]),
    "bb25_cosine_to_probability": (_i32, [_i32, _vp, _i64, _vp, _i64, _vp]),
    "bb25_log_odds_conjunction": (_i32, [_i32, _vp, _i64, _i32, _vp, _dbl, _i32, _dbl, _i32, _dbl, _vp, _vp]),
    "bb25_fuse_bm25_signal": (_i32, [_vp, _PP, _vp, _i32, _dbl, _i32, _dbl, _i32, _vp, _vp]),
    "bb25_fuse_cosine_signal": (_i32, [_i32, _vp, _i64, _dbl, _i32, _dbl, _i32, _vp, _vp]),
    "bb25_fuse_prob_signal": (_i32, [_i32, _vp, _i64, _dbl, _i32, _dbl, _i32, _vp, _vp]),
    "bb25_retrieve_fused_batch": (_i32, [_i32, C.POINTER(FusedField), _vp, _i64, _dbl, _i32, _dbl, _i64, _i32, _vp, _vp, _vp]),
    "bb25_fused_stats": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                                C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_dbl)]),
    "bb25_fused_prune_stats": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "bb25_match_counts": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "bb25_trace_bm25": (_i32, [_i32, _PP, _vp, _vp, _vp, _i64, _vp, _vp]),
    "bb25_attention_weights": (_i32, [_i32, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "bb25_attention_fuse": (_i32, [_i32, _vp, _i64, _i32, _vp, _i64, _dbl, _i32, _dbl, _i32, _vp, _vp]),
    "bb25_balanced_fusion": (_i32, [_i32, _vp, _vp, _i64, _dbl, _vp, _vp]),
    "bb25_cosine_gemm": (_i32, [_i32, _vp, _i32, _vp, _i64, _i32, _vp, _i64, _vp]),
    "bb25_blockmax_dense": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _vp]),
    "bb25_blockmax_csc": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(bayesian_bm25_b200 has no CPU fallback)")
        handle = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().bb25_last_error()
        raise RuntimeError("libbb25: " + (msg.decode() if msg else "unknown error"))


def require_cuda() -> int:
    """Current CUDA device index, or raise (no CPU fallback)."""
    import torch

    if lib().bb25_device_count() < 1 or not torch.cuda.is_available():
        raise RuntimeError("bayesian_bm25_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.cuda.current_device()


def stream_ptr() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)
