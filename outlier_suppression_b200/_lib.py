"""ctypes binding of libosq_b200.so (the C ABI declared in include/osq.h).

No torch types cross this boundary: raw device pointers, sizes and a cudaStream_t only.
The library is built in-tree by ``outlier_suppression_b200.build``; if it is missing and cannot be
built this module raises -- there is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OSQ_LIB_PATH") or os.path.join(_PKG, "libosq_b200.so")  # OSQ_LIB_PATH: A/B builds of the same ABI

EXPORTS = [
    "osq_version", "osq_last_error", "osq_sm_count", "osq_workspace_bytes",
    "osq_fq_per_tensor_f32", "osq_fq_per_tensor_bins_f32", "osq_act_fq_per_tensor_bins_f32", "osq_fq_per_channel_f32", "osq_fq_per_tensor_bins_only_f32",
    "osq_dequant_bins_f32",
    "osq_residual_layernorm_fq_f32", "osq_attn_scores_fq_f32", "osq_attn_context_fq_f32",
    "osq_minmax_masked_f32", "osq_minmax_flat_f32", "osq_token_minmax_f32", "osq_prune_select_f32", "osq_prune_select_unsorted_f32",
    "osq_prune_observe_f32", "osq_prune_observe_many_f32", "osq_token_minmax_hist_f32", "osq_prune_select_cached_f32", "osq_quantile_observe_f32", "osq_replay_average_f32", "osq_replay_average_peer_f32", "osq_replay_exchange_f32",
    "osq_rowwise_minmax_qparams_f32", "osq_calc_qparams_f32",
    "osq_mse_multi_f32", "osq_mse_brent_rows_f32", "osq_mse_brent_tensor_f32", "osq_mse_tensor_scratch_bytes",
    "osq_pack_weight_s8", "osq_fused_fq_linear", "osq_fused_fq_linear_multi", "osq_lsqplus_backward_f32",
]


class Tokens(C.Structure):
    """osq_tokens_t"""
    _fields_ = [(n, C.c_int64) for n in ("B", "S", "F1", "F2", "sb", "ss", "sf1", "sf2")]


class StatEpilogue(C.Structure):
    """osq_stat_epilogue_t"""
    _fields_ = [("mode", C.c_int), ("cnt", C.c_int), ("state_min", C.c_void_p), ("state_max", C.c_void_p),
                ("scale_out", C.c_void_p), ("zp_out", C.c_void_p), ("zp_out_is_int32", C.c_int),
                ("qmin", C.c_int), ("qmax", C.c_int), ("symmetric", C.c_int)]


class ReplayTarget(C.Structure):
    """osq_replay_target_t"""
    _fields_ = [("state_min", C.c_void_p), ("state_max", C.c_void_p), ("scale_out", C.c_void_p), ("zp_out", C.c_void_p),
                ("zp_out_is_int32", C.c_int), ("qmin", C.c_int), ("qmax", C.c_int), ("symmetric", C.c_int)]


class FusedLinearArgs(C.Structure):
    """osq_fused_linear_t"""
    _fields_ = [("A", C.c_void_p), ("M", C.c_int64), ("K", C.c_int64), ("a_scale", C.c_void_p), ("a_zp", C.c_void_p),
                ("a_zp_is_int32", C.c_int), ("lsq_grad_factor", C.c_float), ("a_qmin", C.c_int), ("a_qmax", C.c_int),
                ("w_codes", C.c_void_p), ("w_scale", C.c_void_p), ("w_rowsum", C.c_void_p), ("bias", C.c_void_p),
                ("Y", C.c_void_p), ("N", C.c_int64), ("mma_kind", C.c_int), ("a_codes", C.c_void_p), ("debug_trace", C.c_void_p),
                ("out_act", C.c_int), ("out_scale", C.c_void_p), ("out_zp", C.c_void_p), ("out_zp_is_int32", C.c_int),
                ("out_lsq_grad_factor", C.c_float), ("out_qmin", C.c_int), ("out_qmax", C.c_int), ("out_bins", C.c_void_p)]


class SelectProblem(C.Structure):
    """osq_select_problem_t"""
    _fields_ = [("tmin", C.c_void_p), ("tmax", C.c_void_p), ("n_slots", C.c_int64), ("n_valid", C.c_void_p), ("hist0", C.c_void_p),
                ("cur", C.c_void_p)]


class QuantizerArgs(C.Structure):
    """osq_quantizer_t"""
    _fields_ = [("scale", C.c_void_p), ("zero_point", C.c_void_p), ("zp_is_int32", C.c_int), ("lsq_grad_factor", C.c_float),
                ("qmin", C.c_int), ("qmax", C.c_int)]


_lock = threading.Lock()
_lib = None


class OsqError(RuntimeError):
    pass


def _declare(lib):
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
    lib.osq_version.restype = i32
    lib.osq_last_error.restype = C.c_char_p
    lib.osq_sm_count.restype = i32
    lib.osq_workspace_bytes.restype = i64
    lib.osq_mse_tensor_scratch_bytes.restype = i64
    sig = {
        "osq_fq_per_tensor_f32": [vp, vp, vp, i64, vp, vp, i32, f32, i32, i32, vp],
        "osq_fq_per_tensor_bins_f32": [vp, vp, vp, i64, vp, vp, i32, f32, i32, i32, vp],
        "osq_act_fq_per_tensor_bins_f32": [vp, vp, vp, i64, i32, vp, vp, i32, f32, i32, i32, vp],
        "osq_fq_per_channel_f32": [vp, vp, vp, i64, i64, vp, vp, i32, i32, vp],
        "osq_fq_per_tensor_bins_only_f32": [vp, vp, i64, i32, vp, vp, i32, f32, i32, i32, vp, vp],
        "osq_dequant_bins_f32": [vp, vp, i32, i32, vp, i64, vp],
        "osq_residual_layernorm_fq_f32": [vp, vp, vp, vp, vp, f32, i64, i64, vp, vp, i32, f32, i32, i32, vp, vp, vp, vp],
        "osq_minmax_masked_f32": [vp, C.POINTER(Tokens), vp, i32, vp, C.POINTER(StatEpilogue), vp, vp],
        "osq_minmax_flat_f32": [vp, i64, vp, C.POINTER(StatEpilogue), vp, vp],
        "osq_token_minmax_f32": [vp, C.POINTER(Tokens), vp, i32, vp, vp, vp, vp],
        "osq_prune_select_f32": [vp, vp, vp, vp, i64, vp, f32, vp, C.POINTER(StatEpilogue), vp, vp],
        "osq_prune_select_unsorted_f32": [vp, vp, i64, vp, f32, vp, C.POINTER(StatEpilogue), vp, vp],
        "osq_prune_observe_f32": [vp, C.POINTER(Tokens), vp, i32, f32, vp, vp, vp, vp, vp, C.POINTER(StatEpilogue), vp, vp],
        "osq_prune_observe_many_f32": [vp, i32, C.POINTER(Tokens), vp, i32, f32, vp, vp, vp, vp, vp, C.POINTER(StatEpilogue), vp, vp],
        "osq_token_minmax_hist_f32": [vp, C.POINTER(Tokens), vp, i32, vp, vp, vp, vp, vp],
        "osq_prune_select_cached_f32": [vp, i32, f32, vp, vp],
        "osq_quantile_observe_f32": [vp, C.POINTER(Tokens), vp, i32, i32, C.c_double, vp, vp, C.POINTER(StatEpilogue), vp, vp],
        "osq_replay_average_f32": [vp, i32, i32, i32, vp, vp],
        "osq_replay_average_peer_f32": [vp, i32, i32, i32, i32, vp, vp],
        "osq_replay_exchange_f32": [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp],
        "osq_rowwise_minmax_qparams_f32": [vp, i64, i64, i32, vp, vp, vp, vp, i32, i32, i32, vp],
        "osq_calc_qparams_f32": [vp, vp, i64, i32, i32, i32, vp, vp, vp, vp],
        "osq_mse_multi_f32": [vp, C.POINTER(Tokens), vp, i32, vp, vp, i32, i32, i32, vp, vp, vp],
        "osq_mse_brent_rows_f32": [vp, i64, i64, i32, i32, i32, vp, vp, vp, vp],
        "osq_mse_brent_tensor_f32": [vp, C.POINTER(Tokens), vp, i32, i32, i32, i32, vp, vp, vp, vp, vp],
        "osq_pack_weight_s8": [vp, i64, i64, vp, vp, i32, i32, vp, vp, vp],
        "osq_fused_fq_linear": [C.POINTER(FusedLinearArgs), vp],
        "osq_fused_fq_linear_multi": [C.POINTER(FusedLinearArgs), i32, vp],
        "osq_attn_scores_fq_f32": [vp, vp, i64, i64, i64, i64, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(QuantizerArgs),
                                   C.POINTER(QuantizerArgs), f32, vp, vp, vp],
        "osq_attn_context_fq_f32": [vp, vp, i64, i64, i64, i64, i64, C.POINTER(i64), C.POINTER(QuantizerArgs), C.POINTER(QuantizerArgs),
                                    C.POINTER(QuantizerArgs), vp, vp, vp],
        "osq_lsqplus_backward_f32": [vp, vp, vp, i64, vp, vp, f32, i32, i32, vp, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = i32


def load(build_if_missing: bool = True):
    """Returns the loaded CDLL; builds it in-tree first if absent. Raises OsqError otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise OsqError("%s is missing; run `python -m outlier_suppression_b200.build`" % LIB_PATH)
            from . import build as _build
            _build.build()
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise OsqError("cannot load %s: %s (no CPU fallback exists)" % (LIB_PATH, e)) from e
        _declare(lib)
        if lib.osq_version() != 100:
            raise OsqError("libosq_b200.so version mismatch: %d" % lib.osq_version())
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().osq_last_error().decode("utf-8", "replace")
        raise OsqError("%s failed (code %d): %s" % (what or "libosq_b200 call", rc, msg))
