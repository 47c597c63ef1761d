"""ctypes binding of libwxb_align.so (C ABI: include/wxb_align.h).

There is no CPU fallback: if the library is missing and cannot be built, or no sm_100 device is
present when a context is created, the caller gets a RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WXB_LIB") or os.path.join(_HERE, "libwxb_align.so")

WXB_OK = 0
SEG_OK = 0
SEG_FAILED = 1
F_LOGITS = 1
F_NO_SHORTCUT = 2
F_WILDCARD = 4
F_CLOBBER = 8

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


class Batch(C.Structure):
    _fields_ = [
        ("n_segments", C.c_int64), ("vocab", C.c_int32), ("blank_id", C.c_int32),
        ("flags", C.c_uint32), ("reserved", C.c_uint32),
        ("frame_off", i64p), ("frame_len", i32p), ("tok_off", i64p), ("tokens", i32p),
        ("seg_t1", f64p), ("seg_ratio", f64p), ("seg_round_np", C.POINTER(C.c_uint8)),
        ("n_words", C.c_int64), ("word_off", i64p), ("wchar_off", i64p), ("wchar_idx", i32p),
    ]


class Outputs(C.Structure):
    _fields_ = [
        ("status", C.c_void_p), ("n_frames", C.c_void_p),
        ("char_start", C.c_void_p), ("char_end", C.c_void_p), ("char_score", C.c_void_p),
        ("char_start_s", C.c_void_p), ("char_end_s", C.c_void_p), ("char_score_r", C.c_void_p),
        ("word_start_s", C.c_void_p), ("word_end_s", C.c_void_p), ("word_score", C.c_void_p),
        ("path_token", C.c_void_p), ("path_logp", C.c_void_p), ("final_col", C.c_void_p),
        ("trellis", C.c_void_p), ("logprobs", C.c_void_p),
    ]


class Limits(C.Structure):
    _fields_ = [("max_vocab", C.c_int32), ("max_tokens", C.c_int32),
                ("tokens_per_lane", C.c_int32), ("sm_count", C.c_int32)]


# every symbol include/wxb_align.h declares (tests check the list against the header)
SYMBOLS = (
    "wxb_abi_version", "wxb_last_error", "wxb_create", "wxb_destroy", "wxb_get_limits",
    "wxb_plan_create", "wxb_plan_destroy", "wxb_plan_launch_count", "wxb_plan_stats",
    "wxb_plan_run", "wxb_plan_debug_trace", "wxb_align_device", "wxb_align_host", "wxb_log_softmax_device",
    "wxb_backtrack_trellis_device", "wxb_round3", "wxb_round3_np", "wxb_pairwise_sum", "wxb_tokens_per_lane",
    "wxb_segment_warps", "wxb_set_k_target", "wxb_assign_speakers_device", "wxb_assign_speakers_host",
    "wxb_set_chunk_bytes", "wxb_exact_scores_host", "wxb_plan_run_async", "wxb_ctx_join",
)

_LIB = None


def load(build_if_missing: bool = True):
    """Load the shared library (building it with nvcc first if it is absent)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing; run `python -m whisperx_b200.build`")
        from . import build as _build
        _build.build()
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.wxb_abi_version.restype = C.c_int
    L.wxb_last_error.restype = C.c_char_p
    L.wxb_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.wxb_destroy.argtypes = [vp]
    L.wxb_destroy.restype = None
    L.wxb_get_limits.argtypes = [vp, C.POINTER(Limits)]
    L.wxb_plan_create.argtypes = [vp, C.POINTER(Batch), C.POINTER(vp)]
    L.wxb_plan_destroy.argtypes = [vp]
    L.wxb_plan_destroy.restype = None
    L.wxb_plan_launch_count.argtypes = [vp]
    L.wxb_plan_stats.argtypes = [vp, i64p, i64p, i64p]
    L.wxb_plan_run.argtypes = [vp, vp, C.POINTER(Outputs), vp, f32p]
    L.wxb_plan_run_async.argtypes = [vp, vp, C.POINTER(Outputs), vp]
    L.wxb_plan_run_async.restype = C.c_int
    L.wxb_ctx_join.argtypes = [vp, vp]
    L.wxb_ctx_join.restype = C.c_int
    L.wxb_plan_debug_trace.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int64]
    L.wxb_plan_debug_trace.restype = C.c_int
    L.wxb_align_device.argtypes = [vp, C.POINTER(Batch), vp, C.POINTER(Outputs), vp]
    L.wxb_align_host.argtypes = [vp, C.POINTER(Batch), vp, C.POINTER(Outputs)]
    L.wxb_log_softmax_device.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, vp]
    L.wxb_backtrack_trellis_device.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               vp, vp, vp, vp, vp]
    L.wxb_round3.argtypes = [C.c_double]
    L.wxb_round3.restype = C.c_double
    L.wxb_round3_np.argtypes = [C.c_double]
    L.wxb_round3_np.restype = C.c_double
    L.wxb_pairwise_sum.argtypes = [f64p, C.c_int64]
    L.wxb_pairwise_sum.restype = C.c_double
    L.wxb_tokens_per_lane.argtypes = [C.c_int32]
    L.wxb_segment_warps.argtypes = [C.c_int32]
    L.wxb_segment_warps.restype = C.c_int
    L.wxb_assign_speakers_device.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int32, vp, vp, C.c_int64, C.c_int32, vp, vp]
    L.wxb_assign_speakers_device.restype = C.c_int
    L.wxb_assign_speakers_host.argtypes = [vp, f64p, f64p, i32p, C.c_int64, C.c_int32, f64p, f64p, C.c_int64, C.c_int32, i32p]
    L.wxb_assign_speakers_host.restype = C.c_int
    L.wxb_set_chunk_bytes.argtypes = [vp, C.c_int64]
    L.wxb_set_chunk_bytes.restype = C.c_int64
    L.wxb_exact_scores_host.argtypes = [C.POINTER(Batch), vp, vp, vp, vp, vp, vp, vp]
    L.wxb_exact_scores_host.restype = C.c_int
    L.wxb_set_k_target.argtypes = [C.c_int]
    L.wxb_set_k_target.restype = C.c_int
    for name in ("wxb_create", "wxb_get_limits", "wxb_plan_create", "wxb_plan_launch_count", "wxb_plan_stats",
                 "wxb_plan_run", "wxb_plan_debug_trace", "wxb_align_device", "wxb_align_host", "wxb_log_softmax_device",
                 "wxb_backtrack_trellis_device", "wxb_tokens_per_lane"):
        getattr(L, name).restype = C.c_int
    if L.wxb_abi_version() != 1:
        raise RuntimeError("libwxb_align.so ABI version mismatch")
    _LIB = L
    return L


class NativeError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != WXB_OK:
        msg = load().wxb_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what} failed (code {rc}): {msg}")
