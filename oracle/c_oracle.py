"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/align_oracle.c (see its header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libwxoracle.so")
_LIB = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "align_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(_SO)
        f32p, i32p, i64p, f64p = (C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int64), C.POINTER(C.c_double))
        L.wxo_forward_trellis.argtypes = [f32p, C.c_int, C.c_int, i32p, C.c_int, C.c_int, f32p]
        L.wxo_forward_trellis.restype = None
        L.wxo_walk_back.argtypes = [f32p, f32p, C.c_int, C.c_int, i32p, C.c_int, C.c_int,
                                    i32p, i32p, f32p, f64p, i32p, i32p]
        L.wxo_walk_back.restype = C.c_int
        L.wxo_collapse_repeats.argtypes = [i32p, i32p, f64p, C.c_int, i32p, i32p, i32p, f64p]
        L.wxo_collapse_repeats.restype = C.c_int
        L.wxo_align_segment.argtypes = [f32p, C.c_int, C.c_int, i32p, C.c_int, C.c_int, f32p,
                                        i32p, i32p, f64p, i32p]
        L.wxo_align_segment.restype = C.c_int
        L.wxo_round3.argtypes = [C.c_double]
        L.wxo_round3.restype = C.c_double
        L.wxo_round3_np.argtypes = [C.c_double]
        L.wxo_round3_np.restype = C.c_double
        L.wxo_pairwise_sum.argtypes = [f64p, C.c_long]
        L.wxo_pairwise_sum.restype = C.c_double
        L.wxo_align_batch.argtypes = [f32p, i64p, i32p, i64p, C.c_long, C.c_int, C.c_int,
                                      i32p, i32p, i32p, f64p]
        L.wxo_align_batch.restype = C.c_long
        _LIB = L
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def forward_trellis(E: np.ndarray, tokens, blank: int = 0) -> np.ndarray:
    E = np.ascontiguousarray(E, dtype=np.float32)
    tok = np.ascontiguousarray(tokens, dtype=np.int32)
    T, V = E.shape
    N = tok.shape[0]
    R = np.empty((T + 1, N + 1), dtype=np.float32)
    lib().wxo_forward_trellis(_p(E, C.c_float), T, V, _p(tok, C.c_int32), N, blank, _p(R, C.c_float))
    return R


def walk_back(R: np.ndarray, E: np.ndarray, tokens, blank: int = 0):
    """-> None or dict(tok, time, logp, prob, t_start)."""
    E = np.ascontiguousarray(E, dtype=np.float32)
    R = np.ascontiguousarray(R, dtype=np.float32)
    tok = np.ascontiguousarray(tokens, dtype=np.int32)
    T, V = E.shape
    N = tok.shape[0]
    cap = max(T, 1)
    pt = np.empty(cap, np.int32); pi = np.empty(cap, np.int32)
    pl = np.empty(cap, np.float32); pp = np.empty(cap, np.float64)
    n = C.c_int32(0); ts = C.c_int32(0)
    rc = lib().wxo_walk_back(_p(R, C.c_float), _p(E, C.c_float), T, V, _p(tok, C.c_int32), N, blank,
                             _p(pt, C.c_int32), _p(pi, C.c_int32), _p(pl, C.c_float),
                             _p(pp, C.c_double), C.byref(n), C.byref(ts))
    if rc != 0:
        return None
    k = n.value
    return dict(tok=pt[:k].copy(), time=pi[:k].copy(), logp=pl[:k].copy(), prob=pp[:k].copy(),
                t_start=ts.value)


def align_segment(E: np.ndarray, tokens, blank: int = 0):
    """-> (status, start[N], end[N], score[N], t_start)."""
    E = np.ascontiguousarray(E, dtype=np.float32)
    tok = np.ascontiguousarray(tokens, dtype=np.int32)
    T, V = E.shape
    N = tok.shape[0]
    cs = np.full(max(N, 1), -1, np.int32); ce = np.full(max(N, 1), -1, np.int32)
    sc = np.full(max(N, 1), np.nan, np.float64)
    ts = C.c_int32(0)
    rc = lib().wxo_align_segment(_p(E, C.c_float), T, V, _p(tok, C.c_int32), N, blank, None,
                                 _p(cs, C.c_int32), _p(ce, C.c_int32), _p(sc, C.c_double), C.byref(ts))
    return rc, cs[:N], ce[:N], sc[:N], ts.value


def align_batch(E, frame_off, tokens, tok_off, V: int, blank: int = 0):
    """Concatenated layout (same as the engine C-ABI). -> (status[S], start, end, score)."""
    E = np.ascontiguousarray(E, dtype=np.float32)
    fo = np.ascontiguousarray(frame_off, dtype=np.int64)
    to = np.ascontiguousarray(tok_off, dtype=np.int64)
    tok = np.ascontiguousarray(tokens, dtype=np.int32)
    S = fo.shape[0] - 1
    tot = int(to[-1])
    st = np.zeros(S, np.int32)
    cs = np.full(max(tot, 1), -1, np.int32); ce = np.full(max(tot, 1), -1, np.int32)
    sc = np.full(max(tot, 1), np.nan, np.float64)
    lib().wxo_align_batch(_p(E, C.c_float), _p(fo, C.c_int64), _p(tok, C.c_int32), _p(to, C.c_int64),
                          S, V, blank, _p(st, C.c_int32), _p(cs, C.c_int32), _p(ce, C.c_int32),
                          _p(sc, C.c_double))
    return st, cs[:tot], ce[:tot], sc[:tot]


def round3(x: float) -> float:
    return lib().wxo_round3(float(x))


def round3_np(x: float) -> float:
    return lib().wxo_round3_np(float(x))


def pairwise_sum(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().wxo_pairwise_sum(_p(a, C.c_double), a.shape[0])
