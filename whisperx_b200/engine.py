"""Batched forced-alignment engine: Python face of the C ABI (include/wxb_align.h).

PyTorch is used for device memory and streams only; all arithmetic of the alignment path
(reference: whisperx/alignment.py:233, 359-379, 387-421, 438-454, 252-264, 297-320) runs in the
hand-written sm_100a kernels of libwxb_align.so.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native as N


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a: Optional[np.ndarray], ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


@dataclass
class BatchSpec:
    """Host-side description of a batch of segments in the concatenated layout of the C ABI."""
    frame_off: np.ndarray            # int64 [S+1]
    tok_off: np.ndarray              # int64 [S+1]
    tokens: np.ndarray               # int32 [sum_N]
    vocab: int
    blank_id: int = 0
    logits: bool = False             # emissions are raw logits -> fused log_softmax
    no_shortcut: bool = False
    wildcard: bool = False           # extension: token id -1 = wildcard (max over the non-blank vocabulary)
    clobber: bool = False            # the emission buffer is scratch: logits may be normalised in place
    seg_t1: Optional[np.ndarray] = None      # float64 [S]
    seg_ratio: Optional[np.ndarray] = None   # float64 [S]
    seg_round_np: Optional[np.ndarray] = None  # uint8 [S]: numpy-typed segment times (numpy rounding)
    frame_len: Optional[np.ndarray] = None   # int32 [S]; None = contiguous segments
    word_off: Optional[np.ndarray] = None    # int64 [S+1]
    wchar_off: Optional[np.ndarray] = None   # int64 [W+1]
    wchar_idx: Optional[np.ndarray] = None   # int32 [sum]

    def __post_init__(self):
        self.frame_off = _np(self.frame_off, np.int64)
        self.tok_off = _np(self.tok_off, np.int64)
        self.tokens = _np(self.tokens, np.int32)
        if self.frame_len is not None:
            self.frame_len = _np(self.frame_len, np.int32)
        if self.seg_t1 is not None:
            self.seg_t1 = _np(self.seg_t1, np.float64)
            self.seg_ratio = _np(self.seg_ratio, np.float64)
            if self.seg_round_np is not None:
                self.seg_round_np = _np(self.seg_round_np, np.uint8)
        if self.word_off is not None:
            self.word_off = _np(self.word_off, np.int64)
            self.wchar_off = _np(self.wchar_off, np.int64)
            self.wchar_idx = _np(self.wchar_idx, np.int32)

    @property
    def n_segments(self) -> int:
        return int(self.frame_off.shape[0] - 1)

    @property
    def sum_T(self) -> int:
        return int(self.frame_off[-1]) if self.frame_off.size else 0

    @property
    def sum_N(self) -> int:
        return int(self.tok_off[-1]) if self.tok_off.size else 0

    @property
    def n_words(self) -> int:
        return int(self.wchar_off.shape[0] - 1) if self.wchar_off is not None else 0

    @property
    def T(self) -> np.ndarray:
        return self.frame_len.astype(np.int64) if self.frame_len is not None else np.diff(self.frame_off)

    def trellis_offsets(self) -> np.ndarray:
        T = self.T
        Nn = np.diff(self.tok_off)
        return np.concatenate([[0], np.cumsum((T + 1) * (Nn + 1))]).astype(np.int64)

    def to_c(self) -> N.Batch:
        b = N.Batch()
        b.n_segments = self.n_segments
        b.vocab = int(self.vocab)
        b.blank_id = int(self.blank_id)
        b.flags = ((N.F_LOGITS if self.logits else 0) | (N.F_NO_SHORTCUT if self.no_shortcut else 0)
                   | (N.F_WILDCARD if self.wildcard else 0) | (N.F_CLOBBER if self.clobber else 0))
        b.frame_off = _ptr(self.frame_off, C.c_int64)
        b.frame_len = _ptr(self.frame_len, C.c_int32)
        b.tok_off = _ptr(self.tok_off, C.c_int64)
        b.tokens = _ptr(self.tokens, C.c_int32)
        b.seg_t1 = _ptr(self.seg_t1, C.c_double)
        b.seg_ratio = _ptr(self.seg_ratio, C.c_double)
        b.seg_round_np = _ptr(self.seg_round_np, C.c_uint8)
        b.n_words = self.n_words
        b.word_off = _ptr(self.word_off, C.c_int64)
        b.wchar_off = _ptr(self.wchar_off, C.c_int64)
        b.wchar_idx = _ptr(self.wchar_idx, C.c_int32)
        return b


@dataclass
class BatchResult:
    """Outputs of one batch; tensors live where the call put them (CUDA or host)."""
    status: torch.Tensor
    n_frames: torch.Tensor
    char_start: torch.Tensor
    char_end: torch.Tensor
    char_score: torch.Tensor
    char_start_s: Optional[torch.Tensor] = None
    char_end_s: Optional[torch.Tensor] = None
    char_score_r: Optional[torch.Tensor] = None
    word_start_s: Optional[torch.Tensor] = None
    word_end_s: Optional[torch.Tensor] = None
    word_score: Optional[torch.Tensor] = None
    path_token: Optional[torch.Tensor] = None
    path_logp: Optional[torch.Tensor] = None
    final_col: Optional[torch.Tensor] = None
    trellis: Optional[torch.Tensor] = None
    logprobs: Optional[torch.Tensor] = None
    extra: dict = field(default_factory=dict)

    _FIELDS = ("status", "n_frames", "char_start", "char_end", "char_score", "char_start_s", "char_end_s",
               "char_score_r", "word_start_s", "word_end_s", "word_score", "path_token", "path_logp",
               "final_col", "trellis", "logprobs")

    def to_c(self) -> N.Outputs:
        o = N.Outputs()
        for f in self._FIELDS:
            t = getattr(self, f)
            setattr(o, f, t.data_ptr() if t is not None else None)
        return o

    def cpu(self) -> "BatchResult":
        kw = {f: (getattr(self, f).cpu() if getattr(self, f) is not None else None) for f in self._FIELDS}
        return BatchResult(**kw, extra=self.extra)


def _alloc_outputs(spec: BatchSpec, device, want: Sequence[str], pin: bool = False) -> BatchResult:
    S, sN, sT, W = spec.n_segments, max(spec.sum_N, 1), max(spec.sum_T, 1), max(spec.n_words, 1)
    kw = dict(device=device) if not pin else dict(device="cpu", pin_memory=True)
    e = lambda n, dt: torch.empty(n, dtype=dt, **kw)
    r = BatchResult(status=e(max(S, 1), torch.int32), n_frames=e(max(S, 1), torch.int32),
                    char_start=e(sN, torch.int32), char_end=e(sN, torch.int32), char_score=e(sN, torch.float64))
    if spec.seg_t1 is not None:
        r.char_start_s, r.char_end_s, r.char_score_r = (e(sN, torch.float64) for _ in range(3))
        if spec.n_words > 0:
            r.word_start_s, r.word_end_s, r.word_score = (e(W, torch.float64) for _ in range(3))
    if "path" in want:
        r.path_token = e(sT, torch.int32)
        r.path_logp = e(sT, torch.float32)
    if "final_col" in want:
        r.final_col = e(sT + S, torch.float32)
    if "trellis" in want:
        r.trellis = e(max(int(spec.trellis_offsets()[-1]), 1), torch.float32)
    if "logprobs" in want:
        if not spec.logits:
            raise ValueError("logprobs output needs logits=True")
        r.logprobs = e(sT * spec.vocab, torch.float32)
    return r


class Plan:
    """A batch uploaded and bucketed once (wxb_plan); run it as often as needed."""

    def __init__(self, engine: "AlignEngine", spec: BatchSpec):
        self.engine = engine
        self.spec = spec           # keeps the host arrays alive
        self._h = C.c_void_p()
        cb = spec.to_c()
        with engine._lock:
            N.check(engine._lib.wxb_plan_create(engine._ctx, C.byref(cb), C.byref(self._h)), "wxb_plan_create")
        cells, alg, short = C.c_int64(), C.c_int64(), C.c_int64()
        N.check(engine._lib.wxb_plan_stats(self._h, C.byref(cells), C.byref(alg), C.byref(short)), "wxb_plan_stats")
        self.cells, self.alg_bytes, self.n_shortcut = cells.value, alg.value, short.value
        self.launches = engine._lib.wxb_plan_launch_count(self._h)

    def run(self, emissions: torch.Tensor, out: Optional[BatchResult] = None, want: Sequence[str] = (),
            timed: bool = False) -> BatchResult:
        eng = self.engine
        eng._check_emissions(emissions, self.spec)
        if out is None:
            out = _alloc_outputs(self.spec, eng.device, want)
        co = out.to_c()
        ms = C.c_float(0.0)
        with eng._lock, torch.cuda.device(eng.device):
            stream = torch.cuda.current_stream(eng.device).cuda_stream
            N.check(eng._lib.wxb_plan_run(self._h, emissions.data_ptr(), C.byref(co), stream,
                                          C.byref(ms) if timed else None), "wxb_plan_run")
        if timed:
            out.extra["dp_ms"] = ms.value
        return out

    def run_overlapped(self, emissions: torch.Tensor, out: BatchResult) -> BatchResult:
        """wxb_plan_run_async: enqueue the batch without making the current stream wait for it; up to two such
        runs are in flight per engine (the tail of one batch overlaps the head of the next).  Call
        `engine.join()` before reading `out` (and before handing `out` to a third run)."""
        eng = self.engine
        eng._check_emissions(emissions, self.spec)
        co = out.to_c()
        with eng._lock, torch.cuda.device(eng.device):
            stream = torch.cuda.current_stream(eng.device).cuda_stream
            N.check(eng._lib.wxb_plan_run_async(self._h, emissions.data_ptr(), C.byref(co), stream), "wxb_plan_run_async")
        return out

    def close(self):
        if self._h:
            with self.engine._lock:
                self.engine._lib.wxb_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AlignEngine:
    """One per GPU.  Owns the native workspace (wxb_ctx)."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("whisperx_b200 runs the alignment path on a CUDA (sm_100a) device only; "
                               f"got device={device!r}. There is no CPU fallback.")
        if not torch.cuda.is_available():
            raise RuntimeError("whisperx_b200: no CUDA device is available; the alignment kernels cannot run "
                               "(there is no CPU fallback).")
        self._lib = N.load()
        self._ctx = C.c_void_p()
        # a wxb_ctx owns scratch buffers and side streams: one native call at a time per engine
        self._lock = threading.RLock()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.empty(1, device=self.device)          # make sure the primary context exists
            N.check(self._lib.wxb_create(idx, C.byref(self._ctx)), "wxb_create")
        lim = N.Limits()
        N.check(self._lib.wxb_get_limits(self._ctx, C.byref(lim)), "wxb_get_limits")
        self.limits = lim

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.wxb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers
    def _check_emissions(self, emissions: torch.Tensor, spec: BatchSpec):
        if not (emissions.is_cuda and emissions.device == self.device):
            raise ValueError(f"emissions must live on {self.device}")
        if emissions.dtype != torch.float32 or not emissions.is_contiguous():
            raise ValueError("emissions must be a contiguous float32 tensor")
        if emissions.numel() < spec.sum_T * spec.vocab:
            raise ValueError("emissions is smaller than frame_off[-1] * vocab")

    # ---- hot path
    def plan(self, spec: BatchSpec) -> Plan:
        return Plan(self, spec)

    def join(self) -> None:
        """Make the current stream wait for every overlapped run of this engine (wxb_ctx_join)."""
        with self._lock, torch.cuda.device(self.device):
            N.check(self._lib.wxb_ctx_join(self._ctx, torch.cuda.current_stream(self.device).cuda_stream), "wxb_ctx_join")

    def align_batch(self, emissions: torch.Tensor, spec: BatchSpec, want: Sequence[str] = ()) -> BatchResult:
        """get_trellis + backtrack + merge_repeats (+ timestamps / words) for every segment."""
        with self._lock:
            plan = Plan(self, spec)
            try:
                out = plan.run(emissions, want=want)
                torch.cuda.current_stream(self.device).synchronize()
            finally:
                plan.close()
        return out

    def align_batch_host(self, emissions_host, spec: BatchSpec, want: Sequence[str] = (),
                         out: Optional[BatchResult] = None) -> BatchResult:
        """Same through wxb_align_host: host buffers in, host buffers out (copies inside)."""
        if isinstance(emissions_host, torch.Tensor):
            if emissions_host.is_cuda or emissions_host.dtype != torch.float32 or not emissions_host.is_contiguous():
                raise ValueError("emissions_host must be a contiguous float32 CPU tensor")
            eptr = emissions_host.data_ptr()
        else:
            emissions_host = _np(emissions_host, np.float32)
            eptr = emissions_host.ctypes.data
        if out is None:
            out = _alloc_outputs(spec, "cpu", want, pin=True)
        cb, co = spec.to_c(), out.to_c()
        with self._lock:
            N.check(self._lib.wxb_align_host(self._ctx, C.byref(cb), eptr, C.byref(co)), "wxb_align_host")
        return out

    def log_softmax(self, logits: torch.Tensor) -> torch.Tensor:
        """Kernel (i) alone: the engine's log_softmax(dim=-1) (alignment.py:233)."""
        if not logits.is_cuda or logits.dtype != torch.float32:
            raise ValueError("logits must be a float32 CUDA tensor")
        x = logits.contiguous().view(-1, logits.shape[-1])
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            N.check(self._lib.wxb_log_softmax_device(self._ctx, x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(),
                                                     torch.cuda.current_stream(self.device).cuda_stream),
                    "wxb_log_softmax_device")
        return out.view(logits.shape)

    def backtrack_trellis(self, trellis: torch.Tensor, emission: torch.Tensor, tokens: Sequence[int], blank_id: int):
        """backtrack() from a caller-supplied trellis (alignment.py:387-421).
        -> (status, t_start, path_token[T] cpu, path_logp[T] cpu)"""
        dev = self.device
        R = trellis.to(dev, torch.float32).contiguous()
        E = emission.to(dev, torch.float32).contiguous()
        tok = torch.as_tensor(np.asarray(tokens, dtype=np.int32), device=dev)
        T, V = E.shape
        Nn = tok.numel()
        if R.shape != (T + 1, Nn + 1):
            raise ValueError(f"trellis shape {tuple(R.shape)} does not match emission/tokens ({T + 1}, {Nn + 1})")
        pt = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
        pl = torch.empty(max(T, 1), dtype=torch.float32, device=dev)
        st = torch.zeros(2, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(self._lib.wxb_backtrack_trellis_device(
                self._ctx, R.data_ptr(), E.data_ptr(), tok.data_ptr(), T, Nn, V, int(blank_id), pt.data_ptr(),
                pl.data_ptr(), st.data_ptr(), st.data_ptr() + 4, torch.cuda.current_stream(dev).cuda_stream),
                "wxb_backtrack_trellis_device")
        stc = st.cpu()
        return int(stc[0]), int(stc[1]), pt[:T].cpu(), pl[:T].cpu()


_ENGINES: dict = {}


def get_engine(device="cuda:0") -> AlignEngine:
    """Process-wide engine cache, one per device."""
    d = torch.device(device)
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    key = str(d)
    if key not in _ENGINES:
        _ENGINES[key] = AlignEngine(d)
    return _ENGINES[key]
