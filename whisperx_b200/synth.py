"""Seeded synthetic workloads for the alignment path (SURVEY.md §8d / BASELINE.json configs).

Shapes follow wav2vec2: 16 kHz audio, conv stride 320, receptive field 400 ->
T = (n_samples - 400)//320 + 1 frames; transcripts run at ~13.3 characters per second.
Used by bench.py, __graft_entry__.smoke() and the tests — never by align() itself.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

SAMPLE_RATE = 16000
CHARS_PER_SECOND = 13.3


def frames_for(duration_s: float) -> int:
    n_samples = int(duration_s * SAMPLE_RATE)
    return max((n_samples - 400) // 320 + 1, 1)


def make_tokens(rng: np.random.Generator, n: int, vocab: int, blank: int, sep: Optional[int]) -> np.ndarray:
    """Words of 1-8 symbols drawn from the non-blank, non-separator ids, joined by `sep`."""
    ids = np.array([v for v in range(vocab) if v != blank and v != sep], dtype=np.int32)
    out = np.empty(n, dtype=np.int32)
    i = 0
    while i < n:
        w = int(rng.integers(1, 9))
        w = min(w, n - i)
        out[i:i + w] = rng.choice(ids, w)
        i += w
        if i < n and sep is not None:
            out[i] = sep
            i += 1
    return out


def make_logits(rng: np.random.Generator, T: int, vocab: int, kind: str = "iid",
                tokens: Optional[np.ndarray] = None, blank: int = 0) -> np.ndarray:
    """D1 'iid': N(0,1)*3.  D2 'planted': D1 + 8 on a random monotone token path, + 8 on blank elsewhere."""
    x = rng.standard_normal((T, vocab), dtype=np.float32) * np.float32(3.0)
    if kind == "planted" and tokens is not None and 0 < len(tokens) <= T:
        at = np.sort(rng.choice(T, len(tokens), replace=False))
        boost = np.full(T, blank, dtype=np.int64)
        boost[at] = tokens
        x[np.arange(T), boost] += np.float32(8.0)
    elif kind not in ("iid", "planted"):
        raise ValueError(kind)
    return x


@dataclass
class SynthBatch:
    emissions: np.ndarray      # float32 [sum_T, V] raw logits (or log-probs if as_logprobs)
    frame_off: np.ndarray      # int64 [S+1] row offset of each segment
    frame_len: np.ndarray      # int32 [S]   frames per segment (<= gap to the next offset)
    tok_off: np.ndarray        # int64 [S+1]
    tokens: np.ndarray         # int32 [sum_N]
    durations: np.ndarray      # float64 [S] seconds
    vocab: int
    blank: int

    @property
    def cells(self) -> int:
        return int(np.sum(self.frame_len.astype(np.int64) * np.diff(self.tok_off)))

    @property
    def audio_seconds(self) -> float:
        return float(self.durations.sum())


def make_batch(n_segments: int, config: str = "c2", seed0: int = 1234, vocab: Optional[int] = None,
               kind: str = "iid", align_frames: int = 1, first_segment: int = 0) -> SynthBatch:
    """BASELINE.json workloads: c1/c3/c5 = 30 s segments, c2 = U[5,30] s, c4 = U[40,60] s with long
    transcripts over a ~100-symbol vocabulary whose blank is the last id (HF '<pad>' style).

    `align_frames` rounds every segment's row count up to a multiple of that many rows (4 makes
    every segment start 16-byte aligned for any V, which is what the bulk-copy path wants);
    padding rows are zeros and belong to no segment (frame_len excludes them).
    """
    if config in ("c1", "c3", "c5"):
        V = vocab or 29
        blank, sep = 0, 1
    elif config == "c2":
        V = vocab or 29
        blank, sep = 0, 1
    elif config == "c4":
        V = vocab or 100
        blank, sep = V - 1, 4
    else:
        raise ValueError(config)
    em, toks, durs, lens = [], [], [], []
    frame_off, tok_off = [0], [0]
    for s in range(first_segment, first_segment + n_segments):
        rng = np.random.default_rng(seed0 + s)
        if config == "c2":
            d = float(rng.uniform(5.0, 30.0))
        elif config == "c4":
            d = float(rng.uniform(40.0, 60.0))
        else:
            d = 30.0
        T = frames_for(d)
        if config == "c4":
            N = int(rng.integers(500, min(2000, T) + 1))
        else:
            N = max(int(round(CHARS_PER_SECOND * d)), 1)
        tk = make_tokens(rng, N, V, blank, sep)
        x = make_logits(rng, T, V, kind, tk, blank)
        pad = (-T) % align_frames          # pad rows after the segment keep the next start aligned
        em.append(x)
        if pad:
            em.append(np.zeros((pad, V), np.float32))
        toks.append(tk)
        durs.append(d)
        lens.append(T)
        frame_off.append(frame_off[-1] + T + pad)
        tok_off.append(tok_off[-1] + N)
    return SynthBatch(np.concatenate(em, 0), np.asarray(frame_off, np.int64), np.asarray(lens, np.int32),
                      np.asarray(tok_off, np.int64), np.concatenate(toks), np.asarray(durs), V, blank)


# ----------------------------------------------------------------------------- on-device sweep generator
# BASELINE configs[4] (10,000 audio-hours = 1.2 M x 30 s segments) is 208 GB of float32 emissions: it cannot
# be staged from the host, so the sweep's batches are generated on the GPU, one batch at a time, by a
# counter-based generator: every value is a pure function of (segment id, element index), so any rank of
# any world size produces bit-identical data for the segments it owns.  Same distributions as make_batch
# (D1 iid N(0,9); D2 planted), different random streams.
_M64 = (1 << 64) - 1


def _i64(c: int) -> int:
    """Python int -> the int64 with the same 64 bits (torch integer arithmetic wraps)."""
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


def _mix64(x):
    """splitmix64 finaliser on an int64 tensor (logical shifts emulated by masking the arithmetic ones)."""
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * _i64(0x94D049BB133111EB)
    return x ^ ((x >> 31) & ((1 << 33) - 1))


def _uniform24(h, shift: int):
    """24 bits of the hash -> float32 in (0, 1)."""
    import torch
    return (((h >> shift) & 0xFFFFFF).to(torch.float32) + 0.5) * (1.0 / 16777216.0)


def device_sweep_batch(seg_ids, device, duration_s: float = 30.0, vocab: int = 29, blank: int = 0, sep: int = 1,
                       kind: str = "iid", seed: int = 1234):
    """One batch of the sweep for the given segment ids (all `duration_s` long, N = round(13.3 * d) tokens).

    Returns (emissions [S * Tp, V] float32 CUDA tensor of raw logits, SynthBatch-like host metadata with
    `emissions=None`).  Tp = T rounded up to 4 rows (16-byte aligned segment starts); pad rows are zero.
    Tokens: separator with probability 2/11 (never first/last or doubled), else a uniform non-blank,
    non-separator symbol — generated on the device from the same hash and copied to the host (the C ABI takes
    token ids from host memory)."""
    import torch
    dev = torch.device(device)
    ids = torch.as_tensor(np.asarray(seg_ids, dtype=np.int64), device=dev)
    S = int(ids.numel())
    T = frames_for(duration_s)
    Tp = T + (-T) % 4
    N = max(int(round(CHARS_PER_SECOND * duration_s)), 1)
    key = _mix64(ids * _i64(0x9E3779B97F4A7C15) + _i64(seed))                      # [S] per-segment key
    # tokens
    k = torch.arange(N, device=dev, dtype=torch.int64)
    ht = _mix64(key[:, None] + (k[None, :] + 1) * _i64(0xD1B54A32D192ED03))
    letters = torch.tensor([v for v in range(vocab) if v != blank and v != sep], device=dev, dtype=torch.int64)
    tok = letters[((ht >> 8) & 0x7FFFFFFF) % letters.numel()]
    is_sep = ((ht >> 44) & 0xFFFF) < (0x10000 * 2 // 11)
    is_sep[:, 0] = False
    is_sep[:, -1] = False
    is_sep[:, 1:] &= ~is_sep[:, :-1]
    tok = torch.where(is_sep, torch.full_like(tok, sep), tok)
    # logits, in row blocks to bound the int64 temporaries
    em = torch.zeros((S, Tp, vocab), dtype=torch.float32, device=dev)
    e_idx = torch.arange(T * vocab, device=dev, dtype=torch.int64)
    step = max(1, (1 << 27) // (T * vocab))
    for a in range(0, S, step):
        h = _mix64(key[a:a + step, None] ^ ((e_idx[None, :] + 1) * _i64(0xA0761D6478BD642F)))
        z = torch.sqrt(-2.0 * torch.log(_uniform24(h, 40))) * torch.cos(6.283185307179586 * _uniform24(h, 12))
        em[a:a + step, :T] = (z * 3.0).view(-1, T, vocab)
    if kind == "planted":
        f = torch.arange(T, device=dev, dtype=torch.int64)
        hk = _mix64(key[:, None] + (f[None, :] + 1) * _i64(0xE7037ED1A0B428DB)) & ((1 << 62) - 1)
        thresh = torch.sort(hk, dim=1).values[:, N - 1:N]               # the N frames with the smallest keys enter a token
        entry = hk <= thresh
        which = torch.cumsum(entry, 1) - 1
        col = torch.where(entry, torch.gather(tok, 1, which.clamp(0, N - 1)), torch.full_like(which, blank))
        em[:, :T].scatter_add_(2, col[:, :, None], torch.full((S, T, 1), 8.0, device=dev))
    elif kind != "iid":
        raise ValueError(kind)
    meta = SynthBatch(None, np.arange(S + 1, dtype=np.int64) * Tp, np.full(S, T, np.int32),
                      np.arange(S + 1, dtype=np.int64) * N, tok.to(torch.int32).cpu().numpy().reshape(-1),
                      np.full(S, float(duration_s)), vocab, blank)
    return em.view(S * Tp, vocab), meta


def device_c4_batch(seg_ids, device, vocab: int = 100, kind: str = "iid", seed: int = 4321):
    """BASELINE configs[3] on the device: segments of 40-60 s, 500 .. min(2000, T) tokens, vocabulary `vocab` with the
    blank at the last id (HF '<pad>' style) and separator id 4 — the shape of make_batch(config="c4"), generated by
    the counter-based generator so that a GPU-filling batch (thousands of segments, gigabytes of logits) does not have
    to be drawn on the host.  Every segment occupies Tmax_padded rows of the emission matrix (frame_len holds its real
    length); rows beyond a segment's length are zero.  Returns (emissions [S * Tp, V] CUDA float32, SynthBatch)."""
    import torch
    dev = torch.device(device)
    ids = torch.as_tensor(np.asarray(seg_ids, dtype=np.int64), device=dev)
    S = int(ids.numel())
    blank, sep = vocab - 1, 4
    key = _mix64(ids * _i64(0x9E3779B97F4A7C15) + _i64(seed))
    dur = 40.0 + 20.0 * _uniform24(_mix64(key + 1), 20).double()                       # U[40, 60) seconds
    T = ((dur * SAMPLE_RATE).long() - 400) // 320 + 1
    hi = torch.clamp(T, max=2000)
    N = 500 + (_uniform24(_mix64(key + 2), 20).double() * (hi - 500 + 1).double()).long().clamp(max=hi - 500)
    Tmax = frames_for(60.0)
    Tp = Tmax + (-Tmax) % 4
    Nmax = 2000
    k = torch.arange(Nmax, device=dev, dtype=torch.int64)
    ht = _mix64(key[:, None] + (k[None, :] + 1) * _i64(0xD1B54A32D192ED03))
    letters = torch.tensor([v for v in range(vocab) if v != blank and v != sep], device=dev, dtype=torch.int64)
    tok = letters[((ht >> 8) & 0x7FFFFFFF) % letters.numel()]
    is_sep = ((ht >> 44) & 0xFFFF) < (0x10000 * 2 // 11)
    is_sep[:, 0] = False
    is_sep[:, 1:] &= ~is_sep[:, :-1]
    tok = torch.where(is_sep, torch.full_like(tok, sep), tok)
    last = (N - 1).clamp(min=0)
    fix = tok.gather(1, last[:, None]) == sep                                          # no separator at the end of a transcript
    tok.scatter_(1, last[:, None], torch.where(fix, letters[0].expand_as(fix), tok.gather(1, last[:, None])))
    em = torch.zeros((S, Tp, vocab), dtype=torch.float32, device=dev)
    e_idx = torch.arange(Tmax * vocab, device=dev, dtype=torch.int64)
    step = max(1, (1 << 27) // (Tmax * vocab))
    for a in range(0, S, step):
        h = _mix64(key[a:a + step, None] ^ ((e_idx[None, :] + 1) * _i64(0xA0761D6478BD642F)))
        z = torch.sqrt(-2.0 * torch.log(_uniform24(h, 40))) * torch.cos(6.283185307179586 * _uniform24(h, 12))
        em[a:a + step, :Tmax] = (z * 3.0).view(-1, Tmax, vocab)
    live = torch.arange(Tp, device=dev)[None, :] < T[:, None]
    em *= live[:, :, None]
    if kind != "iid":
        raise ValueError("device_c4_batch generates iid logits only")
    Th, Nh = T.cpu().numpy(), N.cpu().numpy()
    tok_h = tok.to(torch.int32).cpu().numpy()
    tokens = np.concatenate([tok_h[s, :Nh[s]] for s in range(S)]) if S else np.zeros(0, np.int32)
    meta = SynthBatch(None, np.arange(S + 1, dtype=np.int64) * Tp, Th.astype(np.int32),
                      np.concatenate([[0], np.cumsum(Nh)]).astype(np.int64), tokens, dur.cpu().numpy().astype(np.float64), vocab, blank)
    return em.view(S * Tp, vocab), meta
