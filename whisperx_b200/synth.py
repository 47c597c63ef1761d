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
