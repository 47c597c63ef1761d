"""Drop-in for the reference's `assign_word_speakers` (whisperx/diarize.py:35-67): the speaker of
every aligned segment and word is the one whose diarization turns overlap it most.

The reference recomputes two DataFrame columns and a groupby per segment and per word (~1 ms each
in pandas); here all segment and word intervals of the call go to the GPU as one batch of queries
(`wxb_assign_speakers_*`, one thread per interval, per-speaker sums accumulated in row order with
the same compensated summation pandas uses).  The diarization pipeline itself (pyannote) is out of
scope; `diarize_df` is whatever it produced: a DataFrame (or mapping) with `start`, `end`, `speaker`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .engine import get_engine


def assign_word_speakers(diarize_df, transcript_result, fill_nearest=False, device="cuda:0"):
    """Same arguments, in-place effects and return value as the reference: `speaker` is set on every
    segment / word with an overlapping turn (any turn with fill_nearest), `transcript_result` is
    returned, and `diarize_df` ends up with the `intersection` / `union` columns of the last interval
    the reference would have processed (diarize.py:39-40, 52-53)."""
    eng = get_engine(device)
    start = np.ascontiguousarray(np.asarray(diarize_df["start"], dtype=np.float64))
    end = np.ascontiguousarray(np.asarray(diarize_df["end"], dtype=np.float64))
    labels = list(diarize_df["speaker"])
    order = sorted(set(labels))                              # groupby("speaker") sorts its keys
    ids = {lab: i for i, lab in enumerate(order)}
    spk = np.ascontiguousarray([ids[lab] for lab in labels], dtype=np.int32)

    targets, qs, qe = [], [], []
    for seg in transcript_result["segments"]:
        targets.append(seg); qs.append(seg["start"]); qe.append(seg["end"])
        if "words" in seg:
            for word in seg["words"]:
                if "start" in word:
                    targets.append(word); qs.append(word["start"]); qe.append(word["end"])
    if not targets:
        return transcript_result
    q_start = np.ascontiguousarray(qs, dtype=np.float64)
    q_end = np.ascontiguousarray(qe, dtype=np.float64)
    out = np.empty(len(targets), dtype=np.int32)
    f64, i32 = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    N.check(eng._lib.wxb_assign_speakers_host(
        eng._ctx, start.ctypes.data_as(f64), end.ctypes.data_as(f64), spk.ctypes.data_as(i32), len(spk), len(order),
        q_start.ctypes.data_as(f64), q_end.ctypes.data_as(f64), len(targets), 1 if fill_nearest else 0,
        out.ctypes.data_as(i32)), "wxb_assign_speakers_host")
    for tgt, k in zip(targets, out.tolist()):
        if k >= 0:
            tgt["speaker"] = order[k]
    # the reference leaves its scratch columns on the caller's table; keep that observable state
    try:
        diarize_df["intersection"] = np.minimum(end, q_end[-1]) - np.maximum(start, q_start[-1])
        diarize_df["union"] = np.maximum(end, q_end[-1]) - np.minimum(start, q_start[-1])
    except TypeError:
        pass                                                  # an immutable mapping: nothing to keep
    return transcript_result
