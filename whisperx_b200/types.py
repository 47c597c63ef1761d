"""Schema of the dictionaries on the alignment path (reference `whisperx/types.py:4-58`).

Same type names as the reference so annotations keep working.  Keys that `align()` *omits* for
characters / words it could not time (`alignment.py:311-318, 330-334`) are declared optional here
(`total=False` parts), which is what the data really looks like; `check_aligned_result()` is the
executable form of the schema used by the tests."""
from __future__ import annotations

from typing import List, Optional, TypedDict

_Timing = TypedDict("_Timing", {"start": float, "end": float, "score": float}, total=False)


class SingleWordSegment(_Timing):
    word: str


class SingleCharSegment(_Timing):
    char: str


class SingleSegment(TypedDict):
    start: float
    end: float
    text: str


class _AlignedExtras(TypedDict, total=False):
    chars: Optional[List[SingleCharSegment]]          # only with return_char_alignments=True
    speaker: str                                      # added by assign_word_speakers


class SingleAlignedSegment(SingleSegment, _AlignedExtras):
    words: List[SingleWordSegment]


class TranscriptionResult(TypedDict):
    segments: List[SingleSegment]
    language: str


class AlignedTranscriptionResult(TypedDict):
    segments: List[SingleAlignedSegment]
    word_segments: List[SingleWordSegment]


def _check_timed(item: dict, label_key: str, where: str) -> None:
    if not isinstance(item.get(label_key), str):
        raise TypeError(f"{where}: '{label_key}' must be a str")
    extra = set(item) - {label_key, "start", "end", "score", "speaker"}
    if extra:
        raise TypeError(f"{where}: unexpected keys {sorted(extra)}")
    timing = [k for k in ("start", "end", "score") if k in item]
    if timing and len(timing) != 3:
        raise TypeError(f"{where}: start / end / score come together or not at all, got {timing}")
    for k in timing:
        if not isinstance(item[k], float):
            raise TypeError(f"{where}: '{k}' must be a float, got {type(item[k]).__name__}")


def check_aligned_result(result: dict, char_alignments: Optional[bool] = None) -> None:
    """Raise TypeError unless `result` has the shape `align()` returns (`alignment.py:349-354`)."""
    if set(result) - {"segments", "word_segments", "language"} or "segments" not in result or "word_segments" not in result:
        raise TypeError(f"result keys {sorted(result)}")
    flat = []
    for i, seg in enumerate(result["segments"]):
        where = f"segments[{i}]"
        for k in ("start", "end"):
            if not isinstance(seg.get(k), (float, int)):
                raise TypeError(f"{where}: '{k}' missing or not a number")
        if not isinstance(seg.get("text"), str) or not isinstance(seg.get("words"), list):
            raise TypeError(f"{where}: needs 'text' (str) and 'words' (list)")
        for j, w in enumerate(seg["words"]):
            _check_timed(w, "word", f"{where}.words[{j}]")
        flat += seg["words"]
        if char_alignments is not None and (("chars" in seg) != char_alignments) and seg["words"]:
            raise TypeError(f"{where}: 'chars' {'missing' if char_alignments else 'unexpected'}")
        for j, c in enumerate(seg.get("chars") or []):
            _check_timed(c, "char", f"{where}.chars[{j}]")
    if result["word_segments"] != flat:
        raise TypeError("word_segments is not the concatenation of the segments' words")
