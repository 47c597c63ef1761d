"""whisperx_b200 — B200-native forced alignment behind whisperX's alignment API.

    from whisperx_b200 import load_align_model, align
    model, metadata = load_align_model("en", "cuda")
    result = align(segments, model, metadata, audio, "cuda")     # {"segments": [...], "word_segments": [...]}

Same names, arguments and result dict as `whisperx.load_align_model` / `whisperx.align`
(reference: whisperx/__init__.py:2, whisperx/alignment.py:64, :100).  The trellis, backtrack and
timestamp reduction run in hand-written sm_100a CUDA kernels behind a C ABI
(include/wxb_align.h); there is no CPU fallback.
"""
from .alignment import (Point, Segment, align, backtrack, get_trellis, load_align_model, load_audio,  # noqa: F401
                        merge_repeats, merge_words, set_sentence_splitter)
from .diarize import assign_word_speakers  # noqa: F401
from .engine import AlignEngine, BatchResult, BatchSpec, get_engine  # noqa: F401
from .writers import format_timestamp, get_writer  # noqa: F401

__all__ = ["load_align_model", "align", "load_audio", "get_trellis", "backtrack", "merge_repeats", "merge_words",
           "Point", "Segment", "set_sentence_splitter", "assign_word_speakers", "AlignEngine", "BatchSpec", "BatchResult", "get_engine",
           "get_writer", "format_timestamp"]
