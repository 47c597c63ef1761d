"""CPU: the drop-in's public surface mirrors the reference's (whisperx/__init__.py:2-4, alignment.py:64-110)."""
import inspect

import pytest

import whisperx_b200 as wx
from oracle import ref_import


def test_exports_and_signatures():
    for name in ("load_align_model", "align", "load_audio", "assign_word_speakers", "get_trellis", "backtrack",
                 "merge_repeats", "merge_words", "Point", "Segment"):
        assert hasattr(wx, name), name
    assert list(inspect.signature(wx.load_align_model).parameters) == ["language_code", "device", "model_name", "model_dir"]
    p = inspect.signature(wx.align).parameters
    assert list(p) == ["transcript", "model", "align_model_metadata", "audio", "device", "interpolate_method",
                       "return_char_alignments", "print_progress", "combined_progress"]
    assert p["interpolate_method"].default == "nearest" and p["return_char_alignments"].default is False
    assert list(inspect.signature(wx.assign_word_speakers).parameters)[:3] == ["diarize_df", "transcript_result", "fill_nearest"]


@pytest.mark.needs_reference
def test_signatures_match_the_live_reference():
    A = ref_import.load_reference_alignment()
    assert list(inspect.signature(A.align).parameters) == list(inspect.signature(wx.align).parameters)
    assert list(inspect.signature(A.load_align_model).parameters) == list(inspect.signature(wx.load_align_model).parameters)
    for fn in ("get_trellis", "backtrack", "merge_repeats", "merge_words"):
        ref = list(inspect.signature(getattr(A, fn)).parameters)
        assert list(inspect.signature(getattr(wx, fn)).parameters)[:len(ref)] == ref, fn
    assert wx.alignment.DEFAULT_ALIGN_MODELS_TORCH == A.DEFAULT_ALIGN_MODELS_TORCH
    assert wx.alignment.DEFAULT_ALIGN_MODELS_HF == A.DEFAULT_ALIGN_MODELS_HF
    assert tuple(A.LANGUAGES_WITHOUT_SPACES) == tuple(wx.alignment.LANGUAGES_WITHOUT_SPACES)
    assert tuple(A.PUNKT_ABBREVIATIONS) == tuple(wx.alignment.PUNKT_ABBREVIATIONS)
    U = ref_import.load_reference_utils()
    for name in ("ResultWriter", "WriteTXT", "SubtitlesWriter", "WriteVTT", "WriteSRT", "WriteTSV", "WriteAudacity", "WriteJSON"):
        ref_cls, cls = getattr(U, name), getattr(wx.writers, name)
        assert getattr(ref_cls, "extension", None) == getattr(cls, "extension", None), name
        assert list(inspect.signature(ref_cls.write_result).parameters) == list(inspect.signature(cls.write_result).parameters), name
    assert list(inspect.signature(U.get_writer).parameters) == list(inspect.signature(wx.get_writer).parameters)
    assert list(inspect.signature(U.format_timestamp).parameters) == list(inspect.signature(wx.format_timestamp).parameters)
    D = ref_import.load_reference_diarize()
    assert list(inspect.signature(D.assign_word_speakers).parameters) == list(inspect.signature(wx.assign_word_speakers).parameters)[:3]


def test_unknown_language_raises_like_the_reference(capsys):
    with pytest.raises(ValueError, match="No default align-model for language: xx"):
        wx.load_align_model("xx", "cpu")
    assert "There is no default alignment model set for this language (xx)" in capsys.readouterr().out


def test_align_refuses_a_cpu_device():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wx.align([], None, {"language": "en", "dictionary": {}, "type": "torchaudio"}, __import__("numpy").zeros(16000, "float32"), "cpu")
