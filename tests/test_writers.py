"""CPU: the result writers (reference whisperx/utils.py:170-431) against golden
outputs generated from the reference's own classes, plus a live differential against the reference
where it is present (build container)."""
import io
import json
import os

import numpy as np
import pytest

from whisperx_b200 import writers as W

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "writer_cases.json")
CLASSES = {"vtt": W.WriteVTT, "srt": W.WriteSRT, "txt": W.WriteTXT, "tsv": W.WriteTSV, "aud": W.WriteAudacity, "json": W.WriteJSON}


def render(cls, result, options):
    buf = io.StringIO()
    cls(".").write_result(result, file=buf, options=options)
    return buf.getvalue()


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN, encoding="utf-8") as f:
        return json.load(f)


def test_every_format_matches_the_reference_bytes(golden):
    n = 0
    for case in golden["cases"]:
        res = golden["results"][case["result"]]
        for fmt, want in case["out"].items():
            got = render(CLASSES[fmt], res, case["options"])
            assert got == want, (case["result"], case["options"], fmt)
            n += 1
    assert n >= 300


def test_format_timestamp_matches(golden):
    for row in golden["format_timestamp"]:
        assert W.format_timestamp(row["seconds"], row["hours"], row["marker"]) == row["out"], row
    with pytest.raises(AssertionError):
        W.format_timestamp(-0.001)


def test_writers_do_not_modify_the_result(golden):
    res = golden["results"]["speakers"]
    before = json.dumps(res, sort_keys=True)
    render(W.WriteSRT, res, {"max_line_width": 12, "max_line_count": 2, "highlight_words": True})
    assert json.dumps(res, sort_keys=True) == before


def test_get_writer_writes_files(tmp_path, golden):
    res = golden["results"]["plain"]
    opts = {"max_line_width": None, "max_line_count": None, "highlight_words": False}
    W.get_writer("all", str(tmp_path))(res, "/some/dir/talk.final.wav", opts)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["talk.final.json", "talk.final.srt", "talk.final.tsv", "talk.final.txt", "talk.final.vtt"]
    assert (tmp_path / "talk.final.srt").read_text(encoding="utf-8") == render(W.WriteSRT, res, opts)
    W.get_writer("aud", str(tmp_path))(res, "talk.final.wav", opts)
    assert (tmp_path / "talk.final.aud").exists()
    assert isinstance(W.get_writer("vtt", str(tmp_path)), W.WriteVTT)
    with pytest.raises(KeyError):
        W.get_writer("docx", str(tmp_path))
    with pytest.raises(KeyError):                       # the options are read even for an empty result, like the reference
        render(W.WriteVTT, {"segments": [], "language": "en"}, {})
    with pytest.raises(NotImplementedError):
        W.ResultWriter(".").write_result(res, io.StringIO(), opts)


def test_align_output_goes_straight_into_the_writers():
    """The dict `align()` assembles (words without timings omit the keys) is what the writers read."""
    from whisperx_b200.assemble import assemble_segment  # noqa: F401  (the schema producer)
    result = {"language": "en", "segments": [
        {"start": 0.5, "end": 2.0, "text": " hello big world", "words": [
            {"word": "hello", "start": 0.5, "end": 0.9, "score": 0.9}, {"word": "big"}, {"word": "world", "start": 1.2, "end": 2.0, "score": 0.8}]}]}
    srt = render(W.WriteSRT, result, {"max_line_width": None, "max_line_count": None, "highlight_words": True})
    assert srt.count("<u>") == 2 and "00:00:00,900 --> 00:00:01,200\nhello big world" in srt


@pytest.mark.needs_reference
def test_live_differential_against_the_reference():
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("reference tree not present")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "scripts"))
    from make_golden_writers import make_result
    U = ref_import.load_reference_utils()
    ref = {"vtt": U.WriteVTT, "srt": U.WriteSRT, "txt": U.WriteTXT, "tsv": U.WriteTSV, "aud": U.WriteAudacity, "json": U.WriteJSON}
    rng = np.random.default_rng(0)
    for seed in range(100, 160):
        res = make_result(seed, language=("ja" if seed % 7 == 0 else "en"), n_segments=int(rng.integers(1, 9)),
                          speakers=bool(seed % 2), untimed=float(rng.choice([0.0, 0.2, 0.7, 1.0])), pauses=bool(seed % 3 == 0))
        opts = {"max_line_width": [None, 5, 16, 42][int(rng.integers(0, 4))], "max_line_count": [None, 1, 2, 3][int(rng.integers(0, 4))],
                "highlight_words": bool(rng.integers(0, 2))}
        for fmt in ref:
            assert render(CLASSES[fmt], res, opts) == render(ref[fmt], res, opts), (seed, opts, fmt)
