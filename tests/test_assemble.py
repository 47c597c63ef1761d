"""CPU: the host side of align() (text preparation, word groups, record assembly, NaN interpolation)
against the reference's own outputs in tests/golden/align_cases.* (scripts/make_golden.py).

The per-character values come from the oracle here (torch port: bit-identical to the reference),
the exact rounding / numpy-order mean from the library's host helpers; the GPU tests
(test_align_gpu.py) run the same scenarios through the kernels."""
import copy
import json
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import ref_port
from tests.helpers import DICTIONARY, regex_spans, same_records
from whisperx_b200 import _native, assemble
from whisperx_b200.types import check_aligned_result


@pytest.fixture(autouse=True)
def _splitter():
    assemble.set_sentence_splitter(regex_spans)
    yield
    assemble.set_sentence_splitter(None)


def _host_align(meta, emissions, lib):
    """align() with the DP done by the oracle and everything else by the product's host code."""
    lang = meta["lang"]
    transcript = copy.deepcopy(meta["segments"])
    tround = lib.wxb_round3_np if meta.get("numpy_times") else lib.wxb_round3     # np.float64 times round numpy's way
    assemble.prepare_transcript(transcript, DICTIONARY, lang)
    max_dur = meta["audio_samples"] / 16000
    out, k = [], 0
    for seg in transcript:
        plain = {"start": seg["start"], "end": seg["end"], "text": seg["text"], "words": []}
        if meta["return_char_alignments"]:
            plain["chars"] = []
        if not seg["clean_char"] or seg["start"] >= max_dur:
            out.append(plain)
            continue
        E = torch.from_numpy(emissions[k]); k += 1
        tokens = [DICTIONARY[c] for c in seg["clean_char"]]
        chars, R = ref_port.align_tokens(E, tokens, 0, "".join(seg["clean_char"]))
        if chars is None:
            out.append(plain)
            continue
        ratio = (seg["end"] - seg["start"]) * 1 / E.shape[0]
        cs = [tround(c.start * ratio + seg["start"]) for c in chars]
        ce = [tround(c.end * ratio + seg["start"]) for c in chars]
        sc = [lib.wxb_round3(c.score) for c in chars]
        lay = assemble.layout_segment(seg["text"], seg["clean_cdx"], seg["sentence_spans"], lang)
        ws, we, wsc = [], [], []
        for g in range(lay.n_groups):
            idx = lay.group_idx[lay.group_off[g]:lay.group_off[g + 1]]
            al = [i for i in idx if i >= 0]
            if not al:
                ws.append(math.nan); we.append(math.nan); wsc.append(math.nan)
                continue
            z = np.array([sc[i] if i >= 0 else 0.0 for i in idx], np.float64)
            ws.append(cs[al[0]]); we.append(ce[al[-1]])
            wsc.append(lib.wxb_round3_np(lib.wxb_pairwise_sum(z.ctypes.data_as(C.POINTER(C.c_double)), len(z)) / len(al)))
        out += assemble.assemble_segment(seg["text"], seg["clean_cdx"], lay, seg["sentence_spans"], lang, cs, ce, sc,
                                         ws, we, wsc, meta["interpolate_method"], meta["return_char_alignments"])
    return transcript, out


def test_prepare_and_assemble_match_reference(golden_align):
    meta, arr = golden_align
    lib = _native.load()
    for name, m in meta.items():
        em = [arr[f"{name}_E{i}"] for i in range(m["n_emissions"])]
        transcript, segs = _host_align(m, em, lib)
        for t, exp in zip(transcript, m["mutated_inputs"]):
            assert t["clean_char"] == exp["clean_char"] and t["clean_cdx"] == exp["clean_cdx"], name
            assert t["clean_wdx"] == exp["clean_wdx"], name
            assert [list(s) for s in t["sentence_spans"]] == exp["sentence_spans"], name
        diff = same_records(segs, m["result"]["segments"])
        assert diff is None, (name, diff)
        words = [w for s in segs for w in s["words"]]
        assert same_records(words, m["result"]["word_segments"]) is None, name
        # the schema (whisperx/types.py:4-58) holds for the reference's dict and for ours
        check_aligned_result(m["result"], char_alignments=m["return_char_alignments"])
        check_aligned_result({"segments": segs, "word_segments": words}, char_alignments=m["return_char_alignments"])


def test_schema_checker_rejects_malformed_results():
    good = {"segments": [{"start": 0.0, "end": 1.0, "text": "a b", "words": [{"word": "a", "start": 0.0, "end": 0.5, "score": 0.9}, {"word": "b"}]}]}
    good["word_segments"] = list(good["segments"][0]["words"])
    check_aligned_result(good, char_alignments=False)
    for breakage in (lambda r: r.pop("word_segments"), lambda r: r["segments"][0].pop("words"),
                     lambda r: r["segments"][0]["words"][0].pop("end"), lambda r: r["segments"][0]["words"][1].update(start=None),
                     lambda r: r["word_segments"].pop(), lambda r: r["segments"][0]["words"][0].update(score=1)):
        bad = copy.deepcopy(good)
        breakage(bad)
        with pytest.raises(TypeError):
            check_aligned_result(bad, char_alignments=False)
    from whisperx_b200 import types as T
    assert set(T.SingleAlignedSegment.__annotations__) >= {"start", "end", "text", "words", "chars"}
    assert set(T.AlignedTranscriptionResult.__annotations__) == {"segments", "word_segments"}


def test_interpolate_nans_matches_pandas():
    import pandas as pd
    rng = np.random.default_rng(11)
    for method in ("nearest", "linear"):
        for _ in range(400):
            n = int(rng.integers(1, 12))
            v = np.round(rng.random(n) * 30, 3)
            v[rng.random(n) < 0.45] = np.nan
            s = pd.Series(v)
            exp = (s.interpolate(method=method).ffill().bfill() if s.notnull().sum() > 1 else s.ffill().bfill()).tolist()
            got = assemble.interpolate_nans(v.tolist(), method)
            assert len(got) == len(exp)
            for a, b in zip(got, exp):
                assert (a != a and b != b) or a == b, (method, v.tolist(), got, exp)


def test_layout_word_groups():
    text = " ab c.  d"
    lay = assemble.layout_segment(text, [1, 2, 3, 4, 8], [(1, 6), (8, 9)], "en")
    assert lay.word_index == [0, 0, 0, 1, 1, 1, 2, 3, 3]
    assert lay.windows == [(1, 6), (8, 8)]
    # window 1 = "ab c. " (inclusive end): words "ab", " c." ; the lone space that is word 2 is dropped
    assert [w for w, _ in lay.sentence_words[0]] == ["ab", "c."]
    assert [w for w, _ in lay.sentence_words[1]] == ["d"]
    groups = [lay.group_idx[lay.group_off[g]:lay.group_off[g + 1]] for g in range(lay.n_groups)]
    assert groups == [[0, 1], [3, -1], [4]]       # the space before "c" (clean index 2) is not part of the word rows
    ja = assemble.layout_segment("ab c", [0, 1, 3], [(0, 4)], "ja")
    assert [w for w, _ in ja.sentence_words[0]] == ["a", "b", "c"]


def test_fallback_splitter():
    assemble.set_sentence_splitter(assemble._fallback_spans)
    assert assemble.sentence_spans("Hello world. It costs 20! ok") == [(0, 12), (13, 25), (26, 28)]
    assert assemble.sentence_spans("Dr. Smith is here. Yes") == [(0, 18), (19, 22)]
    assert assemble.sentence_spans(" lead and trail ") == [(1, 15)]
    assert assemble.blank_id_of({"a": 1, "<pad>": 7, "b": 2}) == 7 and assemble.blank_id_of({"a": 3}) == 0


def _device_like_words(lay, cs, ce, sc, lib):
    ws, we, wsc = [], [], []
    for g in range(lay.n_groups):
        idx = lay.group_idx[lay.group_off[g]:lay.group_off[g + 1]]
        al = [i for i in idx if i >= 0]
        if not al:
            ws.append(math.nan); we.append(math.nan); wsc.append(math.nan)
            continue
        z = np.array([sc[i] if i >= 0 else 0.0 for i in idx], np.float64)
        ws.append(cs[al[0]]); we.append(ce[al[-1]])
        wsc.append(lib.wxb_round3_np(lib.wxb_pairwise_sum(z.ctypes.data_as(C.POINTER(C.c_double)), len(z)) / len(al)))
    return ws, we, wsc


@pytest.mark.parametrize("lang", ["en", "ja"])
def test_random_texts_match_the_pandas_tables(lang):
    """Differential test of the record assembly (alignment.py:252-347) on random texts: odd spacing,
    unknown characters, words straddling sentence windows, sentences without any aligned character,
    both interpolation methods, with and without char alignments — against the oracle's pandas
    restatement fed the same per-character frames and scores."""
    lib = _native.load()
    rng = np.random.default_rng(17 if lang == "en" else 18)
    alphabet = list("abcdefg hij  klm.!?,'0123 ") + ["é", "-"]
    for case in range(120):
        L = int(rng.integers(1, 60))
        text = "".join(rng.choice(alphabet, L))
        if not text.strip():
            text = text + "a"
        seg = {"start": round(float(rng.uniform(0, 50)), 3), "text": text}
        seg["end"] = round(seg["start"] + float(rng.uniform(0.3, 20)), 3)
        transcript = [seg]
        assemble.prepare_transcript(transcript, DICTIONARY, lang)
        ref_seg = {"start": seg["start"], "end": seg["end"], "text": text}
        ref_port.prepare_segments([ref_seg], DICTIONARY, lang, regex_spans)
        assert ref_seg["clean_char"] == seg["clean_char"] and ref_seg["clean_cdx"] == seg["clean_cdx"], text
        assert ref_seg["clean_wdx"] == seg["clean_wdx"] and [tuple(x) for x in ref_seg["sentence_spans"]] == seg["sentence_spans"]
        N = len(seg["clean_char"])
        if N == 0:
            continue
        T = N + int(rng.integers(0, 80))
        starts = np.sort(rng.choice(T, N, replace=False))
        ends = np.append(starts[1:], min(T, starts[-1] + int(rng.integers(1, 5))))
        scores = np.round(rng.random(N) ** 2, 6)
        chars = [ref_port.Segment(c, int(a), int(b), float(s)) for c, a, b, s in zip(seg["clean_char"], starts, ends, scores)]
        ratio = (seg["end"] - seg["start"]) / T
        method = "nearest" if case % 2 else "linear"
        want_chars = bool(case % 3 == 0)
        exp = ref_port.segment_tables(ref_seg, chars, ratio, lang, method, want_chars)
        cs = [lib.wxb_round3(c.start * ratio + seg["start"]) for c in chars]
        ce = [lib.wxb_round3(c.end * ratio + seg["start"]) for c in chars]
        sc = [lib.wxb_round3(c.score) for c in chars]
        lay = assemble.layout_segment(text, seg["clean_cdx"], seg["sentence_spans"], lang)
        ws, we, wsc = _device_like_words(lay, cs, ce, sc, lib)
        got = assemble.assemble_segment(text, seg["clean_cdx"], lay, seg["sentence_spans"], lang, cs, ce, sc, ws, we, wsc,
                                        method, want_chars)
        exp = json.loads(json.dumps(exp, default=float))
        d = same_records(got, exp)
        assert d is None, (text, method, want_chars, d)
