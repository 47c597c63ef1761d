"""Generates tests/golden/writer_cases.json by running the REAL reference writers
(`whisperx/utils.py`: WriteTXT/VTT/SRT/TSV/Audacity/JSON through `write_result`) on synthetic results
(build container only).

    python scripts/make_golden_writers.py
"""
import io
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from oracle import ref_import

WORDS = ["the", "quick", "brown", "fox", "jumps", "over", "a", "lazy", "dog,", "while", "it's", "raining.", "Really?",
         "-->", "tab\there", "ünïcödé", "x", "supercalifragilistic"]
KANA = list("こんにちは世界今日は良い天気です")


def make_result(seed, language="en", n_segments=6, speakers=False, untimed=0.15, pauses=False, words=True):
    rng = np.random.default_rng(seed)
    t, segs = float(rng.uniform(0, 5)), []
    for i in range(n_segments):
        n = int(rng.integers(1, 14))
        vocab = KANA if language in ("ja", "zh") else WORDS
        ws, w0 = [], t
        for _ in range(n):
            d = float(rng.uniform(0.05, 0.7))
            w = {"word": str(rng.choice(vocab))}
            if rng.random() >= untimed:
                w.update(start=round(w0, 3), end=round(w0 + d, 3), score=round(float(rng.random()), 3))
            ws.append(w)
            w0 += d + (float(rng.uniform(2.5, 5.0)) if pauses and rng.random() < 0.2 else float(rng.uniform(0, 0.3)))
        joiner = "" if language in ("ja", "zh") else " "
        seg = {"start": round(t, 3), "end": round(w0, 3), "text": (" " if i % 2 else "") + joiner.join(w["word"] for w in ws) + " "}
        if words:
            seg["words"] = ws
        if speakers and i % 3 != 2:
            seg["speaker"] = f"SPEAKER_{i % 2:02d}"
        segs.append(seg)
        t = w0 + float(rng.uniform(0, 4.0))
    return {"segments": segs, "language": language}


def render(writer_cls, result, options):
    buf = io.StringIO()
    writer_cls(".").write_result(result, file=buf, options=options)
    return buf.getvalue()


def main():
    U = ref_import.load_reference_utils()
    results = {
        "plain": make_result(1),
        "speakers": make_result(2, speakers=True),
        "pauses": make_result(3, pauses=True, n_segments=8),
        "all_untimed_first": make_result(4, untimed=0.6),
        "japanese": make_result(5, language="ja"),
        "no_words": make_result(6, words=False, speakers=True),
        "one_word": make_result(7, n_segments=1),
        "empty": {"segments": [], "language": "en"},
        "hour_long": {"segments": [{"start": 3599.9996, "end": 7325.0625, "text": "late", "words": [{"word": "late", "start": 3600.0004, "end": 7325.0615}]}], "language": "en"},
    }
    grid = [dict(max_line_width=w, max_line_count=c, highlight_words=h)
            for w, c, h in itertools.product((None, 12, 30), (None, 1, 2), (False, True))]
    cases = []
    for name, res in results.items():
        for opt in grid:
            out = {"vtt": render(U.WriteVTT, res, opt), "srt": render(U.WriteSRT, res, opt)}
            cases.append({"result": name, "options": opt, "out": out})
        plain = {"max_line_width": None, "max_line_count": None, "highlight_words": False}
        cases.append({"result": name, "options": plain, "out": {
            "txt": render(U.WriteTXT, res, plain), "tsv": render(U.WriteTSV, res, plain),
            "aud": render(U.WriteAudacity, res, plain), "json": render(U.WriteJSON, res, plain)}})
    stamps = [0.0, 0.0004, 0.0005, 0.0015, 0.9995, 59.9994, 59.9996, 61.5, 3599.9995, 3600.0, 86399.999, 360000.25]
    fmt = [{"seconds": s, "hours": h, "marker": m, "out": U.format_timestamp(s, h, m)} for s in stamps for h in (False, True) for m in (".", ",")]
    path = os.path.join(ROOT, "tests", "golden", "writer_cases.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"generator": "scripts/make_golden_writers.py (reference whisperx/utils.py writers, unmodified)",
                   "results": results, "cases": cases, "format_timestamp": fmt}, f, ensure_ascii=False, indent=0)
    print(path, len(cases), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
