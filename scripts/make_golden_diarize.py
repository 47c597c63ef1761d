"""Generates tests/golden/diarize_cases.json by running the REAL reference
`whisperx.diarize.assign_word_speakers` (build container only; pyannote stubbed at import).

    python scripts/make_golden_diarize.py
"""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import pandas as pd

from oracle import ref_import


def scenario(seed, n_turns, n_speakers, n_segments, ties=False):
    rng = np.random.default_rng(seed)
    t, turns = 0.0, []
    for _ in range(n_turns):
        gap = float(rng.uniform(-0.5, 1.0))                    # negative gap = overlapping turns
        d = float(rng.uniform(0.2, 6.0))
        s = max(t + gap, 0.0)
        turns.append({"start": round(s, 3), "end": round(s + d, 3), "speaker": f"SPEAKER_{int(rng.integers(0, n_speakers)):02d}"})
        t = s + d
    if ties:                                                   # identical turns for two speakers: equal sums
        turns += [dict(t_, speaker="SPEAKER_98") for t_ in turns[:3]] + [dict(t_, speaker="SPEAKER_00") for t_ in turns[:3]]
    segs, cur = [], 0.0
    for i in range(n_segments):
        d = float(rng.uniform(0.5, 8.0))
        seg = {"start": round(cur, 3), "end": round(cur + d, 3), "text": f"segment {i}", "words": []}
        w = seg["start"]
        while w < seg["end"] - 0.05:
            wd = float(rng.uniform(0.05, 0.6))
            word = {"word": "w"}
            if rng.random() > 0.15:                            # some words carry no timestamps
                word.update(start=round(w, 3), end=round(min(w + wd, seg["end"]), 3), score=0.5)
            seg["words"].append(word)
            w += wd + float(rng.uniform(0, 0.2))
        if i % 5 == 4:
            del seg["words"]                                   # a segment without a word list
        segs.append(seg)
        cur += d + float(rng.uniform(-0.2, 3.0))
    segs.append({"start": t + 50.0, "end": t + 51.0, "text": "far after the last turn", "words": [{"word": "x", "start": t + 50.1, "end": t + 50.2}]})
    return turns, segs


def main():
    D = ref_import.load_reference_diarize()
    out = {}
    for name, kw in (("basic", dict(seed=1, n_turns=40, n_speakers=3, n_segments=12)),
                     ("many_speakers", dict(seed=2, n_turns=120, n_speakers=9, n_segments=25)),
                     ("ties", dict(seed=3, n_turns=20, n_speakers=2, n_segments=8, ties=True)),
                     ("no_turns", dict(seed=4, n_turns=0, n_speakers=1, n_segments=3))):
        turns, segs = scenario(**kw)
        for fill in (False, True):
            df = pd.DataFrame(turns, columns=["start", "end", "speaker"])
            res = D.assign_word_speakers(df, {"segments": copy.deepcopy(segs)}, fill_nearest=fill)
            out[f"{name}_fill{int(fill)}"] = {"turns": turns, "segments": segs, "fill_nearest": fill, "result": res,
                                              "last_intersection": [float(x) for x in df.get("intersection", [])]}
            n = sum("speaker" in s for s in res["segments"]) + sum("speaker" in w for s in res["segments"] for w in s.get("words", []))
            print(name, fill, "speakers assigned:", n)
    with open(os.path.join(ROOT, "tests", "golden", "diarize_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
