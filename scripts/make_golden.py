"""Generates tests/golden/* by running the REAL reference (/root/reference, whisperX 3.1.1).

Only runs in the build container (the GPU box has no /root/reference).  The outputs are small and
committed; tests/ compare the oracle (-m "not gpu") and the CUDA engine (-m gpu) against them.

    python scripts/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import ref_import

GOLD = os.path.join(ROOT, "tests", "golden")


def log_softmax_np(x):
    return torch.log_softmax(torch.from_numpy(x), -1)


def dp_cases():
    """(name, T, V, N, blank, kind) — edge-case matrix of SURVEY §4 item 3."""
    cases = []
    k = 0
    for T, N in ((1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (5, 4), (5, 5), (5, 6), (9, 1), (17, 8), (33, 32),
                 (33, 33), (40, 34), (64, 20), (64, 63), (64, 64), (64, 65), (50, 7), (31, 30), (48, 12)):
        for kind in ("iid", "ties"):
            for V, blank in ((29, 0), (7, 3)):
                cases.append((f"c{k:03d}", T, V, N, blank, kind))
                k += 1
    return cases


def make_dp(A):
    rng = np.random.default_rng(20240517)
    out = {}
    meta = []
    for name, T, V, N, blank, kind in dp_cases():
        x = rng.standard_normal((T, V)).astype(np.float32) * 3
        E = log_softmax_np(x)
        if kind == "ties":
            E = torch.round(E * 4) / 4
        if name.endswith(("3", "7")):
            tok = np.repeat(rng.integers(0, V, (N + 1) // 2), 2)[:N].astype(np.int32)   # repeated tokens
        else:
            tok = rng.integers(0, V, N).astype(np.int32)
        R = A.get_trellis(E, tok.tolist(), blank)
        path = A.backtrack(R, E, tok.tolist(), blank)
        out[f"{name}_E"] = E.numpy()
        out[f"{name}_tok"] = tok
        out[f"{name}_R"] = R.numpy()
        ok = path is not None
        if ok:
            out[f"{name}_ptok"] = np.array([p.token_index for p in path], np.int32)
            out[f"{name}_ptime"] = np.array([p.time_index for p in path], np.int32)
            out[f"{name}_pscore"] = np.array([p.score for p in path], np.float64)
            transcript = "".join(chr(ord("a") + int(t) % 26) if int(t) != 1 else "|" for t in tok)
            segs = A.merge_repeats(path, transcript)
            out[f"{name}_sstart"] = np.array([s.start for s in segs], np.int32)
            out[f"{name}_send"] = np.array([s.end for s in segs], np.int32)
            out[f"{name}_sscore"] = np.array([s.score for s in segs], np.float64)
            words = A.merge_words(segs)
            out[f"{name}_wstart"] = np.array([w.start for w in words], np.int32)
            out[f"{name}_wend"] = np.array([w.end for w in words], np.int32)
            out[f"{name}_wscore"] = np.array([w.score for w in words], np.float64)
            meta.append(dict(name=name, T=T, V=V, N=N, blank=blank, kind=kind, ok=True, transcript=transcript,
                             words=[w.label for w in words]))
        else:
            meta.append(dict(name=name, T=T, V=V, N=N, blank=blank, kind=kind, ok=False))
    np.savez_compressed(os.path.join(GOLD, "dp_cases.npz"), **out)
    with open(os.path.join(GOLD, "dp_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("dp cases:", len(meta), "failed-by-design:", sum(not m["ok"] for m in meta))


# ----------------------------------------------------------------------------- align()
LABELS = ('-', '|', 'E', 'T', 'A', 'O', 'N', 'I', 'H', 'S', 'R', 'D', 'L', 'U', 'M', 'W', 'C', 'F', 'G', 'Y',
          'P', 'B', 'V', 'K', "'", 'X', 'J', 'Q', 'Z')     # torchaudio WAV2VEC2_ASR_BASE_960H labels


def regex_spans(text):
    """Sentence splitter shared by reference and engine in the fixtures (nltk is not installed):
    split after . ! ? followed by whitespace, spans exclude the separating whitespace."""
    import re
    spans, start = [], 0
    for m in re.finditer(r"[.!?]+(?=\s+\S)", text):
        spans.append((start, m.end()))
        nxt = re.compile(r"\S").search(text, m.end())
        start = nxt.start() if nxt else len(text)
    spans.append((start, len(text)))
    # Punkt never yields leading whitespace in a span
    out = []
    for a, b in spans:
        while a < b and text[a].isspace():
            a += 1
        out.append((a, b))
    return out


class ScriptedModel(torch.nn.Module):
    """Returns prepared logits, one tensor per call, in call order (torchaudio-style signature)."""

    def __init__(self, logits_list):
        super().__init__()
        self.logits_list = logits_list
        self.calls = 0

    def forward(self, waveform, lengths=None):
        x = self.logits_list[self.calls]
        self.calls += 1
        return x.unsqueeze(0), None


def frames_for_samples(n):
    return max((max(n, 400) - 400) // 320 + 1, 1)


ALIGN_SCENARIOS = {
    "en_basic": dict(lang="en", audio_s=12.0, segments=[
        dict(start=0.5, end=4.0, text=" Hello world. It costs 20 dollars! ok"),
        dict(start=4.2, end=7.0, text="The quick brown fox jumps over the lazy dog."),
        dict(start=7.0, end=7.01, text="tiny"),
        dict(start=7.5, end=8.0, text=" 2014 "),
        dict(start=8.0, end=8.2, text="this text is far too long for a fifth of a second"),
        dict(start=13.0, end=14.0, text="after the end"),
        dict(start=8.5, end=11.9, text="Dr. Smith isn't here...  Two  spaces, and a trailing one "),
    ]),
    "en_chars": dict(lang="en", audio_s=9.0, return_char_alignments=True, segments=[
        dict(start=0.0, end=3.5, text="Wait, what? No way! Yes."),
        dict(start=3.5, end=8.8, text="a b c d e f g h i j k l m n o p q r s t u v w x y z"),
        dict(start=8.8, end=9.5, text="clipped by the audio end"),
    ]),
    "ja_like": dict(lang="ja", audio_s=6.0, return_char_alignments=True, segments=[
        dict(start=0.2, end=2.9, text="abc de.fg"),
        dict(start=3.0, end=5.5, text="xyz?!q"),
    ]),
    "en_linear": dict(lang="en", audio_s=8.0, interpolate_method="linear", segments=[
        dict(start=0.0, end=6.0, text="One two. 33 44! Five six. 7? eight nine."),
    ]),
    # sentences without any alignable character at the start, in the middle (two in a row) and at the end:
    # NaN bounds -> nearest interpolation, ffill / bfill, and groupby-merge of sentences that end up with equal keys
    "en_nan_sentences": dict(lang="en", audio_s=20.0, return_char_alignments=True, segments=[
        dict(start=0.0, end=7.0, text="2024! Hello there. 123. 456? General Kenobi! 789."),
        dict(start=7.0, end=9.0, text="99. 98. 97."),
        dict(start=9.0, end=15.0, text="Only one sentence here"),
        dict(start=15.0, end=19.5, text="A. B. 1. C."),
    ]),
    "en_nan_linear": dict(lang="en", audio_s=12.0, interpolate_method="linear", segments=[
        dict(start=0.25, end=11.0, text="7. first words. 8. 9. more words here! 10."),
    ]),
    # numpy-typed segment times (np.float64 arithmetic and numpy rounding of the timestamps)
    "en_numpy_times": dict(lang="en", audio_s=10.0, numpy_times=True, return_char_alignments=True, segments=[
        dict(start=0.125, end=4.375, text="Rounding follows numpy here. Not python!"),
        dict(start=4.375, end=9.8125, text="half way cases matter"),
    ]),
    "zh_like": dict(lang="zh", audio_s=7.0, segments=[
        dict(start=0.0, end=3.3, text="ab c。de!f g"),
        dict(start=3.3, end=6.9, text=" xy z? 12 w "),
    ]),
    "en_long": dict(lang="en", audio_s=31.0, segments=[
        dict(start=0.0, end=30.0, text="The birch canoe slid on the smooth planks. Glue the sheet to the dark blue background. "
             "It's easy to tell the depth of a well. These days a chicken leg is a rare dish. Rice is often served in round bowls. "
             "The juice of lemons makes fine punch. The box was thrown beside the parked truck."),
    ]),
}


def make_align(A):
    dictionary = {c.lower(): i for i, c in enumerate(LABELS)}
    out_json = {}
    arrays = {}
    ref_import.set_sentence_splitter(regex_spans)
    for name, sc in ALIGN_SCENARIOS.items():
        rng = np.random.default_rng(sum(ord(c) for c in name) * 7919)
        # the scripted model ignores the samples; only the length of the audio matters
        audio = np.zeros(int(sc["audio_s"] * 16000), np.float32)
        meta = {"language": sc["lang"], "dictionary": dictionary, "type": "torchaudio"}
        # prepared logits per segment that reaches the model (same guards as the reference)
        logits = []
        for seg in sc["segments"]:
            text = seg["text"]
            lead = len(text) - len(text.lstrip()); trail = len(text) - len(text.rstrip())
            clean = [c for i, c in enumerate(text) if not (i < lead or i > len(text) - trail - 1)
                     and (c.lower().replace(" ", "|") if sc["lang"] not in ("ja", "zh") else c.lower()) in dictionary]
            if not clean or seg["start"] >= sc["audio_s"]:
                continue
            f1, f2 = int(seg["start"] * 16000), int(seg["end"] * 16000)
            n = min(f2, audio.shape[0]) - f1
            T = frames_for_samples(n)
            x = rng.standard_normal((T, len(LABELS))).astype(np.float32) * 3
            # planted path so that most chars align with decent scores
            toks = [dictionary[(c.lower().replace(" ", "|") if sc["lang"] not in ("ja", "zh") else c.lower())] for c in clean]
            if len(toks) <= T:
                at = np.sort(rng.choice(T, len(toks), replace=False))
                boost = np.zeros(T, np.int64); boost[at] = toks
                x[np.arange(T), boost] += 6.0
            logits.append(torch.from_numpy(x))
        model = ScriptedModel(logits)
        transcript = [dict(s) for s in sc["segments"]]
        if sc.get("numpy_times"):
            for t in transcript:
                t["start"], t["end"] = np.float64(t["start"]), np.float64(t["end"])
        emissions = []
        # capture the log-probs the reference actually used
        orig = A.get_trellis

        def spy(emission, tokens, blank_id=0):
            emissions.append(emission.numpy().copy())
            return orig(emission, tokens, blank_id)

        A.get_trellis = spy
        try:
            res = A.align(transcript, model, meta, audio, "cpu",
                          interpolate_method=sc.get("interpolate_method", "nearest"),
                          return_char_alignments=sc.get("return_char_alignments", False))
        finally:
            A.get_trellis = orig
        for i, e in enumerate(emissions):
            arrays[f"{name}_E{i}"] = e
        out_json[name] = dict(lang=sc["lang"], audio_s=sc["audio_s"], audio_samples=int(audio.shape[0]),
                              interpolate_method=sc.get("interpolate_method", "nearest"),
                              return_char_alignments=sc.get("return_char_alignments", False),
                              numpy_times=sc.get("numpy_times", False),
                              segments=sc["segments"], n_emissions=len(emissions),
                              mutated_inputs=[{k: v for k, v in t.items() if k in ("clean_char", "clean_cdx", "clean_wdx", "sentence_spans")} for t in transcript],
                              result=json.loads(json.dumps(res, default=float)))
        print(name, "segments out:", len(res["segments"]), "words:", len(res["word_segments"]))
    np.savez_compressed(os.path.join(GOLD, "align_cases.npz"), **arrays)
    with open(os.path.join(GOLD, "align_cases.json"), "w") as f:
        json.dump(out_json, f, indent=1, allow_nan=True)


def main():
    os.makedirs(GOLD, exist_ok=True)
    A = ref_import.load_reference_alignment()
    make_dp(A)
    make_align(A)


if __name__ == "__main__":
    main()
