"""whisperx_b200.align() end to end on BASELINE configs[2]: a 1-hour recording cut into 30 s segments
(random-init WAV2VEC2_ASR_BASE_960H, synthetic audio and transcripts), next to the reference's
align() algorithm (oracle port: same GPU forward, then its per-segment CPU trellis / backtrack /
pandas tables) on a sample of the segments.  Prints one JSON line.

    python scripts/bench_align_e2e.py [--minutes 60] [--ref-segments 8]
"""
import argparse
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=60.0)
    ap.add_argument("--ref-segments", type=int, default=8)
    args = ap.parse_args()
    import torchaudio
    from oracle import ref_port
    from whisperx_b200 import alignment as wxa

    labels = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H.get_labels()
    dictionary = {c.lower(): i for i, c in enumerate(labels)}
    meta = {"language": "en", "dictionary": dictionary, "type": "torchaudio"}
    torch.manual_seed(0)
    model = torchaudio.models.wav2vec2_model(**torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params).eval().to("cuda:0")
    n_seg = int(args.minutes * 2)
    g = torch.Generator().manual_seed(1)
    audio = (torch.randn(n_seg * 30 * 16000, generator=g) * 0.1).numpy()
    rng = np.random.default_rng(2)
    letters = "abcdefghijklmnopqrstuvwxyz'"
    segs = []
    for i in range(n_seg):
        words = []
        while sum(len(w) + 1 for w in words) < 400:
            words.append("".join(rng.choice(list(letters), int(rng.integers(1, 9)))))
        text = " ".join(words)
        text = text[:200] + ". " + text[200:].capitalize()
        segs.append({"start": 30.0 * i, "end": 30.0 * (i + 1), "text": text})
    wxa.set_sentence_splitter(wxa.assemble._fallback_spans)

    out = {"config": f"{args.minutes:.0f} min recording, {n_seg} x 30 s segments, ~400 chars each, random-init WAV2VEC2_ASR_BASE_960H"}
    for batch in (1, 8):
        wxa.EMISSION_BATCH = batch
        wxa.align(copy.deepcopy(segs[:4]), model, meta, audio[: 4 * 30 * 16000], "cuda:0")        # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = wxa.align(copy.deepcopy(segs), model, meta, audio, "cuda:0", return_char_alignments=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[f"ours_emission_batch_{batch}"] = {"seconds": dt, "audio_hours_per_s": args.minutes / 60 / dt,
                                                "segments_out": len(res["segments"]), "words": len(res["word_segments"])}
    wxa.EMISSION_BATCH = 1
    # where the time goes in our align(): forward only
    with torch.inference_mode():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in segs:
            model(torch.from_numpy(audio[int(s["start"] * 16000):int(s["end"] * 16000)]).unsqueeze(0).to("cuda:0"))
        torch.cuda.synchronize()
        out["wav2vec2_forward_batch1_seconds"] = time.perf_counter() - t0
    k = min(args.ref_segments, n_seg)
    sample = copy.deepcopy(segs[:k])
    t0 = time.perf_counter()
    ref = ref_port.align_port(sample, model, meta, audio[: k * 30 * 16000], "cuda:0", return_char_alignments=True,
                              span_fn=wxa.assemble._fallback_spans)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["reference_port"] = {"segments": k, "seconds": dt, "audio_hours_per_s": k * 30 / 3600 / dt,
                             "note": "reference algorithm (oracle/ref_port.py): GPU forward, CPU trellis/backtrack/pandas per segment"}
    out["speedup_vs_reference_port"] = out["ours_emission_batch_1"]["audio_hours_per_s"] / out["reference_port"]["audio_hours_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
