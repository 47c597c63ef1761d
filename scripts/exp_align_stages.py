"""Development aid: per-stage wall times of whisperx_b200.align() on the 1-hour BASELINE configs[2] shape
(WXB_ALIGN_TIMING=1 prints them on stderr).     python scripts/exp_align_stages.py"""
import copy, os, sys
os.environ["WXB_ALIGN_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torchaudio
from whisperx_b200 import alignment as wxa

bundle = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H
meta = {"language": "en", "dictionary": {c.lower(): i for i, c in enumerate(bundle.get_labels())}, "type": "torchaudio"}
torch.manual_seed(0)
model = torchaudio.models.wav2vec2_model(**bundle._params).eval().to("cuda:0")
rng = np.random.default_rng(2)
audio = (torch.randn(3600 * 16000, generator=torch.Generator().manual_seed(1)) * 0.1).numpy()
letters = list("abcdefghijklmnopqrstuvwxyz'")
segs = []
for i in range(120):
    words = []
    while sum(len(w) + 1 for w in words) < 400:
        words.append("".join(rng.choice(letters, int(rng.integers(1, 9)))))
    t = " ".join(words)
    segs.append({"start": 30.0 * i, "end": 30.0 * (i + 1), "text": t[:200] + ". " + t[200:].capitalize()})
wxa.set_sentence_splitter(wxa.assemble._fallback_spans)
for rep in range(3):
    wxa.align(copy.deepcopy(segs), model, meta, audio, "cuda:0", return_char_alignments=True)
