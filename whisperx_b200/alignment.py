"""Drop-in for the reference's alignment stage: `load_align_model()` and `align()` with the
reference's signatures and result dict (whisperx/alignment.py:64-97, :100-354), plus the
functional helpers `get_trellis` / `backtrack` / `merge_repeats` / `merge_words` (:359-470).

What changes is where the work happens.  The reference walks the segments one by one, copies
every emission matrix to the host and runs the trellis, the backtrack and the pandas tables in
Python.  Here the wav2vec2 forward (PyTorch, unchanged) leaves its logits on the GPU, all
segments of the call are concatenated into one [sum_T, V] batch, and one pass of the sm_100a
kernels behind the C ABI (include/wxb_align.h) does log-softmax, trellis, backtrack, repeat
merging, timestamp rounding and the per-word reduction for every segment at once.  Only the
per-character / per-word result arrays come back to the host, where `assemble.py` builds the
records.  There is no CPU fallback: `device` must be a CUDA device.
"""
from __future__ import annotations

import os
import subprocess
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import assemble
from .assemble import LANGUAGES_WITHOUT_SPACES, PUNKT_ABBREVIATIONS, set_sentence_splitter  # noqa: F401
from .engine import BatchSpec, _alloc_outputs, get_engine

SAMPLE_RATE = 16000                       # whisperx/audio.py:13

DEFAULT_ALIGN_MODELS_TORCH = {            # alignment.py:24-30
    "en": "WAV2VEC2_ASR_BASE_960H",
    "fr": "VOXPOPULI_ASR_BASE_10K_FR",
    "de": "VOXPOPULI_ASR_BASE_10K_DE",
    "es": "VOXPOPULI_ASR_BASE_10K_ES",
    "it": "VOXPOPULI_ASR_BASE_10K_IT",
}
_XLSR = "jonatasgrosman/wav2vec2-large-xlsr-53-"
DEFAULT_ALIGN_MODELS_HF = {               # alignment.py:32-61
    **{code: _XLSR + name for code, name in (
        ("ja", "japanese"), ("zh", "chinese-zh-cn"), ("nl", "dutch"), ("pt", "portuguese"), ("ar", "arabic"),
        ("ru", "russian"), ("pl", "polish"), ("hu", "hungarian"), ("fi", "finnish"), ("fa", "persian"),
        ("el", "greek"))},
    "uk": "Yehor/wav2vec2-xls-r-300m-uk-with-small-lm",
    "cs": "comodoro/wav2vec2-xls-r-300m-cs-250",
    "tr": "mpoyraz/wav2vec2-xls-r-300m-cv7-turkish",
    "da": "saattrupdan/wav2vec2-xls-r-300m-ftspeech",
    "he": "imvladikon/wav2vec2-xls-r-300m-hebrew",
    "vi": "nguyenvulebinh/wav2vec2-base-vi",
    "ko": "kresnik/wav2vec2-large-xlsr-korean",
    "ur": "kingabzpro/wav2vec2-large-xls-r-300m-Urdu",
    "te": "anuragshas/wav2vec2-large-xlsr-53-telugu",
    "hi": "theainerd/Wav2Vec2-large-xlsr-hindi",
    "ca": "softcatala/wav2vec2-large-xlsr-catala",
    "ml": "gvs/wav2vec2-large-xlsr-malayalam",
    "no": "NbAiLab/nb-wav2vec2-1b-bokmaal",
    "nn": "NbAiLab/nb-wav2vec2-300m-nynorsk",
    "sk": "comodoro/wav2vec2-xls-r-300m-sk-cv8",
    "sl": "anton-l/wav2vec2-large-xlsr-53-slovenian",
    "hr": "classla/wav2vec2-xls-r-parlaspeech-hr",
}

# How emissions reach the kernels:
#   "torch"  torch.log_softmax on the device first — the very op the reference calls at alignment.py:233 on
#            the model's device — then the kernels take log-probs.  Default: with device="cuda" the reference
#            feeds its DP exactly these floats, so paths, frames and timestamps are the reference's bit for bit.
#   "fused"  raw logits go in and log_softmax(dim=-1) happens while a tile is staged (the batch C ABI's
#            WXB_F_LOGITS; saves one pass over the emissions).  The engine's log-probs differ from torch's by
#            up to ~2e-5, which can flip a near-tie of the float32 trellis: measured divergence from the
#            reference in profiles/r2_divergence.md.  Bit-exact *given identical log-probs*.
#   "logprobs" the model's output already is log-probabilities; nothing is applied
EMISSION_MODE = os.environ.get("WXB_EMISSION_MODE", "torch")
# The reference's point probability is `emission[t, c].exp().item()` on a CPU tensor (alignment.py:409): the host
# torch's exp, whose float32 results depend on the host's instruction set (MKL VML; profiles/r2_exp_host_dependence.md).
# True (default): align() reads the path's log-probs back (bit-exact float32), applies torch.exp on the host — the
# same call, on the same machine, as the reference — and redoes the score reductions in the reference's order
# (wxb_exact_scores_host): every `score` of the returned dict is byte-identical to the reference's.
# False: scores come from the kernels (CUDA expf: <= 1e-6 relative, rounded values may differ by 0.001).
EXACT_SCORES = os.environ.get("WXB_EXACT_SCORES", "1") not in ("", "0")
# The emission forward (alignment.py:212-232; the reference runs it segment by segment, batch 1 — its line 215
# carries the TODO).  It stays a PyTorch module; what this file decides is how segments share forwards:
#   EMISSION_BATCH   > 1: up to that many segments per forward.  Segments with the same number of samples are
#                    stacked as they are — no padding, so normalisation statistics are untouched and the only
#                    difference from batch 1 is the GEMM / convolution algorithm cuDNN and cuBLAS pick for the
#                    larger batch (logits differ in the last bits: measured in profiles/r2_emission_batching.json).
#                    1 (default) keeps the reference's batch-1 forwards: bit-identical emissions.
#   EMISSION_PAD     opt-in, needs EMISSION_BATCH > 1: segments whose lengths are within EMISSION_PAD_SLACK of the
#                    longest one of a bucket are zero-padded to it and go through the model's own `lengths` /
#                    `attention_mask` path.  Only for models whose feature extractor normalises per frame
#                    ("layer_norm": large / XLSR checkpoints): padded samples then never reach a valid frame (the
#                    convolutions are unpadded, the encoder zeroes padded frames and masks them in attention).
#                    Refused for "group_norm" extractors (WAV2VEC2_ASR_BASE_960H): their first layer normalises over
#                    time, so padding would change every frame — those segments keep their own forward.
#   EMISSION_PRECISION "fp32" (default: what the reference computes), or opt-in "tf32" / "bf16" (tensor cores;
#                    emissions then differ from the reference's at the 1e-3 level: outside the parity contract).
EMISSION_BATCH = int(os.environ.get("WXB_EMISSION_BATCH", "1"))
EMISSION_PAD = os.environ.get("WXB_EMISSION_PAD", "0") not in ("", "0")
EMISSION_PAD_SLACK = float(os.environ.get("WXB_EMISSION_PAD_SLACK", "0.1"))
EMISSION_PRECISION = os.environ.get("WXB_EMISSION_PRECISION", "fp32")
# EXTENSION, off by default (the reference fork drops characters its model does not know,
# alignment.py:159-161): keep them as wildcard tokens scored as the best non-blank symbol of each
# frame, so that digits, symbols and foreign letters get timestamps too.
WILDCARD = os.environ.get("WXB_WILDCARD", "0") not in ("", "0")
# Introspection for the parity tests: when True, align() keeps the last batch it ran (spec, device
# results and, in fused mode, the log-probs the kernels computed) in LAST_BATCH.
TIMING = os.environ.get("WXB_ALIGN_TIMING", "0") not in ("", "0")     # per-stage wall times of align() on stderr
KEEP_LAST_BATCH = False
LAST_BATCH: dict = {}


def load_audio(file: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    """Decode `file` to mono float32 at `sr` Hz through the ffmpeg CLI (whisperx/audio.py:25-65)."""
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1",
           "-acodec", "pcm_s16le", "-ar", str(sr), "-"]
    try:
        raw = subprocess.run(cmd, capture_output=True, check=True).stdout
    except subprocess.CalledProcessError as e:
        raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e
    return np.frombuffer(raw, np.int16).flatten().astype(np.float32) / 32768.0


def load_align_model(language_code, device, model_name=None, model_dir=None):
    """alignment.py:64-97: pick (or take) a wav2vec2 CTC model, load it on `device`, and return
    `(model, {"language", "dictionary", "type"})`.  The emission model stays a PyTorch module."""
    import torchaudio
    if model_name is None:
        model_name = DEFAULT_ALIGN_MODELS_TORCH.get(language_code) or DEFAULT_ALIGN_MODELS_HF.get(language_code)
        if model_name is None:
            print(f"There is no default alignment model set for this language ({language_code}).\
                Please find a wav2vec2.0 model finetuned on this language in https://huggingface.co/models, then pass the model name in --align_model [MODEL_NAME]")
            raise ValueError(f"No default align-model for language: {language_code}")
    if model_name in torchaudio.pipelines.__all__:
        kind = "torchaudio"
        bundle = getattr(torchaudio.pipelines, model_name)
        model = bundle.get_model(dl_kwargs={"model_dir": model_dir}).to(device)
        dictionary = {c.lower(): i for i, c in enumerate(bundle.get_labels())}
    else:
        from transformers import Wav2Vec2ForCTC, Wav2Vec2Processor
        try:
            processor = Wav2Vec2Processor.from_pretrained(model_name)
            model = Wav2Vec2ForCTC.from_pretrained(model_name)
        except Exception as e:
            print(e)
            print("Error loading model from huggingface, check https://huggingface.co/models for finetuned wav2vec2.0 models")
            raise ValueError(f'The chosen align_model "{model_name}" could not be found in huggingface (https://huggingface.co/models) or torchaudio (https://pytorch.org/audio/stable/pipelines.html#id14)')
        kind = "huggingface"
        model = model.to(device)
        dictionary = {ch.lower(): code for ch, code in processor.tokenizer.get_vocab().items()}
    return model, {"language": language_code, "dictionary": dictionary, "type": kind}


# --------------------------------------------------------------------------- emissions
def _forward(model, model_type: str, wav: torch.Tensor, lengths):
    if model_type == "torchaudio":
        out, _ = model(wav, lengths=lengths)
    elif model_type == "huggingface":
        out = model(wav).logits
    else:
        raise NotImplementedError(f"Align model of type {model_type} not supported.")
    return out


def frames_for_samples(n_samples: int) -> int:
    """wav2vec2 frame count: conv strides 5 * 2^6 = 320, receptive field 400."""
    return (max(n_samples, 400) - 400) // 320 + 1


def pads_safely(model, model_type: str) -> bool:
    """True when zero-padded samples cannot reach a valid frame: the feature extractor normalises per frame
    (layer norm), not over time (group norm)."""
    if model_type == "torchaudio":
        try:
            first = model.feature_extractor.conv_layers[0].layer_norm
        except AttributeError:
            return False
        return first is not None and not isinstance(first, torch.nn.GroupNorm)
    if model_type == "huggingface":
        return getattr(getattr(model, "config", None), "feat_extract_norm", "group") == "layer"
    return False


class _precision:
    """Context for the emission forward: fp32 (reference), tf32 or bf16 (opt-in, EMISSION_PRECISION)."""

    def __enter__(self):
        self.old = torch.backends.cuda.matmul.allow_tf32
        self.cast = None
        if EMISSION_PRECISION == "tf32":
            torch.backends.cuda.matmul.allow_tf32 = True
        elif EMISSION_PRECISION == "bf16":
            self.cast = torch.autocast("cuda", dtype=torch.bfloat16)
            self.cast.__enter__()
        elif EMISSION_PRECISION != "fp32":
            raise ValueError(f"WXB_EMISSION_PRECISION={EMISSION_PRECISION!r}: expected fp32, tf32 or bf16")
        return self

    def __exit__(self, *exc):
        if self.cast is not None:
            self.cast.__exit__(*exc)
        torch.backends.cuda.matmul.allow_tf32 = self.old
        return False


def _forward_padded(model, model_type: str, wavs: list, device):
    """One forward over segments of different lengths through the model's own padding path.
    -> list of [T_i, V] float32 tensors (valid frames only)."""
    n = [int(w.shape[-1]) for w in wavs]
    batch = torch.zeros((len(wavs), max(n)), dtype=wavs[0].dtype)
    for i, w in enumerate(wavs):
        batch[i, :n[i]] = w
    batch = batch.to(device)
    if model_type == "torchaudio":
        out, _ = model(batch, lengths=torch.as_tensor(n, device=device))
    else:
        mask = (torch.arange(max(n))[None, :] < torch.as_tensor(n)[:, None]).to(device=device, dtype=torch.long)
        out = model(batch, attention_mask=mask).logits
    return [out[i, :frames_for_samples(n[i])].float() for i in range(len(wavs))]


def _emit(model, model_type: str, audio: torch.Tensor, jobs: list, device) -> list:
    """wav2vec2 forward for every runnable segment (alignment.py:212-232); logits stay on `device`.
    jobs: [(f1, f2)] sample ranges.  Returns one [T, V] float32 tensor per job."""
    outs: List[Optional[torch.Tensor]] = [None] * len(jobs)
    with torch.inference_mode(), _precision():
        done = set()
        if EMISSION_BATCH > 1 and audio.shape[0] == 1:
            length = {}
            for i, (f1, f2) in enumerate(jobs):
                n = min(f2, audio.shape[1]) - f1
                if n >= 400:
                    length[i] = n
            # 1. segments with the same number of samples: stacked, no padding (any model)
            groups: dict = {}
            for i, n in length.items():
                groups.setdefault(n, []).append(i)
            for n, members in groups.items():
                for k in range(0, len(members), EMISSION_BATCH):
                    part = members[k:k + EMISSION_BATCH]
                    if len(part) < 2:
                        continue
                    wav = torch.stack([audio[0, jobs[i][0]:jobs[i][0] + n] for i in part]).to(device)
                    logits = _forward(model, model_type, wav, None)
                    for row, i in enumerate(part):
                        outs[i] = logits[row].float()
                        done.add(i)
            # 2. opt-in: length buckets, zero-padded through the model's lengths / attention-mask path
            if EMISSION_PAD and pads_safely(model, model_type):
                rest = sorted((i for i in length if i not in done), key=lambda i: -length[i])
                k = 0
                while k < len(rest):
                    part = [rest[k]]
                    floor = length[rest[k]] * (1.0 - EMISSION_PAD_SLACK)
                    while k + len(part) < len(rest) and len(part) < EMISSION_BATCH and length[rest[k + len(part)]] >= floor:
                        part.append(rest[k + len(part)])
                    k += len(part)
                    if len(part) < 2:
                        continue
                    res = _forward_padded(model, model_type, [audio[0, jobs[i][0]:jobs[i][0] + length[i]] for i in part], device)
                    for i, r in zip(part, res):
                        outs[i] = r
                        done.add(i)
        for i in range(len(jobs)):
            if i in done:
                continue
            f1, f2 = jobs[i]
            wav = audio[:, f1:f2]
            lengths = None
            if wav.shape[-1] < 400:                       # minimum wav2vec2 input (alignment.py:217-221)
                lengths = torch.as_tensor([wav.shape[-1]]).to(device)
                wav = torch.nn.functional.pad(wav, (0, 400 - wav.shape[-1]))
            outs[i] = _forward(model, model_type, wav.to(device), lengths)[0].float()
    return outs


# --------------------------------------------------------------------------- align()
def align(
    transcript: Iterable[dict],
    model: torch.nn.Module,
    align_model_metadata: dict,
    audio: Union[str, np.ndarray, torch.Tensor],
    device: str,
    interpolate_method: str = "nearest",
    return_char_alignments: bool = False,
    print_progress: bool = False,
    combined_progress: bool = False,
) -> dict:
    """Align phoneme recognition predictions to known transcription (alignment.py:100-354).

    Same arguments, side effects (the four keys added to every input segment) and result as the
    reference; the per-segment "Failed to align segment" messages are printed for the same
    three soft failures (:199-207, :245-248)."""
    grouped = align_grouped(transcript, model, align_model_metadata, audio, device, interpolate_method,
                            return_char_alignments, print_progress, combined_progress)
    aligned: List[dict] = [rec for group in grouped for rec in group]
    word_segments: List[dict] = []
    for seg in aligned:
        word_segments += seg["words"]
    return {"segments": aligned, "word_segments": word_segments}


def align_grouped(transcript, model, align_model_metadata, audio, device, interpolate_method="nearest",
                  return_char_alignments=False, print_progress=False, combined_progress=False) -> List[List[dict]]:
    """align() with the output records still grouped by input segment (one list per input segment;
    the reference concatenates them, :347).  Used by align() and by the multi-GPU front end."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("whisperx_b200.align() runs the alignment path on a CUDA (sm_100a) device only; "
                           f"got device={device!r}. There is no CPU fallback.")
    engine = get_engine(dev)
    dev = engine.device

    if not torch.is_tensor(audio):
        if isinstance(audio, str):
            audio = load_audio(audio)
        audio = torch.from_numpy(audio)
    if len(audio.shape) == 1:
        audio = audio.unsqueeze(0)
    max_duration = audio.shape[1] / SAMPLE_RATE

    dictionary = align_model_metadata["dictionary"]
    lang = align_model_metadata["language"]
    model_type = align_model_metadata["type"]
    if model_type not in ("torchaudio", "huggingface"):
        raise NotImplementedError(f"Align model of type {model_type} not supported.")

    # 1. keep only characters the model knows (mutates the input, like the reference)
    import time as _time
    _t = [_time.perf_counter()]
    _mark = (lambda: (torch.cuda.synchronize(dev), _t.append(_time.perf_counter()))) if TIMING else (lambda: None)
    transcript = transcript if isinstance(transcript, list) else list(transcript)
    assemble.prepare_transcript(transcript, dictionary, lang, print_progress, combined_progress, wildcard=WILDCARD)
    _mark()

    # 2a. which segments reach the model (alignment.py:198-207)
    plain: List[dict] = []
    verdict: List[Optional[str]] = []
    run: List[int] = []
    for sdx, seg in enumerate(transcript):
        rec = {"start": seg["start"], "end": seg["end"], "text": seg["text"], "words": []}
        if return_char_alignments:
            rec["chars"] = []
        plain.append(rec)
        if len(seg["clean_char"]) == 0:
            verdict.append("no characters in this segment found in model dictionary, resorting to original...")
        elif seg["start"] >= max_duration:
            verdict.append("original start time longer than audio duration, skipping...")
        else:
            verdict.append(None)
            run.append(sdx)

    results: dict = {}
    if run:
        jobs = [(int(transcript[s]["start"] * SAMPLE_RATE), int(transcript[s]["end"] * SAMPLE_RATE)) for s in run]
        logits = _emit(model, model_type, audio, jobs, dev)
        _mark()
        results = _align_batch(engine, [transcript[s] for s in run], logits, dictionary, lang, audio.shape[0])
        _mark()

    # 3. records, in input order
    grouped: List[List[dict]] = []
    k = 0
    for sdx, seg in enumerate(transcript):
        why = verdict[sdx]
        if why is None:
            r = results["segments"][k]
            k += 1
            if r is None:
                why = "backtrack failed, resorting to original..."
            else:
                lay, cs, ce, sc, ws, we, wsc = r
                grouped.append(assemble.assemble_segment(seg["text"], seg["clean_cdx"], lay, seg["sentence_spans"], lang,
                                                         cs, ce, sc, ws, we, wsc, interpolate_method, return_char_alignments))
                continue
        print(f'Failed to align segment ("{seg["text"]}"): {why}')
        grouped.append([plain[sdx]])
    if TIMING and len(_t) == 4:
        _mark()
        import sys as _sys
        print("[wxb] align(): text prep %.1f ms, emission forward %.1f ms, alignment batch (layout + kernels + exact scores) %.1f ms, "
              "record assembly %.1f ms" % tuple((b - a) * 1e3 for a, b in zip(_t, _t[1:])), file=_sys.stderr)
    return grouped


def _align_batch(engine, segs: Sequence[dict], logits: Sequence[torch.Tensor], dictionary: dict, lang: str,
                 channels: int) -> dict:
    """One pass of the kernels over all runnable segments of the call (alignment.py:233-320)."""
    dev = engine.device
    S = len(segs)
    V = int(logits[0].shape[-1])
    blank = assemble.blank_id_of(dictionary)
    frame_off = np.zeros(S + 1, np.int64)
    frame_len = np.zeros(S, np.int32)
    tok_off = np.zeros(S + 1, np.int64)
    t1 = np.zeros(S, np.float64)
    ratio = np.zeros(S, np.float64)
    round_np = np.zeros(S, np.uint8)
    word_off = np.zeros(S + 1, np.int64)
    tokens: List[int] = []
    wchar_off: List[int] = [0]
    wchar_idx: List[int] = []
    layouts = []
    for s, seg in enumerate(segs):
        T = int(logits[s].shape[0])
        frame_len[s] = T
        frame_off[s + 1] = frame_off[s] + T + (-T) % 4        # 4-row granules keep every segment 16-byte aligned
        tokens += [dictionary.get(c, -1) for c in seg["clean_char"]]        # -1: wildcard entry (extension)
        tok_off[s + 1] = len(tokens)
        a, b = seg["start"], seg["end"]
        t1[s] = a
        ratio[s] = (b - a) * channels / T                      # alignment.py:252-253
        round_np[s] = isinstance(a, np.floating) or isinstance(b, np.floating)
        lay = assemble.layout_segment(seg["text"], seg["clean_cdx"], seg["sentence_spans"], lang)
        layouts.append(lay)
        base = len(wchar_idx)
        wchar_idx += lay.group_idx
        wchar_off += [base + o for o in lay.group_off[1:]]
        word_off[s + 1] = len(wchar_off) - 1

    emis = torch.zeros((int(frame_off[-1]), V), dtype=torch.float32, device=dev)
    for s in range(S):
        emis[int(frame_off[s]):int(frame_off[s]) + int(frame_len[s])] = logits[s]
    if EMISSION_MODE not in ("fused", "torch", "logprobs"):
        raise ValueError(f"WXB_EMISSION_MODE={EMISSION_MODE!r}: expected fused, torch or logprobs")
    fused = EMISSION_MODE == "fused"
    if EMISSION_MODE == "torch":
        emis = torch.log_softmax(emis, dim=-1)
    has_words = len(wchar_off) > 1
    spec = BatchSpec(frame_off=frame_off, frame_len=frame_len, tok_off=tok_off, tokens=np.asarray(tokens, np.int32),
                     vocab=V, blank_id=blank, logits=fused, wildcard=(-1 in tokens), seg_t1=t1, seg_ratio=ratio,
                     seg_round_np=round_np, clobber=(fused and not KEEP_LAST_BATCH),      # emis is our own concatenation
                     word_off=word_off if has_words else None,
                     wchar_off=np.asarray(wchar_off, np.int64) if has_words else None,
                     wchar_idx=np.asarray(wchar_idx, np.int32) if has_words else None)
    want = (("logprobs",) if (KEEP_LAST_BATCH and fused) else ()) + (("path",) if EXACT_SCORES else ())
    res = engine.align_batch(emis, spec, want=want)
    if KEEP_LAST_BATCH:
        LAST_BATCH.clear()
        LAST_BATCH.update(spec=spec, result=res, emissions=emis)
    status = res.status.cpu().tolist()
    cs = res.char_start_s.cpu().tolist()
    ce = res.char_end_s.cpu().tolist()
    sc = res.char_score_r.cpu().tolist()
    if has_words:
        ws, we, wsc = res.word_start_s.cpu().tolist(), res.word_end_s.cpu().tolist(), res.word_score.cpu().tolist()
    else:
        ws = we = wsc = []
    if EXACT_SCORES:
        sc, exact_w = exact_scores(engine, spec, res)
        wsc = exact_w if has_words else wsc
    out = []
    for s in range(S):
        if status[s] != 0:
            out.append(None)
            continue
        a, b = int(tok_off[s]), int(tok_off[s + 1])
        g0, g1 = int(word_off[s]), int(word_off[s + 1])
        out.append((layouts[s], cs[a:b], ce[a:b], sc[a:b], ws[g0:g1], we[g0:g1], wsc[g0:g1]))
    return {"segments": out, "spec": spec, "result": res}


def exact_scores(engine, spec: BatchSpec, res):
    """Scores of a batch with the host's own exp (see EXACT_SCORES).  -> (char_score_r list, word_score list)."""
    import ctypes as C
    from . import _native as N
    prob = torch.exp(res.path_logp.cpu()).contiguous()              # alignment.py:409, the reference's own op
    status = res.status.cpu().contiguous()
    c0, c1 = res.char_start.cpu().contiguous(), res.char_end.cpu().contiguous()
    n_chars, n_words = max(spec.sum_N, 1), max(spec.n_words, 1)
    score = np.empty(n_chars, np.float64)
    score_r = np.empty(n_chars, np.float64)
    wscore = np.empty(n_words, np.float64)
    cb = spec.to_c()
    N.check(engine._lib.wxb_exact_scores_host(C.byref(cb), prob.data_ptr(), status.data_ptr(), c0.data_ptr(), c1.data_ptr(),
                                              score.ctypes.data, score_r.ctypes.data,
                                              wscore.ctypes.data if spec.n_words > 0 else None), "wxb_exact_scores_host")
    return score_r.tolist(), wscore.tolist()


# --------------------------------------------------------------------------- functional API
@dataclass
class Point:                               # alignment.py:381-385
    token_index: int
    time_index: int
    score: float


@dataclass
class Segment:                             # alignment.py:424-436
    label: str
    start: int
    end: int
    score: float

    def __repr__(self):
        return f"{self.label}\t({self.score:4.2f}): [{self.start:5d}, {self.end:5d})"

    @property
    def length(self):
        return self.end - self.start


def _single_spec(emission: torch.Tensor, tokens: Sequence[int], blank_id: int, **kw) -> BatchSpec:
    T, V = emission.shape
    return BatchSpec(frame_off=[0, T], tok_off=[0, len(tokens)], tokens=np.asarray(tokens, np.int32), vocab=V,
                     blank_id=blank_id, **kw)


def get_trellis(emission: torch.Tensor, tokens: Sequence[int], blank_id: int = 0, device="cuda:0") -> torch.Tensor:
    """alignment.py:359-379 — the full (T+1, N+1) float32 table, computed by the kernel's
    introspection path (the batch path never materialises it).  Returned on the CPU like the
    reference's."""
    eng = get_engine(device)
    E = emission.to(eng.device, torch.float32).contiguous()
    spec = _single_spec(E, tokens, blank_id, no_shortcut=True)
    res = eng.align_batch(E, spec, want=("trellis",))
    return res.trellis[: (E.shape[0] + 1) * (len(tokens) + 1)].view(E.shape[0] + 1, len(tokens) + 1).cpu()


def backtrack(trellis: torch.Tensor, emission: torch.Tensor, tokens: Sequence[int], blank_id: int = 0,
              device="cuda:0") -> Optional[List[Point]]:
    """alignment.py:387-421 from a caller-supplied trellis; None when the walk never reaches token 0."""
    eng = get_engine(device)
    status, _, ptok, plogp = eng.backtrack_trellis(trellis, emission, tokens, blank_id)
    if status != 0:
        return None
    prob = torch.exp(plogp).double().tolist()
    tok = ptok.tolist()
    return [Point(tok[t], t, prob[t]) for t in range(len(tok)) if tok[t] >= 0]


def merge_repeats(path: Sequence[Point], transcript: str) -> List[Segment]:
    """alignment.py:438-454: one Segment per run of equal token_index, score = mean point score."""
    out: List[Segment] = []
    i = 0
    while i < len(path):
        j = i
        total = 0
        while j < len(path) and path[j].token_index == path[i].token_index:
            total = total + path[j].score
            j += 1
        out.append(Segment(transcript[path[i].token_index], path[i].time_index, path[j - 1].time_index + 1,
                           total / (j - i)))
        i = j
    return out


def merge_words(segments: Sequence[Segment], separator: str = "|") -> List[Segment]:
    """alignment.py:456-470 (unused by align(); kept for API parity): words are the runs between
    separator labels, score = length-weighted mean of the character scores."""
    words: List[Segment] = []
    run: List[Segment] = []
    for seg in list(segments) + [None]:
        if seg is not None and seg.label != separator:
            run.append(seg)
            continue
        if run:
            weight = sum(s.length for s in run)
            words.append(Segment("".join(s.label for s in run), run[0].start, run[-1].end,
                                 sum(s.score * s.length for s in run) / weight))
            run = []
    return words
