"""GPU: the align() drop-in and the functional API against the reference's outputs.

  * golden scenarios (tests/golden/align_cases.*, produced by the real reference): same
    transcript, same log-probs (scripted model) -> same records;
  * a random-init WAV2VEC2_ASR_BASE_960H end to end against the oracle's align() port run on the
    same device, in both emission modes;
  * get_trellis / backtrack / merge_repeats / merge_words against the golden DP cases.
Everything must be identical, scores included: align() runs in its exact mode (EXACT_SCORES: the path's
log-probs come back bit-exact and the host's own torch.exp — the reference's call at alignment.py:409 —
turns them into probabilities), so comparisons with the oracle executed on this machine use score_tol=0.
Against the committed fixtures the tolerance is 0 when this machine's torch.exp is the one that generated
them (tests/helpers.host_exp_matches_fixtures), else one step of the rounding (MKL's exp is host dependent).
With EXACT_SCORES off the kernels' own scores are within 1e-6 relative (rounded: one step of 0.001)."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_port
from whisperx_b200.types import check_aligned_result
from tests.helpers import DICTIONARY, LABELS, ScriptedModel, host_exp_matches_fixtures, regex_spans, same_records

pytestmark = pytest.mark.gpu


@pytest.fixture()
def wxa(engine):
    from whisperx_b200 import alignment
    alignment.set_sentence_splitter(regex_spans)
    old = alignment.EMISSION_MODE
    yield alignment
    alignment.EMISSION_MODE = old
    alignment.KEEP_LAST_BATCH = False
    alignment.set_sentence_splitter(None)


def test_golden_scenarios(wxa, golden_align, capsys):
    meta, arr = golden_align
    wxa.EMISSION_MODE = "logprobs"
    assert wxa.EXACT_SCORES
    fixture_tol = 0.0 if host_exp_matches_fixtures() else 0.0010001
    for name, m in meta.items():
        model = ScriptedModel([arr[f"{name}_E{i}"] for i in range(m["n_emissions"])], "cuda:0")
        transcript = copy.deepcopy(m["segments"])
        if m.get("numpy_times"):
            for t in transcript:
                t["start"], t["end"] = np.float64(t["start"]), np.float64(t["end"])
        audio = np.zeros(m["audio_samples"], np.float32)
        tr_live = copy.deepcopy(transcript)
        res = wxa.align(transcript, model, {"language": m["lang"], "dictionary": DICTIONARY, "type": "torchaudio"},
                        audio, "cuda:0", interpolate_method=m["interpolate_method"],
                        return_char_alignments=m["return_char_alignments"])
        assert model.calls == m["n_emissions"]
        d = same_records(res["segments"], m["result"]["segments"], score_tol=fixture_tol)
        assert d is None, (name, d)
        assert same_records(res["word_segments"], m["result"]["word_segments"], score_tol=fixture_tol) is None
        check_aligned_result(res, char_alignments=m["return_char_alignments"])
        # the oracle executed on this machine (its torch.exp is the one align() used): byte-identical records
        served = iter(range(m["n_emissions"]))           # the fixtures hold log-probs: hand them to the DP untouched
        live = ref_port.align_port(tr_live, ScriptedModel([arr[f"{name}_E{i}"] for i in range(m["n_emissions"])]),
                                   {"language": m["lang"], "dictionary": DICTIONARY, "type": "torchaudio"}, audio, "cpu",
                                   interpolate_method=m["interpolate_method"], return_char_alignments=m["return_char_alignments"],
                                   span_fn=regex_spans,
                                   emission_hook=lambda _k, _e: torch.from_numpy(arr[f"{name}_E{next(served)}"]))
        d = same_records(res["segments"], live["segments"], score_tol=0.0)
        assert d is None, (name, "live oracle", d)
        for t, exp in zip(transcript, m["mutated_inputs"]):
            assert t["clean_char"] == exp["clean_char"] and t["clean_cdx"] == exp["clean_cdx"]
    out = capsys.readouterr().out
    for why in ("no characters in this segment found in model dictionary", "original start time longer than audio duration",
                "backtrack failed"):
        assert why in out


def _random_model():
    import torchaudio
    torch.manual_seed(0)
    params = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params
    return torchaudio.models.wav2vec2_model(**params).eval().to("cuda:0")


TRANSCRIPT = [
    {"start": 0.0, "end": 4.0, "text": " It was the best of times. It was the worst of times!"},
    {"start": 4.0, "end": 9.5, "text": "We had 2 tickets, and nothing before us... Really?"},
    {"start": 9.5, "end": 9.6, "text": "much too long for a tenth of a second"},
    {"start": np.float64(9.6), "end": np.float64(12.0), "text": "numpy typed times"},
    {"start": 11.0, "end": 15.0, "text": "runs past the end of the audio"},
]


@pytest.mark.parametrize("mode", ["torch", "fused"])
def test_random_model_end_to_end(wxa, mode):
    model = _random_model()
    meta = {"language": "en", "dictionary": DICTIONARY, "type": "torchaudio"}
    g = torch.Generator().manual_seed(5)
    audio = (torch.randn(12 * 16000, generator=g) * 0.1).numpy()
    wxa.EMISSION_MODE = mode
    wxa.KEEP_LAST_BATCH = True
    ours_in = copy.deepcopy(TRANSCRIPT)
    ours = wxa.align(ours_in, model, meta, audio, "cuda:0", return_char_alignments=True)
    last = wxa.LAST_BATCH
    spec = last["spec"]
    lp = last["result"].logprobs if mode == "fused" else last["emissions"]
    lp = lp.view(-1, len(LABELS)).cpu()

    def engine_logprobs(k, emission, _n=[0]):
        # the k-th segment that reached the model, in order
        i = _n[0]; _n[0] += 1
        a, T = int(spec.frame_off[i]), int(spec.frame_len[i])
        mine = lp[a:a + T]
        assert mine.shape == emission.shape
        if int(spec.tok_off[i + 1] - spec.tok_off[i]) > T:
            return None       # N > T is failed by the planner without running the kernels: no log-probs exist
        # same model, same device: torch mode is bit-identical, fused mode within a few ulp
        torch.testing.assert_close(mine, emission, rtol=0, atol=0 if mode == "torch" else 2e-5)
        return mine

    ref_in = copy.deepcopy(TRANSCRIPT)
    ref = ref_port.align_port(ref_in, model, meta, audio, "cuda:0", return_char_alignments=True,
                              span_fn=regex_spans, emission_hook=engine_logprobs)
    d = same_records(ours["segments"], ref["segments"], score_tol=0.0)
    assert d is None, d
    assert same_records(ours["word_segments"], ref["word_segments"], score_tol=0.0) is None
    assert len(ours["segments"]) >= 5 and any(s["words"] == [] for s in ours["segments"])
    # the kernels' own scores (EXACT_SCORES off): one step of the 3-decimal rounding at most
    wxa.EXACT_SCORES = False
    try:
        dev_scores = wxa.align(copy.deepcopy(TRANSCRIPT), model, meta, audio, "cuda:0", return_char_alignments=True)
    finally:
        wxa.EXACT_SCORES = True
    assert same_records(dev_scores["segments"], ref["segments"], score_tol=0.0010001) is None
    for a, b in zip(ours_in, ref_in):
        assert a["clean_char"] == b["clean_char"] and a["sentence_spans"] == [tuple(x) for x in b["sentence_spans"]]
    # unrounded scores of the batch against the oracle DP on the engine's log-probs
    res = last["result"].cpu()
    for i in range(spec.n_segments):
        a, T = int(spec.frame_off[i]), int(spec.frame_len[i])
        sl = slice(int(spec.tok_off[i]), int(spec.tok_off[i + 1]))
        segs, _ = ref_port.align_tokens(lp[a:a + T], spec.tokens[sl].tolist(), 0, "x" * (sl.stop - sl.start))
        if segs is None:
            assert int(res.status[i]) == 1
            continue
        assert res.char_start[sl].tolist() == [s.start for s in segs] and res.char_end[sl].tolist() == [s.end for s in segs]
        np.testing.assert_allclose(res.char_score[sl].numpy(), [s.score for s in segs], rtol=1e-6, atol=0)


def test_batched_emission_forward(wxa):
    """EMISSION_BATCH > 1: segments of equal length are stacked (no padding), and — opt-in, layer-norm extractors
    only — length buckets go zero-padded through the model's `lengths` path.  What batching can change is the
    logits, through the library's choice of convolution / GEMM algorithm for the larger batch: every valid frame
    must agree with the batch-1 forward to float32 round-off of a TF32 convolution stack (1e-3; the alignment of
    identical log-probs is bit-exact, so nothing else can differ), and the records keep their structure.  The
    default stays batch 1: bit-identical emissions."""
    import torchaudio
    meta = {"language": "en", "dictionary": DICTIONARY, "type": "torchaudio"}
    g = torch.Generator().manual_seed(6)
    audio = (torch.randn(20 * 16000, generator=g) * 0.1).numpy()
    tr = [{"start": 4.0 * i, "end": 4.0 * i + 4.0, "text": f"segment number {i} of four"} for i in range(4)]
    tr += [{"start": 16.0, "end": 19.7, "text": "a shorter one"}, {"start": 16.1, "end": 19.9, "text": "and another"}]
    jobs = [(int(t["start"] * 16000), int(t["end"] * 16000)) for t in tr]
    torch.manual_seed(1)
    params = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params
    layer = torchaudio.models.wav2vec2_model(**dict(params, extractor_mode="layer_norm", extractor_conv_bias=True,
                                                    encoder_layer_norm_first=True)).eval().to("cuda:0")
    assert wxa.EMISSION_BATCH == 1 and not wxa.EMISSION_PAD
    for model, pad in ((_random_model(), False), (layer, True)):
        assert wxa.pads_safely(model, "torchaudio") == pad
        wav = torch.from_numpy(audio).unsqueeze(0)
        lg1 = wxa._emit(model, "torchaudio", wav, jobs, torch.device("cuda:0"))
        one = wxa.align(copy.deepcopy(tr), model, meta, audio, "cuda:0")
        wxa.EMISSION_BATCH, wxa.EMISSION_PAD = 4, True
        try:
            lg4 = wxa._emit(model, "torchaudio", wav, jobs, torch.device("cuda:0"))
            four = wxa.align(copy.deepcopy(tr), model, meta, audio, "cuda:0")
        finally:
            wxa.EMISSION_BATCH, wxa.EMISSION_PAD = 1, False
        for a, b in zip(lg1, lg4):
            assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-3
        assert len(one["segments"]) == len(four["segments"]) == 6
        for a, b in zip(one["word_segments"], four["word_segments"]):
            assert a["word"] == b["word"] and ("start" in a) == ("start" in b)


def test_cpu_device_is_refused(wxa):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wxa.align([], None, {"language": "en", "dictionary": DICTIONARY, "type": "torchaudio"}, np.zeros(16000, np.float32), "cpu")


def test_functional_api_golden(wxa, golden_dp):
    meta, arr = golden_dp
    for m in meta:
        a = {k[len(m["name"]) + 1:]: arr[k] for k in arr.files if k.startswith(m["name"] + "_")}
        if m["N"] > 2048:
            continue
        E = torch.from_numpy(a["E"])
        R = wxa.get_trellis(E, a["tok"].tolist(), m["blank"])
        assert R.shape == a["R"].shape and np.array_equal(R.numpy().view(np.int32), a["R"].view(np.int32)), m
        path = wxa.backtrack(torch.from_numpy(a["R"]), E, a["tok"].tolist(), m["blank"])
        if not m["ok"]:
            assert path is None
            continue
        assert [p.token_index for p in path] == a["ptok"].tolist() and [p.time_index for p in path] == a["ptime"].tolist()
        np.testing.assert_allclose([p.score for p in path], a["pscore"], rtol=1e-6, atol=0)
        # merge on the reference's own points: exact
        pts = [wxa.Point(int(t), int(i), float(s)) for t, i, s in zip(a["ptok"], a["ptime"], a["pscore"])]
        segs = wxa.merge_repeats(pts, m["transcript"])
        assert [s.start for s in segs] == a["sstart"].tolist() and [s.end for s in segs] == a["send"].tolist()
        assert [s.score for s in segs] == a["sscore"].tolist()
        words = wxa.merge_words(segs)
        assert [w.label for w in words] == m["words"] and [w.score for w in words] == a["wscore"].tolist()
        assert [w.start for w in words] == a["wstart"].tolist() and [w.end for w in words] == a["wend"].tolist()


def test_wildcard_extension_end_to_end(wxa):
    """EXTENSION (opt-in, not in the reference fork): with WILDCARD on, characters the model does not know
    are aligned as wildcard tokens; compared with the oracle port's restatement of the same rule."""
    rng = np.random.default_rng(33)
    transcript = [
        {"start": 0.0, "end": 4.0, "text": "It costs 2014 dollars & 50 cents."},
        {"start": 4.0, "end": 6.5, "text": " 99 "},
        {"start": 6.5, "end": 9.0, "text": "naïve café, 3 o'clock!"},
    ]
    audio = np.zeros(10 * 16000, np.float32)
    logps = []
    for seg in transcript:
        T = (int(seg["end"] * 16000) - int(seg["start"] * 16000) - 400) // 320 + 1
        logps.append(torch.log_softmax(torch.from_numpy(rng.standard_normal((T, len(LABELS))).astype(np.float32) * 3), -1))
    meta = {"language": "en", "dictionary": DICTIONARY, "type": "torchaudio"}
    wxa.EMISSION_MODE = "logprobs"
    wxa.WILDCARD = True
    try:
        ours_in = copy.deepcopy(transcript)
        ours = wxa.align(ours_in, ScriptedModel(logps, "cuda:0"), meta, audio, "cuda:0", return_char_alignments=True)
    finally:
        wxa.WILDCARD = False
    ref_in = copy.deepcopy(transcript)
    ref = ref_port.align_port(ref_in, ScriptedModel(logps), meta, audio, "cpu", return_char_alignments=True,
                              span_fn=regex_spans, wildcard=True)
    d = same_records(ours["segments"], ref["segments"], score_tol=0.0)
    assert d is None, d
    assert [s["clean_char"] for s in ours_in] == [s["clean_char"] for s in ref_in]
    words = {w["word"]: w for w in ours["word_segments"]}
    assert "start" in words["2014"] and "start" in words["&"] and "start" in words["99"]     # dropped by the reference fork
    # off (default): identical to the reference behaviour — the digits get no timestamps
    plain = wxa.align(copy.deepcopy(transcript), ScriptedModel(logps, "cuda:0"), meta, audio, "cuda:0")
    assert "start" not in {w["word"]: w for w in plain["word_segments"]}["2014"]


class _HFStyle(torch.nn.Module):
    """A HuggingFace-shaped CTC model: forward(waveform) -> object with `.logits` [1, T, V] (alignment.py:229-230)."""

    def __init__(self, vocab):
        super().__init__()
        import torchaudio
        torch.manual_seed(1)
        params = dict(torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params, aux_num_out=vocab)
        self.net = torchaudio.models.wav2vec2_model(**params).eval()

    def forward(self, waveform):
        from types import SimpleNamespace
        return SimpleNamespace(logits=self.net(waveform)[0])


@pytest.mark.parametrize("mode", ["fused", "torch"])
def test_huggingface_style_model_with_a_32_symbol_head(wxa, mode):
    """Typical HF wav2vec2 CTC head: 32 symbols, '<pad>' (the blank, alignment.py:237-240) at id 0, '|' word
    separator.  An even vocabulary takes the normalise-first route inside the engine; results must be the
    reference algorithm's on the engine's own log-probs."""
    specials = ["<pad>", "<s>", "</s>", "<unk>", "|"]
    vocab = {tok: i for i, tok in enumerate(specials + list("abcdefghijklmnopqrstuvwxyz") + ["'"])}
    assert len(vocab) == 32
    model = _HFStyle(32).to("cuda:0")
    meta = {"language": "en", "dictionary": vocab, "type": "huggingface"}
    g = torch.Generator().manual_seed(9)
    audio = (torch.randn(12 * 16000, generator=g) * 0.1).numpy()
    wxa.EMISSION_MODE = mode
    wxa.KEEP_LAST_BATCH = True
    ours = wxa.align(copy.deepcopy(TRANSCRIPT), model, meta, audio, "cuda:0", return_char_alignments=True)
    last = wxa.LAST_BATCH
    spec = last["spec"]
    assert spec.vocab == 32 and spec.blank_id == 0
    lp = (last["result"].logprobs if mode == "fused" else last["emissions"]).view(-1, 32).cpu()

    def engine_logprobs(k, emission, _n=[0]):
        i = _n[0]; _n[0] += 1
        a, T = int(spec.frame_off[i]), int(spec.frame_len[i])
        if int(spec.tok_off[i + 1] - spec.tok_off[i]) > T:
            return None
        torch.testing.assert_close(lp[a:a + T], emission, rtol=0, atol=0 if mode == "torch" else 2e-5)
        return lp[a:a + T]

    ref = ref_port.align_port(copy.deepcopy(TRANSCRIPT), model, meta, audio, "cuda:0", return_char_alignments=True,
                              span_fn=regex_spans, emission_hook=engine_logprobs)
    d = same_records(ours["segments"], ref["segments"], score_tol=0.0)
    assert d is None, d
    assert same_records(ours["word_segments"], ref["word_segments"], score_tol=0.0) is None
    check_aligned_result(ours, char_alignments=True)
    wxa.KEEP_LAST_BATCH = False
    again = wxa.align(copy.deepcopy(TRANSCRIPT), model, meta, audio, "cuda:0", return_char_alignments=True)   # in-place route
    assert same_records(again["segments"], ours["segments"]) is None


@pytest.mark.needs_reference
def test_the_reference_itself_on_cuda_returns_the_same_dict(wxa):
    """The drop-in claim, end to end: the *unmodified reference* `whisperx.alignment.align()` (imported from
    /root/reference, or from its build-time copy oracle/_ref on the GPU box) and `whisperx_b200.align()` run on the same
    random-init WAV2VEC2_ASR_BASE_960H, the same audio and the same transcript with device="cuda" must return the same
    dict — text, timestamps and scores compared for equality, no tolerance (default settings: torch.log_softmax on the
    device like the reference, scores finished with the host's exp like the reference) — and leave the same keys on
    the input segments."""
    from oracle import ref_import
    A = ref_import.load_reference_alignment()
    ref_import.set_sentence_splitter(regex_spans)
    model = _random_model()
    meta = {"language": "en", "dictionary": DICTIONARY, "type": "torchaudio"}
    g = torch.Generator().manual_seed(11)
    audio = (torch.randn(40 * 16000, generator=g) * 0.1).numpy()
    transcript = TRANSCRIPT + [
        {"start": 15.0, "end": 39.5, "text": "A much longer segment follows here, with several sentences in it. Some are short. Others go on for quite "
                                             "a while before they stop! Does the question mark work? It should; and so should 3 numbers like 42 or 1999."},
        {"start": 39.5, "end": 45.0, "text": "clipped by the end of the audio"},
    ]
    assert wxa.EMISSION_MODE == "torch" and wxa.EXACT_SCORES and wxa.EMISSION_BATCH == 1
    for chars in (False, True):
        ref_in, our_in = copy.deepcopy(transcript), copy.deepcopy(transcript)
        ref = A.align(ref_in, model, meta, audio, "cuda:0", return_char_alignments=chars)
        ours = wxa.align(our_in, model, meta, audio, "cuda:0", return_char_alignments=chars)
        d = same_records(ours["segments"], ref["segments"], score_tol=0.0)
        assert d is None, d
        assert same_records(ours["word_segments"], ref["word_segments"], score_tol=0.0) is None
        assert len(ours["segments"]) >= 10 and sum("start" in w for w in ours["word_segments"]) > 60
        for a, b in zip(our_in, ref_in):
            assert a["clean_char"] == b["clean_char"] and a["clean_cdx"] == b["clean_cdx"] and a["clean_wdx"] == b["clean_wdx"]
            assert [tuple(x) for x in a["sentence_spans"]] == [tuple(x) for x in b["sentence_spans"]]


@pytest.mark.needs_reference
@pytest.mark.parametrize("variant", ["ja_linear", "huggingface_32"])
def test_the_reference_itself_on_cuda_other_paths(wxa, variant):
    """Same statement as above for the paths the English torchaudio case does not take: a language without spaces
    (every character its own word, alignment.py:22,277-278) with linear interpolation of un-alignable sentences, and a
    HuggingFace-style model (`.logits`, 32-symbol head with '<pad>' as blank, alignment.py:229-230,237-240)."""
    from oracle import ref_import
    A = ref_import.load_reference_alignment()
    ref_import.set_sentence_splitter(regex_spans)
    g = torch.Generator().manual_seed(12)
    audio = (torch.randn(20 * 16000, generator=g) * 0.1).numpy()
    if variant == "ja_linear":
        model, meta = _random_model(), {"language": "ja", "dictionary": DICTIONARY, "type": "torchaudio"}
        transcript = [{"start": 0.0, "end": 6.0, "text": "abc de。12。fgh!ij k"}, {"start": 6.0, "end": 13.0, "text": " 99。lm nop? 7。qrs "},
                      {"start": 13.0, "end": 19.0, "text": "tuv wxyz"}]
        kw = dict(interpolate_method="linear", return_char_alignments=True)
    else:
        specials = ["<pad>", "<s>", "</s>", "<unk>", "|"]
        vocab = {tok: i for i, tok in enumerate(specials + list("abcdefghijklmnopqrstuvwxyz") + ["'"])}
        model, meta = _HFStyle(32).to("cuda:0"), {"language": "en", "dictionary": vocab, "type": "huggingface"}
        transcript = copy.deepcopy(TRANSCRIPT)
        kw = dict(return_char_alignments=True)
    ref_in, our_in = copy.deepcopy(transcript), copy.deepcopy(transcript)
    ref = A.align(ref_in, model, meta, audio, "cuda:0", **kw)
    ours = wxa.align(our_in, model, meta, audio, "cuda:0", **kw)
    d = same_records(ours["segments"], ref["segments"], score_tol=0.0)
    assert d is None, d
    assert same_records(ours["word_segments"], ref["word_segments"], score_tol=0.0) is None
    assert sum("start" in w for w in ours["word_segments"]) >= 8
