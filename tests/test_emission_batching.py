"""CPU: how segments share wav2vec2 forwards (whisperx_b200.alignment._emit; the reference runs batch 1 and has
batching as a TODO at whisperx/alignment.py:215).  A small random-init wav2vec2 on the CPU: stacked equal-length
segments and zero-padded length buckets must reproduce the batch-1 logits of every valid frame (to float32
round-off: the arithmetic is the same, only the library's batched algorithm differs), and a group-norm feature
extractor — whose first layer normalises over time, so padding would change every frame — must never be padded."""
import numpy as np
import pytest
import torch


def small_model(extractor_mode):
    import torchaudio
    torch.manual_seed(3)
    return torchaudio.models.wav2vec2_model(
        extractor_mode=extractor_mode, extractor_conv_layer_config=[(32, 10, 5), (32, 3, 2), (32, 3, 2), (32, 3, 2), (32, 3, 2), (32, 2, 2), (32, 2, 2)],
        extractor_conv_bias=False, encoder_embed_dim=48, encoder_projection_dropout=0.0, encoder_pos_conv_kernel=16,
        encoder_pos_conv_groups=4, encoder_num_layers=2, encoder_num_heads=4, encoder_attention_dropout=0.0,
        encoder_ff_interm_features=96, encoder_ff_interm_dropout=0.0, encoder_dropout=0.0, encoder_layer_norm_first=extractor_mode == "layer_norm",
        encoder_layer_drop=0.0, aux_num_out=29).eval()


@pytest.fixture()
def wxa():
    from whisperx_b200 import alignment
    old = (alignment.EMISSION_BATCH, alignment.EMISSION_PAD, alignment.EMISSION_PAD_SLACK)
    yield alignment
    alignment.EMISSION_BATCH, alignment.EMISSION_PAD, alignment.EMISSION_PAD_SLACK = old


def jobs_and_audio():
    g = torch.Generator().manual_seed(1)
    audio = torch.randn(1, 16000 * 40, generator=g) * 0.1
    secs = [(0.0, 4.0), (4.0, 8.0), (8.0, 11.7), (11.7, 15.5), (15.6, 15.61), (16.0, 20.0), (20.0, 23.9), (24.0, 39.0), (39.0, 41.0)]
    return [(int(a * 16000), int(b * 16000)) for a, b in secs], audio


def test_frame_count_formula(wxa):
    m = small_model("layer_norm")
    for n in (400, 719, 720, 16000, 62400):
        with torch.inference_mode():
            assert m(torch.zeros(1, n))[0].shape[1] == wxa.frames_for_samples(n)


@pytest.mark.parametrize("mode", ["layer_norm", "group_norm"])
def test_batched_forwards_reproduce_batch_one(wxa, mode):
    model = small_model(mode)
    jobs, audio = jobs_and_audio()
    wxa.EMISSION_BATCH, wxa.EMISSION_PAD = 1, False
    one = wxa._emit(model, "torchaudio", audio, jobs, "cpu")
    calls = []
    hook = model.register_forward_hook(lambda mod, args, kw, out: calls.append(tuple(args[0].shape)), with_kwargs=True)
    wxa.EMISSION_BATCH, wxa.EMISSION_PAD, wxa.EMISSION_PAD_SLACK = 8, True, 0.1
    many = wxa._emit(model, "torchaudio", audio, jobs, "cpu")
    hook.remove()
    assert wxa.pads_safely(model, "torchaudio") == (mode == "layer_norm")
    for a, b in zip(one, many):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 2e-5
    stacked = [s for s in calls if s[0] > 1]
    assert (3, 64000) in stacked                                   # the three 4.0 s segments share a forward, unpadded
    if mode == "layer_norm":
        assert (3, 62400) in stacked                               # 3.9 / 3.8 / 3.7 s: one bucket, zero-padded to 3.9 s
        assert len(calls) == 5
    else:
        assert all(s == (3, 64000) for s in stacked)               # group norm: nothing is ever padded
    # the 160-sample segment still takes the reference's pad-to-400 path with `lengths`
    assert one[4].shape[0] == 1 and torch.equal(one[4], many[4])


def test_precision_knob_rejects_unknown_values(wxa):
    wxa.EMISSION_PRECISION, old = "fp8", wxa.EMISSION_PRECISION
    try:
        with pytest.raises(ValueError):
            wxa._emit(small_model("layer_norm"), "torchaudio", torch.zeros(1, 1000), [(0, 1000)], "cpu")
    finally:
        wxa.EMISSION_PRECISION = old
