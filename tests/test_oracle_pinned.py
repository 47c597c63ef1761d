"""CPU, build container only: pins both oracles against the reference itself (imported unmodified
from /root/reference) on seeded random and edge inputs — the reference ships no tests of its own."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_import, ref_port

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope="module")
def A():
    return ref_import.load_reference_alignment()


def test_dp_functions_match_reference(A):
    rng = np.random.default_rng(11)
    for trial in range(150):
        T = int(rng.integers(1, 70)); V = int(rng.choice([5, 29, 40]))
        N = int(rng.integers(1, T + 3)); blank = int(rng.choice([0, V - 1, 2]))
        x = rng.standard_normal((T, V)).astype(np.float32) * 3
        E = torch.log_softmax(torch.from_numpy(x), -1)
        if trial % 3 == 1:
            E = torch.round(E * 4) / 4
        tok = rng.integers(0, V, N).astype(np.int32)
        Rr = A.get_trellis(E, tok.tolist(), blank)
        assert np.array_equal(Rr.numpy().view(np.int32), c_oracle.forward_trellis(E.numpy(), tok, blank).view(np.int32))
        assert torch.equal(Rr, ref_port.forward_trellis(E, tok.tolist(), blank))
        pr = A.backtrack(Rr, E, tok.tolist(), blank)
        pp = ref_port.walk_back(Rr, E, tok.tolist(), blank)
        rc, cs, ce, sc, ts = c_oracle.align_segment(E.numpy(), tok, blank)
        if pr is None:
            assert pp is None and rc == 1
            continue
        assert [(p.token_index, p.time_index, p.score) for p in pr] == [(p.token_index, p.time_index, p.score) for p in pp]
        text = "".join("|" if t == 1 else chr(97 + t % 26) for t in tok)
        sr = A.merge_repeats(pr, text)
        sp = ref_port.collapse_repeats(pp, text)
        assert [(s.label, s.start, s.end, s.score) for s in sr] == [(s.label, s.start, s.end, s.score) for s in sp]
        assert rc == 0 and [s.start for s in sr] == cs.tolist() and [s.end for s in sr] == ce.tolist()
        np.testing.assert_allclose(sc, [s.score for s in sr], rtol=1e-6, atol=0)
        wr = A.merge_words(sr); wp = ref_port.collapse_words(sp)
        assert [(w.label, w.start, w.end, w.score) for w in wr] == [(w.label, w.start, w.end, w.score) for w in wp]


def test_align_port_matches_reference(A):
    import copy
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import make_golden as mg
    ref_import.set_sentence_splitter(mg.regex_spans)
    dictionary = {c.lower(): i for i, c in enumerate(mg.LABELS)}
    rng = np.random.default_rng(5)
    for name, sc in mg.ALIGN_SCENARIOS.items():
        audio = np.zeros(int(sc["audio_s"] * 16000), np.float32)
        meta = {"language": sc["lang"], "dictionary": dictionary, "type": "torchaudio"}
        logits = [torch.from_numpy(rng.standard_normal((mg.frames_for_samples(
            min(int(s["end"] * 16000), audio.shape[0]) - int(s["start"] * 16000)), 29)).astype(np.float32) * 3)
            for s in sc["segments"] if s["start"] < sc["audio_s"]]
        kw = dict(interpolate_method=sc.get("interpolate_method", "nearest"),
                  return_char_alignments=sc.get("return_char_alignments", False))
        t_ref = copy.deepcopy(sc["segments"]); t_port = copy.deepcopy(sc["segments"])
        # every segment gets a logits tensor here even if a guard skips it: index by call order
        r1 = A.align(t_ref, mg.ScriptedModel(list(logits)), meta, audio, "cpu", **kw)
        r2 = ref_port.align_port(t_port, mg.ScriptedModel(list(logits)), meta, audio, "cpu", span_fn=mg.regex_spans, **kw)
        assert r1 == r2, name
        assert t_ref == t_port
