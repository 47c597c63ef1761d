"""CPU: both oracles (torch/pandas port and scalar C port) against the committed reference outputs
(tests/golden/dp_cases.*, generated from /root/reference by scripts/make_golden.py)."""
import random

import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_port
from tests.helpers import host_exp_matches_fixtures

SAME_EXP = host_exp_matches_fixtures()       # this machine's torch CPU exp is the one that generated the fixtures


def _cases(golden_dp):
    meta, arr = golden_dp
    for m in meta:
        yield m, {k[len(m["name"]) + 1:]: arr[k] for k in arr.files if k.startswith(m["name"] + "_")}


def test_golden_has_edge_cases(golden_dp):
    meta, _ = golden_dp
    assert len(meta) >= 60
    assert any(not m["ok"] for m in meta) and any(m["N"] == m["T"] for m in meta)
    assert any(m["kind"] == "ties" for m in meta) and any(m["blank"] != 0 for m in meta)


def test_c_oracle_trellis_bit_exact(golden_dp):
    for m, a in _cases(golden_dp):
        R = c_oracle.forward_trellis(a["E"], a["tok"], m["blank"])
        assert np.array_equal(R.view(np.int32), a["R"].view(np.int32)), m


def test_torch_port_trellis_bit_exact(golden_dp):
    for m, a in _cases(golden_dp):
        R = ref_port.forward_trellis(torch.from_numpy(a["E"]), a["tok"].tolist(), m["blank"])
        assert np.array_equal(R.numpy().view(np.int32), a["R"].view(np.int32)), m


def test_c_oracle_path_and_merge(golden_dp):
    for m, a in _cases(golden_dp):
        p = c_oracle.walk_back(a["R"], a["E"], a["tok"], m["blank"])
        if not m["ok"]:
            assert p is None, m
            assert c_oracle.align_segment(a["E"], a["tok"], m["blank"])[0] == 1
            continue
        assert p["tok"].tolist() == a["ptok"].tolist() and p["time"].tolist() == a["ptime"].tolist(), m
        np.testing.assert_allclose(p["prob"], a["pscore"], rtol=1e-6, atol=0)
        rc, cs, ce, sc, ts = c_oracle.align_segment(a["E"], a["tok"], m["blank"])
        assert rc == 0 and cs.tolist() == a["sstart"].tolist() and ce.tolist() == a["send"].tolist(), m
        np.testing.assert_allclose(sc, a["sscore"], rtol=1e-6, atol=0)
        assert ts == int(a["ptime"][-1]) + 1


def test_torch_port_path_merge_words_exact(golden_dp):
    for m, a in _cases(golden_dp):
        E = torch.from_numpy(a["E"])
        R = torch.from_numpy(a["R"])
        path = ref_port.walk_back(R, E, a["tok"].tolist(), m["blank"])
        if not m["ok"]:
            assert path is None
            continue
        assert [p.token_index for p in path] == a["ptok"].tolist()
        assert [p.time_index for p in path] == a["ptime"].tolist()
        if SAME_EXP:                                                     # same torch exp: bit-exact
            assert [p.score for p in path] == a["pscore"].tolist()
        else:                                                            # another MKL code path: 1 ulp of float32
            np.testing.assert_allclose([p.score for p in path], a["pscore"], rtol=2e-7, atol=0)
        segs = ref_port.collapse_repeats(path, m["transcript"])
        assert [s.start for s in segs] == a["sstart"].tolist() and [s.end for s in segs] == a["send"].tolist()
        if SAME_EXP:
            assert [s.score for s in segs] == a["sscore"].tolist()
        else:
            np.testing.assert_allclose([s.score for s in segs], a["sscore"], rtol=2e-7, atol=0)
        assert "".join(s.label for s in segs) == m["transcript"]
        words = ref_port.collapse_words(segs)
        assert [w.label for w in words] == m["words"]
        assert [w.start for w in words] == a["wstart"].tolist() and [w.end for w in words] == a["wend"].tolist()
        if SAME_EXP:
            assert [w.score for w in words] == a["wscore"].tolist()
        else:
            np.testing.assert_allclose([w.score for w in words], a["wscore"], rtol=2e-7, atol=0)


def test_round3_is_python_round():
    rnd = random.Random(7)
    for _ in range(50000):
        x = rnd.random() * rnd.choice([1, 10, 100, 3000])
        assert c_oracle.round3(x) == round(x, 3)
    for k in range(4000):                    # exact decimal ties (odd multiples of 1/16, 1/2000 grid)
        assert c_oracle.round3(k / 16) == round(k / 16, 3)
        assert c_oracle.round3(k / 2000) == round(k / 2000, 3)


def test_pairwise_sum_is_numpy_sum():
    rng = np.random.default_rng(3)
    for n in list(range(0, 300)) + [1000, 4097]:
        a = np.round(rng.random(n), 3)
        assert c_oracle.pairwise_sum(a) == np.sum(a), n


def test_pandas_mean_is_zero_filled_pairwise():
    """Series.mean() with NaN rows == numpy pairwise sum of the zero-filled vector / count."""
    import pandas as pd
    rng = np.random.default_rng(5)
    for n in (1, 3, 7, 8, 9, 15, 16, 17, 40, 129, 200):
        for _ in range(20):
            a = np.round(rng.random(n), 3)
            mask = rng.random(n) < 0.3
            if mask.all():
                mask[0] = False
            s = pd.Series(np.where(mask, np.nan, a))
            z = np.where(mask, 0.0, a)
            assert s.mean() == c_oracle.pairwise_sum(z) / int((~mask).sum())


def test_two_restatements_agree_on_random_cases():
    """The scalar C port and the torch port are independent restatements of the same reference lines
    (and of the wildcard extension): on random shapes, vocabularies, blank ids, ties and wildcards they
    must produce the same trellis bits, the same path and the same merged segments."""
    rng = np.random.default_rng(99)
    for case in range(60):
        V = int(rng.choice([2, 5, 29, 40]))
        blank = int(rng.integers(0, V))
        T = int(rng.integers(1, 70))
        N = int(rng.integers(1, T + 3))
        x = rng.standard_normal((T, V)).astype(np.float32) * 3
        E = torch.log_softmax(torch.from_numpy(x), -1)
        if case % 3 == 0:
            E = torch.round(E * 4) / 4
        tok = rng.integers(0, V, N).astype(np.int32)
        if case % 2:
            tok[rng.random(N) < 0.3] = -1            # wildcard extension
        Rc = c_oracle.forward_trellis(E.numpy(), tok, blank)
        Rt = ref_port.forward_trellis(E, tok.tolist(), blank)
        assert np.array_equal(Rc.view(np.int32), Rt.numpy().view(np.int32)), (case, T, N, V)
        segs, _ = ref_port.align_tokens(E, tok.tolist(), blank, "x" * N)
        rc, cs, ce, sc, ts = c_oracle.align_segment(E.numpy(), tok, blank)
        assert (segs is None) == (rc != 0), (case, T, N)
        if segs is not None:
            assert [s.start for s in segs] == cs.tolist() and [s.end for s in segs] == ce.tolist()
            np.testing.assert_allclose([s.score for s in segs], sc, rtol=1e-6, atol=0)
