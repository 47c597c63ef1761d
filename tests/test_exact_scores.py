"""wxb_exact_scores_host (include/wxb_align.h): the host-side score reductions of the exact mode — Python-order
float64 sum / count per character (alignment.py:444), round(.,3) (:264), numpy-pairwise word mean (:308) —
against the reference's own outputs in tests/golden/dp_cases.* (no GPU needed: plain C in libwxb_align.so)."""
import ctypes as C

import numpy as np

from whisperx_b200 import _native as N
from whisperx_b200.engine import BatchSpec


def _run(spec, prob, status, c0, c1, words=False):
    lib = N.load()
    sN, W = max(spec.sum_N, 1), max(spec.n_words, 1)
    score, score_r, wscore = np.empty(sN), np.empty(sN), np.empty(W)
    cb = spec.to_c()
    N.check(lib.wxb_exact_scores_host(C.byref(cb), prob.ctypes.data, status.ctypes.data, c0.ctypes.data, c1.ctypes.data,
                                      score.ctypes.data, score_r.ctypes.data, wscore.ctypes.data if words else None),
            "wxb_exact_scores_host")
    return score, score_r, wscore


def test_char_scores_equal_the_reference_merge_repeats(golden_dp):
    meta, arr = golden_dp
    checked = 0
    for m in meta:
        if not m["ok"]:
            continue
        a = {k[len(m["name"]) + 1:]: arr[k] for k in arr.files if k.startswith(m["name"] + "_")}
        T, Nn = a["E"].shape[0], len(a["tok"])
        prob = np.ones(T, np.float32)
        p32 = a["pscore"].astype(np.float32)
        assert np.array_equal(p32.astype(np.float64), a["pscore"])        # Point.score is the f64 image of a float32
        prob[a["ptime"]] = p32
        spec = BatchSpec(frame_off=[0, T], tok_off=[0, Nn], tokens=a["tok"], vocab=a["E"].shape[1], blank_id=m["blank"])
        score, score_r, _ = _run(spec, prob, np.zeros(1, np.int32), a["sstart"].astype(np.int32), a["send"].astype(np.int32))
        assert score.tolist() == a["sscore"].tolist(), m["name"]
        assert score_r.tolist() == [round(float(x), 3) for x in a["sscore"]], m["name"]
        checked += 1
    assert checked >= 40


def test_word_scores_and_failed_segments():
    import pandas as pd
    rng = np.random.default_rng(3)
    T, n = 400, 150
    starts = np.sort(rng.choice(T - 1, n, replace=False)).astype(np.int32)
    ends = np.append(starts[1:], T - 1).astype(np.int32)
    prob = rng.uniform(0, 1, 2 * T).astype(np.float32)
    # words of 1..12 characters; every 7th character is one the model could not align (-1)
    wchar_off, wchar_idx, k = [0], [], 0
    while k < n:
        w = int(rng.integers(1, 13))
        for i in range(k, min(k + w, n)):
            wchar_idx.append(i if i % 7 else -1)
        k += w
        wchar_off.append(len(wchar_idx))
    W = len(wchar_off) - 1
    spec = BatchSpec(frame_off=[0, T, 2 * T], tok_off=[0, n, 2 * n], tokens=np.zeros(2 * n, np.int32), vocab=5,
                     seg_t1=np.zeros(2), seg_ratio=np.ones(2), word_off=[0, W, 2 * W],
                     wchar_off=np.concatenate([wchar_off, np.asarray(wchar_off[1:]) + wchar_off[-1]]),
                     wchar_idx=np.asarray(wchar_idx * 2, np.int32))
    status = np.array([0, 1], np.int32)
    score, score_r, wscore = _run(spec, prob, status, np.tile(starts, 2), np.tile(ends, 2), words=True)
    exp = [sum(float(x) for x in prob[a:b]) / (b - a) for a, b in zip(starts, ends)]
    assert score[:n].tolist() == exp and np.isnan(score[n:]).all()
    assert score_r[:n].tolist() == [round(x, 3) for x in exp]
    for w in range(W):
        idx = wchar_idx[wchar_off[w]:wchar_off[w + 1]]
        col = pd.Series([score_r[i] if i >= 0 else np.nan for i in idx], dtype=float)
        want = round(col.mean(), 3) if col.notna().any() else np.nan       # alignment.py:308
        assert (np.isnan(want) and np.isnan(wscore[w])) or wscore[w] == want, w
    assert np.isnan(wscore[W:]).all()
