"""GPU: the CUDA engine against the oracle through the C ABI — bit-exact trellis / final column /
paths / frame indices, scores within 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests.helpers import concat_cases, logprobs
from whisperx_b200 import synth
from whisperx_b200.engine import BatchSpec

pytestmark = pytest.mark.gpu

from tests.helpers import host_exp_matches_fixtures
EXP_IS_FIXTURE_EXP = host_exp_matches_fixtures()
SCORE_RTOL = 1e-6          # north_star: per-char scores within 1e-6 relative


def run(engine, cases, V, blank, align=1, want=("path", "final_col", "trellis"), host=False, **kw):
    E, fo, fl, to, tok = concat_cases(cases, V, align)
    spec = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=blank, **kw)
    if host:
        res = engine.align_batch_host(E, spec, want=want)
    else:
        res = engine.align_batch(torch.from_numpy(E).cuda(), spec, want=want).cpu()
    return spec, res


def check_against_oracle(spec, res, cases, blank, full=True):
    fo, to = spec.frame_off, spec.tok_off
    troff = spec.trellis_offsets()
    for s, (E, tk) in enumerate(cases):
        T, N = E.shape[0], len(tk)
        rc, cs, ce, sc, ts = c_oracle.align_segment(E, tk, blank)
        assert int(res.status[s]) == rc, (s, T, N)
        sl = slice(int(to[s]), int(to[s + 1]))
        if rc == 0:
            assert np.array_equal(res.char_start[sl].numpy(), cs), (s, T, N)
            assert np.array_equal(res.char_end[sl].numpy(), ce), (s, T, N)
            np.testing.assert_allclose(res.char_score[sl].numpy(), sc, rtol=SCORE_RTOL, atol=0)
            assert int(res.n_frames[s]) == ts
        else:
            assert int(res.n_frames[s]) == 0
            assert (res.char_start[sl] == -1).all() and torch.isnan(res.char_score[sl]).all()
        if not full:
            continue
        R = c_oracle.forward_trellis(E, tk, blank)
        if res.trellis is not None:
            Rg = res.trellis[troff[s]:troff[s + 1]].numpy().reshape(T + 1, N + 1)
            assert np.array_equal(Rg.view(np.int32), R.view(np.int32)), (s, T, N)
        if res.final_col is not None:
            fc = res.final_col[fo[s] + s: fo[s] + s + T + 1].numpy()
            assert np.array_equal(fc.view(np.int32), R[:, N].view(np.int32))
        if res.path_token is not None:
            pt = res.path_token[fo[s]:fo[s] + T].numpy()
            pl = res.path_logp[fo[s]:fo[s] + T].numpy()
            p = c_oracle.walk_back(R, E, tk, blank)
            if p is None:
                assert (pt == -1).all()
            else:
                on = np.nonzero(pt >= 0)[0]
                assert on.tolist() == p["time"].tolist()
                assert pt[on].tolist() == p["tok"].tolist()
                assert np.array_equal(pl[on].view(np.int32), p["logp"].view(np.int32))   # bit-exact log-probs


@pytest.mark.parametrize("V,blank,align", [(29, 0, 1), (29, 0, 4), (29, 7, 4), (100, 99, 1), (100, 99, 4), (5, 2, 1), (2, 1, 4), (32, 0, 1), (64, 63, 4)])
def test_random_segments_bit_exact(engine, V, blank, align):
    rng = np.random.default_rng(100 + V + align)
    cases = []
    for c in range(150):
        T = int(rng.integers(1, 160))
        N = int(rng.integers(1, T + 1)) if c % 5 else int(rng.integers(1, T + 4))
        cases.append((logprobs(rng, T, V, ties=(c % 3 == 1)), rng.integers(0, V, N).astype(np.int32)))
    spec, res = run(engine, cases, V, blank, align, no_shortcut=True)
    check_against_oracle(spec, res, cases, blank)


def test_golden_reference_vectors(engine, golden_dp):
    """Committed outputs of the real reference (tests/golden/dp_cases.*)."""
    meta, arr = golden_dp
    for V, blank in ((29, 0), (7, 3)):
        ms = [m for m in meta if m["V"] == V and m["blank"] == blank]
        cases = [(arr[m["name"] + "_E"], arr[m["name"] + "_tok"]) for m in ms]
        spec, res = run(engine, cases, V, blank, no_shortcut=True)
        troff = spec.trellis_offsets()
        for s, m in enumerate(ms):
            n = m["name"]
            T, N = m["T"], m["N"]
            Rg = res.trellis[troff[s]:troff[s + 1]].numpy().reshape(T + 1, N + 1)
            assert np.array_equal(Rg.view(np.int32), arr[n + "_R"].view(np.int32)), m
            assert int(res.status[s]) == (0 if m["ok"] else 1), m
            if not m["ok"]:
                continue
            sl = slice(int(spec.tok_off[s]), int(spec.tok_off[s + 1]))
            assert res.char_start[sl].tolist() == arr[n + "_sstart"].tolist(), m
            assert res.char_end[sl].tolist() == arr[n + "_send"].tolist(), m
            np.testing.assert_allclose(res.char_score[sl].numpy(), arr[n + "_sscore"], rtol=SCORE_RTOL, atol=0)
            pt = res.path_token[spec.frame_off[s]:spec.frame_off[s] + T].numpy()
            on = np.nonzero(pt >= 0)[0]
            assert on.tolist() == arr[n + "_ptime"].tolist() and pt[on].tolist() == arr[n + "_ptok"].tolist()
            pl = res.path_logp[spec.frame_off[s]:spec.frame_off[s] + T][on]
            # the engine's path log-probs are the reference's emission[t, c] floats, bit for bit (host independent) ...
            E, tok = arr[n + "_E"], arr[n + "_tok"]
            moved = np.diff(np.concatenate([[-1], arr[n + "_ptok"]])) != 0           # the frame that entered a token
            want = np.where(moved, E[arr[n + "_ptime"], tok[arr[n + "_ptok"]]], E[arr[n + "_ptime"], 0])
            assert np.array_equal(pl.numpy().view(np.int32), want.astype(np.float32).view(np.int32))
            # ... and their torch-CPU exp is the reference's Point.score wherever torch's exp is the one that generated
            # the fixtures (MKL's float32 exp depends on the host's instruction set: tests/helpers.host_exp_matches_fixtures)
            if EXP_IS_FIXTURE_EXP:
                assert torch.exp(pl).double().tolist() == arr[n + "_pscore"].tolist()
            else:
                np.testing.assert_allclose(torch.exp(pl).double().numpy(), arr[n + "_pscore"], rtol=2e-7, atol=0)


@pytest.mark.parametrize("N", [1, 31, 32, 33, 64, 65, 95, 97, 191, 192, 193, 256, 257, 384, 385, 416, 417, 512, 513, 576, 577,
                               700, 1023, 1025, 1500, 2047, 2048, 3100])
def test_every_token_bucket_boundary(engine, N):
    """(warps, tokens-per-lane) classes incl. the r=1 / r=32W lane splits and the 1->2->3 warp steps."""
    rng = np.random.default_rng(N)
    T = N + int(rng.integers(0, 120))
    cases = [(logprobs(rng, T, 29, ties=bool(N % 2)), rng.integers(0, 29, N).astype(np.int32))]
    small = N <= 700
    spec, res = run(engine, cases, 29, 0, align=4, want=("path", "final_col", "trellis") if small else ("path",))
    check_against_oracle(spec, res, cases, 0, full=small)


@pytest.mark.parametrize("k_target", [1, 2, 3, 5, 7, 8])
def test_k_target_sweep(engine, k_target):
    """Every instantiated tokens-per-lane class (K = 1..8) and many warp counts, by moving the planner knob."""
    from whisperx_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(100 + k_target)
    cases = []
    for N in (1, 5, 32, 33, 70, 128, 200, 257, 300, 449, 640, 900):
        T = N + int(rng.integers(0, 150))
        cases.append((logprobs(rng, T, 29, ties=bool(N % 2)), rng.integers(0, 29, N).astype(np.int32)))
    old = lib.wxb_set_k_target(k_target)
    try:
        spec, res = run(engine, cases, 29, 0, align=4, want=("path", "final_col", "trellis"))
    finally:
        lib.wxb_set_k_target(old if old > 0 else -1)
    check_against_oracle(spec, res, cases, 0, full=True)


def test_all_K_classes_one_batch(engine):
    rng = np.random.default_rng(9)
    cases = []
    for K in range(1, 65):
        N = 32 * (K - 1) + int(rng.integers(1, 33))
        T = N + int(rng.integers(0, 40))
        cases.append((logprobs(rng, T, 29), rng.integers(0, 29, N).astype(np.int32)))
    spec, res = run(engine, cases, 29, 0, align=4, want=())
    check_against_oracle(spec, res, cases, 0, full=False)


def test_edge_shapes_and_failures(engine):
    rng = np.random.default_rng(4)
    V = 29
    cases = []
    for T, N in ((1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (3, 5), (40, 40), (40, 41), (40, 39), (33, 1), (0, 3), (5, 0)):
        cases.append((logprobs(rng, T, V) if T else np.zeros((0, V), np.float32), rng.integers(0, V, N).astype(np.int32)))
    # with the host shortcut (default) and with the kernel deciding on its own
    for no_shortcut in (False, True):
        runnable = [c for c in cases if not (no_shortcut and (c[0].shape[0] == 0 or len(c[1]) == 0))] if no_shortcut else cases
        spec, res = run(engine, runnable, V, 0, want=("path",), no_shortcut=no_shortcut)
        for s, (E, tk) in enumerate(runnable):
            T, N = E.shape[0], len(tk)
            if N == 0 or T == 0:
                assert int(res.status[s]) == 1
                continue
            rc, cs, ce, sc, ts = c_oracle.align_segment(E, tk, 0)
            assert int(res.status[s]) == rc, (T, N, no_shortcut)
            if rc == 0:
                sl = slice(int(spec.tok_off[s]), int(spec.tok_off[s + 1]))
                assert res.char_start[sl].tolist() == cs.tolist() and res.char_end[sl].tolist() == ce.tolist()


def test_empty_batch(engine):
    spec = BatchSpec(frame_off=np.zeros(1, np.int64), tok_off=np.zeros(1, np.int64), tokens=np.zeros(0, np.int32), vocab=29)
    res = engine.align_batch(torch.zeros((0, 29), device="cuda"), spec)
    assert res.status.numel() >= 0


def test_invalid_arguments_raise(engine):
    from whisperx_b200._native import NativeError
    E = torch.zeros((4, 29), device="cuda")
    with pytest.raises(NativeError):
        engine.align_batch(E, BatchSpec(frame_off=[0, 4], tok_off=[0, 2], tokens=[1, 29], vocab=29))      # token out of range
    with pytest.raises(NativeError):
        engine.align_batch(E, BatchSpec(frame_off=[0, 4], tok_off=[0, 2], tokens=[1, 2], vocab=29, blank_id=29))
    with pytest.raises(NativeError):
        engine.align_batch(E, BatchSpec(frame_off=[4, 0], tok_off=[0, 2], tokens=[1, 2], vocab=29))
    with pytest.raises(ValueError):
        engine.align_batch(torch.zeros((2, 29), device="cuda"), BatchSpec(frame_off=[0, 4], tok_off=[0, 2], tokens=[1, 2], vocab=29))


def test_logits_mode_is_path_exact_on_own_logprobs(engine):
    """Fused log-softmax: feeding the engine's own log-probs to the oracle DP reproduces its paths
    exactly, and those log-probs are within 2 ulp-ish of torch's (SURVEY §7.2 hard part 1)."""
    rng = np.random.default_rng(21)
    for V, blank in ((29, 0), (100, 99), (33, 4), (32, 0), (48, 47), (64, 1), (1, 0), (2, 0)):   # odd <= 64: fused; others: split
        raw, toks = [], []
        for c in range(60):
            T = int(rng.integers(2, 300)); N = int(rng.integers(1, T + 1))
            raw.append(rng.standard_normal((T, V)).astype(np.float32) * 3)
            toks.append(rng.integers(0, V, N).astype(np.int32))
        for align in (1, 4):
            E, fo, fl, to, tok = concat_cases(list(zip(raw, toks)), V, align)
            spec = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=blank, logits=True)
            res = engine.align_batch(torch.from_numpy(E).cuda(), spec, want=("path", "logprobs")).cpu()
            lp = res.logprobs.view(-1, V).numpy()
            alone = engine.log_softmax(torch.from_numpy(E).cuda()).cpu().numpy()
            cases = []
            for s in range(len(raw)):
                mine = lp[fo[s]:fo[s] + fl[s]]
                assert np.array_equal(mine.view(np.int32), alone[fo[s]:fo[s] + fl[s]].view(np.int32))   # same arithmetic in both kernels
                ref = torch.log_softmax(torch.from_numpy(raw[s]), -1).numpy()
                np.testing.assert_allclose(mine, ref, rtol=0, atol=4e-6)
                cases.append((mine.copy(), toks[s]))
            spec2 = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=blank)
            check_against_oracle(spec2, res, cases, blank, full=False)
            # normalising first and aligning the log-probs gives the fused call's bytes (V <= 64: same arithmetic)
            two_step = engine.align_batch(torch.from_numpy(alone).cuda(), spec2, want=("path",)).cpu()
            for f in ("status", "n_frames", "char_start", "char_end", "char_score"):
                assert torch.equal(getattr(res, f).view(torch.uint8), getattr(two_step, f).view(torch.uint8)), (V, f)
            # path log-probs are taken from the same values
            for s, (mine, tk) in enumerate(cases):
                p = c_oracle.walk_back(c_oracle.forward_trellis(mine, tk, blank), mine, tk, blank)
                pt = res.path_token[fo[s]:fo[s] + fl[s]].numpy()
                on = np.nonzero(pt >= 0)[0]
                assert p is not None and on.tolist() == p["time"].tolist()
                assert np.array_equal(res.path_logp[fo[s]:fo[s] + fl[s]].numpy()[on].view(np.int32), p["logp"].view(np.int32))


def test_host_entry_point_matches_device(engine):
    rng = np.random.default_rng(33)
    cases = [(logprobs(rng, int(t), 29), rng.integers(0, 29, max(int(t) // 3, 1)).astype(np.int32))
             for t in rng.integers(5, 400, 40)]
    spec, a = run(engine, cases, 29, 0, align=4, want=("path",))
    _, b = run(engine, cases, 29, 0, align=4, want=("path",), host=True)
    for f in ("status", "n_frames", "char_start", "char_end", "char_score"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    for s in range(len(cases)):          # per-frame outputs: rows between segments are alignment padding
        sl = slice(int(spec.frame_off[s]), int(spec.frame_off[s] + spec.frame_len[s]))
        assert torch.equal(a.path_token[sl], b.path_token[sl]) and torch.equal(a.path_logp[sl], b.path_logp[sl])
    check_against_oracle(spec, b, cases, 0, full=False)


def test_timestamps_and_words_exact(engine):
    """Device-side round(x, 3) and numpy-order word means against Python / pandas arithmetic."""
    import pandas as pd
    rng = np.random.default_rng(77)
    cases, t1s, ratios, np_flags, word_off, wchar_off, wchar_idx = [], [], [], [], [0], [0], []
    for c in range(40):
        T = int(rng.integers(20, 500)); N = int(rng.integers(1, min(T, 200)))
        cases.append((logprobs(rng, T, 29), rng.integers(0, 29, N).astype(np.int32)))
        t1 = round(float(rng.random() * 1000), 3); dur = float(rng.random() * 30 + 0.1)
        t1s.append(t1); ratios.append(dur * 1 / T); np_flags.append(c % 2)
        # random word groups over the chars, with unalignable (-1) entries sprinkled in
        k = 0
        nw = 0
        while k < N:
            L = int(rng.integers(1, 20))
            idx = list(range(k, min(k + L, N)))
            k += L
            ent = []
            for i in idx:
                if rng.random() < 0.15:
                    ent.append(-1)
                ent.append(i)
            if rng.random() < 0.1:
                ent = [-1] * len(ent)
            wchar_idx += ent
            wchar_off.append(len(wchar_idx))
            nw += 1
        word_off.append(word_off[-1] + nw)
    E, fo, fl, to, tok = concat_cases(cases, 29, 4)
    spec = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=29, seg_t1=t1s, seg_ratio=ratios, seg_round_np=np_flags,
                     word_off=word_off, wchar_off=wchar_off, wchar_idx=wchar_idx)
    res = engine.align_batch(torch.from_numpy(E).cuda(), spec).cpu()
    for s in range(len(cases)):
        assert int(res.status[s]) == 0
        for k in range(int(to[s]), int(to[s + 1])):
            # odd segments emulate numpy-typed segment times: np.float64 arithmetic and rounding
            cast = np.float64 if np_flags[s] else float
            assert float(res.char_start_s[k]) == round(int(res.char_start[k]) * cast(ratios[s]) + cast(t1s[s]), 3)
            assert float(res.char_end_s[k]) == round(int(res.char_end[k]) * cast(ratios[s]) + cast(t1s[s]), 3)
            assert float(res.char_score_r[k]) == round(float(res.char_score[k]), 3)
        for w in range(word_off[s], word_off[s + 1]):
            ent = wchar_idx[wchar_off[w]:wchar_off[w + 1]]
            df = pd.DataFrame({
                "start": [float(res.char_start_s[to[s] + i]) if i >= 0 else None for i in ent],
                "end": [float(res.char_end_s[to[s] + i]) if i >= 0 else None for i in ent],
                "score": [float(res.char_score_r[to[s] + i]) if i >= 0 else None for i in ent]}, dtype=float)
            ws, we, wsc = df["start"].min(), df["end"].max(), round(df["score"].mean(), 3)
            for got, exp in ((res.word_start_s[w], ws), (res.word_end_s[w], we), (res.word_score[w], wsc)):
                assert (np.isnan(exp) and np.isnan(float(got))) or float(got) == exp, (s, w, float(got), exp)


def test_full_size_properties(engine):
    """BASELINE config 2 at full size (4,096 segments of 5-30 s): size-independent properties +
    a sampled oracle comparison."""
    b = synth.make_batch(4096, "c2", align_frames=4, kind="planted")
    E = torch.log_softmax(torch.from_numpy(b.emissions).cuda(), -1)
    spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens,
                     vocab=b.vocab, blank_id=b.blank)
    res = engine.align_batch(E, spec).cpu()
    assert int(res.status.sum()) == 0
    cs, ce = res.char_start.numpy().astype(np.int64), res.char_end.numpy().astype(np.int64)
    to = b.tok_off
    seg_of = np.repeat(np.arange(4096), np.diff(to))
    assert (ce > cs).all() and (cs >= 0).all()
    assert (ce <= b.frame_len[seg_of]).all()
    # inside a segment: token k+1 starts where token k ends (monotone, gap-free path)
    inner = np.ones(len(cs), bool); inner[to[1:] - 1] = False
    assert np.array_equal(cs[1:][inner[:-1]], ce[:-1][inner[:-1]])
    assert np.array_equal(ce[to[1:] - 1], res.n_frames.numpy())
    sc = res.char_score.numpy()
    assert ((sc > 0) & (sc <= 1.0 + 1e-12)).all()
    Ec = E.cpu().numpy()
    rng = np.random.default_rng(0)
    for s in rng.choice(4096, 48, replace=False):
        Es = Ec[b.frame_off[s]:b.frame_off[s] + b.frame_len[s]]
        tk = b.tokens[to[s]:to[s + 1]]
        rc, ocs, oce, osc, ts = c_oracle.align_segment(Es, tk, b.blank)
        assert rc == 0 and np.array_equal(ocs, cs[to[s]:to[s + 1]]) and np.array_equal(oce, ce[to[s]:to[s + 1]])
        np.testing.assert_allclose(sc[to[s]:to[s + 1]], osc, rtol=SCORE_RTOL, atol=0)


def test_backtrack_from_trellis_kernel(engine, golden_dp):
    meta, arr = golden_dp
    for m in meta[::3]:
        n = m["name"]
        st, ts, pt, pl = engine.backtrack_trellis(torch.from_numpy(arr[n + "_R"]), torch.from_numpy(arr[n + "_E"]),
                                                  arr[n + "_tok"].tolist(), m["blank"])
        assert st == (0 if m["ok"] else 1)
        if m["ok"]:
            on = np.nonzero(pt.numpy() >= 0)[0]
            assert on.tolist() == arr[n + "_ptime"].tolist() and pt.numpy()[on].tolist() == arr[n + "_ptok"].tolist()
            if EXP_IS_FIXTURE_EXP:
                assert torch.exp(pl[on]).double().tolist() == arr[n + "_pscore"].tolist()
            else:
                np.testing.assert_allclose(torch.exp(pl[on]).double().numpy(), arr[n + "_pscore"], rtol=2e-7, atol=0)


def test_repeated_runs_are_bitwise_identical(engine):
    """The wavefront hand-offs (mbarriers between producer and compute warps) must not leave room for a
    race: the same plan run many times, interleaved with other plans, returns the same bytes."""
    from whisperx_b200 import synth
    b = synth.make_batch(384, "c2", align_frames=4, kind="planted")
    spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=b.vocab,
                     blank_id=b.blank, logits=True, seg_t1=np.zeros(384), seg_ratio=b.durations / b.frame_len)
    E = torch.from_numpy(b.emissions).cuda()
    plan = engine.plan(spec)
    other = synth.make_batch(64, "c4", align_frames=4)
    ospec = BatchSpec(frame_off=other.frame_off, frame_len=other.frame_len, tok_off=other.tok_off, tokens=other.tokens,
                      vocab=other.vocab, blank_id=other.blank, logits=True)
    oE = torch.from_numpy(other.emissions).cuda()
    first = None
    valid = np.zeros(int(b.frame_off[-1]), bool)          # rows between segments are padding: never written
    for s0, T in zip(b.frame_off[:-1], b.frame_len):
        valid[int(s0):int(s0) + int(T)] = True
    valid = torch.from_numpy(valid)
    for it in range(12):
        res = plan.run(E, want=("path",))
        if it % 3 == 1:
            engine.align_batch(oE, ospec)           # a different shape class mix in between
        torch.cuda.synchronize()
        got = [getattr(res, f).cpu() for f in ("status", "n_frames", "char_start", "char_end", "char_score", "char_start_s",
                                               "char_end_s", "char_score_r")]
        got += [res.path_token.cpu()[valid], res.path_logp.cpu()[valid]]
        if first is None:
            first = got
            assert int(got[0].sum()) == 0
        else:
            for a, c in zip(first, got):
                assert torch.equal(a.view(torch.uint8), c.view(torch.uint8))
    # same buffers every call: from the third call on the plan is replayed as a CUDA graph; a larger plan in
    # between regrows the context's scratch, which must invalidate the captured pointers
    from whisperx_b200.engine import _alloc_outputs
    out = _alloc_outputs(spec, engine.device, ("path",))
    big = synth.make_batch(700, "c5", align_frames=4)
    bspec = BatchSpec(frame_off=big.frame_off, frame_len=big.frame_len, tok_off=big.tok_off, tokens=big.tokens,
                      vocab=big.vocab, blank_id=big.blank, logits=True)
    bE = torch.from_numpy(big.emissions).cuda()
    for it in range(10):
        for f in ("status", "char_start", "char_end", "path_token"):
            getattr(out, f).fill_(-7)
        plan.run(E, out=out)
        torch.cuda.synchronize()
        got = [getattr(out, f).cpu() for f in ("status", "n_frames", "char_start", "char_end", "char_score", "char_start_s",
                                               "char_end_s", "char_score_r")]
        got += [out.path_token.cpu()[valid], out.path_logp.cpu()[valid]]
        for a, c in zip(first, got):
            assert torch.equal(a.view(torch.uint8), c.view(torch.uint8)), it
        if it == 5:
            engine.align_batch(bE, bspec)               # needs far more choice-bit scratch than `plan`
    plan.close()


@pytest.mark.parametrize("V,blank", [(29, 0), (100, 99), (7, 3), (32, 0)])
def test_wildcard_tokens_extension(engine, V, blank):
    """EXTENSION (not in the reference fork; parity unpinned by the reference): token id -1 is
    a wildcard whose emission is the best non-blank log-prob of the frame.  Checked against the oracle's
    restatement of the same rule: trellis, path and frames bit-exact, scores 1e-6."""
    from oracle import ref_port
    rng = np.random.default_rng(500 + V)
    cases = []
    for N in (1, 3, 20, 33, 64, 100, 161, 257, 400, 700):
        T = N + int(rng.integers(0, 200))
        tok = rng.integers(0, V, N).astype(np.int32)
        tok[rng.random(N) < 0.25] = -1
        if N >= 3:
            tok[0] = -1; tok[-1] = -1
        cases.append((logprobs(rng, T, V, ties=bool(N % 2)), tok))
    spec, res = run(engine, cases, V, blank, align=4, wildcard=True)
    check_against_oracle(spec, res, cases, blank, full=True)
    hopeless = [(logprobs(rng, 50, V), np.full(60, -1, np.int32))]           # N > T, all wildcards: fails
    for no_shortcut in (False, True):
        spec_h, res_h = run(engine, hopeless, V, blank, align=4, want=("path",), wildcard=True, no_shortcut=no_shortcut)
        check_against_oracle(spec_h, res_h, hopeless, blank, full=False)
        assert int(res_h.status[0]) == 1
    # the two oracles agree with each other on the extension
    E, tk = cases[3]
    segs, _ = ref_port.align_tokens(torch.from_numpy(E), tk.tolist(), blank, "x" * len(tk))
    rc, cs, ce, sc, _ = c_oracle.align_segment(E, tk, blank)
    assert rc == 0 and [s.start for s in segs] == cs.tolist() and [s.end for s in segs] == ce.tolist()
    np.testing.assert_allclose([s.score for s in segs], sc, rtol=SCORE_RTOL, atol=0)
    # fused log-softmax + wildcards: path-exact on the engine's own log-probs
    raw = [(rng.standard_normal((c[0].shape[0], V)).astype(np.float32) * 3, c[1]) for c in cases[:8]]
    spec2, res2 = run(engine, raw, V, blank, align=4, want=("path", "logprobs"), wildcard=True, logits=True)
    lp = res2.logprobs.view(-1, V)
    own = [(lp[spec2.frame_off[s]:spec2.frame_off[s] + raw[s][0].shape[0]].numpy(), raw[s][1]) for s in range(len(raw))]
    check_against_oracle(spec2, res2, own, blank, full=False)
    # without the flag a -1 token is rejected
    from whisperx_b200._native import NativeError
    with pytest.raises(NativeError, match="outside"):
        run(engine, cases[:2], V, blank)


@pytest.mark.parametrize("chunk_bytes", [4096, 150_000, 3_000_000])
def test_host_path_with_tiny_chunks(engine, chunk_bytes):
    """wxb_align_host streams the emissions in chunks and lets every class wait only for the chunk holding its
    segments' last rows; chunks smaller than one segment (4 KB = 35 rows) and chunks of a few segments must give
    the device path's bytes, word reduction and per-chunk result copies included."""
    b = synth.make_batch(96, "c2", align_frames=4, kind="planted")
    S = 96
    rng = np.random.default_rng(5)
    word_off, wchar_off, wchar_idx = [0], [0], []
    for s in range(S):
        N = int(b.tok_off[s + 1] - b.tok_off[s]); k = 0; nw = 0
        while k < N:
            L = int(rng.integers(1, 9))
            wchar_idx += list(range(k, min(k + L, N))); wchar_off.append(len(wchar_idx)); k += L; nw += 1
        word_off.append(word_off[-1] + nw)
    spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=b.vocab,
                     blank_id=b.blank, logits=True, seg_t1=np.arange(S) * 30.0, seg_ratio=b.durations / b.frame_len,
                     word_off=word_off, wchar_off=wchar_off, wchar_idx=wchar_idx)
    dev = engine.align_batch(torch.from_numpy(b.emissions).cuda(), spec, want=("path",)).cpu()
    old = engine._lib.wxb_set_chunk_bytes(engine._ctx, chunk_bytes)
    try:
        host = engine.align_batch_host(b.emissions, spec, want=("path",))
    finally:
        engine._lib.wxb_set_chunk_bytes(engine._ctx, old)
    valid = np.zeros(int(b.frame_off[-1]), bool)
    for s0, T in zip(b.frame_off[:-1], b.frame_len):
        valid[int(s0):int(s0) + int(T)] = True
    for f in ("status", "n_frames", "char_start", "char_end", "char_score", "char_start_s", "char_end_s", "char_score_r",
              "word_start_s", "word_end_s", "word_score"):
        a, c = getattr(dev, f), getattr(host, f)
        assert torch.equal(a.view(torch.uint8), c.view(torch.uint8)), f
    assert torch.equal(dev.path_token[torch.from_numpy(valid)], host.path_token[torch.from_numpy(valid)])
    assert int(dev.status.sum()) == 0


def test_maximum_transcript_length(engine):
    """The largest transcript the build accepts: 31 compute warps x 32 lanes x 8 columns = 7,936 tokens in one CTA
    of 1,024 threads; one more token is refused with a clear error."""
    from whisperx_b200._native import NativeError
    assert engine.limits.max_tokens == 31 * 32 * 8
    rng = np.random.default_rng(8)
    N = engine.limits.max_tokens
    T = N + 64
    cases = [(logprobs(rng, T, 29), rng.integers(0, 29, N).astype(np.int32)),
             (logprobs(rng, 300, 29), rng.integers(0, 29, 120).astype(np.int32))]
    spec, res = run(engine, cases, 29, 0, align=4, want=("path",))
    check_against_oracle(spec, res, cases, 0, full=False)
    assert int(res.status[0]) == 0
    too_long = [(logprobs(rng, N + 70, 29), rng.integers(0, 29, N + 1).astype(np.int32))]
    with pytest.raises(NativeError, match="at most"):
        run(engine, too_long, 29, 0, want=())


@pytest.mark.parametrize("V,blank", [(512, 0), (2304, 1), (3503, 3502), (7001, 0)])
def test_wide_vocabularies_shorten_the_tiles(engine, V, blank):
    """The reference's default zh / ja models (alignment.py:34-57) have vocabularies in the thousands:
    the planner then picks 16-, 8- or 4-frame tiles so the ring still fits shared memory.  Same bit-exact
    bar, log-prob and fused-logit input, one- and many-warp segments, unaligned row offsets."""
    rng = np.random.default_rng(V)
    cases = []
    for T, N in ((1, 1), (7, 3), (50, 20), (90, 90), (130, 70), (200, 199), (160, 161)):
        cases.append((logprobs(rng, T, V, ties=(T % 2 == 0)), rng.integers(0, V, N).astype(np.int32)))
    for align in (1, 4):
        spec, res = run(engine, cases, V, blank, align, no_shortcut=True)
        check_against_oracle(spec, res, cases, blank)
    # fused log-softmax on raw logits: paths must be those of the DP over the engine's own log-probs
    raw = [(rng.standard_normal(E.shape).astype(np.float32) * 4, tk) for E, tk in cases]
    E, fo, fl, to, tok = concat_cases(raw, V, 1)
    spec = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=blank, logits=True, no_shortcut=True)
    res = engine.align_batch(torch.from_numpy(E).cuda(), spec, want=("path", "logprobs")).cpu()
    lp = res.logprobs.view(-1, V).numpy()
    mine = [(lp[fo[s]:fo[s] + fl[s]].copy(), raw[s][1]) for s in range(len(raw))]
    for (m, _), (r, _) in zip(mine, raw):
        np.testing.assert_allclose(m, torch.log_softmax(torch.from_numpy(r), -1).numpy(), rtol=0, atol=2e-5)   # a few ulp at |logp| ~ 20-30 (fp32 sum of V terms)
    check_against_oracle(BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=blank), res, mine, blank, full=False)
    # the same logits without asking for the log-probs (scratch buffer), replayed (CUDA graph from the 3rd run),
    # and through the host entry point in small chunks (normalised in place chunk by chunk): same bytes
    Ed = torch.from_numpy(E).cuda()
    plan = engine.plan(spec)
    outs = None
    for _ in range(4):
        outs = plan.run(Ed, out=outs, want=("path",))
    torch.cuda.synchronize()
    again = outs.cpu()
    plan.close()
    old = engine._lib.wxb_set_chunk_bytes(engine._ctx, 40 * V * 4)
    try:
        host = engine.align_batch_host(E, spec, want=("path", "logprobs"))
        host2 = engine.align_batch_host(E, spec, want=("path",))
    finally:
        engine._lib.wxb_set_chunk_bytes(engine._ctx, old)
    # WXB_F_CLOBBER: the caller's buffer is scratch, the rows are normalised in place (no copy)
    scratch = Ed.clone()
    spec_c = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=blank, logits=True, no_shortcut=True, clobber=True)
    inplace = engine.align_batch(scratch, spec_c, want=("path",)).cpu()
    assert torch.equal(scratch.cpu().view(torch.int32), res.logprobs.view(-1, V).view(torch.int32))
    for f in ("status", "n_frames", "char_start", "char_end", "char_score", "path_token", "path_logp"):
        for other in (again, host, host2, inplace):
            assert torch.equal(getattr(res, f).view(torch.uint8), getattr(other, f).view(torch.uint8)), f
    assert torch.equal(res.logprobs.view(torch.int32), host.logprobs.view(torch.int32))


def test_vocabulary_beyond_the_limit_is_refused(engine):
    from whisperx_b200._native import NativeError
    V = int(engine.limits.max_vocab) + 1
    E = torch.zeros((4, V), device="cuda")
    with pytest.raises(NativeError):
        engine.align_batch(E, BatchSpec(frame_off=[0, 4], tok_off=[0, 2], tokens=[1, 2], vocab=V))


@pytest.mark.needs_reference
def test_engine_against_the_reference_functions_live(engine):
    """The reference's own get_trellis / backtrack / merge_repeats (unmodified whisperx/alignment.py, from /root/reference
    or its build-time copy oracle/_ref) on this machine against the CUDA path on the same log-probs: random shapes incl.
    N = T and N > T, blank != 0, quantised (tie-heavy) emissions.  Trellis and frames bit-exact; scores exact once the
    path's log-probs go through the host's exp and the reference-order reductions (wxb_exact_scores_host)."""
    import ctypes as C
    from oracle import ref_import
    from whisperx_b200 import _native as N
    A = ref_import.load_reference_alignment()
    rng = np.random.default_rng(2026)
    cases, blanks = [], []
    for i in range(24):
        V = int(rng.choice([5, 29, 32, 100]))
        T = int(rng.integers(2, 320))
        Nn = int(rng.integers(1, T + 2)) if i % 5 else T
        cases.append((logprobs(rng, T, V, ties=bool(i % 3 == 0)), rng.integers(0, V, Nn).astype(np.int32)))
        blanks.append(V)
    for V in sorted(set(blanks)):
        sub = [c for c, v in zip(cases, blanks) if v == V]
        blank = 0 if V == 29 else V - 1
        spec, res = run(engine, sub, V, blank, align=4, want=("path", "trellis"), no_shortcut=True)
        toff = spec.trellis_offsets()
        prob = torch.exp(res.path_logp).contiguous()
        score, score_r = np.empty(max(spec.sum_N, 1)), np.empty(max(spec.sum_N, 1))
        cb = spec.to_c()
        N.check(engine._lib.wxb_exact_scores_host(C.byref(cb), prob.data_ptr(), res.status.data_ptr(), res.char_start.data_ptr(),
                                                  res.char_end.data_ptr(), score.ctypes.data, score_r.ctypes.data, None), "exact")
        for s, (E, tk) in enumerate(sub):
            Et = torch.from_numpy(E)
            R = A.get_trellis(Et, tk.tolist(), blank)
            mine = res.trellis[toff[s]:toff[s + 1]].view(R.shape)
            assert torch.equal(mine.view(torch.int32), R.view(torch.int32)), (V, s)
            path = A.backtrack(R, Et, tk.tolist(), blank)
            sl = slice(int(spec.tok_off[s]), int(spec.tok_off[s + 1]))
            if path is None:
                assert int(res.status[s]) == 1
                continue
            segs = A.merge_repeats(path, "x" * len(tk))
            assert int(res.status[s]) == 0
            assert res.char_start[sl].tolist() == [g.start for g in segs] and res.char_end[sl].tolist() == [g.end for g in segs]
            assert score[sl].tolist() == [g.score for g in segs]
            np.testing.assert_allclose(res.char_score[sl].numpy(), [g.score for g in segs], rtol=SCORE_RTOL, atol=0)
