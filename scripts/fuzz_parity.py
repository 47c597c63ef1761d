"""Soak test: random batches (shapes, vocabularies, blank ids, ties, padding, wildcards, planner knobs) through
the CUDA path and the C oracle; every status / frame index must be identical, scores within 1e-6.

    python scripts/fuzz_parity.py [--seconds 120] [--seed 0]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from oracle import c_oracle
from whisperx_b200.engine import BatchSpec, get_engine


def fuzz(seconds: float, seed: int, eng=None):
    """Runs random batches for `seconds`; returns (batches, segments, cells).  Raises on the first difference."""
    import types
    args = types.SimpleNamespace(seconds=seconds, seed=seed)
    eng = eng or get_engine("cuda:0")
    rng = np.random.default_rng(args.seed)
    t_end = time.time() + args.seconds
    batches = segs = cells = 0
    while time.time() < t_end:
        V = int(rng.choice([1, 2, 5, 29, 29, 29, 32, 32, 33, 48, 64, 100, 131, 300, 1100]))
        blank = int(rng.integers(0, V))
        wildcard = bool(rng.random() < 0.3)
        logits = bool(rng.random() < 0.5)
        align = int(rng.choice([1, 4]))
        kt = int(rng.choice([-1, -1, 1, 2, 3, 4, 5, 6, 7, 8]))
        S = int(rng.integers(1, 40))
        em, toks, fo, fl, to = [], [], [0], [], [0]
        for _ in range(S):
            T = int(rng.choice([rng.integers(1, 40), rng.integers(1, 400), rng.integers(200, 1600)]))
            N = int(rng.integers(0, T + 3)) if rng.random() < 0.7 else int(rng.integers(0, min(T, 60) + 1))
            x = rng.standard_normal((T, V)).astype(np.float32) * np.float32(rng.choice([0.5, 3, 10]))
            if not logits:
                x = torch.log_softmax(torch.from_numpy(x), -1).numpy()
                if rng.random() < 0.3:
                    x = np.round(x * 4) / 4
            tk = rng.integers(0, V, N).astype(np.int32)
            if wildcard and N:
                tk[rng.random(N) < 0.3] = -1
            pad = (-T) % align
            em.append(x)
            if pad:
                em.append(np.zeros((pad, V), np.float32))
            toks.append(tk); fl.append(T); fo.append(fo[-1] + T + pad); to.append(to[-1] + N)
        E = np.concatenate(em) if em else np.zeros((0, V), np.float32)
        tok = np.concatenate(toks) if toks else np.zeros(0, np.int32)
        spec = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=blank, logits=logits,
                         wildcard=wildcard, no_shortcut=bool(rng.random() < 0.3))
        old = eng._lib.wxb_set_k_target(kt)
        try:
            if rng.random() < 0.3:
                res = eng.align_batch_host(E, spec, want=("logprobs",) if logits else ())
            else:
                res = eng.align_batch(torch.from_numpy(E).cuda(), spec, want=("logprobs",) if logits else ()).cpu()
        finally:
            eng._lib.wxb_set_k_target(old if old > 0 else -1)
        lp = res.logprobs.view(-1, V).numpy() if logits else E
        for s in range(S):
            a, T = fo[s], fl[s]
            sl = slice(to[s], to[s + 1])
            rc, cs, ce, sc, ts = c_oracle.align_segment(lp[a:a + T], tok[sl], blank)
            ctx = dict(V=V, blank=blank, T=T, N=sl.stop - sl.start, wildcard=wildcard, logits=logits, kt=kt, seed=args.seed, batch=batches)
            assert int(res.status[s]) == rc, ctx
            if rc == 0:
                assert np.array_equal(res.char_start[sl].numpy(), cs) and np.array_equal(res.char_end[sl].numpy(), ce), ctx
                np.testing.assert_allclose(res.char_score[sl].numpy(), sc, rtol=1e-6, atol=0, err_msg=str(ctx))
                assert int(res.n_frames[s]) == ts, ctx
            cells += T * (sl.stop - sl.start)
        batches += 1
        segs += S
    return batches, segs, cells


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    batches, segs, cells = fuzz(args.seconds, args.seed)
    print(f"fuzz ok: {batches} batches, {segs} segments, {cells/1e6:.1f} Mcells, seed {args.seed}")


if __name__ == "__main__":
    main()
