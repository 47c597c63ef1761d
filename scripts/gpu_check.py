"""Ad-hoc GPU sanity run (development aid; the judged checks live in tests/ and bench.py).

    python scripts/gpu_check.py [--big]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from oracle import c_oracle
from whisperx_b200 import synth
from whisperx_b200.engine import AlignEngine, BatchSpec


def random_cases(rng, n_cases, V, maxT=200, align=1):
    em, toks, fo, fl, to = [], [], [0], [], [0]
    for c in range(n_cases):
        T = int(rng.integers(1, maxT))
        N = int(rng.integers(1, T + 1)) if c % 7 else int(rng.integers(1, T + 4))
        x = rng.standard_normal((T, V)).astype(np.float32) * 3
        m = x.max(1, keepdims=True)
        lp = (x - m) - np.log(np.exp(x - m).sum(1, keepdims=True))
        lp = lp.astype(np.float32)
        if c % 3 == 1:
            lp = np.round(lp * 4) / 4          # tie stress
        pad = (-T) % align
        em.append(lp)
        if pad:
            em.append(np.zeros((pad, V), np.float32))
        toks.append(rng.integers(0, V, N).astype(np.int32))
        fl.append(T)
        fo.append(fo[-1] + T + pad)
        to.append(to[-1] + N)
    return np.concatenate(em), np.array(fo, np.int64), np.array(fl, np.int32), np.array(to, np.int64), np.concatenate(toks)


def compare(eng, E, fo, fl, to, tok, V, blank, tag, want=("path", "final_col", "trellis")):
    spec = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=blank, no_shortcut=True)
    res = eng.align_batch(torch.from_numpy(E).cuda(), spec, want=want).cpu()
    bad = 0
    troff = spec.trellis_offsets()
    for s in range(len(fl)):
        T, N = int(fl[s]), int(to[s + 1] - to[s])
        Es = E[fo[s]:fo[s] + T]
        tk = tok[to[s]:to[s + 1]]
        rc, cs, ce, sc, ts = c_oracle.align_segment(Es, tk, blank)
        ok = int(res.status[s]) == rc
        if rc == 0:
            ok &= np.array_equal(res.char_start[to[s]:to[s + 1]].numpy(), cs)
            ok &= np.array_equal(res.char_end[to[s]:to[s + 1]].numpy(), ce)
            ok &= np.allclose(res.char_score[to[s]:to[s + 1]].numpy(), sc, rtol=1e-6, atol=0)
            ok &= int(res.n_frames[s]) == ts
        if "trellis" in want:
            R = c_oracle.forward_trellis(Es, tk, blank)
            Rg = res.trellis[troff[s]:troff[s + 1]].numpy().reshape(T + 1, N + 1)
            ok &= np.array_equal(R.view(np.int32), Rg.view(np.int32))
            fc = res.final_col[fo[s] + s: fo[s] + s + T + 1].numpy()
            ok &= np.array_equal(fc.view(np.int32), R[:, N].view(np.int32))
        if not ok:
            bad += 1
            if bad <= 5:
                print(f"  MISMATCH {tag} seg={s} T={T} N={N} rc={rc} got_status={int(res.status[s])}")
    print(f"{tag}: {len(fl)} segments, {bad} mismatches")
    return bad


def main():
    eng = AlignEngine("cuda:0")
    print("limits", eng.limits.max_vocab, eng.limits.max_tokens, eng.limits.sm_count)
    rng = np.random.default_rng(0)
    bad = 0
    for V, blank, align in ((29, 0, 1), (29, 0, 4), (29, 5, 4), (100, 99, 1), (100, 99, 4), (5, 2, 4)):
        E, fo, fl, to, tok = random_cases(rng, 200, V, align=align)
        bad += compare(eng, E, fo, fl, to, tok, V, blank, f"V={V} blank={blank} align={align}")
    # larger K classes
    for N in (33, 64, 65, 400, 416, 417, 700, 1000, 1500, 2048):
        T = N + int(rng.integers(0, 300))
        x = rng.standard_normal((T, 29)).astype(np.float32) * 3
        lp = torch.log_softmax(torch.from_numpy(x), -1).numpy()
        tok = rng.integers(0, 29, N).astype(np.int32)
        bad += compare(eng, lp, np.array([0, T], np.int64), np.array([T], np.int32), np.array([0, N], np.int64), tok,
                       29, 0, f"N={N} T={T}", want=("path", "final_col", "trellis") if N <= 700 else ("path",))
    print("TOTAL MISMATCHES", bad)

    # rough timing, C2-like
    nseg = 4096 if "--big" in sys.argv else 512
    b = synth.make_batch(nseg, "c2", align_frames=4)
    for logits in (False, True):
        spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens,
                         vocab=b.vocab, blank_id=b.blank, logits=logits)
        E = torch.from_numpy(b.emissions).cuda()
        if not logits:
            E = torch.log_softmax(E, -1)
        plan = eng.plan(spec)
        out = None
        for it in range(5):
            out = plan.run(E, out=out, timed=True)
            print(f"logits={logits} it={it} dp_ms={out.extra['dp_ms']:.3f}  cells={plan.cells/1e9:.3f}G "
                  f"-> {plan.cells/out.extra['dp_ms']/1e6:.1f} Gcells/s  launches={plan.launches}")
        st = out.status.cpu()
        print("failed:", int(st.sum()))
        plan.close()
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
