"""Development aid: is the C2 slowdown caused by mixing tokens-per-lane classes on an SM?
Runs (a) C2 as is, (b) C2's frame counts with every transcript forced to one K class."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisperx_b200 import synth
from whisperx_b200.engine import AlignEngine, BatchSpec

eng = AlignEngine("cuda:0")
b = synth.make_batch(4096, "c2", align_frames=4)
E = torch.from_numpy(b.emissions).cuda()

def run(tag, tok_off, tokens):
    spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=tok_off, tokens=tokens, vocab=b.vocab,
                     blank_id=b.blank, logits=True)
    plan = eng.plan(spec)
    out = None
    ms = []
    for it in range(8):
        out = plan.run(E, out=out, timed=True)
        ms.append(out.extra["dp_ms"])
    print(f"{tag}: cells={plan.cells/1e6:.0f}M launches={plan.launches} dp_ms={np.median(ms[2:]):.3f} -> {plan.cells/np.median(ms[2:])/1e6:.1f} Gcells/s  failed={int(out.status.sum())}")
    plan.close()

run("c2 as is", b.tok_off, b.tokens)
rng = np.random.default_rng(0)
for n in (64, 128, 224):
    N = np.minimum(n, b.frame_len).astype(np.int64)
    to = np.concatenate([[0], np.cumsum(N)])
    tk = rng.integers(1, b.vocab, int(to[-1])).astype(np.int32)
    run(f"all N={n} (K={-(-n//32)})", to, tk)
