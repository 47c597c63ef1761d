"""The C2 batch with other small vocabularies (HF wav2vec2 CTC heads are often 32 wide): fused-logits and
log-prob input.  usage: python scripts/exp_vocab_c2.py [V ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisperx_b200 import synth
from whisperx_b200.engine import AlignEngine, BatchSpec

eng = AlignEngine("cuda:0")
vs = [int(a) for a in sys.argv[1:]] or [29, 31, 32, 33, 36, 40, 48, 64]
for V in vs:
    b = synth.make_batch(2048, "c2", vocab=V, align_frames=4)
    E = torch.from_numpy(b.emissions).cuda()
    for logits in (True, False):
        x = E if logits else eng.log_softmax(E)
        spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=V, blank_id=b.blank, logits=logits)
        plan = eng.plan(spec)
        out = None
        ms = []
        for it in range(8):
            out = plan.run(x, out=out, timed=True)
            ms.append(out.extra["dp_ms"])
        print(f"V={V:3d} {'logits  ' if logits else 'logprobs'} {np.median(ms[3:]):.3f} ms  ({plan.cells / np.median(ms[3:]) / 1e9:.2f} Tcells/s)", flush=True)
        plan.close()
