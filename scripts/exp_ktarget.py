"""Development aid: C2 step time against the planner's k_target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisperx_b200 import synth
from whisperx_b200.engine import AlignEngine, BatchSpec

eng = AlignEngine("cuda:0")
wl = os.environ.get("WXB_EXP_WORKLOAD", "c2")
b = synth.make_batch(int(os.environ.get("WXB_EXP_SEGMENTS", "4096")), wl, align_frames=4)
E = torch.from_numpy(b.emissions).cuda()
spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=b.vocab,
                 blank_id=b.blank, logits=True)
for kt in (2, 3, 4, 5, 6, 7, 8):
    eng._lib.wxb_set_k_target(kt)
    plan = eng.plan(spec)
    out, ms = None, []
    for it in range(8):
        out = plan.run(E, out=out, timed=True)
        ms.append(out.extra["dp_ms"])
    print(f"{wl} k_target={kt}: launches={plan.launches} dp_ms={np.median(ms[2:]):.3f} -> {plan.cells/np.median(ms[2:])/1e6:.1f} Gcells/s failed={int(out.status.sum())}")
    plan.close()
eng._lib.wxb_set_k_target(-1)
