"""Development aid: the C2 / C5 step with raw logits in (fused log-softmax in the producer warp) against the same
batch with log-probs in (the producer only stages tiles and builds column 0): how much of the step is the producer?
    python scripts/exp_modes.py [c2|c5]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from whisperx_b200 import synth
from whisperx_b200.engine import AlignEngine, BatchSpec, _alloc_outputs


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
    eng = AlignEngine("cuda:0")
    b = synth.make_batch(4096, cfg, align_frames=4)
    E = torch.from_numpy(b.emissions).cuda()
    LP = torch.log_softmax(E, -1)
    for name, em, logits in (("logits (fused log-softmax)", E, True), ("log-probs in", LP, False)):
        spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=b.vocab,
                         blank_id=b.blank, logits=logits, seg_t1=np.zeros(4096), seg_ratio=b.durations / b.frame_len)
        plan = eng.plan(spec)
        outs = [_alloc_outputs(spec, eng.device, ()) for _ in range(2)]
        for i in range(10):
            plan.run_overlapped(em, outs[i % 2])
        eng.join(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50):
            plan.run_overlapped(em, outs[i % 2])
        eng.join(); e1.record(); torch.cuda.synchronize()
        ov = e0.elapsed_time(e1) / 50
        for i in range(10):
            plan.run(em, out=outs[0])
        torch.cuda.synchronize(); e0.record()
        for i in range(50):
            plan.run(em, out=outs[0])
        e1.record(); torch.cuda.synchronize()
        print(f"{cfg} {name:28s} overlapped {ov:.4f} ms  serial {e0.elapsed_time(e1) / 50:.4f} ms")
        plan.close()


if __name__ == "__main__":
    main()
