import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from whisperx_b200 import synth
from whisperx_b200.engine import AlignEngine, BatchSpec, _alloc_outputs
b = synth.make_batch(4096, "c2", align_frames=4)
eng = AlignEngine("cuda:0")
spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=b.vocab, blank_id=b.blank, logits=True,
                 seg_t1=np.zeros(4096), seg_ratio=b.durations / b.frame_len)
host_em = torch.from_numpy(b.emissions).pin_memory()
out = _alloc_outputs(spec, "cpu", (), pin=True)
for f in ("n_frames", "char_start", "char_end", "char_score"): setattr(out, f, None)
for i in range(6):
    t0 = time.perf_counter(); eng.align_batch_host(host_em, spec, out=out); print("call %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
