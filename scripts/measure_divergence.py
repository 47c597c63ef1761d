"""How often does the fused log-softmax (WXB_F_LOGITS: the engine's own exp2-based formula) lead to a different
alignment than the reference's torch.log_softmax?  The DP is bit-exact given identical log-probs; the two
log-softmax implementations differ by up to ~2e-5, which can flip a near-tie of the float32 trellis.

For BASELINE configs[1] (C2) and configs[3] (C4), iid and planted emissions, three pipelines on the same raw logits:
  fused   raw logits -> engine (WXB_F_LOGITS)
  cuda    torch.log_softmax on the GPU (what the reference does at alignment.py:233 with device="cuda") -> engine
  cpu     torch.log_softmax on the CPU (the reference with device="cpu") -> engine
(the engine's DP equals the reference's DP on the same floats — tests/ — so differences are the log-softmax's).
Reports the fraction of segments and of characters whose frames differ, and the largest frame shift.

    python scripts/measure_divergence.py [--segments 4096] > profiles/r2_divergence.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from whisperx_b200 import synth
from whisperx_b200.engine import BatchSpec, get_engine


def run(eng, b, em, logits):
    spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=b.vocab,
                     blank_id=b.blank, logits=logits)
    r = eng.align_batch(em, spec)
    return r.status.cpu().numpy(), r.char_start.cpu().numpy().astype(np.int64), r.char_end.cpu().numpy().astype(np.int64)


def compare(b, a, c):
    seg_of = np.repeat(np.arange(len(b.frame_len)), np.diff(b.tok_off))
    diff = (a[1] != c[1]) | (a[2] != c[2])
    segs = np.unique(seg_of[diff])
    shift = np.abs(a[1] - c[1])[diff]
    return {"segments_differing": int(len(segs)), "segments": int(len(b.frame_len)), "segment_fraction": len(segs) / len(b.frame_len),
            "chars_differing": int(diff.sum()), "chars": int(diff.size), "char_fraction": float(diff.mean()),
            "max_frame_shift": int(shift.max()) if shift.size else 0, "median_frame_shift": float(np.median(shift)) if shift.size else 0.0,
            "status_differing": int((a[0] != c[0]).sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=4096)
    args = ap.parse_args()
    eng = get_engine("cuda:0")
    out = {"what": __doc__.split("\n\n")[0]}
    for cfg, n in (("c2", args.segments), ("c4", max(args.segments // 8, 64))):
        for kind in ("iid", "planted"):
            b = synth.make_batch(n, cfg, align_frames=4, kind=kind)
            raw = torch.from_numpy(b.emissions)
            dev = raw.cuda()
            fused = run(eng, b, dev, True)
            cuda = run(eng, b, torch.log_softmax(dev, -1), False)
            cpu = run(eng, b, torch.log_softmax(raw, -1).cuda(), False)
            err = (eng.log_softmax(dev) - torch.log_softmax(dev, -1)).abs().max().item()
            out[f"{cfg}_{kind}"] = {"fused_vs_torch_cuda": compare(b, fused, cuda), "fused_vs_torch_cpu": compare(b, fused, cpu),
                                    "torch_cuda_vs_torch_cpu": compare(b, cuda, cpu), "max_abs_logprob_difference_fused_vs_torch_cuda": err}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
