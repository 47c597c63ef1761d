"""Development aid: per-CTA timeline of one C2 step (WXB_TRACE=1) + side-stream sweep.
    python scripts/exp_timeline.py [out.npy]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("WXB_TRACE", "1")
if len(sys.argv) > 2:
    os.environ["WXB_SIDE_STREAMS"] = sys.argv[2]

import numpy as np
import torch

from whisperx_b200 import synth
from whisperx_b200.engine import AlignEngine, BatchSpec


def main():
    eng = AlignEngine("cuda:0")
    b = synth.make_batch(4096, os.environ.get("WXB_EXP_WORKLOAD", "c2"), align_frames=4)
    spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens,
                     vocab=b.vocab, blank_id=b.blank, logits=True)
    E = torch.from_numpy(b.emissions).cuda()
    plan = eng.plan(spec)
    out = None
    for it in range(6):
        out = plan.run(E, out=out, timed=True)
        print(f"it={it} dp_ms={out.extra['dp_ms']:.3f}")
    rows = np.zeros((4096, 8), np.uint64)
    n = eng._lib.wxb_plan_debug_trace(plan._h, rows.ctypes.data_as(C.POINTER(C.c_uint64)), 4096)
    rows = rows[:n].astype(np.int64)
    t0 = rows[:, 0].min()
    beg, fwd, end = (rows[:, 0] - t0) / 1e3, (rows[:, 1] - t0) / 1e3, (rows[:, 2] - t0) / 1e3
    print(f"rows={n} span={end.max():.1f} us  last begin={beg.max():.1f} us  last forward pass done={fwd.max():.1f} us  "
          f"first walk done={end.min():.1f} us")
    K, T = rows[:, 6], rows[:, 4]
    Wn = -(-rows[:, 5] // (32 * np.maximum(K, 1)))            # compute warps of the segment's CTA: ceil(N / 32K)
    KW = Wn * 100 + K
    for kw in sorted(set(KW.tolist()), reverse=True):
        m = KW == kw
        k = kw % 100
        print(f"W={kw // 100} K={k:2d} n={m.sum():4d} begin[{beg[m].min():7.1f},{beg[m].max():7.1f}] end max={end[m].max():7.1f} "
              f"dur mean={np.mean(end[m]-beg[m]):7.1f} max={np.max(end[m]-beg[m]):7.1f}  fwd share={np.mean((fwd[m]-beg[m])/(end[m]-beg[m])):.2f} "
              f"ns/frame={np.mean((fwd[m]-beg[m])*1e3/T[m]):.0f}")
    # concurrency over time
    grid = np.linspace(0, end.max(), 25)
    act = [(int(((beg <= g) & (end > g)).sum())) for g in grid]
    print("active CTAs over time:", list(zip(np.round(grid).astype(int).tolist(), act)))
    sm = rows[:, 3]
    per_sm_end = np.array([end[sm == s].max() for s in range(148) if (sm == s).any()])
    print(f"per-SM finish: min={per_sm_end.min():.0f} mean={per_sm_end.mean():.0f} max={per_sm_end.max():.0f} us")
    longest = np.argsort(-(end - beg))[:8]
    for i in longest:
        print(f"  long CTA: K={K[i]} T={T[i]} sm={sm[i]} begin={beg[i]:.0f} fwd_done={fwd[i]:.0f} end={end[i]:.0f}")
    if len(sys.argv) > 1:
        np.save(sys.argv[1], rows)


if __name__ == "__main__":
    main()
