"""Development aid: throughput of one (W, K) class against the number of resident CTAs per SM
(all segments identical: T frames, N tokens), to see where issue rate stops growing with occupancy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisperx_b200.engine import AlignEngine, BatchSpec

eng = AlignEngine("cuda:0")
V, T = 29, 1000
rng = np.random.default_rng(0)
for N in (160, 320):
    for per_sm in (1, 2, 4, 8, 12, 16, 24, 32):
        S = 148 * per_sm
        Tp = T + (-T) % 4
        fo = np.arange(S + 1, dtype=np.int64) * Tp
        fl = np.full(S, T, np.int32)
        to = np.arange(S + 1, dtype=np.int64) * N
        tok = rng.integers(1, V, S * N).astype(np.int32)
        E = torch.randn(S * Tp, V, device="cuda") * 3
        spec = BatchSpec(frame_off=fo, frame_len=fl, tok_off=to, tokens=tok, vocab=V, blank_id=0, logits=True)
        eng._lib.wxb_set_k_target(5)
        plan = eng.plan(spec)
        out, ms = None, []
        for it in range(6):
            out = plan.run(E, out=out, timed=True)
            ms.append(out.extra["dp_ms"])
        m = float(np.median(ms[2:]))
        warps = eng._lib.wxb_segment_warps(N)
        instr = S * T * (warps * (5.3 * eng._lib.wxb_tokens_per_lane(N) + 11) + 15)
        print(f"N={N} W={warps} K={eng._lib.wxb_tokens_per_lane(N)} CTAs/SM={per_sm:2d}: {m:.3f} ms  {plan.cells/m/1e6:7.1f} Gcells/s  "
              f"est. IPC/SMSP={instr/(m*1e-3*1.965e9*592):.2f}")
        plan.close()
        del E
eng._lib.wxb_set_k_target(-1)
