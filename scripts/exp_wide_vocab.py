"""Wide vocabularies (the reference's default ja / zh models): time of the DP on log-probs, of the fused
logits mode, and of the stand-alone log-softmax kernel.  usage: PYTHONPATH=. python scripts/exp_wide_vocab.py"""
import numpy as np, torch
from whisperx_b200.engine import AlignEngine, BatchSpec

eng = AlignEngine(0)
rng = np.random.default_rng(0)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for V, S, T, N in ((29, 256, 1500, 300), (512, 64, 1500, 150), (2304, 64, 1500, 100), (3503, 64, 1500, 100)):
    x = torch.randn((S * T, V), device="cuda") * 3
    fo = np.arange(S + 1, dtype=np.int64) * T
    to = np.arange(S + 1, dtype=np.int64) * N
    tok = rng.integers(1, V, S * N).astype(np.int32)
    gb = x.numel() * 4 / 1e9
    lp = eng.log_softmax(x)
    t_ls = timed(lambda: eng.log_softmax(x))
    t_torch = timed(lambda: torch.log_softmax(x, -1))
    t_dp = timed(lambda: eng.align_batch(lp, BatchSpec(frame_off=fo, tok_off=to, tokens=tok, vocab=V), want=()))
    t_fused = timed(lambda: eng.align_batch(x, BatchSpec(frame_off=fo, tok_off=to, tokens=tok, vocab=V, logits=True), want=()))
    print(f"V={V:5d} S={S} T={T} N={N} ({gb:.2f} GB): log_softmax {t_ls:.3f} ms ({2 * gb / t_ls * 1e3:.0f} GB/s; torch {t_torch:.3f} ms)  "
          f"DP on log-probs {t_dp:.3f} ms ({gb / t_dp * 1e3:.0f} GB/s)  fused logits {t_fused:.3f} ms ({gb / t_fused * 1e3:.0f} GB/s)", flush=True)
