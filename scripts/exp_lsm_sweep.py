"""log_softmax_kernel across vocabulary widths (thread-per-row for V <= 64, warp-per-row above)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from whisperx_b200.engine import get_engine
eng = get_engine("cuda:0")
for V in (29, 32, 48, 64, 65, 100, 128, 200, 256, 512, 1024, 2304, 3503, 7001, 9000):
    rows = 1_200_000_000 // 4 // V
    x = torch.randn(rows, V, device="cuda") * 3
    res = []
    for fn in (eng.log_softmax, lambda t: torch.log_softmax(t, -1)):
        for _ in range(3):
            fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res.append(2 * x.numel() * 4 / ms / 1e6)
    print(f"V={V:5d} rows={rows:9d}  wxb {res[0]:7.0f} GB/s   torch {res[1]:7.0f} GB/s", flush=True)
