"""Kernel (i) on its own: wxb_log_softmax_device over the C2 emission matrix ([3.56 M, 29] f32), achieved HBM
GB/s against MEASURED_PEAKS.json, next to torch.log_softmax on the same tensor."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from whisperx_b200.engine import get_engine

eng = get_engine("cuda:0")
peak = 6545.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {}
for V in (29, 100):
    rows = 412_579_056 // 4 // 29 if V == 29 else 1_000_000
    x = torch.randn(rows, V, device="cuda") * 3
    for name, fn in (("wxb_log_softmax_device", eng.log_softmax), ("torch.log_softmax", lambda t: torch.log_softmax(t, -1))):
        for _ in range(3):
            y = fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            y = fn(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gbs = 2 * x.numel() * 4 / ms / 1e6
        out[f"{name} V={V}"] = {"rows": rows, "ms": ms, "GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak}
    ref = torch.log_softmax(x, -1)
    out[f"max_abs_diff V={V}"] = float((eng.log_softmax(x) - ref).abs().max())
print(json.dumps(out, indent=1))
