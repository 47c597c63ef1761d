"""One launch each of the log-softmax shapes (V = 29 / 32: staged thread-per-row; V = 100: warp-per-row in
registers; V = 2304: CTA-per-row) for an ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from whisperx_b200.engine import get_engine
eng = get_engine("cuda:0")
for V in (29, 32, 100, 2304):
    x = torch.randn(800_000_000 // 4 // V, V, device="cuda") * 3
    for _ in range(2):
        y = eng.log_softmax(x)
    torch.cuda.synchronize()
    print(V, float(y[0, 0]))
