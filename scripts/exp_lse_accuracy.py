"""Accuracy of the engine's log-softmax against float64 truth, next to torch's own fp32 kernels."""
import numpy as np, torch
from whisperx_b200.engine import AlignEngine
eng = AlignEngine(0)
rng = np.random.default_rng(0)
for V in (29, 100, 512, 2304, 7001):
    for scale in (1.0, 4.0, 10.0):
        x = torch.from_numpy(rng.standard_normal((2000, V)).astype(np.float32) * scale)
        truth = torch.log_softmax(x.double(), -1)
        mine = eng.log_softmax(x.cuda()).cpu().double()
        tg = torch.log_softmax(x.cuda(), -1).cpu().double()
        tc = torch.log_softmax(x, -1).double()
        f = lambda a: ((a - truth).abs().max().item(), (a - truth).mean().item())
        print(f"V={V:5d} scale={scale:4.1f}  mine max/mean {f(mine)[0]:.2e} {f(mine)[1]:+.2e}   torch-cuda {f(tg)[0]:.2e} {f(tg)[1]:+.2e}   torch-cpu {f(tc)[0]:.2e} {f(tc)[1]:+.2e}")
