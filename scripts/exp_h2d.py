"""Development aid: raw pinned host<->device copy bandwidth on this box (the floor of the e2e leg)."""
import torch, time
n = 412_579_056 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10): d.copy_(h, non_blocking=True)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
print(f"H2D {n*4/1e6:.0f} MB: {ms:.2f} ms -> {n*4/ms/1e6:.1f} GB/s")
h2 = torch.empty(37_844_448 // 4, dtype=torch.float32).pin_memory()
ev0.record()
for _ in range(10): h2.copy_(d[: h2.numel()], non_blocking=True)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
print(f"D2H {h2.numel()*4/1e6:.0f} MB: {ms:.2f} ms -> {h2.numel()*4/ms/1e6:.1f} GB/s")
