"""H2D bandwidth from ordinary pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined)."""
import ctypes as C, sys, os
import numpy as np, torch
rt = C.CDLL("libcudart.so.12")
n = 413 * 1000 * 1000
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
def bw(ptr, label):
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), C.c_void_p(ptr), C.c_size_t(n), 1, C.c_void_p(s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), C.c_void_p(ptr), C.c_size_t(n), 1, C.c_void_p(s))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label}: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s", flush=True)
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
bw(pin.data_ptr(), "torch pinned")
for flags, name in ((0, "cudaHostAlloc default"), (4, "write-combined"), (1, "portable"), (5, "portable+WC")):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags))
    assert rc == 0, rc
    C.memset(p, 1, n)
    bw(p.value, name)
    rt.cudaFreeHost(p)
