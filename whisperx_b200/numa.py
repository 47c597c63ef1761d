"""Host-side placement for one-process-per-GPU runs: pin the process (and therefore the pinned host
buffers it allocates afterwards) to the CPU cores that are local to its GPU.  With 8 ranks each
streaming hundreds of MB per step over PCIe, buffers on the wrong socket halve the copy rate."""
from __future__ import annotations

import os
from typing import Optional, Set


def gpu_local_cpus(index: int) -> Optional[Set[int]]:
    """CPU ids NVML reports as local to GPU `index` (honours CUDA_VISIBLE_DEVICES order only when it
    is a plain list of indices), or None when NVML is unavailable."""
    try:
        import pynvml
    except Exception:
        return None
    try:
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                phys = int(ids[index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        return cpus or None
    except Exception:
        return None
    finally:
        try:
            pynvml.nvmlShutdown()
        except Exception:
            pass


def bind_process_to_gpu(index: int) -> Optional[Set[int]]:
    """Restrict this process to the cores local to GPU `index` (intersected with what it is allowed
    to use).  Returns the cpu set applied, or None if nothing was changed."""
    cpus = gpu_local_cpus(index)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return target
    except Exception:
        return None
