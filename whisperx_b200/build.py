"""Builds libwxb_align.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m whisperx_b200.build [--force] [--verbose]

Two translation units (kernel instantiations, host side of the C ABI) compiled in parallel.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libwxb_align.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: the reference adds and compares float32 values; no contraction anywhere.
FLAGS = ARCH + ["-lineinfo", "-O3", "-fmad=false", "-std=c++17", "-Xcompiler", "-fPIC", "-I", INCLUDE]
if os.environ.get("WXB_EXTRA_NVCC_FLAGS"):            # experiments only (e.g. -DWXB_REGS(K)=64)
    FLAGS += os.environ["WXB_EXTRA_NVCC_FLAGS"].split()
if os.environ.get("WXB_LIB_OUT"):
    LIB = os.environ["WXB_LIB_OUT"]
    OBJ = LIB + ".objs"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libwxb_align.so cannot be built (there is no CPU fallback)")


def _sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "wxb_align.h"), __file__]
    return max(os.path.getmtime(p) for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _sources():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    inst_obj = os.path.join(OBJ, "wxb_inst.o")
    jobs.append((inst_obj, [nvcc] + FLAGS + ["-c", os.path.join(CSRC, "wxb_inst.cu"), "-o", inst_obj]))
    main_obj = os.path.join(OBJ, "wxb_align.o")
    jobs.append((main_obj, [nvcc] + FLAGS + ["-c", os.path.join(CSRC, "wxb_align.cu"), "-o", main_obj]))

    def run(job):
        obj, cmd = job
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {os.path.basename(obj)}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(run, jobs))
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
