"""The C ABI from plain C: examples/align_host.c compiles as C99 against include/wxb_align.h, links to the
library without any CUDA header, fails loudly without a GPU, and on a GPU prints what the oracle computes."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from whisperx_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    N.load()                                   # builds the library if it is not there yet
    out = str(tmp_path_factory.mktemp("cex") / "align_host")
    libdir = os.path.join(ROOT, "whisperx_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "align_host.c"), "-L", libdir, "-lwxb_align", "-lm", f"-Wl,-rpath,{libdir}", "-o", out]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return out


def test_header_is_plain_c_and_the_program_fails_loudly_without_a_gpu(binary):
    p = subprocess.run([binary], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert p.returncode == 0, p.stderr
    else:
        assert p.returncode == 2 and "wxb_create" in p.stderr and len(p.stderr.strip()) > len("wxb_create:")


@pytest.mark.gpu
def test_c_program_matches_the_oracle(binary):
    from oracle import c_oracle
    raw = np.array([[2.0, 0.1, 0.0, -1.0], [0.2, 2.5, 0.0, -0.5], [0.1, 2.0, 0.3, 0.0], [1.5, 0.0, 0.2, 0.1],
                    [0.0, 0.1, 2.2, 0.0], [0.3, 0.0, 1.9, 0.4], [0.1, 0.0, 0.2, 2.4], [1.8, 0.0, 0.1, 0.6]], np.float32).astype(np.float64)
    m = raw.max(1, keepdims=True)
    E = (raw - m - np.log(np.exp(raw - m).sum(1, keepdims=True))).astype(np.float32)
    tok = np.array([1, 2, 3], np.int32)
    rc, cs, ce, sc, ts = c_oracle.align_segment(E, tok, 0)
    p = subprocess.run([binary], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.strip().splitlines()
    assert lines[0].split() == ["status", str(rc), "t_start", str(ts)]
    ratio = (13.3 - 12.5) / 8
    for k, line in enumerate(lines[1:]):
        f = line.split()
        assert int(f[0]) == tok[k] and int(f[1]) == cs[k] and int(f[2]) == ce[k]
        assert abs(float(f[3]) - sc[k]) <= 1e-6 * abs(sc[k])
        assert f[4] == f"{round(cs[k] * ratio + 12.5, 3):.3f}" and f[5] == f"{round(ce[k] * ratio + 12.5, 3):.3f}"
