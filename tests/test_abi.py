"""CPU: the C-ABI library loads, exports every symbol include/wxb_align.h declares, and its
host-only helpers are exact.  No GPU work is issued here."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from whisperx_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "wxb_align.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wxb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in wxb_align.h but not exported"
    assert sorted(N.SYMBOLS) == names
    assert lib.wxb_abi_version() == 1


def test_struct_layouts_match_header():
    # field order/size sanity: pointers are 8 bytes, the structs have no surprises
    assert C.sizeof(N.Batch) == 8 + 4 + 4 + 4 + 4 + 8 * 7 + 8 + 8 * 3
    assert C.sizeof(N.Outputs) == 8 * 16
    assert C.sizeof(N.Limits) == 16


def test_round3_matches_python_round():
    lib = N.load()
    rnd = random.Random(1)
    for _ in range(100000):
        x = rnd.random() * rnd.choice([1, 10, 100, 3000, 1e6])
        assert lib.wxb_round3(x) == round(x, 3), x
        assert lib.wxb_round3(-x) == round(-x, 3), x
    for k in range(20000):
        for x in (k / 16, k / 2000, k / 1000, k * 0.001, k * 0.0005):
            assert lib.wxb_round3(x) == round(x, 3), x
    # timestamps as the reference forms them: frame*ratio + t1 (alignment.py:262)
    for _ in range(50000):
        T = rnd.randint(1, 3000); dur = rnd.random() * 30; t1 = round(rnd.random() * 3600, 3)
        ratio = dur * 1 / T
        f = rnd.randint(0, T)
        x = f * ratio + t1
        assert lib.wxb_round3(x) == round(x, 3)
    assert np.isnan(lib.wxb_round3(float("nan"))) and lib.wxb_round3(float("inf")) == float("inf")


def test_round3_np_matches_numpy_scalar_round():
    lib = N.load()
    rnd = random.Random(2)
    for _ in range(100000):
        k = rnd.randint(1, 30)
        m = np.float64(sum(rnd.randint(0, 1000) / 1000 for _ in range(k)) / k)
        assert lib.wxb_round3_np(float(m)) == round(m, 3)
    assert lib.wxb_round3_np(0.0005) == 0.0 and lib.wxb_round3(0.0005) == 0.001      # the two rules differ


def test_pairwise_sum_matches_numpy():
    lib = N.load()
    rng = np.random.default_rng(0)
    for n in list(range(0, 300)) + [1000, 5000]:
        a = np.round(rng.random(n), 3)
        got = lib.wxb_pairwise_sum(a.ctypes.data_as(N.f64p), n)
        assert got == np.sum(a), n


def test_tokens_per_lane_buckets():
    lib = N.load()
    shape = lambda n: (lib.wxb_segment_warps(n), lib.wxb_tokens_per_lane(n))
    assert lib.wxb_set_k_target(0) == 0                    # default: chosen per batch
    lib.wxb_set_k_target(6)
    assert shape(0) == (0, 0)
    # every full warp holds K = k_target columns per lane; the last warp of a segment what is left
    assert shape(1) == (1, 1) and shape(32) == (1, 1) and shape(33) == (1, 2) and shape(192) == (1, 6)
    assert shape(193) == (2, 6) and shape(384) == (2, 6) and shape(400) == (3, 6) and shape(2048) == (11, 6)
    assert shape(31 * 32 * 6) == (31, 6) and shape(31 * 32 * 6 + 1) == (27, 7)      # too long for 31 warps of 6: a K of its own
    assert shape(31 * 32 * 8) == (31, 8) and shape(31 * 32 * 8 + 1) == (-1, -1)
    assert lib.wxb_set_k_target(4) == 6
    assert shape(128) == (1, 4) and shape(129) == (2, 4) and shape(400) == (4, 4)
    assert lib.wxb_set_k_target(-1) == 4 and lib.wxb_set_k_target(0) == 0
    assert shape(128) == (1, 4) and shape(161) == (2, 4) and shape(400) == (4, 4)      # unforced: the helpers answer for K = 4


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to run instead of falling back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from whisperx_b200.engine import AlignEngine
    with pytest.raises(RuntimeError):
        AlignEngine("cuda:0")
    with pytest.raises(RuntimeError):
        AlignEngine("cpu")
    ctx = C.c_void_p()
    rc = N.load().wxb_create(0, C.byref(ctx))
    assert rc != 0 and not ctx.value
    assert len(N.load().wxb_last_error()) > 0
