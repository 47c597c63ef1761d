"""CPU: the synthetic workloads (whisperx_b200/synth.py) — shapes of the BASELINE configs and the counter-based
sweep generator (a segment's data must not depend on the batch it is generated in: that is what makes the
strong-scaling sweep process identical data at every world size)."""
import numpy as np
import torch

from oracle import c_oracle
from whisperx_b200 import synth


def test_frames_formula():
    assert [synth.frames_for(d) for d in (30.0, 5.0, 1.0, 0.025)] == [1499, 249, 49, 1]


def test_sweep_generator_is_a_function_of_the_segment_id():
    for kind in ("iid", "planted"):
        em, m = synth.device_sweep_batch(np.arange(100, 132), "cpu", kind=kind)
        em2, m2 = synth.device_sweep_batch(np.array([131, 7, 110]), "cpu", kind=kind)
        rows = int(m.frame_off[1])
        assert rows % 4 == 0 and int(m.frame_len[0]) == 1499 and m.tokens.size == 32 * 399
        a, b = em.view(32, rows, 29), em2.view(3, rows, 29)
        assert torch.equal(a[31], b[0]) and torch.equal(a[10], b[2])
        assert np.array_equal(m.tokens.reshape(32, -1)[31], m2.tokens.reshape(3, -1)[0])
        assert not torch.equal(a[0], a[1]) and float(a[:, 1499:].abs().max()) == 0.0
        tok = m.tokens.reshape(32, -1)
        assert tok.min() >= 1 and tok.max() <= 28 and (tok[:, 0] != 1).all() and (tok[:, -1] != 1).all()
        assert not ((tok[:, 1:] == 1) & (tok[:, :-1] == 1)).any()
    # planted: the path the generator planted is (nearly) the path the reference DP finds
    lp = torch.log_softmax(em.view(32, rows, 29)[0, :1499], -1).numpy()
    rc, cs, ce, sc, ts = c_oracle.align_segment(lp, m.tokens[:399], 0)
    assert rc == 0 and ts >= 1400 and np.median(sc) > 0.5
    assert abs(float(em.view(32, rows, 29)[:, :1499].std()) - 3.0) < 0.5


def test_c4_device_generator_shapes():
    em, m = synth.device_c4_batch(np.arange(40, 56), "cpu")
    n = np.diff(m.tok_off)
    rows = int(m.frame_off[1])
    assert m.vocab == 100 and m.blank == 99 and rows % 4 == 0 and rows >= int(m.frame_len.max())
    assert (n >= 500).all() and (n <= 2000).all() and (n <= m.frame_len).all()
    assert (m.frame_len >= synth.frames_for(40.0)).all() and (m.frame_len <= synth.frames_for(60.0)).all()
    assert m.tokens.min() >= 0 and m.tokens.max() <= 98 and (m.tokens[m.tok_off[1:] - 1] != 4).all()
    e = em.view(16, rows, 100)
    assert float(e[3, int(m.frame_len[3]):].abs().max()) == 0.0
    em2, m2 = synth.device_c4_batch(np.array([45, 41]), "cpu")
    assert torch.equal(em2.view(2, rows, 100)[0], e[5]) and int(m2.frame_len[1]) == int(m.frame_len[1])


def test_make_ref_copy_is_byte_identical():
    """oracle/make_ref.py: the build-time copy of the reference modules (git-ignored) equals the checkout."""
    import hashlib
    import json
    import os
    from oracle import make_ref, ref_import
    if not os.path.isdir("/root/reference/whisperx"):
        import pytest
        pytest.skip("/root/reference is not present on this machine")
    assert make_ref.make()
    man = json.load(open(os.path.join(make_ref.DST, "MANIFEST.json")))
    for name, digest in man["sha256"].items():
        assert hashlib.sha256(open(os.path.join("/root/reference/whisperx", name), "rb").read()).hexdigest() == digest
    assert ref_import.reference_available()
