"""GPU: the BASELINE configurations at their full sizes — the batches bench.py times — with size-independent
properties over every segment and a sampled comparison with the CPU oracle (oracle/align_oracle.c) on the
engine's own log-probs; plus a time-boxed slice of the random soak (scripts/fuzz_parity.py), contexts on two
devices in one process, and the sharded align() on two GPUs.  Frames bit-exact, scores <= 1e-6 relative."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import c_oracle
from whisperx_b200 import synth
from whisperx_b200.engine import BatchSpec

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCORE_RTOL = 1e-6


def check_batch(engine, meta, em_dev, n_sample, seed, planted):
    """Fused-logits run of a whole batch; properties on all segments, oracle on a sample."""
    S = len(meta.frame_len)
    spec = BatchSpec(frame_off=meta.frame_off, frame_len=meta.frame_len, tok_off=meta.tok_off, tokens=meta.tokens,
                     vocab=meta.vocab, blank_id=meta.blank, logits=True, seg_t1=np.zeros(S), seg_ratio=meta.durations / meta.frame_len)
    res = engine.align_batch(em_dev, spec, want=("logprobs",))
    lp_dev = res.logprobs.view(-1, meta.vocab)
    st = res.status.cpu().numpy()
    assert int(st.sum()) == 0
    cs, ce = res.char_start.cpu().numpy().astype(np.int64), res.char_end.cpu().numpy().astype(np.int64)
    sc = res.char_score.cpu().numpy()
    to = meta.tok_off
    seg_of = np.repeat(np.arange(S), np.diff(to))
    assert (ce > cs).all() and (cs >= 0).all() and (ce <= meta.frame_len[seg_of]).all()
    inner = np.ones(len(cs), bool); inner[to[1:] - 1] = False
    assert np.array_equal(cs[1:][inner[:-1]], ce[:-1][inner[:-1]])                 # gap-free, monotone paths
    assert np.array_equal(ce[to[1:] - 1], res.n_frames.cpu().numpy())
    assert ((sc > 0) & (sc <= 1.0 + 1e-12)).all()
    if planted:     # the planted path dominates: nearly every frame of the segment is on the path
        cover = (ce[to[1:] - 1] - cs[to[:-1]]) / meta.frame_len
        assert np.median(cover) > 0.9
    # timestamps: round(frame * ratio + 0, 3), Python rounding
    ratio = (meta.durations / meta.frame_len)[seg_of]
    k = np.random.default_rng(seed).choice(len(cs), 2000, replace=False)
    css = res.char_start_s.cpu().numpy()
    assert all(css[i] == round(int(cs[i]) * float(ratio[i]) + 0.0, 3) for i in k)
    order = np.argsort(-(meta.frame_len.astype(np.int64) * np.diff(to)))
    pick = {int(order[0]), int(order[-1])} | {int(x) for x in np.random.default_rng(seed).choice(S, n_sample, replace=False)}
    for s in sorted(pick):
        a, T = int(meta.frame_off[s]), int(meta.frame_len[s])
        tk = meta.tokens[to[s]:to[s + 1]]
        rc, ocs, oce, osc, ts = c_oracle.align_segment(lp_dev[a:a + T].cpu().numpy(), tk, meta.blank)
        assert rc == 0 and np.array_equal(ocs, cs[to[s]:to[s + 1]]) and np.array_equal(oce, ce[to[s]:to[s + 1]]), s
        np.testing.assert_allclose(sc[to[s]:to[s + 1]], osc, rtol=SCORE_RTOL, atol=0)
    return res


@pytest.mark.parametrize("kind", ["iid", "planted"])
def test_c2_fused_logits_at_size(engine, kind):
    """BASELINE configs[1] exactly as bench.py times it: 4,096 VAD-cut segments, raw logits, fused log-softmax."""
    b = synth.make_batch(4096, "c2", align_frames=4, kind=kind)
    check_batch(engine, b, torch.from_numpy(b.emissions).cuda(), 40, 1, kind == "planted")


def test_c4_long_transcripts_wide_vocab_at_size(engine):
    """BASELINE configs[3]: vocab 100 (blank = last id), 40-60 s segments, 500-2,000 tokens (up to 8 warps x 8 columns)."""
    b = synth.make_batch(256, "c4", align_frames=4)
    assert int(np.diff(b.tok_off).max()) > 1900 and b.vocab == 100 and b.blank == 99
    check_batch(engine, b, torch.from_numpy(b.emissions).cuda(), 10, 2, False)


@pytest.mark.parametrize("kind", ["iid", "planted"])
def test_c5_sweep_batch_at_size(engine, kind):
    """BASELINE configs[4]: one 4,096-segment batch of the 10,000-hour sweep, generated on the device."""
    em, meta = synth.device_sweep_batch(np.arange(7000, 7000 + 4096), "cuda:0", kind=kind)
    res = check_batch(engine, meta, em, 24, 3, kind == "planted")
    # the generator is a pure function of the segment id: another batch composition gives the same segments
    em2, meta2 = synth.device_sweep_batch(np.arange(7000 + 4000, 7000 + 4200), "cuda:0", kind=kind)
    rows = meta.frame_off[1] - meta.frame_off[0]
    assert torch.equal(em.view(4096, -1)[4000:4096], em2.view(200, -1)[:96]) and rows == meta2.frame_off[1]
    assert np.array_equal(meta.tokens.reshape(4096, -1)[4000:], meta2.tokens.reshape(200, -1)[:96])
    del res


def test_fuzz_slice(engine):
    """60 s of scripts/fuzz_parity.py (random vocabularies, blanks, ties, padding, wildcards, planner knobs,
    device and host entry points) against the C oracle."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import fuzz_parity
    batches, segs, cells = fuzz_parity.fuzz(60.0, 20261018, engine)
    assert batches >= 20 and segs >= 300


def test_contexts_on_two_devices_in_one_process():
    """Kernel attributes (dynamic shared memory opt-in) are per device: a V=100 batch (> 48 KB of shared memory)
    must run on the second device after the first one has been used."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from whisperx_b200.engine import get_engine
    b = synth.make_batch(24, "c4", align_frames=4)
    spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=b.vocab,
                     blank_id=b.blank, logits=True)
    outs = []
    for d in ("cuda:0", "cuda:1", "cuda:0"):
        eng = get_engine(d)
        outs.append(eng.align_batch(torch.from_numpy(b.emissions).to(d), spec).cpu())
    for o in outs[1:]:
        for f in ("status", "char_start", "char_end", "char_score"):
            assert torch.equal(getattr(outs[0], f).view(torch.uint8), getattr(o, f).view(torch.uint8)), f
    assert int(outs[0].status.sum()) == 0


def test_sharded_align_on_two_gpus_equals_single_gpu():
    """align_sharded over 2 ranks (torchrun, NCCL, LPT partition, host-side gather) == single-GPU align()."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "scripts", "run_sharded_align.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert "identical to 1 GPU" in p.stdout


def test_overlapped_runs_equal_serial_runs(engine):
    """wxb_plan_run_async keeps two runs in flight (two scratch sets, two groups of side streams): a stream of
    different batches pushed through it back to back must give, batch for batch, the bytes of the serial path."""
    from whisperx_b200.engine import _alloc_outputs
    specs, ems, serial = [], [], []
    for i, (cfg, n, kind) in enumerate((("c2", 700, "planted"), ("c5", 300, "iid"), ("c2", 64, "iid"), ("c4", 48, "iid"), ("c2", 1500, "iid"))):
        b = synth.make_batch(n, cfg, align_frames=4, kind=kind, first_segment=1000 * i)
        spec = BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens, vocab=b.vocab,
                         blank_id=b.blank, logits=True, seg_t1=np.zeros(n), seg_ratio=b.durations / b.frame_len)
        specs.append(spec)
        ems.append(torch.from_numpy(b.emissions).cuda())
        serial.append(engine.align_batch(ems[-1], spec).cpu())
        assert int(serial[-1].status.sum()) == 0
    fields = ("status", "n_frames", "char_start", "char_end", "char_score", "char_start_s", "char_end_s", "char_score_r")
    plans = [engine.plan(s) for s in specs]
    outs = [[_alloc_outputs(s, engine.device, ()) for s in specs] for _ in range(3)]
    order = [0, 1, 2, 3, 4, 4, 0, 3, 1, 2, 0, 0, 1, 4, 3]
    for rep in range(3):
        for o in outs[rep]:
            for f in ("status", "char_start", "char_end"):
                getattr(o, f).fill_(-9)
        for i in order:
            plans[i].run_overlapped(ems[i], outs[rep][i])
    engine.join()
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(len(specs)):
            got = outs[rep][i].cpu()
            for f in fields:
                assert torch.equal(getattr(serial[i], f).view(torch.uint8), getattr(got, f).view(torch.uint8)), (rep, i, f)
    # mixing the two entry points: a serial run in between waits for nothing it should not and corrupts nothing
    plans[0].run_overlapped(ems[0], outs[0][0])
    mid = plans[1].run(ems[1])
    plans[4].run_overlapped(ems[4], outs[0][4])
    engine.join()
    torch.cuda.synchronize()
    for i, got in ((0, outs[0][0].cpu()), (1, mid.cpu()), (4, outs[0][4].cpu())):
        for f in fields:
            assert torch.equal(getattr(serial[i], f).view(torch.uint8), getattr(got, f).view(torch.uint8)), (i, f)
    for p in plans:
        p.close()


@pytest.mark.parametrize("kind", ["iid", "planted"])
def test_fused_log_softmax_against_torch_log_softmax(engine, kind):
    """The DP is bit-exact given identical log-probs; the fused mode (WXB_F_LOGITS) computes its own log-probs, which
    differ from torch.log_softmax by a few ulp.  What a user of the fused mode sees, measured: the fraction of
    characters whose frames differ from the pipeline the reference runs (torch.log_softmax on the device, then the
    DP) — reported, and bounded at 0.1 % (measured on these inputs: 0; scripts/measure_divergence.py has the full
    sizes and the CPU log-softmax as well)."""
    b = synth.make_batch(1024, "c2", align_frames=4, kind=kind)
    raw = torch.from_numpy(b.emissions).cuda()
    spec = lambda logits: BatchSpec(frame_off=b.frame_off, frame_len=b.frame_len, tok_off=b.tok_off, tokens=b.tokens,
                                    vocab=b.vocab, blank_id=b.blank, logits=logits)
    fused = engine.align_batch(raw, spec(True))
    ref = engine.align_batch(torch.log_softmax(raw, -1), spec(False))
    err = float((engine.log_softmax(raw) - torch.log_softmax(raw, -1)).abs().max())
    diff = ((fused.char_start != ref.char_start) | (fused.char_end != ref.char_end)).float().mean().item()
    print(f"fused vs torch.log_softmax ({kind}): max |dlogp| {err:.2e}, characters with different frames {100 * diff:.4f} %")
    assert err <= 2e-5 and diff <= 1e-3 and torch.equal(fused.status, ref.status)
