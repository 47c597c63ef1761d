#!/usr/bin/env python
"""Benchmark of the forced-alignment hot path (BASELINE.json metric: aligned audio-hours/s and
trellis cells/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (fused log-softmax -> trellis -> backtrack -> merge ->
timestamps) over one batch of synthetic emissions: BASELINE.json configs[1], 4,096 VAD-cut segments
of 5-30 s at 50 frames/s, vocabulary 29 — per GPU (weak scaling: every rank aligns its own,
differently seeded, 4,096 segments; segments are independent so there is no collective on the
data path, only a barrier and a max/sum of scalars for the report).

`value`   : whole-job audio-hours/s with the emissions already resident in HBM (CUDA events).
`e2e`     : the same metric through the host-buffer C-ABI call wxb_align_host(): planning, the
            host->device copy of the emissions from pinned memory, the kernels and the
            device->host copy of the per-character results are all inside the timed region;
            `e2e.link_floor_ms` is a bare copy of the same bytes timed in the same run on every rank at once.
`configs` : (N=1) the other BASELINE configs, each device-timed and each with `parity_checked` = the number
            of segments of the *timed* batch compared with the CPU oracle inside this run.
`strong`  : BASELINE configs[4], the 10,000-audio-hour sweep, partitioned over the ranks by
            whisperx_b200.shard.partition_lpt (strong scaling: the pool is fixed, every rank streams its share).
`--impl reference` times the reference's own CPU implementation of the path (the unmodified
whisperx/alignment.py functions from oracle/_ref — a build-time copy, /root/reference does not exist on the
GPU box — or, if that copy is missing, the torch port in oracle/ref_port.py) on all host cores, on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOAD = "c2"
SEGMENTS_PER_GPU = 4096
METRIC = "aligned audio-hours/sec"
UNIT = "audio-h/s"


# ----------------------------------------------------------------------------- CPU reference arm
_REF = {"mod": None, "kind": "port"}


def _load_cpu_reference():
    """The reference's own functions when the build-time copy (or /root/reference) is importable, else the port."""
    if _REF["mod"] is None:
        from oracle import ref_import
        if ref_import.reference_available():
            _REF["mod"] = ref_import.load_reference_alignment()
            _REF["kind"] = "reference"
        else:
            from oracle import ref_port
            _REF["mod"] = ref_port
            _REF["kind"] = "port"
    return _REF["mod"], _REF["kind"]


def _cpu_worker(args):
    """One segment through get_trellis + backtrack + merge_repeats (whisperx/alignment.py:359-454)."""
    import torch
    torch.set_num_threads(1)
    mod, kind = _load_cpu_reference()
    E, tok, blank = args
    E = torch.from_numpy(E)
    tokens = tok.tolist()
    t0 = time.perf_counter()
    if kind == "reference":
        trellis = mod.get_trellis(E, tokens, blank)
        path = mod.backtrack(trellis, E, tokens, blank)
        ok = path is not None
        if ok:
            mod.merge_repeats(path, "x" * len(tokens))
    else:
        segs, _ = mod.align_tokens(E, tokens, blank, "x" * len(tokens))
        ok = segs is not None
    return time.perf_counter() - t0, ok


def _host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_pass(batch, n_sample: int, pool, logp=None):
    """Align the first n_sample segments of `batch` with the CPU reference on all cores.
    Returns (wall_s, audio_seconds, cells)."""
    import torch
    if logp is None:
        logp = torch.log_softmax(torch.from_numpy(batch.emissions), -1).numpy()
    jobs = []
    cells = 0
    for s in range(n_sample):
        a, T = int(batch.frame_off[s]), int(batch.frame_len[s])
        tk = batch.tokens[batch.tok_off[s]:batch.tok_off[s + 1]]
        jobs.append((logp[a:a + T], tk, batch.blank))
        cells += T * len(tk)
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    assert all(ok for _, ok in res)
    return wall, float(batch.durations[:n_sample].sum()), cells


def c_port_pass(batch, n_sample: int, logp):
    """Single-thread scalar C port (oracle/align_oracle.c) on the same sample, for context."""
    from oracle import c_oracle
    c_oracle.lib()
    # contiguous view: the C batch driver takes T from consecutive offsets, so copy out the sample
    em, fo2 = [], [0]
    for s in range(n_sample):
        a, T = int(batch.frame_off[s]), int(batch.frame_len[s])
        em.append(logp[a:a + T]); fo2.append(fo2[-1] + T)
    E = np.concatenate(em)
    to = batch.tok_off[:n_sample + 1]
    t0 = time.perf_counter()
    c_oracle.align_batch(E, np.asarray(fo2, np.int64), batch.tokens[:to[-1]], to, batch.vocab, batch.blank)
    return time.perf_counter() - t0


_KIND_TEXT = {"reference": "the reference's own get_trellis/backtrack/merge_repeats (unmodified whisperx/alignment.py, "
                           "imported from the build-time copy oracle/_ref)",
              "port": "torch port of the reference (oracle/ref_port.py; the oracle/_ref copy is absent)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    import torch
    from whisperx_b200 import synth
    _, kind = _load_cpu_reference()            # before the fork: workers inherit the imported module
    cores = _host_cores()
    n_sample = max(8 * cores, 64)
    batch = synth.make_batch(n_sample, WORKLOAD, align_frames=4)
    logp = torch.log_softmax(torch.from_numpy(batch.emissions), -1).numpy()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_reference_pass(batch, n_sample, pool, logp)
        t0 = time.perf_counter()
        audio = cells = 0.0
        for _ in range(args.steps):
            _, a, c = cpu_reference_pass(batch, n_sample, pool, logp)
            audio += a; cells += c
        wall = time.perf_counter() - t0
    value = audio / 3600.0 / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "cells_per_s": cells / wall, "cells_per_s_per_core": cells / wall / cores,
        "config": {"workload": f"BASELINE configs[1] sample: first {n_sample} of the 4,096 VAD-cut 5-30 s segments, vocab 29, "
                               "log-prob emissions, get_trellis+backtrack+merge_repeats per segment"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n_sample} segments per step x {args.steps} steps, {_KIND_TEXT[kind]}, "
                                   f"one process per core, torch threads=1"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per step from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


# ----------------------------------------------------------------------------- in-run oracle parity
def parity_sample(eng, emissions_dev, meta, res, n: int, logits: bool, seed: int = 0) -> int:
    """Compare `n` segments of a batch that has just been timed with the CPU oracle (oracle/align_oracle.c),
    outside every timed region.  The sampled segments' emission rows are copied into a small batch that goes
    through the same kernels with the `logprobs` output on (per-row arithmetic: the same bits as in the big
    batch); the oracle DP runs on exactly those log-probs.  Checked, per segment: status; every character's
    first / last+1 frame (exact) and score (<= 1e-6 relative) — of the small batch AND of the timed batch `res`.
    Returns the number of segments compared; raises AssertionError on the first difference."""
    import torch
    from oracle import c_oracle
    from whisperx_b200.engine import BatchSpec
    S = len(meta.frame_len)
    rng = np.random.default_rng(seed)
    order = np.argsort(-(meta.frame_len.astype(np.int64) * np.diff(meta.tok_off)))
    pick = sorted({int(order[0]), int(order[-1])} | {int(x) for x in rng.choice(S, min(max(n - 2, 0), S), replace=False)})[:max(n, 1)]
    V = meta.vocab
    rows, fo, fl, to, toks = [], [0], [], [0], []
    for s in pick:
        a, T = int(meta.frame_off[s]), int(meta.frame_len[s])
        Tp = T + (-T) % 4
        rows.append(torch.arange(a, a + Tp, device=emissions_dev.device).clamp_(max=emissions_dev.shape[0] - 1))
        fo.append(fo[-1] + Tp); fl.append(T)
        tk = meta.tokens[meta.tok_off[s]:meta.tok_off[s + 1]]
        toks.append(tk); to.append(to[-1] + len(tk))
    sub = emissions_dev.view(-1, V)[torch.cat(rows)].contiguous()
    spec = BatchSpec(frame_off=np.asarray(fo, np.int64), frame_len=np.asarray(fl, np.int32), tok_off=np.asarray(to, np.int64),
                     tokens=np.concatenate(toks), vocab=V, blank_id=meta.blank, logits=logits)
    small = eng.align_batch(sub, spec, want=("logprobs",) if logits else ()).cpu()
    lp = (small.logprobs if logits else sub.cpu()).view(-1, V).numpy()
    big_status = res.status.cpu().numpy()
    big_cs, big_ce, big_sc = res.char_start.cpu().numpy(), res.char_end.cpu().numpy(), res.char_score.cpu().numpy()
    for i, s in enumerate(pick):
        T = fl[i]
        rc, cs, ce, sc, _ = c_oracle.align_segment(lp[fo[i]:fo[i] + T], toks[i], meta.blank)
        sl_small = slice(to[i], to[i + 1])
        sl_big = slice(int(meta.tok_off[s]), int(meta.tok_off[s + 1]))
        for name, st, c0, c1, sco in (("sub-batch", int(small.status[i]), small.char_start[sl_small].numpy(), small.char_end[sl_small].numpy(),
                                       small.char_score[sl_small].numpy()),
                                      ("timed batch", int(big_status[s]), big_cs[sl_big], big_ce[sl_big], big_sc[sl_big])):
            assert st == rc, f"parity: segment {s} ({name}) status {st} != oracle {rc}"
            if rc != 0:
                continue
            assert np.array_equal(c0, cs) and np.array_equal(c1, ce), f"parity: segment {s} ({name}) frames differ from the oracle"
            np.testing.assert_allclose(sco, sc, rtol=1e-6, atol=0, err_msg=f"parity: segment {s} ({name}) scores")
    return len(pick)


# ----------------------------------------------------------------------------- our arm
def _spec_of(meta, logits=True):
    from whisperx_b200.engine import BatchSpec
    n = len(meta.frame_len)
    return BatchSpec(frame_off=meta.frame_off, frame_len=meta.frame_len, tok_off=meta.tok_off, tokens=meta.tokens,
                     vocab=meta.vocab, blank_id=meta.blank, logits=logits,
                     seg_t1=np.zeros(n), seg_ratio=meta.durations / meta.frame_len)


def time_device_batch(eng, dev, meta, em_dev, steps, warmup, peak, n_parity, logits=True):
    """K device-timed passes over one resident batch (the overlapped serving loop of `value`: two steps in
    flight) + an in-run parity sample of that batch."""
    import torch
    from whisperx_b200.engine import _alloc_outputs
    spec = _spec_of(meta, logits)
    plan = eng.plan(spec)
    outs = [_alloc_outputs(spec, dev, ()) for _ in range(2)]
    for i in range(warmup):
        plan.run_overlapped(em_dev, outs[i % 2])
    eng.join()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        plan.run_overlapped(em_dev, outs[i % 2])
    eng.join()
    ev1.record()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1) / steps
    out = outs[(steps - 1) % 2]
    checked = parity_sample(eng, em_dev, meta, out, n_parity, logits) if n_parity else 0
    rec = {"segments": len(meta.frame_len), "audio_hours": meta.audio_seconds / 3600.0, "cells": int(plan.cells),
           "steps": steps, "ms_per_step": ms, "value": meta.audio_seconds / 3600.0 / (ms * 1e-3), "unit": UNIT,
           "cells_per_s": plan.cells / (ms * 1e-3), "hbm_frac": plan.alg_bytes / (ms * 1e-3) / 1e9 / peak,
           "algorithmic_bytes": int(plan.alg_bytes), "failed_segments": int(out.status.sum().item()),
           "gpu_launches_per_step": int(plan.launches), "parity_checked": checked}
    plan.close()
    return rec


def align_api_configs(dev, steps):
    """BASELINE configs[0] and configs[2] through the drop-in whisperx_b200.align() with a random-init
    WAV2VEC2_ASR_BASE_960H: one 30 s segment (latency) and a 1-hour recording cut into 30 s segments with
    char-level alignment and wildcard tokens.  The unchanged PyTorch wav2vec2 forward is inside these times.
    parity_checked = input segments whose output records equal the oracle port's (oracle/ref_port.py run on the
    same model / device, its CPU DP and pandas tables), compared exactly — scores included."""
    import copy
    import torch
    import torchaudio
    from oracle import ref_import, ref_port
    from whisperx_b200 import alignment as wxa

    labels = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H.get_labels()
    dictionary = {c.lower(): i for i, c in enumerate(labels)}
    meta = {"language": "en", "dictionary": dictionary, "type": "torchaudio"}
    torch.manual_seed(0)
    model = torchaudio.models.wav2vec2_model(**torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params).eval().to(dev)
    n_seg = 120
    g = torch.Generator().manual_seed(1)
    audio = (torch.randn(n_seg * 30 * 16000, generator=g) * 0.1).numpy()
    rng = np.random.default_rng(2)
    letters = list("abcdefghijklmnopqrstuvwxyz'")
    segs = []
    for i in range(n_seg):
        words = []
        while sum(len(w) + 1 for w in words) < 400:
            words.append("".join(rng.choice(letters, int(rng.integers(1, 9)))) if rng.random() > 0.04 else str(int(rng.integers(0, 2030))))
        text = " ".join(words)
        segs.append({"start": 30.0 * i, "end": 30.0 * (i + 1), "text": text[:200] + ". " + text[200:].capitalize()})
    spans = wxa.assemble._fallback_spans
    wxa.set_sentence_splitter(spans)
    recs = []
    try:
        def same(a, b):
            return json.dumps(a, sort_keys=True, default=float) == json.dumps(b, sort_keys=True, default=float)

        def timed(transcript, wav, reps, **kw):
            best = []
            res = None
            for _ in range(reps):
                tr = copy.deepcopy(transcript)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                res = wxa.align(tr, model, meta, wav, str(dev), **kw)
                torch.cuda.synchronize(dev)
                best.append(time.perf_counter() - t0)
            return float(np.median(best)), res

        # ---- configs[0]: one 30 s segment
        one, wav1 = segs[:1], audio[: 30 * 16000]
        timed(one, wav1, 2)
        s1, res1 = timed(one, wav1, max(3, min(steps, 10)))
        ref1 = ref_port.align_port(copy.deepcopy(one), model, meta, wav1, str(dev), span_fn=spans)
        assert same(res1["segments"], ref1["segments"]), "c1: align() records differ from the oracle port"
        cpu_s = None
        if ref_import.reference_available():          # the reference's own align() on the host cores (its CPU-runnable case)
            A = ref_import.load_reference_alignment()
            ref_import.set_sentence_splitter(spans)
            cpu_model = copy.deepcopy(model).to("cpu")
            t0 = time.perf_counter()
            A.align(copy.deepcopy(one), cpu_model, meta, wav1, "cpu")
            cpu_s = time.perf_counter() - t0
        n_chars = sum(len(r.get("words", [])) for r in res1["segments"])
        recs.append({"id": "c1", "workload": "BASELINE configs[0]: whisperx_b200.align() on one 30 s segment, ~400-char transcript, "
                                             "random-init WAV2VEC2_ASR_BASE_960H (wav2vec2 forward included), emission mode torch, exact scores",
                     "segments": 1, "seconds": s1, "ms_per_step": s1 * 1e3, "value": 30.0 / 3600.0 / s1, "unit": UNIT, "words": n_chars,
                     "reference_cpu_align_seconds": cpu_s, "reference_cpu_cores": _host_cores(), "parity_checked": 1})
        # ---- configs[2]: one hour, char-level alignment, wildcard tokens (extension, off by default)
        wxa.WILDCARD = True
        timed(segs[:4], audio[: 4 * 30 * 16000], 1, return_char_alignments=True)
        s3, res3 = timed(segs, audio, 3, return_char_alignments=True)            # median of three whole-hour calls
        k = 3
        ref3 = ref_port.align_port(copy.deepcopy(segs[:k]), model, meta, audio[: k * 30 * 16000], str(dev), return_char_alignments=True,
                                   span_fn=spans, wildcard=True)
        assert same(res3["segments"][:len(ref3["segments"])], ref3["segments"]), "c3: align() records differ from the oracle port"
        recs.append({"id": "c3", "workload": "BASELINE configs[2]: whisperx_b200.align() on a 1-hour recording, 120 x 30 s segments, ~400 chars each "
                                             "(4 % of the words are numbers -> wildcard tokens, extension), return_char_alignments=True, "
                                             "random-init WAV2VEC2_ASR_BASE_960H (batch-1 fp32 wav2vec2 forwards included), exact scores",
                     "segments": n_seg, "seconds": s3, "ms_per_step": s3 * 1e3, "value": 1.0 / s3, "unit": UNIT,
                     "segments_out": len(res3["segments"]), "words": len(res3["word_segments"]), "parity_checked": k})
    finally:
        wxa.WILDCARD = False
        wxa.set_sentence_splitter(None)
    return recs


def strong_sweep(eng, dev, rank, world, hours, kind, barrier, peak, group=16):
    """BASELINE configs[4]: a fixed pool of 30 s segments (10,000 audio-hours = 1.2 M segments by default),
    partitioned over the ranks with whisperx_b200.shard.partition_lpt; every rank generates its segments on its
    GPU (counter-based generator: the data of a segment does not depend on the rank or world size), aligns them
    batch by batch and keeps an order-independent checksum of every character's frames.
    The rank's batches go through in groups of `group`: the group's emissions are generated and its plans made
    (the data source and the host-side planning: untimed, as for `value`), then the group is aligned back to back
    through the overlapped serving loop between two CUDA events (emissions resident when the clock starts).
    Timed = the sum over groups.  Returns this rank's figures; the caller reduces them over ranks."""
    import torch
    from whisperx_b200 import shard, synth
    from whisperx_b200.engine import _alloc_outputs
    n_pool = int(round(hours * 120))
    T = synth.frames_for(30.0)
    N = max(int(round(synth.CHARS_PER_SECOND * 30.0)), 1)
    t0 = time.perf_counter()
    costs = np.full(n_pool, float(T) * N)
    parts = shard.partition_lpt(costs, world)
    mine = np.asarray(parts[rank], np.int64)
    partition_s = time.perf_counter() - t0
    loads = [len(p) * float(T) * N for p in parts]
    n_batches = max(1, -(-len(mine) // SEGMENTS_PER_GPU))
    per = -(-len(mine) // n_batches)
    chunks = [mine[i:i + per] for i in range(0, len(mine), per)]
    evs = []
    checksum = torch.zeros((), dtype=torch.int64, device=dev)
    failed = torch.zeros((), dtype=torch.int64, device=dev)
    launches = checked = cells = 0
    outs = []
    # untimed warm-up on the first batch, twice (both run slots: scratch allocation, kernel attributes, module load)
    em, meta = synth.device_sweep_batch(chunks[0], dev, kind=kind)
    spec = _spec_of(meta, True)
    wplan = eng.plan(spec)
    outs.append(_alloc_outputs(spec, dev, ()))
    for _ in range(2):
        wplan.run_overlapped(em, outs[0])
    eng.join()
    torch.cuda.synchronize(dev)
    wplan.close()
    del em
    barrier()
    w0 = time.perf_counter()
    for g0 in range(0, len(chunks), group):
        work = []
        for ids in chunks[g0:g0 + group]:
            em, meta = synth.device_sweep_batch(ids, dev, kind=kind)
            spec = _spec_of(meta, True)
            work.append((em, meta, spec, eng.plan(spec)))
        while len(outs) < len(work):
            outs.append(_alloc_outputs(work[0][2], dev, ()))
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i, (em, meta, spec, plan) in enumerate(work):
            plan.run_overlapped(em, outs[i])
        eng.join()
        e1.record()
        evs.append((e0, e1))
        for i, (em, meta, spec, plan) in enumerate(work):
            o = outs[i]
            checksum += o.char_start[:spec.sum_N].sum(dtype=torch.int64) * 3 + o.char_end[:spec.sum_N].sum(dtype=torch.int64)
            failed += o.status[:spec.n_segments].sum(dtype=torch.int64)
            launches += plan.launches
            cells += plan.cells
        torch.cuda.synchronize(dev)
        if g0 == 0 or g0 + group >= len(chunks):
            em, meta, spec, plan = work[-1]
            checked += parity_sample(eng, em, meta, outs[len(work) - 1], 4, True, seed=g0)
        for _, _, _, plan in work:
            plan.close()
        del work
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - w0
    group_ms = [a.elapsed_time(b) for a, b in evs]
    align_ms = sum(group_ms)
    print(f"[bench] rank {rank}: strong sweep group times (ms): " + " ".join(f"{g:.1f}" for g in group_ms), file=sys.stderr, flush=True)
    return {"align_ms": align_ms, "wall_s": wall, "segments": len(mine), "batches": len(chunks), "cells": cells,
            "checksum": int(checksum.item()), "failed": int(failed.item()), "launches": launches, "parity_checked": checked,
            "partition_s": partition_s, "lpt_imbalance": max(loads) / (sum(loads) / len(loads)) - 1.0, "n_pool": n_pool}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    from whisperx_b200 import synth
    nseg = args.segments
    batch = synth.make_batch(nseg, args.workload, align_frames=4, first_segment=rank * nseg)

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised, so that fork is safe
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        import torch
        _, kind = _load_cpu_reference()
        cores = _host_cores()
        n_sample = min(max(16 * cores, 256), nseg)          # ~10-30 core-seconds of the Python reference path
        logp_s = torch.log_softmax(torch.from_numpy(batch.emissions[: int(batch.frame_off[n_sample])]), -1).numpy()
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_reference_pass(batch, min(cores, n_sample), pool, logp_s)           # warm the workers
            wall, audio, cells = cpu_reference_pass(batch, n_sample, pool, logp_s)
        c_wall = c_port_pass(batch, n_sample, logp_s)
        cpu_baseline = {"value": audio / 3600.0 / wall, "unit": UNIT, "cores": cores, "kind": kind,
                        "cells_per_s": cells / wall, "cells_per_s_per_core": cells / wall / cores,
                        "sample": f"first {n_sample} of the {nseg} segments ({cells/1e6:.0f} Mcells, {wall:.1f} s wall), "
                                  f"{_KIND_TEXT[kind]}, one process per core",
                        "c_port_1thread": {"value": audio / 3600.0 / c_wall, "unit": UNIT, "cells_per_s": cells / c_wall,
                                           "note": "scalar C restatement (oracle/align_oracle.c), 1 thread, same sample"}}

    import torch
    import torch.distributed as dist
    # one process per GPU: run on (and allocate the pinned staging buffers from) the GPU's own socket
    from whisperx_b200 import numa
    bound = numa.bind_process_to_gpu(local) if not args.no_numa_bind else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from whisperx_b200.engine import AlignEngine, _alloc_outputs
    eng = AlignEngine(dev)
    spec = _spec_of(batch, True)
    host_em = torch.from_numpy(batch.emissions).pin_memory()
    dev_em = host_em.to(dev, non_blocking=True)
    plan = eng.plan(spec)
    out = _alloc_outputs(spec, dev, ())
    host_out = _alloc_outputs(spec, "cpu", (), pin=True)
    # what comes back to the host in the e2e leg: per segment its status, per character start / end in seconds and the
    # score, as align() consumes them — 24 B per character (the frame-valued and unrounded twins stay on the device)
    D2H_FIELDS = ("status", "char_start_s", "char_end_s", "char_score_r")
    for f in ("n_frames", "char_start", "char_end", "char_score"):
        setattr(host_out, f, None)
    peak, peak_src = measured_peak_gbs()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- resident-in-HBM timing (value) ----
    # A serving loop: step after step over resident batches through wxb_plan_run_async — up to two steps in flight
    # (two scratch sets, two groups of side streams), so the tail of a step, when its last long segments leave most
    # SMs idle, is filled by the head of the next.  Every step is a complete pass; K steps are bracketed by CUDA
    # events and a join.  `value_serial` is the same loop with one step in flight at a time (wxb_plan_run).
    out2 = _alloc_outputs(spec, dev, ())
    for i in range(args.warmup):
        plan.run_overlapped(dev_em, out if i % 2 == 0 else out2)
    eng.join()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        plan.run_overlapped(dev_em, out if i % 2 == 0 else out2)
    eng.join()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    for _ in range(args.warmup):
        plan.run(dev_em, out=out)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        plan.run(dev_em, out=out)
    ev1.record()
    barrier()
    ms_serial = ev0.elapsed_time(ev1) / args.steps
    assert torch.equal(out.char_start, out2.char_start) and torch.equal(out.char_score, out2.char_score)
    # dominant kernel alone (all align_kernel launches of one step run concurrently on side streams;
    # the library brackets them with CUDA events on the launching stream)
    dp_ms = []
    for _ in range(args.steps):
        plan.run(dev_em, out=out, timed=True)
        dp_ms.append(out.extra["dp_ms"])
    torch.cuda.synchronize(dev)
    failed = int(out.status.sum().item())
    parity_checked = parity_sample(eng, dev_em, batch, out, args.parity, True)       # the timed batch, after the timed region

    # ---- the same workload on planted (realistic, full-length paths) emissions ----
    planted = None
    if not args.no_planted:
        pb = synth.make_batch(nseg, args.workload, align_frames=4, first_segment=rank * nseg, kind="planted")
        pem = torch.from_numpy(pb.emissions).to(dev)
        planted = time_device_batch(eng, dev, pb, pem, args.steps, args.warmup, peak, args.parity)
        del pem

    # ---- end to end through the host-buffer C-ABI call ----
    e2e_s = float("nan")
    floor = [float("nan"), float("nan")]
    if not args.no_e2e:
        # link floor, measured in this run: a bare copy of the same pinned bytes, every rank at once
        d2h_fields = D2H_FIELDS
        scratch = torch.empty_like(dev_em)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 10))
        scratch.copy_(host_em, non_blocking=True)
        barrier()
        f0.record()
        for _ in range(reps):
            scratch.copy_(host_em, non_blocking=True)
        f1.record()
        barrier()
        floor[0] = f0.elapsed_time(f1) / reps
        f0.record()
        for _ in range(reps):
            for f in d2h_fields:
                getattr(host_out, f).copy_(getattr(out, f), non_blocking=True)
        f1.record()
        barrier()
        floor[1] = f0.elapsed_time(f1) / reps
        del scratch
        for _ in range(max(args.warmup, 1)):
            eng.align_batch_host(host_em, spec, out=host_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.align_batch_host(host_em, spec, out=host_out)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        assert np.array_equal(host_out.char_start_s.numpy(), out.char_start_s.cpu().numpy()), "e2e results differ from the resident run"
        assert np.array_equal(host_out.char_score_r.numpy(), out.char_score_r.cpu().numpy()) and int(host_out.status.sum()) == failed
    clocks = sampler.stop()

    # ---- the other BASELINE configs (N=1 only: they are single-GPU cases) ----
    configs = []
    if world == 1 and not args.no_configs:
        recs = align_api_configs(dev, args.steps)
        em4, c4 = synth.device_c4_batch(np.arange(args.c4_segments), dev)
        rec4 = time_device_batch(eng, dev, c4, em4, max(3, args.steps // 2), 3, peak, 4)
        rec4.update(id="c4", workload=f"BASELINE configs[3]: {args.c4_segments:,} segments of 40-60 s, vocab 100 (blank = last id), "
                                       "500-2,000-char transcripts, raw logits in (log_softmax_kernel, then the recurrence on log-probs), "
                                       "iid N(0,9) logits generated on the device")
        del em4
        em5, m5 = synth.device_sweep_batch(np.arange(nseg), dev, kind="iid")
        rec5 = time_device_batch(eng, dev, m5, em5, args.steps, 3, peak, args.parity)
        rec5.update(id="c5", workload=f"BASELINE configs[4] batch shape: {nseg:,} segments of 30 s, vocab 29, 399-char transcripts, raw logits in "
                                       "(fused log-softmax), iid N(0,9) logits generated on the device")
        del em5
        configs = [recs[0], {"id": "c2", "see": "top level (value / value_serial / value_planted / e2e)"}, recs[1], rec4, rec5]

    # ---- BASELINE configs[4]: strong scaling over a fixed pool, through the sharder ----
    strong = None
    if not args.no_strong and args.sweep_hours > 0:
        strong = strong_sweep(eng, dev, rank, world, args.sweep_hours, args.sweep_kind, barrier, peak)

    t = torch.tensor([ms_total, e2e_s * 1e3, floor[0], floor[1], planted["ms_per_step"] if planted else 0.0,
                      strong["align_ms"] if strong else 0.0, strong["wall_s"] if strong else 0.0, ms_serial], device=dev, dtype=torch.float64)
    units = torch.tensor([batch.audio_seconds, float(plan.cells), float(planted["parity_checked"]) if planted else 0.0,
                          float(strong["cells"]) if strong else 0.0, float(strong["segments"]) if strong else 0.0,
                          float(strong["parity_checked"]) if strong else 0.0, float(strong["failed"]) if strong else 0.0,
                          float(strong["launches"]) if strong else 0.0, float(parity_checked)], device=dev, dtype=torch.float64)
    csum = torch.tensor([strong["checksum"] if strong else 0], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        dist.all_reduce(csum, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, floor_h2d, floor_d2h, planted_ms, strong_ms, strong_wall, ms_serial_max = (float(x) for x in t)
    audio_all, cells_all = float(units[0]), float(units[1])

    if rank == 0:
        ms_step = ms_total / args.steps
        dp = float(np.mean(dp_ms))
        # the chain's average duration over the timed region of `value` (K overlapped steps between two CUDA events)
        achieved = plan.alg_bytes / (ms_step * 1e-3) / 1e9
        h2d = host_em.numel() * 4 + batch.tokens.nbytes + batch.frame_off.nbytes + batch.tok_off.nbytes + batch.frame_len.nbytes
        d2h = sum(getattr(host_out, f).numel() * getattr(host_out, f).element_size() for f in D2H_FIELDS)
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": audio_all / 3600.0 / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "cells_per_s": cells_all / (ms_step * 1e-3),
            "value_serial": {"value": audio_all / 3600.0 / (ms_serial_max * 1e-3), "ms_per_step": ms_serial_max,
                             "note": "one step in flight at a time (wxb_plan_run, CUDA-graph replay); `value` keeps two in flight (wxb_plan_run_async)"},
            "parity_checked": int(units[8]),
            "config": {"workload": WORKLOAD_TEXT[args.workload].format(n=nseg),
                       "segments_per_gpu": nseg, "workload_id": args.workload, "audio_hours_per_gpu": batch.audio_seconds / 3600.0,
                       "cells_per_gpu": int(plan.cells), "emission_bytes_per_gpu": int(host_em.numel() * 4),
                       "l2": f"inputs ({host_em.numel() * 4 / 1e6:.0f} MB per step) larger than L2 (126 MB); no explicit flush",
                       "failed_segments": failed,
                       "parity": "parity_checked = segments of the timed batch compared with the CPU oracle (oracle/align_oracle.c) on the engine's own "
                                 "log-probs, inside this run, outside the timed region: frames exact, scores <= 1e-6 relative",
                       "host_cpus_bound_to_gpu_socket": (len(bound) if bound else None)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic.get("dram_bytes_per_step"), "traffic_whole_chain": traffic.get("dram_bytes_per_step_whole_chain"),
                         "peak_source": peak_src,
                         "kernel": "wxb::align_kernel<K,false,false> followed by its backtrack_kernel and finalize_kernel: one such chain per "
                                   "warps-per-segment class, the chains run concurrently on side streams; kernel_ms = their average duration "
                                   "per step over the timed region of `value` (CUDA events on the launching stream around K overlapped "
                                   "steps); kernel_ms_one_step_in_flight = the chain alone, bracketed by events inside wxb_plan_run",
                         "kernel_ms": ms_step, "kernel_ms_one_step_in_flight": dp, "algorithmic_bytes_per_step": int(plan.alg_bytes),
                         "issue_slot_frac": (traffic.get("warp_instructions_per_step") / (148 * 4 * 1.965e9 * ms_step * 1e-3)
                                             if traffic.get("warp_instructions_per_step") else None),
                         "note": "algorithmic bytes = sum(4*T*V + 20*N) per step; traffic = ncu dram read+write summed over the "
                                 "step's launches (profiles/); the recurrence is issue-bound, not HBM-bound (DESIGN.md 4.1): "
                                 "issue_slot_frac = ncu warp instructions per step / (148 SMs x 4 slots x 1.965 GHz x kernel_ms)"},
            "e2e": {"value": audio_all / 3600.0 / (e2e_ms * 1e-3 / args.steps), "unit": UNIT,
                    "ms_per_step": e2e_ms / args.steps, "cells_per_s": cells_all / (e2e_ms * 1e-3 / args.steps),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "link_floor_ms": floor_h2d, "link_floor_d2h_ms": floor_d2h,
                    "frac_of_link": floor_h2d / (e2e_ms / args.steps) if e2e_ms == e2e_ms and e2e_ms > 0 else None,
                    "link_floor_note": "bare cudaMemcpyAsync of the same pinned emission buffer (h2d) / result buffers (d2h), timed with CUDA "
                                       "events on every rank at once after a barrier, max over ranks; the two directions overlap in wxb_align_host",
                    "api": "wxb_align_host (whisperx_b200.engine.AlignEngine.align_batch_host): pinned host buffers in and out, "
                           "planning + chunked H2D + kernels + D2H inside the timed region"},
            "gpu_launches": int(plan.launches) * args.steps,
            "clocks": clocks,
        }
        if planted:
            line["value_planted"] = {"value": audio_all / 3600.0 / (planted_ms * 1e-3), "unit": UNIT, "ms_per_step": planted_ms,
                                     "cells_per_s": cells_all / (planted_ms * 1e-3), "parity_checked": int(units[2]),
                                     "workload": "same segments, planted emissions (+8 on a random monotone token path, +8 on blank elsewhere): "
                                                 "full-length paths, the realistic case for backtrack_kernel / finalize_kernel"}
        if configs:
            configs[1].update(workload=line["config"]["workload"], segments=nseg, ms_per_step=ms_step, value=line["value"], unit=UNIT,
                              cells_per_s=line["cells_per_s"], hbm_frac=achieved / peak, failed_segments=failed,
                              gpu_launches_per_step=int(plan.launches), parity_checked=int(units[8]))
            line["configs"] = configs
        if strong:
            hours = strong["n_pool"] / 120.0
            line["strong"] = {
                "scaling": "strong", "workload": f"BASELINE configs[4]: {hours:,.0f} audio-hours = {strong['n_pool']:,} x 30 s segments, vocab 29, "
                                                 f"399-char transcripts, {args.sweep_kind} logits generated on the device per batch (counter-based: "
                                                 "a segment's data does not depend on rank or world size), raw logits in (fused log-softmax)",
                "value": hours / (strong_ms * 1e-3), "unit": UNIT, "cells_per_s": float(units[3]) / (strong_ms * 1e-3),
                "align_ms_max_over_ranks": strong_ms, "wall_s_max_over_ranks_with_generation": strong_wall,
                "segments": int(units[4]), "segments_rank0": strong["segments"], "batches_rank0": strong["batches"],
                "partition": "whisperx_b200.shard.partition_lpt over per-segment cost T*N", "partition_s": strong["partition_s"],
                "lpt_imbalance": strong["lpt_imbalance"], "failed_segments": int(units[6]), "gpu_launches": int(units[7]),
                "checksum": int(csum[0]), "checksum_note": "sum over all characters of 3*start_frame + end_frame, summed over ranks: equal at every N",
                "parity_checked": int(units[5]),
                "timed": "the rank's batches go through in groups of 16: per group, emissions generated and plans made (untimed), then the group "
                         "aligned back to back (wxb_plan_run_async, two batches in flight) between two CUDA events; sum over groups, max over ranks",
            }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
            if strong:
                line["strong"]["cpu_reference_extrapolated_hours"] = (strong["n_pool"] / 120.0) / cpu_baseline["value"] / 3600.0
        emit(line)
    plan.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


WORKLOAD_TEXT = {
    "c2": "BASELINE configs[1]: {n:,} VAD-cut segments of 5-30 s per GPU, 50 frames/s, vocab 29, "
          "~13.3 chars/s transcripts, raw logits in (fused log-softmax), iid N(0,9) logits",
    "c5": "profiling workload (BASELINE configs[4] shape): {n:,} segments of 30 s per GPU, 50 frames/s, vocab 29, "
          "~400-char transcripts, raw logits in (fused log-softmax), iid N(0,9) logits",
    "c4": "profiling workload (BASELINE configs[3] shape): {n:,} segments of 40-60 s per GPU, vocab 100 (blank = last id), "
          "500-2,000-char transcripts, raw logits in (log_softmax_kernel, then the recurrence on log-probs), iid N(0,9) logits",
}


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries write banners to file descriptor 1 (NCCL prints its version there when NCCL_DEBUG is set): send
    everything but the result line to stderr, so that stdout carries exactly one JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-planted", action="store_true", help="skip the planted-emission variant of the main workload")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (c1, c3, c4, c5)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sweep (BASELINE configs[4])")
    ap.add_argument("--only-main", action="store_true", help="profiling runs: the resident main workload only")
    ap.add_argument("--sweep-hours", type=float, default=10000.0, help="audio-hours in the strong-scaling pool")
    ap.add_argument("--sweep-kind", default="planted", choices=["iid", "planted"])
    ap.add_argument("--c4-segments", type=int, default=2048)
    ap.add_argument("--parity", type=int, default=16, help="segments of every timed batch compared with the CPU oracle")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's local CPU cores")
    ap.add_argument("--workload", default=WORKLOAD, choices=["c2", "c5", "c4"],
                    help="c2 = the benchmark line; c5 (all 30 s segments, one K class) / c4 are profiling workloads")
    ap.add_argument("--segments", type=int, default=SEGMENTS_PER_GPU)
    args = ap.parse_args()
    if args.only_main:
        args.no_cpu_baseline = args.no_e2e = args.no_planted = args.no_configs = args.no_strong = True
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
