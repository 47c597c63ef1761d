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

`value`  : whole-job audio-hours/s with the emissions already resident in HBM (CUDA events).
`e2e`    : the same metric through the host-buffer C-ABI call wxb_align_host(): planning, the
           host->device copy of the emissions from pinned memory, the kernels and the
           device->host copy of the per-character results are all inside the timed region.
`--impl reference` times the reference's CPU implementation of the path (the torch/pandas port in
oracle/ref_port.py — the reference is pure Python and /root/reference does not exist on the GPU
box) on all host cores, on a bounded sample of the same workload.
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
def _cpu_worker(args):
    """One segment through the reference-faithful port: get_trellis + backtrack + merge_repeats."""
    import torch
    from oracle import ref_port
    torch.set_num_threads(1)
    E, tok, blank = args
    t0 = time.perf_counter()
    segs, _ = ref_port.align_tokens(torch.from_numpy(E), tok.tolist(), blank, "x" * len(tok))
    return time.perf_counter() - t0, (segs is not None)


def _host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_pass(batch, n_sample: int, pool, logp=None):
    """Align the first n_sample segments of `batch` with the CPU port on all cores.
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
    fo = np.concatenate([batch.frame_off[:n_sample], batch.frame_off[n_sample:n_sample + 1]])
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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    import torch
    from whisperx_b200 import synth
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
        "cells_per_s": cells / wall,
        "config": {"workload": f"BASELINE configs[1] sample: first {n_sample} of the 4,096 VAD-cut 5-30 s segments, vocab 29, "
                               "log-prob emissions, get_trellis+backtrack+merge_repeats per segment"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_sample} segments per step x {args.steps} steps, torch/pandas port of the reference "
                                   f"(oracle/ref_port.py), one process per core, torch threads=1"},
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


def ncu_traffic_bytes():
    """dram bytes per step from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["dram_bytes_per_step"]
    except Exception:
        return None


# ----------------------------------------------------------------------------- our arm
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
        cores = _host_cores()
        n_sample = min(max(16 * cores, 256), nseg)          # ~10-30 core-seconds of the Python reference path
        logp_s = torch.log_softmax(torch.from_numpy(batch.emissions[: int(batch.frame_off[n_sample])]), -1).numpy()
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_reference_pass(batch, min(cores, n_sample), pool, logp_s)           # warm the workers
            wall, audio, cells = cpu_reference_pass(batch, n_sample, pool, logp_s)
        c_wall = c_port_pass(batch, n_sample, logp_s)
        cpu_baseline = {"value": audio / 3600.0 / wall, "unit": UNIT, "cores": cores, "kind": "port",
                        "cells_per_s": cells / wall,
                        "sample": f"first {n_sample} of the {nseg} segments ({cells/1e6:.0f} Mcells, {wall:.1f} s wall), "
                                  "torch/pandas port of the reference (oracle/ref_port.py), one process per core",
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
    from whisperx_b200.engine import AlignEngine, BatchSpec, _alloc_outputs
    eng = AlignEngine(dev)
    spec = BatchSpec(frame_off=batch.frame_off, frame_len=batch.frame_len, tok_off=batch.tok_off, tokens=batch.tokens,
                     vocab=batch.vocab, blank_id=batch.blank, logits=True,
                     seg_t1=np.zeros(nseg), seg_ratio=batch.durations / batch.frame_len)
    host_em = torch.from_numpy(batch.emissions).pin_memory()
    dev_em = host_em.to(dev, non_blocking=True)
    plan = eng.plan(spec)
    out = _alloc_outputs(spec, dev, ())
    host_out = _alloc_outputs(spec, "cpu", (), pin=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- resident-in-HBM timing (value) ----
    for _ in range(args.warmup):
        plan.run(dev_em, out=out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        plan.run(dev_em, out=out)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    # dominant kernel alone (all align_kernel launches of one step run concurrently on side streams;
    # the library brackets them with CUDA events on the launching stream)
    dp_ms = []
    for _ in range(args.steps):
        plan.run(dev_em, out=out, timed=True)
        dp_ms.append(out.extra["dp_ms"])
    torch.cuda.synchronize(dev)
    failed = int(out.status.sum().item())

    # ---- end to end through the host-buffer C-ABI call ----
    e2e_s = float("nan")
    if not args.no_e2e:
        for _ in range(max(args.warmup, 1)):
            eng.align_batch_host(host_em, spec, out=host_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.align_batch_host(host_em, spec, out=host_out)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    t = torch.tensor([ms_total, e2e_s * 1e3], device=dev, dtype=torch.float64)
    units = torch.tensor([batch.audio_seconds, float(plan.cells)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    audio_all, cells_all = float(units[0]), float(units[1])

    if rank == 0:
        ms_step = ms_total / args.steps
        peak, peak_src = measured_peak_gbs()
        dp = float(np.mean(dp_ms))
        achieved = plan.alg_bytes / (dp * 1e-3) / 1e9
        h2d = host_em.numel() * 4 + batch.tokens.nbytes + batch.frame_off.nbytes + batch.tok_off.nbytes + batch.frame_len.nbytes
        d2h = sum(getattr(host_out, f).numel() * getattr(host_out, f).element_size()
                  for f in ("status", "n_frames", "char_start", "char_end", "char_score", "char_start_s", "char_end_s", "char_score_r"))
        line = {
            "metric": METRIC, "value": audio_all / 3600.0 / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "cells_per_s": cells_all / (ms_step * 1e-3),
            "config": {"workload": WORKLOAD_TEXT[args.workload].format(n=nseg),
                       "segments_per_gpu": nseg, "workload_id": args.workload, "audio_hours_per_gpu": batch.audio_seconds / 3600.0,
                       "cells_per_gpu": int(plan.cells), "emission_bytes_per_gpu": int(host_em.numel() * 4),
                       "l2": f"inputs ({host_em.numel() * 4 / 1e6:.0f} MB per step) larger than L2 (126 MB); no explicit flush",
                       "failed_segments": failed,
                       "host_cpus_bound_to_gpu_socket": (len(bound) if bound else None)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(), "peak_source": peak_src,
                         "kernel": "wxb::align_kernel<K,false,false> followed by its backtrack_kernel and finalize_kernel: one such chain per "
                                   "(warps, tokens-per-lane) class, the chains run concurrently on side streams and are timed together "
                                   "with CUDA events on the launching stream (ncu share of the three: 69 / 23 / 8 %, profiles/)",
                         "kernel_ms": dp, "algorithmic_bytes_per_step": int(plan.alg_bytes),
                         "note": "algorithmic bytes = sum(4*T*V + 20*N) per step; traffic = ncu dram read+write summed over the "
                                 "step's align_kernel launches (profiles/); the recurrence is issue-bound, not HBM-bound (DESIGN.md 4.1)"},
            "e2e": {"value": audio_all / 3600.0 / (e2e_ms * 1e-3 / args.steps), "unit": UNIT,
                    "ms_per_step": e2e_ms / args.steps, "cells_per_s": cells_all / (e2e_ms * 1e-3 / args.steps),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "wxb_align_host (whisperx_b200.engine.AlignEngine.align_batch_host): pinned host buffers in and out, "
                           "planning + chunked H2D + kernels + D2H inside the timed region"},
            "gpu_launches": int(plan.launches) * args.steps,
            "clocks": clocks,
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
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
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's local CPU cores")
    ap.add_argument("--workload", default=WORKLOAD, choices=["c2", "c5", "c4"],
                    help="c2 = the benchmark line; c5 (all 30 s segments, one K class) / c4 are profiling workloads")
    ap.add_argument("--segments", type=int, default=SEGMENTS_PER_GPU)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
