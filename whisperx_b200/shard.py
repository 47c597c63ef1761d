"""Multi-GPU front end of the alignment path: segments are independent (the reference's loop at
whisperx/alignment.py:182-347 reads only the segment, the shared audio and the model), so they are
partitioned over ranks with no collective on the data path; only the result records are gathered,
on the host, at the end.

One process per GPU (`torchrun`), `torch.distributed` for the plumbing:

    result = align_sharded(segments, model, metadata, audio, f"cuda:{local_rank}")   # dict on rank 0
"""
from __future__ import annotations

import heapq
from typing import Callable, List, Optional, Sequence

SAMPLE_RATE = 16000


def partition_lpt(costs: Sequence[float], n_parts: int) -> List[List[int]]:
    """Longest-processing-time-first greedy partition: items sorted by cost (descending, ties by
    index) each go to the currently lightest part.  Returns, per part, the item indices in ascending
    order.  Deterministic, so every rank computes the same partition without communicating."""
    if n_parts < 1:
        raise ValueError("n_parts must be >= 1")
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    heap = [(0.0, p) for p in range(n_parts)]
    heapq.heapify(heap)
    for i in sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i)):
        load, p = heapq.heappop(heap)
        parts[p].append(i)
        heapq.heappush(heap, (load + float(costs[i]), p))
    for p in parts:
        p.sort()
    return parts


def segment_cost(segment: dict) -> float:
    """Trellis cells of a segment before the model has run: frames (wav2vec2: one per 320 samples after
    the first 400) x characters (an upper bound of the aligned tokens)."""
    n_samples = max(int(segment["end"] * SAMPLE_RATE) - int(segment["start"] * SAMPLE_RATE), 0)
    frames = max((max(n_samples, 400) - 400) // 320 + 1, 1)
    return float(frames) * max(len(segment["text"]), 1)


def align_sharded(transcript, model, align_model_metadata, audio, device, interpolate_method="nearest",
                  return_char_alignments=False, print_progress=False, combined_progress=False, group=None,
                  dst: int = 0, align_grouped_fn: Optional[Callable] = None) -> Optional[dict]:
    """`align()` over all ranks of `group`: every rank aligns its LPT share of the segments on its
    own GPU, rank `dst` receives all records (host-side `gather_object`) and returns the same dict
    a single-GPU `align()` would; the other ranks return None.  Without an initialised process
    group this is `align()`.  `align_grouped_fn` replaces the per-rank aligner (tests)."""
    import torch.distributed as dist
    if align_grouped_fn is None:
        from .alignment import align_grouped as align_grouped_fn
    transcript = transcript if isinstance(transcript, list) else list(transcript)
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    mine = partition_lpt([segment_cost(s) for s in transcript], world)[rank]
    grouped = align_grouped_fn([transcript[i] for i in mine], model, align_model_metadata, audio, device,
                               interpolate_method, return_char_alignments, print_progress, combined_progress)
    if len(grouped) != len(mine):
        raise RuntimeError("aligner returned a different number of segment groups than it was given")
    payload = list(zip(mine, grouped))
    if distributed and world > 1:
        # `dst` is a global rank (that is what gather_object takes); `rank` above is group-local
        is_dst = dist.get_rank() == dst
        gathered = [None] * world if is_dst else None
        dist.gather_object(payload, gathered, dst=dst, group=group)
        if not is_dst:
            return None
        payload = [item for part in gathered for item in part]
    payload.sort(key=lambda item: item[0])
    segments = [rec for _, group_recs in payload for rec in group_recs]
    words: list = []
    for seg in segments:
        words += seg["words"]
    return {"segments": segments, "word_segments": words}
