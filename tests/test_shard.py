"""CPU: the multi-GPU front end (LPT partition + host-side gather) on a world-size-2 gloo group.
The per-rank aligner is replaced by a deterministic stand-in: what is under test is that every
segment is aligned exactly once and that the gathered dict is in input order."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from whisperx_b200.shard import align_sharded, partition_lpt, segment_cost


def test_partition_lpt_balances_and_covers():
    rng = np.random.default_rng(0)
    costs = (rng.integers(249, 1500, 4096) * rng.integers(66, 400, 4096)).astype(float)
    for n in (1, 2, 4, 8):
        parts = partition_lpt(costs, n)
        assert sorted(i for p in parts for i in p) == list(range(4096))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / n) < 1.01            # within ~1 %
        assert all(p == sorted(p) for p in parts)
    assert partition_lpt([], 3) == [[], [], []]
    assert partition_lpt([5.0, 1.0, 1.0], 2) == [[0], [1, 2]]


def test_segment_cost_matches_frame_formula():
    assert segment_cost({"start": 0.0, "end": 30.0, "text": "x" * 400}) == 1499 * 400
    assert segment_cost({"start": 1.0, "end": 1.01, "text": "ab"}) == 1 * 2       # < 400 samples -> one frame


def _fake_aligner(segs, model, meta, audio, device, *a):
    # one or two records per input segment, tagged with the text
    out = []
    for s in segs:
        recs = [{"start": s["start"], "end": s["end"], "text": s["text"], "words": [{"word": s["text"]}]}]
        if len(s["text"]) % 2:
            recs.append({"start": s["end"], "end": s["end"], "text": s["text"] + "+", "words": []})
        out.append(recs)
    return out


def _transcript(n=23):
    rng = np.random.default_rng(1)
    t, out = 0.0, []
    for i in range(n):
        d = float(rng.uniform(0.5, 30))
        out.append({"start": t, "end": t + d, "text": "w" * int(rng.integers(1, 300)) + str(i)})
        t += d
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = align_sharded(_transcript(), None, {}, None, "cpu", align_grouped_fn=_fake_aligner)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_align_sharded_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[1] is None
    single = align_sharded(_transcript(), None, {}, None, "cpu", align_grouped_fn=_fake_aligner)
    assert got[0] == single
    assert [s["text"] for s in single["segments"] if not s["text"].endswith("+")] == [t["text"] for t in _transcript()]
