"""Multi-GPU check of the alignment front end (one process per GPU, torchrun):
every rank aligns its LPT share of a synthetic recording with a random-init WAV2VEC2_ASR_BASE_960H,
rank 0 gathers the records (host side, no collective on the data path) and compares them with a
single-GPU align() of the whole transcript.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/run_sharded_align.py
"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    import torchaudio
    from whisperx_b200 import align, alignment
    from whisperx_b200.shard import align_sharded
    alignment.set_sentence_splitter(alignment.assemble._fallback_spans)
    labels = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H.get_labels()
    meta = {"language": "en", "dictionary": {c.lower(): i for i, c in enumerate(labels)}, "type": "torchaudio"}
    torch.manual_seed(0)                                   # same random weights on every rank
    model = torchaudio.models.wav2vec2_model(**torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params).eval().to(dev)
    rng = np.random.default_rng(3)
    g = torch.Generator().manual_seed(4)
    n_seg = 24
    audio = (torch.randn(n_seg * 20 * 16000, generator=g) * 0.1).numpy()
    segs, t = [], 0.0
    for i in range(n_seg):
        d = float(rng.uniform(3, 20))
        words = ["".join(rng.choice(list("abcdefghijklmnopqrstuvwxyz"), int(rng.integers(1, 8)))) for _ in range(int(d * 2.5))]
        segs.append({"start": round(t, 3), "end": round(t + d, 3), "text": " ".join(words) + "."})
        t += d
    res = align_sharded(copy.deepcopy(segs), model, meta, audio, dev)
    if rank == 0:
        single = align(copy.deepcopy(segs), model, meta, audio, dev)
        assert res is not None and len(res["segments"]) == len(single["segments"])
        assert res == single, "sharded result differs from the single-GPU result"
        print(f"sharded align ok on {world} GPUs: {len(res['segments'])} records, {len(res['word_segments'])} words identical to 1 GPU")
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
