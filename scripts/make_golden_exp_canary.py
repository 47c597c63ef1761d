"""Writes tests/golden/exp_canary.npz: 8,192 float32 log-probs and torch.exp of them *on this machine's CPU*.

The reference's point probability is a CPU `Tensor.exp()` (whisperx/alignment.py:409); torch dispatches it to
MKL VML, whose float32 results depend on the host's instruction set (profiles/r2_exp_host_dependence.md).  The
golden `score` fields under tests/golden/ were produced in the build container; a test machine whose torch.exp
reproduces this canary bit for bit runs the same code path and can compare scores with the fixtures at zero
tolerance — otherwise only against the oracle executed live on that machine.

    python scripts/make_golden_exp_canary.py
"""
import os

import numpy as np
import torch

rng = np.random.default_rng(409)
x = np.concatenate([-np.abs(rng.standard_normal(4096)) * 4, -np.exp(rng.uniform(-12, 3.5, 4096))]).astype(np.float32)
y = torch.exp(torch.from_numpy(x)).numpy()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "exp_canary.npz")
np.savez_compressed(out, x=x, y=y)
print(out, torch.backends.cpu.get_cpu_capability())
