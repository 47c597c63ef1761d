"""Shared builders for the parity tests."""
import numpy as np
import torch


def logprobs(rng, T, V, ties=False, scale=3.0):
    x = rng.standard_normal((T, V)).astype(np.float32) * np.float32(scale)
    E = torch.log_softmax(torch.from_numpy(x), -1)
    if ties:
        E = torch.round(E * 4) / 4
    return E.numpy()


def concat_cases(cases, V, align=1):
    """cases: list of (E[T,V], tokens[N]) -> (emissions, frame_off, frame_len, tok_off, tokens)."""
    em, toks, fo, fl, to = [], [], [0], [], [0]
    for E, tk in cases:
        T = E.shape[0]
        pad = (-T) % align
        em.append(np.ascontiguousarray(E, np.float32))
        if pad:
            em.append(np.zeros((pad, V), np.float32))
        toks.append(np.asarray(tk, np.int32))
        fl.append(T)
        fo.append(fo[-1] + T + pad)
        to.append(to[-1] + len(tk))
    if not em:
        em = [np.zeros((0, V), np.float32)]
        toks = [np.zeros(0, np.int32)]
    return (np.concatenate(em), np.array(fo, np.int64), np.array(fl, np.int32), np.array(to, np.int64),
            np.concatenate(toks))


def regex_spans(text):
    """The sentence splitter the golden align() fixtures were generated with (scripts/make_golden.py)."""
    import re
    spans, start = [], 0
    for m in re.finditer(r"[.!?]+(?=\s+\S)", text):
        spans.append((start, m.end()))
        nxt = re.compile(r"\S").search(text, m.end())
        start = nxt.start() if nxt else len(text)
    spans.append((start, len(text)))
    out = []
    for a, b in spans:
        while a < b and text[a].isspace():
            a += 1
        out.append((a, b))
    return out


class ScriptedModel(torch.nn.Module):
    """Returns prepared emissions, one tensor per call, in call order (torchaudio-style signature)."""

    def __init__(self, outputs, device="cpu"):
        super().__init__()
        self.outputs = [torch.as_tensor(o).to(device) for o in outputs]
        self.calls = 0

    def forward(self, waveform, lengths=None):
        x = self.outputs[self.calls]
        self.calls += 1
        return x.unsqueeze(0), None


LABELS = ('-', '|', 'E', 'T', 'A', 'O', 'N', 'I', 'H', 'S', 'R', 'D', 'L', 'U', 'M', 'W', 'C', 'F', 'G', 'Y',
          'P', 'B', 'V', 'K', "'", 'X', 'J', 'Q', 'Z')     # torchaudio WAV2VEC2_ASR_BASE_960H labels
DICTIONARY = {c.lower(): i for i, c in enumerate(LABELS)}


def same_records(got, exp, score_tol=0.0):
    """Compare align() records: text / timestamps exactly, scores within score_tol.  Returns a
    description of the first difference or None."""
    if len(got) != len(exp):
        return f"{len(got)} records, expected {len(exp)}"
    for i, (g, e) in enumerate(zip(got, exp)):
        if set(g) != set(e):
            return f"record {i}: keys {sorted(g)} != {sorted(e)}"
        for k in e:
            if k in ("words", "chars"):
                d = same_records(g[k], e[k], score_tol)
                if d:
                    return f"record {i}.{k}: {d}"
            elif k == "score":
                if abs(float(g[k]) - float(e[k])) > score_tol:
                    return f"record {i}: score {g[k]} != {e[k]}"
            elif isinstance(e[k], float) or isinstance(g[k], float):
                if float(g[k]) != float(e[k]):
                    return f"record {i}: {k} {g[k]!r} != {e[k]!r}"
            elif g[k] != e[k]:
                return f"record {i}: {k} {g[k]!r} != {e[k]!r}"
    return None


def host_exp_matches_fixtures() -> bool:
    """True when this machine's torch CPU exp reproduces tests/golden/exp_canary.npz, i.e. runs the same MKL code
    path as the container that generated the golden `score` fields (scripts/make_golden_exp_canary.py)."""
    import os
    c = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exp_canary.npz"))
    return bool(np.array_equal(torch.exp(torch.from_numpy(c["x"])).numpy().view(np.int32), c["y"].view(np.int32)))
