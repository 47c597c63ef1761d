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
