"""fp32 torch-CPU port of the reference head (TEST INFRASTRUCTURE / CPU baseline).

A functional restatement of the head part of
`MultiHeadGatedAttentionMIL.mc_inference` (`/root/reference/model.py:280-316`)
that issues the same ATen ops in the same order (native_dropout on the
stride-0 expanded features, addmm, tanh, sigmoid, mul, per-head addmm,
native_dropout on the logits, softmax over patches, bmm, per-head mm), so that
timing it on the GPU box's host cores is a like-for-like stand-in for the
reference's CPU path (`/root/reference` itself does not travel to the GPU box).
`tests/golden/make_golden.py` checks, in the build container, that with the
same `torch.manual_seed` this port and the real reference module return
bit-identical `(Y, A)` (same RNG consumption order: feature mask, then one
logit mask per head — SURVEY.md §8b).

Used by: `bench.py --impl reference`, `bench.py` cpu_baseline, tests.  Never by
the product path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _drop(x: torch.Tensor, p: float, keep: torch.Tensor | None) -> torch.Tensor:
    if keep is None:
        return F.dropout(x, p, training=True)          # model.py:281 / :291 (nn.Dropout in train mode)
    if p >= 1.0:
        return x * 0.0
    return x * keep.reshape(x.shape).to(x.dtype) / (1.0 - p)


def mc_head_torch(sd: dict, H: torch.Tensor, T: int, p_f: float, p_a: float,
                  keep_f: torch.Tensor | None = None, keep_a: torch.Tensor | None = None):
    """H (N,L) fp32 CPU tensor, sd = reference-keyed dict of tensors.
    keep_f (T,N,L) / keep_a (T,C,N) inject masks; None -> torch's own dropout.
    Returns (Y (T,1,C) logits, A (T,1,C,N)) exactly like model.py:328."""
    shared = "attention_V.0.weight" in sd
    C = 0
    while f"classifiers.{C}.weight" in sd:
        C += 1
    with torch.no_grad():
        Hb = H.unsqueeze(0)                                            # (1,N,L)   bs == 1
        Hx = Hb.unsqueeze(0).expand(T, -1, -1, -1)                     # model.py:280
        Hd = _drop(Hx, p_f, keep_f)                                    # model.py:281
        heads = []
        if shared:
            g = torch.tanh(F.linear(Hd, sd["attention_V.0.weight"], sd["attention_V.0.bias"])) * \
                torch.sigmoid(F.linear(Hd, sd["attention_U.0.weight"], sd["attention_U.0.bias"]))
            raw = torch.stack([F.linear(g, sd[f"attention_weights.{c}.weight"],
                                        sd[f"attention_weights.{c}.bias"]) for c in range(C)], dim=2).squeeze(-1)
            for c in range(C):                                          # model.py:291
                heads.append(_drop(raw[:, :, c, :], p_a, None if keep_a is None else keep_a[:, c]))
            logits = torch.stack(heads, dim=2)
        else:
            for c in range(C):                                          # model.py:293-303
                g = torch.tanh(F.linear(Hd, sd[f"attention_V.{c}.0.weight"], sd[f"attention_V.{c}.0.bias"])) * \
                    torch.sigmoid(F.linear(Hd, sd[f"attention_U.{c}.0.weight"], sd[f"attention_U.{c}.0.bias"]))
                a = F.linear(g, sd[f"attention_weights.{c}.weight"],
                             sd[f"attention_weights.{c}.bias"]).transpose(-1, -2)
                heads.append(_drop(a, p_a, None if keep_a is None else keep_a[:, c]))
            logits = torch.cat(heads, dim=2)
        A = F.softmax(logits, dim=-1)                                   # model.py:305
        Hs = Hd.squeeze(1)
        M = torch.stack([torch.bmm(A[:, :, c, :], Hs) for c in range(C)], dim=2)        # model.py:308-311
        Y = torch.stack([F.linear(M[:, :, c], sd[f"classifiers.{c}.weight"]) for c in range(C)], dim=-1)
        return Y.squeeze(-2), A                                         # model.py:317, :328


def sd_to_torch(sd_np: dict) -> dict:
    import numpy as np
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()}
