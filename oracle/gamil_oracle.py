"""fp64 numpy restatement of the MC-dropout GA-MIL head (TEST INFRASTRUCTURE).

Restates, step by step, the head part of
`MultiHeadGatedAttentionMIL.mc_inference` (`/root/reference/model.py:280-316`)
and the MC statistics its callers take (`/root/reference/infer.py:195,212-219`,
`/root/reference/net_utils.py:207-208`).  All arithmetic in float64 so it can
arbitrate between the fp32 reference and the fp16-operand/fp32-accumulate
B200 kernels.  Pinned against the live reference by `tests/golden/` (see
`tests/golden/make_golden.py`, `tests/test_oracle.py`).

Weights are passed as a dict keyed exactly like the reference state_dict
(`/root/reference/model.py:181-203`):
  shared   : attention_V.0.{weight,bias}, attention_U.0.{weight,bias}
  separate : attention_V.{c}.0.{weight,bias}, attention_U.{c}.0.{weight,bias}
  always   : attention_weights.{c}.{weight,bias}, classifiers.{c}.weight
"""
from __future__ import annotations

import math

import numpy as np

L_FEAT = 512
D_HID = 128


# ----------------------------------------------------------------------------
# deterministic, platform-independent synthetic inputs (numpy PCG64)
# ----------------------------------------------------------------------------
def make_weights(seed: int, num_classes: int = 2, shared: bool = True, peaky: float = 1.0,
                 L: int = L_FEAT, D: int = D_HID) -> dict:
    """nn.Linear-style init, U(-1/sqrt(fan_in), 1/sqrt(fan_in)), float32
    (`/root/reference/model.py:181-203` builds the same parameter set).
    `peaky` multiplies the attention_weights (SURVEY.md §7 precision probe)."""
    rng = np.random.default_rng(seed)

    def lin(out_f, in_f, bias=True):
        b = 1.0 / math.sqrt(in_f)
        w = rng.uniform(-b, b, size=(out_f, in_f)).astype(np.float32)
        bb = rng.uniform(-b, b, size=(out_f,)).astype(np.float32) if bias else None
        return w, bb

    sd = {}
    if shared:
        for name in ("attention_V", "attention_U"):
            w, b = lin(D, L)
            sd[f"{name}.0.weight"], sd[f"{name}.0.bias"] = w, b
    else:
        for name in ("attention_V", "attention_U"):
            for c in range(num_classes):
                w, b = lin(D, L)
                sd[f"{name}.{c}.0.weight"], sd[f"{name}.{c}.0.bias"] = w, b
    for c in range(num_classes):
        w, b = lin(1, D)
        sd[f"attention_weights.{c}.weight"] = (w * np.float32(peaky)).astype(np.float32)
        sd[f"attention_weights.{c}.bias"] = b
    for c in range(num_classes):
        w, _ = lin(1, L, bias=False)
        sd[f"classifiers.{c}.weight"] = w
    return sd


def make_features(seed: int, N: int, L: int = L_FEAT, scale: float = 1.0) -> np.ndarray:
    """relu(normal) features: ResNet avg-pool outputs are >= 0."""
    rng = np.random.default_rng(seed)
    return (np.maximum(rng.standard_normal((N, L)), 0.0) * scale).astype(np.float32)


def num_classes_of(sd: dict) -> int:
    c = 0
    while f"classifiers.{c}.weight" in sd:
        c += 1
    return c


def is_shared(sd: dict) -> bool:
    return "attention_V.0.weight" in sd


def _proj(sd, name, c, shared):
    key = f"{name}.0" if shared else f"{name}.{c}.0"
    return np.asarray(sd[key + ".weight"], np.float64), np.asarray(sd[key + ".bias"], np.float64)


# ----------------------------------------------------------------------------
# the head
# ----------------------------------------------------------------------------
def mc_head_oracle(sd: dict, H: np.ndarray, keep_f: np.ndarray, keep_a: np.ndarray,
                   p_f: float, p_a: float, t_chunk: int = 8) -> dict:
    """H (N,L) float; keep_f (T,N,L) {0,1}; keep_a (T,C,N) {0,1}.

    Returns float64 arrays:
      Y (T,C) logits        model.py:313-316
      A (T,C,N) attention   model.py:305
      P (T,C) softmax_c(Y)  infer.py:195
      prob_mean/prob_m2 (C), attn_mean/attn_m2 (C,N)  (M2 = sum of squared
      deviations over the T samples; var = M2/(T-ddof))
    """
    H = np.asarray(H, np.float64)
    N, L = H.shape
    T = keep_f.shape[0]
    C = num_classes_of(sd)
    shared = is_shared(sd)
    sf = 0.0 if p_f >= 1.0 else 1.0 / (1.0 - p_f)   # nn.Dropout(p=1) yields zeros
    sa = 0.0 if p_a >= 1.0 else 1.0 / (1.0 - p_a)

    wa = [np.asarray(sd[f"attention_weights.{c}.weight"], np.float64).reshape(-1) for c in range(C)]
    ba = [float(np.asarray(sd[f"attention_weights.{c}.bias"]).reshape(-1)[0]) for c in range(C)]
    wc = [np.asarray(sd[f"classifiers.{c}.weight"], np.float64).reshape(-1) for c in range(C)]

    Y = np.empty((T, C))
    A = np.empty((T, C, N))
    for t0 in range(0, T, t_chunk):
        t1 = min(T, t0 + t_chunk)
        Hd = H[None] * keep_f[t0:t1].astype(np.float64) * sf          # model.py:280-281
        gate = {}
        for c in range(C):
            if shared and c > 0:
                gate[c] = gate[0]
                continue
            Wv, bv = _proj(sd, "attention_V", c, shared)
            Wu, bu = _proj(sd, "attention_U", c, shared)
            av = np.tanh(Hd @ Wv.T + bv)                               # model.py:285 / :297
            au = 1.0 / (1.0 + np.exp(-(Hd @ Wu.T + bu)))               # model.py:286 / :298
            gate[c] = av * au                                          # model.py:287 / :299
        for c in range(C):
            logit = gate[c] @ wa[c] + ba[c]                            # model.py:289 / :299
            logit = logit * keep_a[t0:t1, c].astype(np.float64) * sa   # model.py:291 / :301  (dropped -> 0, not -inf)
            logit = logit - logit.max(axis=-1, keepdims=True)
            e = np.exp(logit)
            a = e / e.sum(axis=-1, keepdims=True)                      # model.py:305
            A[t0:t1, c] = a
            M = np.einsum("tn,tnl->tl", a, Hd)                         # model.py:308-311
            Y[t0:t1, c] = M @ wc[c]                                    # model.py:313-316
    return finish_stats(Y, A)


def forward_oracle(sd: dict, H: np.ndarray) -> dict:
    """Eval-mode forward of the head (model.py:216-240, dropout modules inactive): Y (C), A (C,N)."""
    H = np.asarray(H, np.float64)
    N, L = H.shape
    C = num_classes_of(sd)
    st = mc_head_oracle(sd, H, np.ones((1, N, L), np.uint8), np.ones((1, C, N), np.uint8), 0.0, 0.0)
    return {"Y": st["Y"][0], "A": st["A"][0]}


def aux_pairwise_loss(a_pos: np.ndarray, a_neg: np.ndarray, is_positive: bool, margin: float = 1.0,
                      scale: float = 0.5, eps: float = 1e-6) -> np.ndarray:
    """scale * AuxiliaryLoss.pairwise_distance_loss (model.py:415-426; scale/margin of model.py:149-151) over
    the last axis; F.pairwise_distance(x1, x2) = ||x1 - x2 + eps||_2."""
    d = np.sqrt(((np.asarray(a_pos, np.float64) - np.asarray(a_neg, np.float64) + eps) ** 2).sum(axis=-1))
    return scale * (np.maximum(margin - d, 0.0) if is_positive else d)


def finish_stats(Y: np.ndarray, A: np.ndarray) -> dict:
    Y = np.asarray(Y, np.float64)
    A = np.asarray(A, np.float64)
    z = Y - Y.max(axis=-1, keepdims=True)
    P = np.exp(z)
    P /= P.sum(axis=-1, keepdims=True)                                 # infer.py:195
    pm = P.mean(axis=0)
    am = A.mean(axis=0)
    return {
        "Y": Y, "A": A, "P": P, "count": Y.shape[0],
        "prob_mean": pm, "prob_m2": ((P - pm) ** 2).sum(axis=0),       # net_utils.py:208, infer.py:49-52
        "attn_mean": am, "attn_m2": ((A - am) ** 2).sum(axis=0),       # infer.py:216-219
    }


def welford_sumform(count, mean, m2):
    """[n, n*mu, M2 + n*mu^2] — the additive form allreduced across MC shards."""
    mean = np.asarray(mean, np.float64)
    return float(count), count * mean, np.asarray(m2, np.float64) + count * mean * mean


def welford_from_sumform(n, s1, s2):
    mean = s1 / n
    return n, mean, s2 - n * mean * mean
