#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by executing the REAL reference.

Run in the build container only (`/root/reference` does not exist on the GPU box):

    python tests/golden/make_golden.py

For each case it
  1. builds deterministic, platform-independent inputs (numpy PCG64 weights and
     features, `oracle/gamil_oracle.py`; Philox masks, `oracle/philox.py`),
  2. imports `/root/reference/model.py`, swaps `feature_extractor` for
     `nn.Flatten()` so the untouched head runs on given features, swaps the three
     `nn.Dropout` modules for an `nn.Dropout` subclass that applies an injected
     keep-mask, and calls the reference's own `mc_inference`
     (`/root/reference/model.py:256-328`) on CPU,
  3. cross-checks with the reference's `mc_inference_serial` (`model.py:330-401`)
     and with `oracle/torch_port.py`,
  4. stores ONLY the reference's outputs (inputs are regenerated from seeds by the
     tests), plus — for the `native` cases — the masks torch's own RNG drew.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import gamil_oracle as G  # noqa: E402
from oracle import philox as PX       # noqa: E402
from oracle import torch_port as TP   # noqa: E402

import model as ref_model             # noqa: E402  (/root/reference/model.py)

OUT = os.path.dirname(os.path.abspath(__file__))


class InjectedDropout(nn.Dropout):
    """nn.Dropout subclass (so `isinstance(m, nn.Dropout)` at model.py:269 holds) that
    multiplies by an injected keep mask; successive calls pop successive masks (the
    serial path calls once per MC pass)."""

    def __init__(self, p, masks):
        super().__init__(p)
        self.masks = list(masks)
        self.calls = 0

    def forward(self, x):
        keep = self.masks[self.calls % len(self.masks)]
        self.calls += 1
        if self.p >= 1.0:
            return x * 0.0
        return x * keep.reshape(x.shape).to(x.dtype) / (1.0 - self.p)


class RecordingDropout(nn.Dropout):
    """Draws the mask exactly as native dropout does (bernoulli on a dense tensor of the
    same numel) and records it."""

    def __init__(self, p):
        super().__init__(p)
        self.recorded = []

    def forward(self, x):
        noise = F.dropout(torch.ones(x.shape, dtype=x.dtype), self.p, True)
        self.recorded.append((noise != 0).to(torch.uint8))
        return x * noise


def build_reference(sd_np, C, shared, p_f, p_a):
    m = ref_model.MultiHeadGatedAttentionMIL(num_classes=C, pretrained=False, feature_dropout=p_f,
                                             attention_dropout=p_a, shared_attention=shared)
    m.feature_extractor = nn.Flatten()
    missing, unexpected = m.load_state_dict(TP.sd_to_torch(sd_np), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("feature_extractor") for k in missing), missing
    return m


def run_reference_injected(sd_np, H, keep_f, keep_a, p_f, p_a, serial=False):
    C, shared = G.num_classes_of(sd_np), G.is_shared(sd_np)
    T, N = keep_f.shape[0], H.shape[0]
    m = build_reference(sd_np, C, shared, p_f, p_a)
    kf = torch.from_numpy(keep_f)
    ka = torch.from_numpy(keep_a)
    if serial:
        m.feature_dropout = InjectedDropout(p_f, [kf[t] for t in range(T)])
        m.attention_dropouts = nn.ModuleList(
            [InjectedDropout(p_a, [ka[t, c] for t in range(T)]) for c in range(C)])
        fn = m.mc_inference_serial
    else:
        m.feature_dropout = InjectedDropout(p_f, [kf])
        m.attention_dropouts = nn.ModuleList([InjectedDropout(p_a, [ka[:, c]]) for c in range(C)])
        fn = m.mc_inference
    x = torch.from_numpy(H).view(1, N, -1, 1, 1)
    Y, A = fn(x, N=T, device="cpu")
    return Y.reshape(T, C).numpy(), A.reshape(T, C, N).numpy()


def case_philox(name, N, T, C, shared, wseed, hseed, mseed, p_f=0.1, p_a=0.1, peaky=1.0,
                bag=0, t0=0, store_A_stride=1, check_serial=True, hscale=1.0):
    sd = G.make_weights(wseed, C, shared, peaky=peaky)
    H = G.make_features(hseed, N, scale=hscale)
    keep_f = PX.feature_keep(mseed, bag, t0, T, N, p_f)
    keep_a = PX.attn_keep(mseed, bag, t0, T, N, C, p_a)
    Y, A = run_reference_injected(sd, H, keep_f, keep_a, p_f, p_a)
    if check_serial:
        Ys, As = run_reference_injected(sd, H, keep_f, keep_a, p_f, p_a, serial=True)
        assert np.abs(Ys - Y).max() < 1e-5 and np.abs(As - A).max() < 1e-6, "batched vs serial reference disagree"
    Yp, Ap = TP.mc_head_torch(TP.sd_to_torch(sd), torch.from_numpy(H), T, p_f, p_a,
                              torch.from_numpy(keep_f), torch.from_numpy(keep_a))
    assert np.abs(Yp.reshape(T, C).numpy() - Y).max() < 1e-6, "torch port disagrees with the reference"
    assert np.abs(Ap.reshape(T, C, N).numpy() - A).max() < 1e-7
    st = G.finish_stats(Y, A)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=np.array([N, T, C, int(shared), wseed, hseed, mseed, bag, t0, store_A_stride], np.int64),
        fmeta=np.array([p_f, p_a, peaky, hscale], np.float64),
        Y=Y.astype(np.float32), A=A[::store_A_stride].astype(np.float32),
        prob_mean=st["prob_mean"], prob_m2=st["prob_m2"],
        attn_mean=st["attn_mean"], attn_m2=st["attn_m2"])
    print(f"{name}: N={N} T={T} C={C} shared={shared}  Y[0]={Y[0]}  max A={A.max():.3e}")


def case_native(name, N, T, C, shared, wseed, hseed, tseed, p_f=0.1, p_a=0.1):
    """Masks drawn by torch's own generator inside the reference (the 'masks injected from
    the reference' leg of north_star), recorded and stored bit-packed."""
    sd = G.make_weights(wseed, C, shared)
    H = G.make_features(hseed, N)
    x = torch.from_numpy(H).view(1, N, -1, 1, 1)
    # (a) the unmodified reference
    torch.set_num_threads(1)  # MKL's bernoulli stream is chunked per thread; pin for replay
    m = build_reference(sd, C, shared, p_f, p_a)
    torch.manual_seed(tseed)
    Y0, A0 = m.mc_inference(x, N=T, device="cpu")
    # (b) same seed, recording dropouts
    m = build_reference(sd, C, shared, p_f, p_a)
    m.feature_dropout = RecordingDropout(p_f)
    m.attention_dropouts = nn.ModuleList([RecordingDropout(p_a) for _ in range(C)])
    torch.manual_seed(tseed)
    Y1, A1 = m.mc_inference(x, N=T, device="cpu")
    assert torch.equal(Y0, Y1) and torch.equal(A0, A1), "recorded masks are not the reference's masks"
    # (c) the port with torch's own dropout consumes the generator identically
    torch.manual_seed(tseed)
    Y2, A2 = TP.mc_head_torch(TP.sd_to_torch(sd), torch.from_numpy(H), T, p_f, p_a)
    assert torch.equal(Y0, Y2) and torch.equal(A0, A2), "torch port is not bit-identical to the reference"
    torch.set_num_threads(os.cpu_count())
    keep_f = m.feature_dropout.recorded[0].reshape(T, N, -1).numpy()
    keep_a = np.stack([m.attention_dropouts[c].recorded[0].reshape(T, N).numpy() for c in range(C)], axis=1)
    Y = Y0.reshape(T, C).numpy()
    A = A0.reshape(T, C, N).numpy()
    st = G.finish_stats(Y, A)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=np.array([N, T, C, int(shared), wseed, hseed, tseed, 0, 0, 1], np.int64),
        fmeta=np.array([p_f, p_a, 1.0, 1.0], np.float64),
        keep_f_bits=PX.pack_bits(keep_f), keep_a_bits=PX.pack_bits(keep_a),
        Y=Y.astype(np.float32), A=A.astype(np.float32),
        prob_mean=st["prob_mean"], prob_m2=st["prob_m2"],
        attn_mean=st["attn_mean"], attn_m2=st["attn_m2"])
    print(f"{name}: native torch masks, keep rate {keep_f.mean():.4f}")


def main():
    torch.manual_seed(0)
    # config 1 (BASELINE.json configs[0]): N=64, T=10
    for shared in (True, False):
        tag = "shared" if shared else "separate"
        case_philox(f"c1_{tag}_s0", 64, 10, 2, shared, wseed=0, hseed=100, mseed=7)
        case_philox(f"c1_{tag}_s1", 64, 10, 2, shared, wseed=1, hseed=101, mseed=8)
        case_native(f"c1_{tag}_native", 64, 10, 2, shared, wseed=0, hseed=100, tseed=1234)
    # edge cases: ragged sizes, one patch, non-default dropout, 3 and 1 heads, t/bag offsets
    case_philox("edge_n77", 77, 12, 2, True, wseed=2, hseed=102, mseed=9, bag=2, t0=5)
    case_philox("edge_n200_sep", 200, 12, 2, False, wseed=3, hseed=103, mseed=9, bag=0, t0=0)
    case_philox("edge_n333", 333, 12, 2, True, wseed=3, hseed=104, mseed=9, bag=1, t0=0)
    case_philox("edge_n1", 1, 4, 2, True, wseed=4, hseed=105, mseed=10)
    case_philox("edge_p0", 64, 3, 2, True, wseed=4, hseed=106, mseed=11, p_f=0.0, p_a=0.0)
    case_philox("edge_p35", 130, 6, 2, False, wseed=5, hseed=107, mseed=12, p_f=0.35, p_a=0.5)
    case_philox("edge_c3", 96, 6, 3, True, wseed=6, hseed=108, mseed=13)
    case_philox("edge_c3_sep", 96, 6, 3, False, wseed=6, hseed=108, mseed=13)
    case_philox("edge_c1", 40, 5, 1, True, wseed=7, hseed=109, mseed=14)
    # config 2 (headline shape): N=1024, T=100 — store every 10th attention sample
    for shared in (True, False):
        tag = "shared" if shared else "separate"
        case_philox(f"c2_{tag}", 1024, 100, 2, shared, wseed=10, hseed=110, mseed=21,
                    store_A_stride=10, check_serial=False)
    case_philox("c2_shared_peaky", 1024, 100, 2, True, wseed=10, hseed=110, mseed=21, peaky=5.0,
                store_A_stride=10, check_serial=False, hscale=2.0)


if __name__ == "__main__":
    main()
