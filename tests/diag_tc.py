#!/usr/bin/env python
"""Bring-up diagnostic for the tcgen05 projection (not a pytest; run on the GPU box):
   python tests/diag_tc.py > gpurun_out/diag.log
1. dumps the raw TMEM accumulators of the first (tile, sample) of each CTA and checks them
   against the expected fp16 x fp16 -> fp32 products under the assumed 2-CTA layout; on a
   mismatch it searches where each expected element actually landed;
2. compares logits / scores / statistics with the fp32 CUDA-core path;
3. times both implementations at BASELINE config 2.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mcmil_b200 as mm                      # noqa: E402
from mcmil_b200 import head as HD            # noqa: E402
from oracle import gamil_oracle as G         # noqa: E402


def expected_acc(sd, H):
    """(X_v, X_u): H16 @ W16^T for the tanh and the sigmoid projection, each (N, 128)."""
    H16 = torch.from_numpy(H).half().float().numpy().astype(np.float64)
    Wv = torch.from_numpy(sd["attention_V.0.weight"]).half().float().numpy().astype(np.float64)
    Wu = torch.from_numpy(sd["attention_U.0.weight"]).half().float().numpy().astype(np.float64)
    return H16 @ Wv.T, H16 @ Wu.T


def layout_check():
    dev = torch.device("cuda")
    N, T = 128, 2
    sd = G.make_weights(3, 2, True)
    H = G.make_features(7, N)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    dbg, lg, sc = HD.debug_proj_tc(w, torch.from_numpy(H).to(dev), T, seed=1, p_f=0.0, p_a=0.0)
    torch.cuda.synchronize()
    dbg = dbg.cpu().numpy()
    Xv, Xu = expected_acc(sd, H)  # (128, 128) each
    ok = True
    for rank in range(2):
        got = dbg[rank]           # CTA rank of pair 0: (128 lanes, 136)
        exp = np.zeros((128, 128))
        for lane in range(128):
            half, r = lane // 64, lane % 64
            row = 64 * rank + r
            for part in range(2):   # dump columns: [64 part + j] tanh unit 64 part + 32 half + j, [64 part + 32 + j] sigmoid
                d0 = 64 * part + 32 * half
                exp[lane, 64 * part:64 * part + 32] = Xv[row, d0:d0 + 32]
                exp[lane, 64 * part + 32:64 * part + 64] = Xu[row, d0:d0 + 32]
        err = np.abs(got[:, :128] - exp).max()
        print(f"[layout] rank {rank}: max |acc - expected| under the assumed 2x2 layout = {err:.3e}")
        if not err < 1e-2:
            ok = False
            bad = np.argwhere(np.abs(got[:, :128] - exp) > 1e-2)
            print(f"[layout]   {len(bad)} mismatching (lane, col) entries, first: {bad[:8].tolist()}")
    # score columns: expected s_c[m] = H16[m] . (hi + lo of classifier c), unscaled (p_f = 0)
    H16 = torch.from_numpy(H).half().float().numpy().astype(np.float64)
    for c in range(2):
        wc = sd[f"classifiers.{c}.weight"].reshape(-1)
        hi = torch.from_numpy(wc).half().float().numpy().astype(np.float64)
        lo = torch.from_numpy((wc - hi).astype(np.float32)).half().float().numpy().astype(np.float64)
        for rank in range(2):
            got_hi = dbg[rank][:64, 128 + c]
            got_lo = dbg[rank][:64, 128 + 4 + c]
            e_hi = np.abs(got_hi - H16[64 * rank:64 * rank + 64] @ hi).max()
            e_lo = np.abs(got_lo - H16[64 * rank:64 * rank + 64] @ lo).max()
            print(f"[layout] score head {c} rank {rank}: hi err {e_hi:.3e}, lo err {e_lo:.3e}")
    return ok


def compare_and_time():
    dev = torch.device("cuda")
    for shared in (True, False):
        N, T = 1024, 100
        sd = G.make_weights(31, 2, shared)
        w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
        H = torch.from_numpy(G.make_features(400, N)).to(dev)
        out = {}
        for impl in ("simt_fp32", "tcgen05"):
            try:
                r = mm.mc_head(w, H, T, seed=5, impl=impl, return_attention=True)
                torch.cuda.synchronize()
                out[impl] = r
            except Exception as e:  # noqa: BLE001
                print(f"[compare] shared={shared} {impl}: FAILED {e!r}")
        if len(out) == 2:
            a, b = out["tcgen05"], out["simt_fp32"]
            print(f"[compare] shared={shared}: max|dY|={float((a.Y - b.Y).abs().max()):.3e} "
                  f"max rel dP={float((a.probs() / b.probs() - 1).abs().max()):.3e} "
                  f"max|dA|={float((a.A - b.A).abs().max()):.3e} "
                  f"max rel dAmean={float((a.attn_mean / b.attn_mean - 1).abs().max()):.3e} "
                  f"max rel dM2={float(((a.attn_m2 - b.attn_m2).abs().max() / b.attn_m2.abs().max())):.3e}")
        for impl, r in out.items():
            for _ in range(3):
                mm.mc_head(w, H, T, seed=5, impl=impl)
            torch.cuda.synchronize()
            reps = 20 if impl == "tcgen05" else 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                mm.mc_head(w, H, T, seed=5 + i, impl=impl)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"[time] shared={shared} {impl}: {ms * 1e3:.1f} us per bag (N={N}, T={T}), {1e3 / ms:.0f} bags/s, launches={r.launches}")


if __name__ == "__main__":
    t0 = time.time()
    print("device:", torch.cuda.get_device_name(0))
    try:
        layout_check()
    except Exception as e:  # noqa: BLE001
        print("[layout] FAILED:", repr(e))
    try:
        compare_and_time()
    except Exception as e:  # noqa: BLE001
        print("[compare] FAILED:", repr(e))
    print(f"done in {time.time() - t0:.1f}s")
