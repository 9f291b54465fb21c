#!/usr/bin/env python
"""Randomised cross-check on the GPU: tcgen05 path vs the fp32 CUDA-core path (same Philox masks) over random
ragged batches, sample counts, head counts, attention modes, offsets and dropout rates.  Not a pytest (minutes)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import mcmil_b200 as mm  # noqa: E402
from oracle import gamil_oracle as G  # noqa: E402

dev = torch.device("cuda")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
worst = 0.0
for it in range(iters):
    C = int(rng.integers(1, 5))
    shared = bool(rng.integers(0, 2))
    n_bags = int(rng.integers(1, 7))
    lens = [int(rng.choice([1, 2, 63, 64, 65, 127, 128, 129, 200, 333, 1024, int(rng.integers(1, 1500))])) for _ in range(n_bags)]
    T = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 16, 33, int(rng.integers(1, 60))]))
    p_f, p_a = float(rng.choice([0.0, 0.1, 0.35])), float(rng.choice([0.0, 0.1, 0.5]))
    sd = G.make_weights(int(rng.integers(0, 1000)), C, shared)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    H = torch.from_numpy(np.concatenate([G.make_features(int(rng.integers(0, 10000)), n) for n in lens])).to(dev)
    cu = np.concatenate([[0], np.cumsum(lens)])
    kw = dict(seed=int(rng.integers(0, 2 ** 40)), p_f=p_f, p_a=p_a, cu_seqlens=cu, return_attention=True,
              t_offset=int(rng.integers(0, 5)), bag_offset=int(rng.integers(0, 3)))
    a = mm.mc_head(w, H, T, impl="tcgen05", **kw)
    b = mm.mc_head(w, H, T, impl="simt_fp32", **kw)
    torch.cuda.synchronize()
    dA = float((a.A - b.A).abs().max())
    rA = float(((a.A - b.A).abs() / b.A.clamp_min(1e-12)).max())
    dY = float((a.Y - b.Y).abs().max())
    dm = float((a.attn_mean - b.attn_mean).abs().max())
    worst = max(worst, rA)
    ok = dA < 1e-4 and dY < 5e-3 * max(1.0, float(b.Y.abs().max())) and dm < 1e-4 and rA < 2e-2 and bool(torch.isfinite(a.Y).all())
    print(f"{it:3d} C={C} shared={int(shared)} lens={lens} T={T} p=({p_f},{p_a}) dA={dA:.2e} relA={rA:.2e} dY={dY:.2e} {'ok' if ok else 'FAIL'}", flush=True)
    if not ok:
        sys.exit(1)
print("stress ok, worst relative attention difference", worst)

# ---- reductions across their dispatch regimes (warp / chunked-warp / CTA row kernels; column kernel with 1..16 sample
# groups, partial tiles, many bags): the statistics must be the statistics of the returned samples, every softmax row
# sums to one, Y equals the pooled scores recomputed from A by the fp32 path
shapes = [([1024] * 40, 60), ([3000, 200, 1500] * 12, 50), ([16384], 200), ([5000, 70], 33), ([1] * 300, 16), ([129] * 9, 1000),
          ([2049, 4097], 17), ([33] * 70, 2)]
for lens, T in shapes:
    C = int(rng.integers(1, 5))
    sd = G.make_weights(int(rng.integers(0, 1000)), C, bool(rng.integers(0, 2)))
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    cu = np.concatenate([[0], np.cumsum(lens)])
    g = torch.Generator(device=dev).manual_seed(int(rng.integers(0, 1 << 30)))
    H = torch.relu(torch.randn(int(cu[-1]), 512, generator=g, device=dev))
    r = mm.mc_head(w, H, T, seed=int(rng.integers(0, 1 << 30)), cu_seqlens=cu, return_attention=True)
    A = r.A.double()
    seg = torch.from_numpy(np.repeat(np.arange(len(lens)), lens)).to(dev)
    row_sums = torch.zeros(T, C, len(lens), dtype=torch.float64, device=dev).index_add_(2, seg, A)
    e_sum = float((row_sums - 1).abs().max())
    e_mean = float((A.mean(0) - r.attn_mean.double()).abs().max() / A.mean(0).abs().max())
    m2 = ((A - A.mean(0)) ** 2).sum(0)
    e_m2 = float((m2 - r.attn_m2.double()).abs().max() / m2.abs().max().clamp_min(1e-300)) if T > 1 else 0.0
    P = torch.softmax(r.Y.double(), -1)
    e_pm = float((P.mean(1) - r.prob_mean.double()).abs().max())
    e_pq = float((((P - P.mean(1, keepdim=True)) ** 2).sum(1) - r.prob_m2.double()).abs().max())
    ok = e_sum < 2e-5 and e_mean < 2e-6 and e_m2 < 2e-4 and e_pm < 1e-6 and e_pq < 1e-5 and bool(torch.isfinite(r.Y).all())
    print(f"reduce n_bags={len(lens)} max_n={max(lens)} T={T} C={C}: rowsum {e_sum:.1e} mean {e_mean:.1e} m2 {e_m2:.1e} "
          f"pmean {e_pm:.1e} pm2 {e_pq:.1e} {'ok' if ok else 'FAIL'}", flush=True)
    if not ok:
        sys.exit(1)
    del r, A
print("reduction stress ok")
