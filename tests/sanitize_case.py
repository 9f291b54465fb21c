"""Smallest end-to-end invocations of the hot path (every row-kernel form, ragged bag lengths): written for
compute-sanitizer; on pools where the sanitizer is closed it still runs as a plain finite / row-sum check."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import mcmil_b200 as mm
from oracle import gamil_oracle as G
dev = torch.device("cuda")
for shared in (True, False):
    sd = G.make_weights(1, 2, shared)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    lens = [77, 130, 1]
    H = torch.from_numpy(np.concatenate([G.make_features(5 + i, n) for i, n in enumerate(lens)])).to(dev)
    cu = np.concatenate([[0], np.cumsum(lens)])
    r = mm.mc_head(w, H, 3, seed=1, cu_seqlens=cu, return_attention=True)
    torch.cuda.synchronize()
    assert torch.isfinite(r.Y).all() and abs(float(r.A.sum(-1).mean()) - len(lens)) < 1e-3  # softmax per bag
# the row-kernel forms: one bag per call (CTA per row, 128- / 256- / 512- / 1024-patch chunks per warp) and a batch
# with >= 2048 rows (warp per row, one chunk / chunked); bag lengths off the 32-column grid on purpose
sd = G.make_weights(2, 2, True)
w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
for n in (701, 1499, 3001, 4999):
    H = torch.from_numpy(G.make_features(n, n)).to(dev)
    r = mm.mc_head(w, H, 2, seed=2, return_attention=True)
    torch.cuda.synchronize()
    assert torch.isfinite(r.Y).all() and abs(float(r.A.sum(-1).mean()) - 1.0) < 1e-3
for big in (63, 1301):
    lens = [big] + [33] * 40
    H = torch.from_numpy(np.concatenate([G.make_features(9 + i, n) for i, n in enumerate(lens)])).to(dev)
    cu = np.concatenate([[0], np.cumsum(lens)])
    r = mm.mc_head(w, H, 26, seed=3, cu_seqlens=cu, return_attention=True)
    torch.cuda.synchronize()
    assert torch.isfinite(r.Y).all() and abs(float(r.A.sum(-1).mean()) - len(lens)) < 1e-2
print("sanitize case ok")
