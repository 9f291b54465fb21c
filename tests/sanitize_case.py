"""Smallest end-to-end invocation of the hot path, for compute-sanitizer (one tool per gpurun call)."""
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
print("sanitize case ok")
