"""Measured error of the CUDA paths on every golden case (run on the GPU box): the numbers the relative bounds in
tests/test_gpu_parity.py (REL) are set from.  Prints one line per implementation with the maxima over all cases."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmil_b200 as mm                                    # noqa: E402
from oracle import gamil_oracle as G                       # noqa: E402
from tests.cases import Case, golden_names                 # noqa: E402

dev = torch.device("cuda")
for impl in ("simt_fp32", "tcgen05"):
    worst = {"attn_mean_rel": (0, ""), "attn_m2_rel_to_max": (0, ""), "logit_abs_over_scale": (0, ""), "A_rel": (0, ""),
             "prob_rel": (0, "")}
    for name in golden_names():
        if name.endswith("native"):
            continue
        c = Case(name)
        w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in c.sd.items()}, dev)
        res = mm.mc_head(w, torch.from_numpy(c.H).to(dev), c.T, seed=c.mseed, p_f=c.p_f, p_a=c.p_a, t_offset=c.t0,
                         bag_offset=c.bag, return_attention=True, impl=impl)
        ref = c.ref
        Y = res.Y[0].double().cpu().numpy()
        am, aq = res.attn_mean.double().cpu().numpy(), res.attn_m2.double().cpu().numpy()
        A = res.A.double().cpu().numpy()[::c.A_stride]
        P = G.finish_stats(Y, np.zeros((c.T, Y.shape[1], 1)))["P"]
        Pr = G.finish_stats(np.asarray(ref["Y"], np.float64), np.zeros((c.T, Y.shape[1], 1)))["P"]
        vals = {"attn_mean_rel": np.abs(am / ref["attn_mean"] - 1).max(),
                "attn_m2_rel_to_max": (np.abs(aq - ref["attn_m2"]).max() / max(np.abs(ref["attn_m2"]).max(), 1e-300)) if c.T > 1 else 0.0,
                "logit_abs_over_scale": np.abs(Y - ref["Y"]).max() / max(1.0, np.abs(ref["Y"]).max()),
                "A_rel": np.abs(A / ref["A"].astype(np.float64) - 1).max(),
                "prob_rel": np.abs(P / Pr - 1).max()}
        for k, v in vals.items():
            if v > worst[k][0]:
                worst[k] = (float(v), name)
    print(impl, {k: ("%.3e" % v[0], v[1]) for k, v in worst.items()}, flush=True)
