"""Shared helpers: rebuild the inputs of a golden case from its seeds."""
import glob
import os

import numpy as np

from oracle import gamil_oracle as G
from oracle import philox as PX

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if not n.startswith(("patcher_", "fwd_"))]   # patcher_*: test_patcher.py, fwd_*: FwdCase


def forward_golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "fwd_*.npz")))


class FwdCase:
    """Deterministic forward + auxiliary loss outputs of the live reference (make_golden_forward.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.N, self.C, shared, self.wseed, self.hseed, self.mseed, self.T = (int(v) for v in z["meta"])
        self.shared = bool(shared)
        self.peaky, self.hscale = (float(v) for v in z["fmeta"])
        self.ref = {k: z[k] for k in z.files if k not in ("meta", "fmeta")}
        self.sd = G.make_weights(self.wseed, self.C, self.shared, peaky=self.peaky)
        self.H = G.make_features(self.hseed, self.N, scale=self.hscale)


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        (self.N, self.T, self.C, shared, self.wseed, self.hseed, self.mseed,
         self.bag, self.t0, self.A_stride) = (int(v) for v in z["meta"])
        self.shared = bool(shared)
        self.p_f, self.p_a, self.peaky, self.hscale = (float(v) for v in z["fmeta"])
        self.native = "keep_f_bits" in z.files
        self.ref = {k: z[k] for k in ("Y", "A", "prob_mean", "prob_m2", "attn_mean", "attn_m2")}
        self.sd = G.make_weights(self.wseed, self.C, self.shared, peaky=self.peaky)
        self.H = G.make_features(self.hseed, self.N, scale=self.hscale)
        if self.native:
            self.keep_f = PX.unpack_bits(z["keep_f_bits"], 512).reshape(self.T, self.N, 512)
            self.keep_a = PX.unpack_bits(z["keep_a_bits"], self.N).reshape(self.T, self.C, self.N)
        else:
            self.keep_f = PX.feature_keep(self.mseed, self.bag, self.t0, self.T, self.N, self.p_f)
            self.keep_a = PX.attn_keep(self.mseed, self.bag, self.t0, self.T, self.N, self.C, self.p_a)
