"""B200-native MC-dropout gated-attention MIL head (drop-in for the hot path of
xkuubix/MonteCarlo-Gated-MIL, `MultiHeadGatedAttentionMIL.mc_inference`).

Import as `mcmil_b200` (the repo-root alias package; this directory's name has a hyphen).
"""
from .head import (HeadWeights, MCHeadResult, MCHeadRunner, mc_head, export_masks, head_forward_eval,  # noqa: F401
                   aux_pairwise_loss)
from .model import MultiHeadGatedAttentionMIL, deactivate_batchnorm  # noqa: F401
from .patcher import ImagePatcher, AttentionMapStats  # noqa: F401
from . import distributed  # noqa: F401

__all__ = ["HeadWeights", "MCHeadResult", "MCHeadRunner", "mc_head", "export_masks", "head_forward_eval", "aux_pairwise_loss",
           "MultiHeadGatedAttentionMIL",
           "deactivate_batchnorm", "distributed", "ImagePatcher", "AttentionMapStats"]
