#!/usr/bin/env python
"""BASELINE.json configs[4]: end-to-end infer.py path on one B200 with synthetic data.

    synthetic 2294x1914 mammogram -> ImagePatcher tiling / bag selection (CUDA kernels)
    -> ResNet-18 features (torch, `deactivate_batchnorm` applied, extractor runs ONCE per bag)
    -> fused MC-dropout head, T=100 -> attention-map statistics per cell (CUDA kernels)

Mirrors /root/reference/infer.py:187-219 (without DICOM loading, Neptune and plotting).  Prints one JSON
line with the stage times (CUDA events)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmil_b200 as mm  # noqa: E402


def synth_mammogram(seed, h, w, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    yy = torch.arange(h, device=dev).view(-1, 1).float()
    xx = torch.arange(w, device=dev).view(1, -1).float()
    inside = ((yy - h / 2) / (0.45 * h)) ** 2 + (xx / (0.8 * w)) ** 2 < 1.0
    img = (torch.rand((1, h, w), generator=g, device=dev) * 0.9 + 0.1) * inside
    return img.expand(3, h, w).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=2294)
    ap.add_argument("--width", type=int, default=1914)
    ap.add_argument("--overlap", type=float, default=0.75)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--separate", action="store_true")
    ap.add_argument("--extractor-mode", default="channels_last", choices=["eager", "channels_last", "graph"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = mm.MultiHeadGatedAttentionMIL(pretrained=False, shared_attention=not args.separate)
    model.apply(mm.deactivate_batchnorm)                      # infer.py:105-109,154
    model.to(dev).eval()
    model.extractor_mode = args.extractor_mode                # torch extractor, channels-last + CUDA graph
    patcher = mm.ImagePatcher(patch_size=224, overlap=args.overlap, bag_size=-1, empty_thresh=0.75)
    patcher.get_tiles(args.height, args.width)
    img = synth_mammogram(0, args.height, args.width, dev)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    times = []
    for rep in range(args.reps + 1):
        e0 = ev()
        bag, idx, _ = patcher.convert_img_to_bag(img)                       # infer.py dataset step
        e1 = ev()
        with torch.no_grad():
            H = model.extract_features(bag.unsqueeze(0))                    # model.py:276-277, once per bag
        e2 = ev()
        res = mm.mc_head(model._head_weights(dev), H, args.T, seed=rep, return_attention=True,
                         p_f=model.feature_dropout.p, p_a=model.attention_dropouts[0].p)   # model.py:280-316
        e3 = ev()
        st = patcher.attention_map_stats(res.A, idx, (args.height, args.width))          # infer.py:197-219
        mean_map, std_map = st.mean_map(), st.std_map()
        e4 = ev()
        torch.cuda.synchronize()
        if rep > 0:
            times.append([e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4)])
    t = np.array(times).mean(0)
    probs = res.probs()[0]
    out = {"workload": f"config5: {args.height}x{args.width} image, overlap {args.overlap}, {len(idx)} of {len(patcher.tiles)} tiles, "
                       f"ResNet-18 (batch-stat BN, extractor_mode={args.extractor_mode}), T={args.T}",
           "ms": {"tiling_and_bag": t[0], "resnet18_features": t[1], "mc_head": t[2], "attention_map_stats": t[3],
                  "total": float(t.sum())},
           "prob_mean": res.prob_mean[0].tolist(), "prob_std": res.prob_var(0)[0].sqrt().tolist(),
           "entropy_mean": float((-(probs * torch.log(probs + 1e-10)).sum(-1)).mean()),       # infer.py:56-57
           "map_shape": list(mean_map.shape), "map_mean_max": float(mean_map.max()), "map_std_max": float(std_map.max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
