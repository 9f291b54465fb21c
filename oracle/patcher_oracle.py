"""numpy restatement of the reference's tiling and attention-map reconstruction (TEST INFRASTRUCTURE).

Follows `/root/reference/image_patcher.py`: `_start_points` :16-28, `get_tiles` :30-41,
`convert_img_to_bag` :43-59 with `_select_bag` :115-131 (bag_size == -1 path, without the final random
shuffle), `reconstruct_attention_map` :83-110, and the caller's statistics `/root/reference/infer.py:212-219`
(mean and UNBIASED std over the MC passes).  Pinned by `tests/golden/patcher_*.npz`, which
`tests/golden/make_golden_patcher.py` generates by running the reference's own ImagePatcher.
"""
from __future__ import annotations

import numpy as np


def synth_image(seed, c, h, w):
    """zero background + a non-zero half ellipse (a crude breast silhouette), SURVEY §8d config 5."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    inside = ((yy - h / 2) / (0.45 * h)) ** 2 + (xx / (0.8 * w)) ** 2 < 1.0
    img = (rng.random((c, h, w)).astype(np.float32) * 0.9 + 0.1) * inside[None]
    return img.astype(np.float32)


def synth_attention(seed, T, Cn, n):
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((T, Cn, n))
    A = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    return A.astype(np.float32)


def start_points(size, split, overlap):
    stride = int(split * (1 - overlap))
    pts, k = [0], 1
    while True:
        pt = stride * k
        if pt + split >= size:
            pts.append(size - split)
            break
        pts.append(pt)
        k += 1
    return pts


def get_tiles(h, w, patch, overlap):
    xs, ys = start_points(w, patch, overlap), start_points(h, patch, overlap)
    return np.array([(y, x, patch, patch, i, j) for i, y in enumerate(ys) for j, x in enumerate(xs)], dtype=np.int64)


def nonzero_pct(image, tiles):
    """image (c,H,W); percentage of channel-0 pixels > 0 per tile (image_patcher.py:51-53)."""
    out = np.zeros(len(tiles), np.float32)
    for i, (y, x, dh, dw, _, _) in enumerate(tiles):
        out[i] = np.float32((image[0, y:y + dh, x:x + dw] > 0).astype(np.float32).mean() * 100)
    return out


def select_bag(pct, empty_thresh):
    """set of selected tile ids for bag_size == -1 (image_patcher.py:55-56,125-127)."""
    return np.flatnonzero(pct > empty_thresh * 100)


def attention_maps(A, tiles, instances_ids, image_shape):
    """A (T,C,n) -> (T,C,c,h,w) per-pass max-normalised maps (image_patcher.py:83-110)."""
    T, Cn, n = A.shape
    c, h, w = image_shape
    rec = np.zeros((T, Cn, c, h, w), np.float64)
    cnt = np.zeros((c, h, w), np.float64)
    for k in range(n):
        y, x, dh, dw, _, _ = tiles[instances_ids[k]]
        rec[:, :, :, y:y + dh, x:x + dw] += A[:, :, k][:, :, None, None, None]
        cnt[:, y:y + dh, x:x + dw] += 1
    cnt = np.where(cnt == 0, 1, cnt)
    rec /= cnt
    mx = rec.reshape(T, Cn, -1).max(-1)
    return rec / mx[:, :, None, None, None]


def attention_map_stats(A, tiles, instances_ids, image_shape):
    """mean and unbiased std over the passes per class (infer.py:212-219)."""
    maps = attention_maps(A, tiles, instances_ids, image_shape)
    return maps.mean(0), maps.std(0, ddof=1)
