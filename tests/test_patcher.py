"""Tiling / bag selection and attention-map statistics: oracle vs the reference
outputs stored by tests/golden/make_golden_patcher.py (CPU), CUDA kernels vs both (GPU)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import patcher_oracle as PO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "patcher_*.npz")))


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    seed, c, h, w, patch, T, Cn = (int(v) for v in z["meta"])
    overlap, thresh = (float(v) for v in z["fmeta"])
    return z, seed, c, h, w, patch, T, Cn, overlap, thresh


@pytest.mark.parametrize("name", NAMES)
def test_patcher_oracle_matches_reference(name):
    z, seed, c, h, w, patch, T, Cn, overlap, thresh = _load(name)
    img = PO.synth_image(seed, c, h, w)
    tiles = PO.get_tiles(h, w, patch, overlap)
    pct = PO.nonzero_pct(img, tiles)
    assert np.allclose(pct, z["pct"], atol=1e-4)
    sel = PO.select_bag(pct, thresh)
    assert np.array_equal(sel, z["selected"])
    A = PO.synth_attention(seed + 1, T, Cn, len(sel)).astype(np.float64)
    mean, std = PO.attention_map_stats(A, tiles, sel, (1, h, w))
    assert np.abs(mean[:, 0] - z["map_mean"]).max() < 2e-6
    assert np.abs(std[:, 0] - z["map_std"]).max() < 2e-6


def test_start_points_edge_cases():
    # the reference clamps the last tile to the border (image_patcher.py:23-24)
    assert PO.start_points(2294, 224, 0.75)[-1] == 2294 - 224
    assert len(PO.start_points(2294, 224, 0.75)) == 38 and len(PO.start_points(1914, 224, 0.75)) == 32
    assert len(PO.get_tiles(2294, 1914, 224, 0.5)) == 340
    assert PO.start_points(224, 224, 0.5) == [0, 0]       # image == patch: the reference emits the tile twice


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_patcher_and_attention_maps(name):
    import mcmil_b200 as mm
    z, seed, c, h, w, patch, T, Cn, overlap, thresh = _load(name)
    dev = torch.device("cuda")
    img = PO.synth_image(seed, c, h, w)
    pt = mm.ImagePatcher(patch_size=patch, overlap=overlap, bag_size=-1, empty_thresh=thresh)
    tiles = pt.get_tiles(h, w)
    assert np.array_equal(tiles, PO.get_tiles(h, w, patch, overlap))
    pct = pt.tile_nonzero_pct(torch.from_numpy(img).to(dev)).cpu().numpy()
    assert np.allclose(pct, z["pct"], atol=1e-4)
    bag, idx, cords = pt.convert_img_to_bag(torch.from_numpy(img).to(dev))
    assert np.array_equal(np.sort(idx), z["selected"])
    bag = bag.cpu().numpy()
    for k in range(len(idx)):
        y, x = tiles[idx[k], 0], tiles[idx[k], 1]
        assert np.array_equal(bag[k], img[:, y:y + patch, x:x + patch])
        assert np.array_equal(cords[k], tiles[idx[k], 4:6])
    sel = z["selected"]
    A = PO.synth_attention(seed + 1, T, Cn, len(sel))
    # embed the bag at a row offset inside a larger packed tensor, as the head would return it
    Apacked = np.zeros((T, Cn, len(sel) + 7), np.float32)
    Apacked[:, :, 5:5 + len(sel)] = A
    st = pt.attention_map_stats(torch.from_numpy(Apacked).to(dev), sel, (h, w), row0=5, n_patches=len(sel))
    assert np.abs(st.mean_map().cpu().numpy() - z["map_mean"]).max() < 5e-6
    assert np.abs(st.std_map().cpu().numpy() - z["map_std"]).max() < 5e-6
    assert st.count == T
