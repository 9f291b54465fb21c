"""GPU versions of the two steps either side of the head, behind the
interface of the reference's `ImagePatcher` (/root/reference/image_patcher.py:7-131).

* `get_tiles` / `convert_img_to_bag` — tiling grid and bag selection by non-empty-pixel fraction
  (image_patcher.py:16-59,115-131): the per-tile Python loop becomes two kernels
  (`mcmil_tile_nonzero_pct`, `mcmil_gather_tiles`).
* `attention_map_stats` — `reconstruct_attention_map` (image_patcher.py:83-110) + mean / unbiased std
  over the MC passes (infer.py:212-219) computed per tile-boundary CELL straight from the head's
  attention tensor in HBM (`mcmil_attnmap_stats`); pixel maps are a gather of the cell values.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@dataclass
class AttentionMapStats:
    """cell_mean / cell_m2: (C, n_cells) fp32 CUDA; count = number of MC passes;
    row_cell (H,), col_cell (W,) int64 CUDA map pixels to cells (cell = row_cell[y] * ncx + col_cell[x])."""
    cell_mean: torch.Tensor
    cell_m2: torch.Tensor
    count: int
    row_cell: torch.Tensor
    col_cell: torch.Tensor
    ncy: int
    ncx: int

    def _to_pixels(self, cells):
        grid = cells.view(cells.shape[0], self.ncy, self.ncx)
        return grid[:, self.row_cell][:, :, self.col_cell]

    def mean_map(self):                          # infer.py:216,218
        return self._to_pixels(self.cell_mean)

    def std_map(self, ddof: int = 1):            # infer.py:217,219 (torch.std: unbiased)
        return self._to_pixels(torch.sqrt(self.cell_m2 / max(self.count - ddof, 1)))


class ImagePatcher:
    def __init__(self, patch_size=224, overlap=0.5, bag_size=-1, empty_thresh=0.8):
        self.patch_size = patch_size
        self.overlap = overlap
        self.bag_size = bag_size
        self.empty_thresh = empty_thresh
        self.tiles = None
        self._tiles_dev = None
        self._tiles_src = None        # the `self.tiles` array the device copy was made from

    # ---- grid (image_patcher.py:16-41) -------------------------------------------------------
    def _start_points(self, size, split_size):
        stride = int(split_size * (1 - self.overlap))
        pts, k = [0], 1
        while True:
            pt = stride * k
            if pt + split_size >= size:
                pts.append(size - split_size)     # the last tile is clamped to the border
                break
            pts.append(pt)
            k += 1
        return pts

    def get_tiles(self, h, w):
        xs = self._start_points(w, self.patch_size)
        ys = self._start_points(h, self.patch_size)
        tiles = np.array([(y, x, self.patch_size, self.patch_size, i, j)
                          for i, y in enumerate(ys) for j, x in enumerate(xs)], dtype=np.int64)
        self.tiles = tiles
        self._tiles_dev = None
        return tiles

    def _tiles_on(self, dev):
        # `self.tiles` may be assigned directly (the reference's dataset sets patcher.tiles per image,
        # dataset.py:60-66): the device copy follows the identity of the host array
        if self._tiles_dev is None or self._tiles_dev.device != dev or self._tiles_src is not self.tiles:
            self._tiles_dev = torch.from_numpy(np.asarray(self.tiles).astype(np.int32)).to(dev).contiguous()
            self._tiles_src = self.tiles
        return self._tiles_dev

    def _check_image(self, h, w):
        """The reference slices `image[:, y:y+p, x:x+p]` and fails on a shape mismatch (image_patcher.py:52); the
        kernels index raw memory, so the grid must fit the image."""
        if self.tiles is None or len(self.tiles) == 0:
            raise RuntimeError("ImagePatcher: call get_tiles(h, w) first")
        t = np.asarray(self.tiles)
        if t[:, 0].min() < 0 or t[:, 1].min() < 0 or t[:, 0].max() + self.patch_size > h or t[:, 1].max() + self.patch_size > w:
            raise ValueError(f"ImagePatcher: the tile grid does not fit a {h}x{w} image (grid built for another size?)")

    # ---- bag selection (image_patcher.py:43-59, 115-131) -----------------------------------------
    def tile_nonzero_pct(self, image: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        if image.device.type != "cuda":
            raise RuntimeError("ImagePatcher (B200): the image must be a CUDA tensor")
        image = image.float().contiguous()
        c, h, w = image.shape
        self._check_image(h, w)
        tiles = self._tiles_on(image.device)
        pct = torch.empty(tiles.shape[0], dtype=torch.float32, device=image.device)
        with torch.cuda.device(image.device):
            _lib.check(lib.mcmil_tile_nonzero_pct(_p(image), w, _p(tiles), tiles.shape[0], self.patch_size, _p(pct),
                                                  _stream(image.device)), "mcmil_tile_nonzero_pct")
        return pct

    def convert_img_to_bag(self, image: torch.Tensor, shuffle: bool = False, generator=None, random_state=None):
        """image (c,H,W) CUDA -> (instances (n,c,p,p), instances_idx (n,), instances_cords (n,2)).

        Order of the returned tiles: descending non-empty fraction (ties by tile index) by default.  The reference
        returns them shuffled (sklearn.utils.shuffle, image_patcher.py:131): `shuffle=True` permutes on the device
        (`generator`), `shuffle="reference"` applies sklearn's own permutation (`random_state`, host side) like the
        reference.  Tie order: the reference sorts with numpy's unstable argsort (image_patcher.py:56), so WHICH of
        several equally non-empty tiles survive a `bag_size` cut is unspecified there; here ties go to the lower
        tile index (stable sort).  With bag_size = -1 the selected set is identical.
        One host read (the number of non-empty tiles sizes the bag tensor) — the reference's output shape is data
        dependent too."""
        lib = _lib.load()
        image = image.float().contiguous()
        c, h, w = image.shape
        pct = self.tile_nonzero_pct(image)
        order = torch.argsort(-pct, stable=True)
        n_ok = int((pct > self.empty_thresh * 100).sum().item())
        if self.bag_size > 0:
            n_sel = min(self.bag_size, n_ok)
        elif self.bag_size == -1:
            n_sel = n_ok
        else:
            raise ValueError("Invalid bag size")
        sel = order[:n_sel]
        if shuffle == "reference" and n_sel > 1:
            from sklearn.utils import shuffle as sk_shuffle
            perm = sk_shuffle(np.arange(n_sel), random_state=random_state)
            sel = sel[torch.from_numpy(np.asarray(perm)).to(sel.device)]
        elif shuffle and n_sel > 1:
            sel = sel[torch.randperm(n_sel, generator=generator, device=sel.device)]
        sel32 = sel.to(torch.int32).contiguous()
        bag = torch.empty((n_sel, c, self.patch_size, self.patch_size), dtype=torch.float32, device=image.device)
        with torch.cuda.device(image.device):
            _lib.check(lib.mcmil_gather_tiles(_p(image), c, h, w, _p(self._tiles_on(image.device)), _p(sel32), n_sel,
                                              self.patch_size, _p(bag), _stream(image.device)), "mcmil_gather_tiles")
        idx = sel.cpu().numpy()
        return bag, idx, self.tiles[idx, 4:6]

    # ---- attention maps (image_patcher.py:83-110 + infer.py:212-219) ----------------------------
    def build_cells(self, instances_ids, image_hw):
        """CSR of the selected patches covering each tile-boundary cell (host, integer logic)."""
        h, w = image_hw
        tiles = self.tiles[np.asarray(instances_ids, dtype=np.int64)]
        ps = self.patch_size
        yb = np.unique(np.concatenate([[0, h], self.tiles[:, 0], self.tiles[:, 0] + ps]))
        xb = np.unique(np.concatenate([[0, w], self.tiles[:, 1], self.tiles[:, 1] + ps]))
        ncy, ncx = len(yb) - 1, len(xb) - 1
        iy0, iy1 = np.searchsorted(yb, tiles[:, 0]), np.searchsorted(yb, tiles[:, 0] + ps)
        ix0, ix1 = np.searchsorted(xb, tiles[:, 1]), np.searchsorted(xb, tiles[:, 1] + ps)
        cells, owners = [], []
        for k in range(len(tiles)):
            cy, cx = np.meshgrid(np.arange(iy0[k], iy1[k]), np.arange(ix0[k], ix1[k]), indexing="ij")
            cid = (cy * ncx + cx).reshape(-1)
            cells.append(cid)
            owners.append(np.full(cid.shape, k, np.int32))
        cells = np.concatenate(cells) if cells else np.zeros(0, np.int64)
        owners = np.concatenate(owners) if owners else np.zeros(0, np.int32)
        order = np.argsort(cells, kind="stable")
        counts = np.bincount(cells, minlength=ncy * ncx)
        ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        row_cell = np.searchsorted(yb, np.arange(h), side="right") - 1
        col_cell = np.searchsorted(xb, np.arange(w), side="right") - 1
        return ptr, owners[order].astype(np.int32), ncy, ncx, row_cell, col_cell

    def attention_map_stats(self, A: torch.Tensor, instances_ids, image_hw, row0: int = 0,
                            n_patches: Optional[int] = None) -> AttentionMapStats:
        """A: (T, C, R) CUDA fp32 (MCHeadResult.A) or the reference-shaped (T,1,C,n).  Patch k of the
        bag is packed row row0 + k and tile instances_ids[k]."""
        lib = _lib.load()
        if A.dim() == 4:
            A = A[:, 0]
        if A.device.type != "cuda" or A.dtype != torch.float32:
            raise RuntimeError("attention_map_stats: A must be a CUDA float32 tensor")
        A = A.contiguous()
        T, Cn, R = A.shape
        n = len(instances_ids) if n_patches is None else n_patches
        if row0 + n > R:
            raise ValueError("attention_map_stats: bag rows exceed A")
        ptr, idx, ncy, ncx, row_cell, col_cell = self.build_cells(instances_ids, image_hw)
        dev = A.device
        n_cells = ncy * ncx
        ptr_d, idx_d = torch.from_numpy(ptr).to(dev), torch.from_numpy(idx).to(dev)
        cellv = torch.empty((T, Cn, n_cells), dtype=torch.float32, device=dev)
        vmax = torch.empty((T, Cn), dtype=torch.float32, device=dev)
        mean = torch.empty((Cn, n_cells), dtype=torch.float32, device=dev)
        m2 = torch.empty((Cn, n_cells), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mcmil_attnmap_stats(_p(A), T, Cn, R, int(row0), _p(ptr_d), _p(idx_d), n_cells, _p(cellv),
                                               _p(vmax), _p(mean), _p(m2), _stream(dev)), "mcmil_attnmap_stats")
        return AttentionMapStats(mean, m2, T, torch.from_numpy(row_cell).to(dev), torch.from_numpy(col_cell).to(dev),
                                 ncy, ncx)
