"""Philox4x32-10 and the dropout-mask convention of the B200 path (TEST INFRASTRUCTURE).

Integer, bit-exact numpy restatement of the counter-based RNG the CUDA kernels
use for the feature dropout (`/root/reference/model.py:281`) and the per-head
logit dropout (`/root/reference/model.py:291`, `:301`).  The reference itself
draws from torch's global generator; a counter-based stream cannot (and is not
expected to) reproduce it — SURVEY.md §8b "RNG".  Bit-exact comparisons with the
reference therefore always go through *injected* masks; this module is what
lets the tests regenerate, on the host, exactly the masks the kernels draw.

Philox4x32-10 follows Salmon et al., "Parallel random numbers: as easy as
1, 2, 3" (SC'11); the known-answer vectors of Random123 are checked in
`tests/test_oracle.py`.

Mask convention (mirrored by `csrc/philox.cuh`):

* key  = (seed & 0xffffffff, seed >> 32)
* feature mask of bag `b`, MC sample `t` (global index), patch `n`, feature `l` (chunk `q = l // 8`,
  K-slice `s = l // 64`): a 15-bit lane built from a *primary* byte P and a *refinement* byte R,
      P = byte 8*((n>>2)&1) + l%8          of philox((q, n & ~4, t, b), key)                    # 16 bytes, LE
      R = byte 4*((n>>2)&3) + (s>>1)       of philox((128 + (s&1)*8 + q%8, n & ~12, t, b), key)
      lane = (P << 8) | R
  (one primary call serves 8 features of the two rows n, n^4; R is shared by the 8 features of a
  (row, chunk) and only matters when the 7 primary bits equal the top bits of the threshold)
* logit mask of (b, t, n), head c:
      ctr = (64 + c // 4, n, t, b);  lane = out[c % 4] & 0xffff
* keep  <=>  (lane & 0x7fff) >= thr,   thr = floor(p * 32768 + 0.5) in float32   (15-bit Bernoulli,
  p_eff = thr / 32768; p = 0.1 -> 3277 / 32768 = 0.100006)
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

ATTN_CHUNK_BASE = 64  # ctr[0] in [64, 128): logit-dropout slots
REF_CHUNK_BASE = 128  # ctr[0] in [128, 144): refinement bytes of the feature masks


def philox4x32(ctr, key, rounds: int = 10):
    """ctr: 4 broadcastable uint32 arrays, key: 2 python ints. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in ctr)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def drop_threshold(p: float) -> int:
    """15-bit threshold: an element is DROPPED iff (lane & 0x7fff) < thr."""
    if p <= 0.0:
        return 0
    if p >= 1.0:
        return 32768
    # same float32 arithmetic as csrc/common.cuh: (int)(p * 32768.0f + 0.5f)
    return int(np.float32(p) * np.float32(32768.0) + np.float32(0.5))


def _key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def _bytes_le(words):
    """4 uint32 arrays [...]: -> uint32 array [..., 16] of their little-endian bytes."""
    out = np.empty(words[0].shape + (16,), dtype=np.uint32)
    for w in range(4):
        for k in range(4):
            out[..., 4 * w + k] = (words[w] >> np.uint32(8 * k)) & np.uint32(0xFF)
    return out


def feature_keep(seed: int, bag: int, t0: int, T: int, N: int, p: float, L: int = 512, rounds: int = 10) -> np.ndarray:
    """keep[t, n, l] in {0,1} (uint8) for t in [t0, t0+T), n in [0, N)."""
    assert L % 64 == 0
    thr = drop_threshold(p)
    Q = L // 8
    keep = np.empty((T, N, L), dtype=np.uint8)
    q = np.arange(Q, dtype=np.uint32)[None, :]
    n = np.arange(N, dtype=np.uint32)[:, None]
    refq = np.uint32(REF_CHUNK_BASE) + ((q >> np.uint32(3)) & np.uint32(1)) * np.uint32(8) + (q & np.uint32(7))
    half = ((n >> np.uint32(2)) & np.uint32(1)).astype(np.int64)              # which 8 bytes of the primary call
    ridx = (4 * ((n >> np.uint32(2)) & np.uint32(3)) + (q >> np.uint32(4))).astype(np.int64)   # [N, Q] refinement byte index
    e = np.arange(8, dtype=np.int64)[None, None, :]
    for i in range(T):
        t = np.uint32(t0 + i)
        pr = _bytes_le(philox4x32((q, n & np.uint32(~np.uint32(4)), t, np.uint32(bag)), _key(seed), rounds))       # [N, Q, 16]
        rf = _bytes_le(philox4x32((refq, n & np.uint32(~np.uint32(12)), t, np.uint32(bag)), _key(seed), rounds))   # [N, Q, 16]
        P = np.take_along_axis(pr, 8 * half[:, :, None] + e, axis=2)          # [N, Q, 8]
        R = np.take_along_axis(rf, ridx[:, :, None], axis=2)                  # [N, Q, 1]
        lanes = (P << np.uint32(8)) | R
        keep[i] = ((lanes & np.uint32(0x7FFF)) >= np.uint32(thr)).reshape(N, L)
    return keep


def attn_keep(seed: int, bag: int, t0: int, T: int, N: int, C: int, p: float, rounds: int = 10) -> np.ndarray:
    """keep[t, c, n] in {0,1} (uint8)."""
    thr = drop_threshold(p)
    keep = np.empty((T, C, N), dtype=np.uint8)
    n = np.arange(N, dtype=np.uint32)[None, :]
    t = (t0 + np.arange(T, dtype=np.uint32))[:, None]
    for g in range((C + 3) // 4):
        out = philox4x32((np.uint32(ATTN_CHUNK_BASE + g), n, t, np.uint32(bag)), _key(seed), rounds)
        for w in range(4):
            c = 4 * g + w
            if c < C:
                keep[:, c, :] = (out[w] & np.uint32(0x7FFF)) >= np.uint32(thr)
    return keep


def pack_bits(keep: np.ndarray) -> np.ndarray:
    """Pack the last axis of a {0,1} array into little-endian uint32 words
    (bit i of word w <-> element 32*w + i); the last axis is zero-padded to a
    multiple of 32."""
    keep = np.asarray(keep, dtype=np.uint8)
    n = keep.shape[-1]
    pad = (-n) % 32
    if pad:
        keep = np.concatenate([keep, np.zeros(keep.shape[:-1] + (pad,), np.uint8)], axis=-1)
    by = np.packbits(keep, axis=-1, bitorder="little")
    return np.ascontiguousarray(by).view(np.uint32)


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    by = np.ascontiguousarray(words.astype(np.uint32)).view(np.uint8)
    return np.unpackbits(by, axis=-1, bitorder="little")[..., :n]
