"""Multi-GPU partitioning of the MC-dropout head (the reference itself is
single-device).  One process per GPU over torch.distributed.

* bags are independent units -> `lpt_assign` shards a batch of bags by length, every rank runs
  `mc_head` on its own bags, NO data-path collective;
* MC samples of one bag are iid and the Philox masks are keyed by the GLOBAL sample index ->
  rank r computes samples [t0, t0+T_r) and ONE all-reduce(sum) of the additive Welford form
  [n, n*mean, M2 + n*mean^2] (fp64) merges the statistics: the merged result is the statistics
  of exactly the samples a single GPU would have drawn.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .head import HeadWeights, MCHeadResult, mc_head


def lpt_assign(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of bags to ranks (cost ~ ceil(N_b/128) tiles)."""
    cost = [-(-int(n) // 128) for n in lengths]
    order = sorted(range(len(lengths)), key=lambda i: (-cost[i], i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += cost[i]
    return [sorted(v) for v in out]


def mc_shard(T: int, rank: int, world: int) -> Tuple[int, int]:
    """(t_offset, T_local): contiguous, sizes differ by at most one, every sample owned once."""
    base, rem = divmod(int(T), world)
    t0 = rank * base + min(rank, rem)
    return t0, base + (1 if rank < rem else 0)


def welford_pack(mean: torch.Tensor, m2: torch.Tensor, count: float) -> torch.Tensor:
    """fp64 [1+2n] additive form; CUDA tensors go through the library kernels."""
    mean, m2 = mean.reshape(-1).contiguous(), m2.reshape(-1).contiguous()
    n = mean.numel()
    out = torch.empty(1 + 2 * n, dtype=torch.float64, device=mean.device)
    if mean.device.type == "cuda":
        lib = _lib.load()
        with torch.cuda.device(mean.device):
            st = C.c_void_p(torch.cuda.current_stream(mean.device).cuda_stream)
            _lib.check(lib.mcmil_welford_pack(C.c_void_p(mean.data_ptr()), C.c_void_p(m2.data_ptr()), float(count),
                                              n, C.c_void_p(out.data_ptr()), st), "mcmil_welford_pack")
    else:  # host-side logic tests (gloo)
        mu = mean.double()
        out[0] = count
        out[1:1 + n] = count * mu
        out[1 + n:] = m2.double() + count * mu * mu
    return out


def welford_unpack(packed: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    mean = torch.empty(n, dtype=torch.float32, device=packed.device)
    m2 = torch.empty(n, dtype=torch.float32, device=packed.device)
    if packed.device.type == "cuda":
        lib = _lib.load()
        with torch.cuda.device(packed.device):
            st = C.c_void_p(torch.cuda.current_stream(packed.device).cuda_stream)
            _lib.check(lib.mcmil_welford_unpack(C.c_void_p(packed.data_ptr()), n, C.c_void_p(mean.data_ptr()),
                                                C.c_void_p(m2.data_ptr()), st), "mcmil_welford_unpack")
    else:
        cnt = packed[0]
        mu = packed[1:1 + n] / cnt
        mean.copy_(mu)
        m2.copy_(torch.clamp(packed[1 + n:] - cnt * mu * mu, min=0))
    return mean, m2


def allreduce_welford(means: Sequence[torch.Tensor], m2s: Sequence[torch.Tensor], count: int, group=None,
                      total_count: Optional[int] = None):
    """ONE all-reduce for any number of statistic tensors. Returns (means, m2s, total_count).
    `total_count`: the merged sample count when the caller knows it (e.g. the job's T): the merged count is then
    not read back from the device, so the call does not synchronise the host."""
    flat_mean = torch.cat([m.reshape(-1) for m in means])
    flat_m2 = torch.cat([m.reshape(-1) for m in m2s])
    packed = welford_pack(flat_mean, flat_m2, count)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    mean, m2 = welford_unpack(packed, flat_mean.numel())
    total = int(total_count) if total_count is not None else int(round(float(packed[0].item())))
    outs_mean, outs_m2, o = [], [], 0
    for m in means:
        outs_mean.append(mean[o:o + m.numel()].view_as(m))
        outs_m2.append(m2[o:o + m.numel()].view_as(m))
        o += m.numel()
    return outs_mean, outs_m2, total


def mc_head_sample_sharded(weights: HeadWeights, H: torch.Tensor, T_total: int, seed: int = 0,
                           p_f: float = 0.1, p_a: float = 0.1, group=None, gather_Y: bool = False,
                           impl: str = "tcgen05") -> MCHeadResult:
    """Config 4 of BASELINE.json: one (large) bag, MC samples sharded over the ranks, one
    all-reduce of the Welford partials.  H is replicated on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    t0, Tl = mc_shard(T_total, rank, world)
    if Tl < 1:
        raise ValueError("mc_head_sample_sharded: fewer MC samples than ranks")
    res = mc_head(weights, H, Tl, seed=seed, p_f=p_f, p_a=p_a, t_offset=t0, impl=impl)
    (am, pm), (aq, pq), total = allreduce_welford([res.attn_mean, res.prob_mean], [res.attn_m2, res.prob_m2],
                                                   Tl, group, total_count=T_total)
    Y = res.Y
    if gather_Y:  # only callers that want medians / IQR of the per-sample probabilities (infer.py:50-55)
        sizes = [mc_shard(T_total, r, world)[1] for r in range(world)]
        padded = torch.zeros((Y.shape[0], max(sizes), Y.shape[2]), dtype=Y.dtype, device=Y.device)
        padded[:, :Tl] = Y
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        Y = torch.cat([p[:, :s] for p, s in zip(parts, sizes)], dim=1)
    return MCHeadResult(Y, pm, pq, am, aq, None, total, res.cu_seqlens, res.launches)


def mc_head_bag_sharded(weights: HeadWeights, bags: Sequence[torch.Tensor], bag_ids: Sequence[int], T: int,
                        seed: int = 0, p_f: float = 0.1, p_a: float = 0.1,
                        impl: str = "tcgen05") -> Optional[MCHeadResult]:
    """Config 3 of BASELINE.json: this rank's share of a variable-length batch (chosen with
    `lpt_assign`), given as a list of (N_b, 512) feature tensors and their GLOBAL bag ids.  One
    packed launch, NO data-path collective; the masks are keyed by the global ids, so a bag's
    result does not depend on the number of ranks."""
    if len(bags) == 0:
        return None
    lens = [int(b.shape[0]) for b in bags]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    H = bags[0] if len(bags) == 1 else torch.cat(list(bags), dim=0)
    return mc_head(weights, H.contiguous(), T, seed=seed, p_f=p_f, p_a=p_a, cu_seqlens=cu,
                   bag_ids=list(bag_ids), impl=impl)
