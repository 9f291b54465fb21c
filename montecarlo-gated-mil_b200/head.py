"""Features-level entry points of the fused MC-dropout GA-MIL head.

`mc_head` replaces the head part of `MultiHeadGatedAttentionMIL.mc_inference`
(/root/reference/model.py:280-316) plus the MC statistics its callers take
(/root/reference/infer.py:195,212-219; net_utils.py:207-208) by ONE call into the CUDA
library (C-ABI in include/mcmil_b200.h).  torch is used for device memory and the stream only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

L_FEAT = 512
D_HID = 128
MAX_CLASSES = 4
FP16_MAX = 65504.0


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _raw_stream(device) -> int:
    """Handle of torch's current stream on `device` (the raw query: no Stream object is built)."""
    return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(_raw_stream(device))


class HeadWeights:
    """Device-side packed copy of the head parameters (reference state_dict keys,
    /root/reference/model.py:181-203).  Re-create after the parameters change."""

    def __init__(self, state_dict: dict, device="cuda"):
        lib = _lib.load()
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("the MC-dropout head runs on CUDA only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        sd = {k: v for k, v in state_dict.items() if not k.startswith("feature_extractor")}
        C_ = 0
        while f"classifiers.{C_}.weight" in sd:
            C_ += 1
        if not 1 <= C_ <= MAX_CLASSES:
            raise NotImplementedError(f"num_classes must be in [1,{MAX_CLASSES}], got {C_}")
        self.num_classes = C_
        self.shared = "attention_V.0.weight" in sd
        self.device = device

        def dev(x):
            return torch.as_tensor(x).detach().to(device=device, dtype=torch.float32).contiguous()

        if self.shared:
            Vw, Vb = dev(sd["attention_V.0.weight"])[None], dev(sd["attention_V.0.bias"])[None]
            Uw, Ub = dev(sd["attention_U.0.weight"])[None], dev(sd["attention_U.0.bias"])[None]
        else:
            Vw = torch.stack([dev(sd[f"attention_V.{c}.0.weight"]) for c in range(C_)])
            Vb = torch.stack([dev(sd[f"attention_V.{c}.0.bias"]) for c in range(C_)])
            Uw = torch.stack([dev(sd[f"attention_U.{c}.0.weight"]) for c in range(C_)])
            Ub = torch.stack([dev(sd[f"attention_U.{c}.0.bias"]) for c in range(C_)])
        ww = torch.stack([dev(sd[f"attention_weights.{c}.weight"]).reshape(-1) for c in range(C_)])
        bw = torch.stack([dev(sd[f"attention_weights.{c}.bias"]).reshape(()) for c in range(C_)])
        cw = torch.stack([dev(sd[f"classifiers.{c}.weight"]).reshape(-1) for c in range(C_)])
        S = 1 if self.shared else C_
        if tuple(Vw.shape) != (S, D_HID, L_FEAT) or tuple(Uw.shape) != (S, D_HID, L_FEAT):
            raise ValueError(f"attention_V/U weights must be ({D_HID},{L_FEAT}) (L=512, D=128), got {tuple(Vw.shape[1:])}")
        if tuple(ww.shape) != (C_, D_HID) or tuple(cw.shape) != (C_, L_FEAT):
            raise ValueError("attention_weights must be (1,128) and classifiers (1,512)")
        tensors = [t.contiguous() for t in (Vw, Vb, Uw, Ub, ww, bw, cw)]
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.mcmil_weights_create(C.byref(handle), C_, int(self.shared), *[_ptr(t) for t in tensors],
                                                _stream_ptr(device)), "mcmil_weights_create")
        self._h = handle
        self._lib = lib

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.mcmil_weights_destroy(h)
            except Exception:
                pass


class _Plan:
    def __init__(self, cu: np.ndarray, T: int, C_: int, device, bag_ids=None):
        lib = _lib.load()
        self.cu = np.ascontiguousarray(cu, dtype=np.int32)
        ids = None
        if bag_ids is not None:
            self.bag_ids = np.ascontiguousarray(bag_ids, dtype=np.int32)
            if self.bag_ids.shape != (len(self.cu) - 1,):
                raise ValueError("bag_ids must have one entry per bag")
            ids = self.bag_ids.ctypes.data_as(C.POINTER(C.c_int32))
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.mcmil_plan_create(C.byref(handle), self.cu.ctypes.data_as(C.POINTER(C.c_int32)), ids,
                                             len(self.cu) - 1, int(T), int(C_), _stream_ptr(device)),
                       "mcmil_plan_create")
        self._h, self._lib = handle, lib
        self.ws_bytes = int(lib.mcmil_plan_workspace_bytes(handle))
        self.R = int(lib.mcmil_plan_total_rows(handle))
        self.n_bags, self.T, self.C = len(self.cu) - 1, int(T), int(C_)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.mcmil_plan_destroy(h)
            except Exception:
                pass


_plan_cache: dict = {}
_workspaces: dict = {}


_one_bag_cu: dict = {}


def _check_cu(what: str, cu_seqlens, R: int) -> np.ndarray:
    """Bag boundaries of a packed batch: host ints, start at 0, end at R, strictly increasing."""
    if cu_seqlens is None and R > 0:                     # one bag: nothing to check (single-bag call loops)
        cu = _one_bag_cu.get(R)
        if cu is None:
            if len(_one_bag_cu) > 4096:
                _one_bag_cu.clear()
            cu = _one_bag_cu[R] = np.array([0, R], np.int32)
            cu.setflags(write=False)
        return cu
    cu = np.array([0, R], np.int64) if cu_seqlens is None else np.asarray(cu_seqlens, dtype=np.int64)
    if cu.ndim != 1 or len(cu) < 2 or cu[0] != 0 or cu[-1] != R or np.any(np.diff(cu) <= 0):
        raise ValueError(f"{what}: cu_seqlens must start at 0, end at the number of packed rows ({R}) and be "
                         "strictly increasing")
    return cu.astype(np.int32)


def _get_plan(cu: np.ndarray, T: int, C_: int, device, bag_ids=None) -> _Plan:
    ids_key = None if bag_ids is None else np.asarray(bag_ids, np.int32).tobytes()
    key = (cu.tobytes(), ids_key, int(T), int(C_), str(device))
    p = _plan_cache.get(key)
    if p is None:
        if len(_plan_cache) > 256:
            _plan_cache.clear()
        p = _plan_cache[key] = _Plan(cu, T, C_, device, bag_ids)
    return p


def _get_workspace(nbytes: int, device, stream: Optional[int] = None) -> int:
    """Address of a 1024-byte aligned scratch area of `nbytes` private to (device, stream); grown on demand and kept."""
    key = (device.index, _raw_stream(device) if stream is None else stream)
    ent = _workspaces.get(key)
    if ent is None or ent[1] < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20) + 1024, dtype=torch.uint8, device=device)
        base = ws.data_ptr()
        ent = _workspaces[key] = (ws, ws.numel() - 1024, base + (-base) % 1024)
    return ent[2]


@dataclass
class MCHeadResult:
    """All tensors on the CUDA device, fp32.

    Y          (n_bags, T, C)  per-sample logits            model.py:313-316
    prob_mean  (n_bags, C)     mean_t softmax_c(Y)          net_utils.py:207-208
    prob_m2    (n_bags, C)     sum_t (P - mean)^2
    attn_mean  (C, R)          mean_t A[t,c,n]              infer.py:216,218 (patch level)
    attn_m2    (C, R)          sum_t (A - mean)^2
    A          (T, C, R) or None                            model.py:305
    count      number of MC samples behind the statistics
    cu_seqlens (n_bags+1,) numpy int32: bag b owns packed rows [cu[b], cu[b+1])
    """
    Y: torch.Tensor
    prob_mean: torch.Tensor
    prob_m2: torch.Tensor
    attn_mean: torch.Tensor
    attn_m2: torch.Tensor
    A: Optional[torch.Tensor]
    count: int
    cu_seqlens: np.ndarray
    launches: int = 0
    stream: Optional[torch.cuda.Stream] = None     # MCHeadRunner throughput mode: the stream the call runs on

    def prob_var(self, ddof: int = 0):      # infer.py:52 uses np.std (ddof=0)
        return self.prob_m2 / max(self.count - ddof, 1)

    def attn_var(self, ddof: int = 1):      # infer.py:217,219 use torch.std (ddof=1)
        return self.attn_m2 / max(self.count - ddof, 1)

    def probs(self):
        return torch.softmax(self.Y, dim=-1)   # infer.py:195 (tiny, caller-side)


def mc_head(weights: HeadWeights, H: torch.Tensor, T: int, seed: int = 0,
            p_f: float = 0.1, p_a: float = 0.1, cu_seqlens: Optional[Sequence[int]] = None,
            keep_f_bits: Optional[torch.Tensor] = None, keep_a_bits: Optional[torch.Tensor] = None,
            return_attention: bool = False, t_offset: int = 0, bag_offset: int = 0,
            bag_ids: Optional[Sequence[int]] = None, impl: str = "tcgen05",
            philox_rounds: int = 10, validate: bool = False) -> MCHeadResult:
    """Run T MC-dropout passes of the GA-MIL head on packed features.

    H            (R, 512) fp32 (or fp16) CUDA, contiguous: one bag (cu_seqlens=None) or a packed batch
    cu_seqlens   bag boundaries (host ints), len n_bags+1
    bag_ids      global id of each bag (keys the Philox masks); default 0..n_bags-1 (+ bag_offset)
    keep_f_bits  optional injected feature keep-mask, uint32/int32 (T, R, 16) CUDA
    keep_a_bits  optional injected logit keep-mask, (T, C, ceil(R/32)) CUDA   (both or neither)
    philox_rounds  10 (Philox4x32-10, default) or 7 (Philox4x32-7, ~25 % faster)
    validate     True: check (one reduction over H + a host sync) that the features are finite and inside the fp16
                 range the tensor-core path rounds them to (|h| <= 65504; ResNet avg-pool features are O(1-10)), and
                 raise ValueError otherwise.  Off by default: without it such features give inf / NaN outputs.
    """
    lib = _lib.load()
    if not isinstance(H, torch.Tensor) or H.device.type != "cuda":
        raise RuntimeError("mc_head: H must be a CUDA tensor (no CPU fallback)")
    if H.dim() != 2 or H.shape[1] != L_FEAT:
        raise ValueError(f"mc_head: H must be (R, {L_FEAT}), got {tuple(H.shape)}")
    if H.dtype not in (torch.float32, torch.float16) or not H.is_contiguous():
        raise ValueError("mc_head: H must be contiguous float32 (or float16: features that already are half precision)")
    if H.device != weights.device:
        raise ValueError("mc_head: H and the weights live on different devices")
    if impl not in _lib.IMPLS:
        raise ValueError(f"mc_head: impl must be one of {sorted(_lib.IMPLS)}")
    R = H.shape[0]
    cu = _check_cu("mc_head", cu_seqlens, R)
    if validate and R > 0:
        amax = float(H.abs().max())
        if not np.isfinite(amax) or (impl == "tcgen05" and amax > FP16_MAX):
            raise ValueError(f"mc_head: features must be finite and within the fp16 range (|h| <= {FP16_MAX:g}) of the "
                             f"tensor-core path; max |h| = {amax:g}.  Use impl='simt_fp32' or rescale the features.")
    if T < 1:
        raise ValueError("mc_head: T must be >= 1")
    dev = H.device
    C_ = weights.num_classes
    plan = _get_plan(cu, T, C_, dev, bag_ids)
    n_bags = plan.n_bags
    if (keep_f_bits is None) != (keep_a_bits is None):
        raise ValueError("mc_head: inject both masks or neither")
    if keep_f_bits is not None:
        Rw = (R + 31) // 32
        if tuple(keep_f_bits.shape) != (T, R, 16) or tuple(keep_a_bits.shape) != (T, C_, Rw):
            raise ValueError(f"mc_head: injected masks must be (T,R,16) and (T,C,{Rw}) words")
        for m in (keep_f_bits, keep_a_bits):
            if m.device != dev or m.element_size() != 4 or not m.is_contiguous():
                raise ValueError("mc_head: injected masks must be contiguous 32-bit CUDA tensors")

    # (no `with torch.cuda.device` context manager on the common path: host overhead matters for single-bag calls)
    switch = torch.cuda.current_device() != dev.index
    if switch:
        prev = torch.cuda.current_device()
        torch.cuda.set_device(dev)
    try:
        # three allocations (per-sample logits | both probability statistics | both attention statistics), views for
        # the rest: every torch.empty costs ~2.5 us of host time, which counts on single-bag call loops
        Y = torch.empty((n_bags, T, C_), dtype=torch.float32, device=dev)
        A = torch.empty((T, C_, R), dtype=torch.float32, device=dev) if return_attention else None
        pmq = torch.empty((2, n_bags, C_), dtype=torch.float32, device=dev)
        amq = torch.empty((2, C_, R), dtype=torch.float32, device=dev)
        pm, pq, am, aq = pmq[0], pmq[1], amq[0], amq[1]
        stream = _raw_stream(dev)
        ws_ptr = _get_workspace(plan.ws_bytes, dev, stream)
        fwd = lib.mcmil_head_forward if H.dtype == torch.float32 else lib.mcmil_head_forward_f16
        p_pm, p_am = pmq.data_ptr(), amq.data_ptr()
        code = fwd(weights._h, plan._h, H.data_ptr(), int(t_offset), int(bag_offset),
                   int(seed) & 0xFFFFFFFFFFFFFFFF, int(philox_rounds), float(p_f), float(p_a),
                   _ptr(keep_f_bits), _ptr(keep_a_bits), _lib.IMPLS[impl],
                   Y.data_ptr(), _ptr(A), p_pm, p_pm + 4 * n_bags * C_, p_am, p_am + 4 * C_ * R,
                   ws_ptr, plan.ws_bytes, stream)
        _lib.check(code, "mcmil_head_forward")
        launches = int(lib.mcmil_last_launch_count())
    finally:
        if switch:
            torch.cuda.set_device(prev)
    return MCHeadResult(Y, pm, pq, am, aq, A, T, cu, launches)


class _RunnerSlot:
    """One in-flight call of an MCHeadRunner: its own plan (SM limit), workspace, outputs and stream."""

    def __init__(self, runner, cu, sm_limit: int, stream):
        lib, dev, T, C_, n_rows = runner.lib, runner.dev, runner.T, runner.w.num_classes, runner.R
        self.plan = _Plan(cu, T, C_, dev) if sm_limit else _get_plan(cu, T, C_, dev)     # a limited plan is private
        if sm_limit:
            _lib.check(lib.mcmil_plan_set_sm_limit(self.plan._h, int(sm_limit)), "mcmil_plan_set_sm_limit")
        nb = self.plan.n_bags
        with torch.cuda.device(dev):
            f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)   # noqa: E731
            self.Y, self.pm, self.pq = f(nb, T, C_), f(nb, C_), f(nb, C_)
            self.am, self.aq = f(C_, n_rows), f(C_, n_rows)
            self.A = f(T, C_, n_rows) if runner.return_attention else None
            self.ws = torch.empty(self.plan.ws_bytes + 2048, dtype=torch.uint8, device=dev)
        off = (-self.ws.data_ptr()) % 1024
        self.tail = (runner.philox_rounds, runner.p_f, runner.p_a, None, None, runner.impl,
                     _ptr(self.Y), _ptr(self.A), _ptr(self.pm), _ptr(self.pq), _ptr(self.am), _ptr(self.aq),
                     C.c_void_p(self.ws.data_ptr() + off), self.plan.ws_bytes)
        self.stream = stream
        self.stream_handle = stream.cuda_stream if stream is not None else None
        self.result = MCHeadResult(self.Y, self.pm, self.pq, self.am, self.aq, self.A, T, cu, 0)
        self.result.stream = stream


class MCHeadRunner:
    """Low-overhead repeated calls for ONE fixed shape (the reference's bs == 1 serving loop, infer.py:187-196):
    plan, workspace and output tensors are created once, `run(H, seed)` is a single C-ABI call (no allocation,
    no validation beyond the shape).  Same kernels and results as `mc_head`.

    n_streams = 1 (default): every call runs on the caller's current stream; the returned MCHeadResult aliases the
    runner's buffers and is valid until the next `run`.

    n_streams = k > 1 (throughput mode): consecutive calls go round-robin to k private streams, each with its own
    buffers, and the projection kernel of a call is limited to 1/k of the SMs (`mcmil_plan_set_sm_limit`), so k bags
    are in flight side by side and the fixed per-kernel cost of a single-bag call (~10 of ~33 us) overlaps with the
    other bags' steady state.  Per-bag latency grows ~k times, bags/s reach the packed-batch rate (N=1024, T=100,
    k=8: 27 us per bag from a Python loop, 25 us from CUDA graphs, against 44 us on one stream).  SM-limited plans
    are launched without programmatic dependent launch (an early-launched projection kernel would hold the SMs its
    own stream's reduction kernels need).  The result
    of a call is valid until k further calls; wait for it with `result.stream.synchronize()` (or `synchronize()`
    for all), the input H must stay untouched until then."""

    def __init__(self, weights: HeadWeights, n_rows: int, T: int, p_f: float = 0.1, p_a: float = 0.1,
                 cu_seqlens: Optional[Sequence[int]] = None, return_attention: bool = False,
                 philox_rounds: int = 10, impl: str = "tcgen05", n_streams: int = 1, reserve_sms: int = 0):
        self.lib = _lib.load()
        self.w, self.dev, self.T, self.R = weights, weights.device, int(T), int(n_rows)
        if impl not in _lib.IMPLS:
            raise ValueError(f"MCHeadRunner: impl must be one of {sorted(_lib.IMPLS)}")
        if T < 1 or n_rows < 1:
            raise ValueError("MCHeadRunner: T and n_rows must be >= 1")
        if philox_rounds not in (7, 10) or not (0.0 <= p_f <= 1.0 and 0.0 <= p_a <= 1.0):
            raise ValueError("MCHeadRunner: philox_rounds must be 10 or 7 and the dropout probabilities in [0, 1]")
        if not 1 <= int(n_streams) <= 16:
            raise ValueError("MCHeadRunner: n_streams must be in [1, 16]")
        cu = _check_cu("MCHeadRunner", cu_seqlens, int(n_rows))
        self.philox_rounds, self.p_f, self.p_a, self.impl = int(philox_rounds), float(p_f), float(p_a), _lib.IMPLS[impl]
        self.return_attention = bool(return_attention)
        self.n_streams = int(n_streams)
        if self.n_streams == 1:
            self.slots = [_RunnerSlot(self, cu, 0, None)]
        else:
            # the projection kernels of the k streams share the SMs; `reserve_sms` of them can be kept free for the
            # (small) row / column kernels of the other bags (measured: 0 is best, tools/stream_probe.py)
            sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
            share = max(2, ((sms - max(0, int(reserve_sms))) // self.n_streams) & ~1)
            with torch.cuda.device(self.dev):
                self.slots = [_RunnerSlot(self, cu, share, torch.cuda.Stream(self.dev)) for _ in range(self.n_streams)]
        self._next = 0
        self.plan = self.slots[0].plan
        self.result = self.slots[0].result

    def run(self, H: torch.Tensor, seed: int = 0, t_offset: int = 0, bag_offset: int = 0,
            sync_input: bool = True) -> MCHeadResult:
        """sync_input (throughput mode only): make the private stream wait for the caller's current stream, on which H
        was produced (one event record + wait, ~3 us of host time); pass False when H is known to be ready."""
        if H.device != self.dev or H.dtype != torch.float32 or tuple(H.shape) != (self.R, L_FEAT) or not H.is_contiguous():
            raise ValueError(f"MCHeadRunner.run: H must be a contiguous float32 ({self.R}, {L_FEAT}) tensor on {self.dev}")
        slot = self.slots[self._next]
        if self.n_streams == 1:
            stream = _raw_stream(self.dev)
        else:
            self._next = (self._next + 1) % self.n_streams
            if sync_input:
                slot.stream.wait_stream(torch.cuda.current_stream(self.dev))   # H was produced on the caller's stream
            stream = slot.stream_handle
        code = self.lib.mcmil_head_forward(self.w._h, slot.plan._h, H.data_ptr(), int(t_offset),
                                           int(bag_offset), int(seed) & 0xFFFFFFFFFFFFFFFF, *slot.tail, stream)
        if code:
            _lib.check(code, "mcmil_head_forward")
        return slot.result

    def synchronize(self):
        """Wait for every call issued so far (throughput mode: all private streams)."""
        for slot in self.slots:
            (slot.stream or torch.cuda.current_stream(self.dev)).synchronize()


def head_forward_eval(weights: HeadWeights, H: torch.Tensor, cu_seqlens: Optional[Sequence[int]] = None,
                      impl: str = "tcgen05", validate: bool = False):
    """Deterministic (eval-mode) forward of the head, /root/reference/model.py:216-240 with the dropout
    modules inactive: one pass of the same fused kernels with every mask element kept.
    Returns (Y (n_bags, C) logits, A (C, R) attention)."""
    res = mc_head(weights, H, 1, seed=0, p_f=0.0, p_a=0.0, cu_seqlens=cu_seqlens, return_attention=True, impl=impl,
                  validate=validate)
    return res.Y[:, 0, :], res.A[0]


def aux_pairwise_loss(A: torch.Tensor, is_positive: bool, cu_seqlens: Optional[Sequence[int]] = None,
                      pos_head: int = 1, neg_head: int = 0, margin: float = 1.0, scale: float = 0.5,
                      eps: float = 1e-6) -> torch.Tensor:
    """scale * AuxiliaryLoss('pairwise')(A[:, pos], A[:, neg], is_positive) per (bag, MC pass):
    /root/reference/model.py:243-248 (forward) and :318-326 (one value per pass), loss of :415-426.
    A (T, C, R) fp32 CUDA (the `A` of mc_head); returns (n_bags, T) fp32 CUDA."""
    lib = _lib.load()
    if not isinstance(A, torch.Tensor) or A.device.type != "cuda" or A.dim() != 3:
        raise RuntimeError("aux_pairwise_loss: A must be a (T, C, R) CUDA tensor (no CPU fallback)")
    if A.dtype != torch.float32 or not A.is_contiguous():
        raise ValueError("aux_pairwise_loss: A must be contiguous float32")
    T, C_, R = A.shape
    if not (0 <= pos_head < C_ and 0 <= neg_head < C_):
        raise ValueError("aux_pairwise_loss: needs the two heads it compares (the reference uses heads 1 and 0)")
    cu = _check_cu("aux_pairwise_loss", cu_seqlens, R)
    dev = A.device
    plan = _get_plan(cu, T, C_, dev)
    with torch.cuda.device(dev):
        out = torch.empty((plan.n_bags, T), dtype=torch.float32, device=dev)
        _lib.check(lib.mcmil_aux_pairwise_loss(plan._h, _ptr(A), int(T), int(R), int(pos_head), int(neg_head),
                                               int(bool(is_positive)),
                                               float(margin), float(scale), float(eps), _ptr(out), _stream_ptr(dev)),
                   "mcmil_aux_pairwise_loss")
    return out


def export_masks(T: int, R_or_cu, num_classes: int, seed: int, p_f: float, p_a: float,
                 t_offset: int = 0, bag_offset: int = 0, device="cuda", philox_rounds: int = 10):
    """The keep-bits the in-kernel Philox draws: (feat (T,R,16) int32, attn (T,C,ceil(R/32)) int32)."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if np.isscalar(R_or_cu):
        cu = _check_cu("export_masks", None, int(R_or_cu))
    else:
        cu = _check_cu("export_masks", R_or_cu, int(np.asarray(R_or_cu)[-1]))
    if T < 1 or not 1 <= num_classes <= MAX_CLASSES or philox_rounds not in (7, 10):
        raise ValueError("export_masks: T >= 1, num_classes in [1,4], philox_rounds 10 or 7")
    plan = _get_plan(cu, T, num_classes, dev)
    R = plan.R
    with torch.cuda.device(dev):
        fb = torch.zeros((T, R, 16), dtype=torch.int32, device=dev)
        ab = torch.zeros((T, num_classes, (R + 31) // 32), dtype=torch.int32, device=dev)
        _lib.check(lib.mcmil_export_masks(plan._h, int(t_offset), int(bag_offset), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                          int(philox_rounds), float(p_f), float(p_a), _ptr(fb), _ptr(ab),
                                          _stream_ptr(dev)),
                   "mcmil_export_masks")
    return fb, ab


def debug_proj_tc(weights: HeadWeights, H: torch.Tensor, T: int, seed: int, p_f: float, p_a: float,
                  cu_seqlens=None, keep_f_bits=None, keep_a_bits=None, t_offset=0, bag_offset=0):
    """Tests only: raw TMEM dump + logits/scores planes of the tcgen05 projection."""
    lib = _lib.load()
    dev = H.device
    R = H.shape[0]
    cu = _check_cu("debug_proj_tc", cu_seqlens, R)
    plan = _get_plan(cu, T, weights.num_classes, dev)
    Rp = int(lib.mcmil_plan_plane_cols(plan._h))
    with torch.cuda.device(dev):
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        dbg = torch.zeros((sms, 128, 136), dtype=torch.float32, device=dev)
        lg = torch.zeros((T, weights.num_classes, Rp), dtype=torch.float32, device=dev)
        sc = torch.zeros_like(lg)
        _lib.check(lib.mcmil_debug_proj_tc(weights._h, plan._h, _ptr(H), int(t_offset), int(bag_offset),
                                           int(seed), float(p_f), float(p_a), _ptr(keep_f_bits), _ptr(keep_a_bits),
                                           _ptr(dbg), _ptr(lg), _ptr(sc), _get_workspace(plan.ws_bytes, dev),
                                           plan.ws_bytes, _stream_ptr(dev)),
                   "mcmil_debug_proj_tc")
    return dbg, lg, sc
