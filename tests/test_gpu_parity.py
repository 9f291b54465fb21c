"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every call goes through the C-ABI
of montecarlo-gated-mil_b200/lib/libmcmil_b200.so (ctypes), never through the oracle.

Tolerances are north_star's: probabilities within 1e-3 relative, attention mean / variance within
1e-4 absolute (fp32 accumulation) — plus tighter relative bounds so regressions stay visible
(attention values are O(1/N)).
"""
import numpy as np
import pytest
import torch

from oracle import gamil_oracle as G
from oracle import philox as PX
from tests.cases import Case, FwdCase, forward_golden_names, golden_names

pytestmark = pytest.mark.gpu

PROB_RTOL = 1e-3
ATTN_ATOL = 1e-4
# relative bounds per implementation: (attention mean, attention M2 relative to its maximum, logits abs / scale).
# About 3-6x the maxima measured over every golden case on a B200 (tools/measure_tolerances.py, round 2):
#   simt_fp32  4.6e-7 / 1.5e-6 / 4.3e-7        tcgen05  1.0e-3 (the "peaky" x5 attention weights; 1.1e-4 at the default
#   weights) / 6.5e-4 / 1.1e-4 — fp16 operand rounding of features and weights, tanh.approx in the epilogue.
REL = {"simt_fp32": (5e-6, 2e-5, 5e-6), "tcgen05": (3e-3, 4e-3, 6e-4)}
IMPLS = ["simt_fp32", "tcgen05"]


@pytest.fixture(scope="module")
def mm():
    import mcmil_b200
    assert torch.cuda.is_available()
    return mcmil_b200


def _bits(keep, dev):
    return torch.from_numpy(PX.pack_bits(keep).view(np.int32)).to(dev)


def _check(res, ref, impl, T, A_ref=None, A_stride=1):
    """res: MCHeadResult for ONE bag; ref: dict with Y (T,C), prob_mean/m2, attn_mean/m2 (fp64-ish)."""
    rel_mean, rel_m2, y_abs = REL[impl]
    Y = res.Y[0].double().cpu().numpy()
    assert np.isfinite(Y).all()
    P = G.finish_stats(Y, np.zeros((T, Y.shape[1], 1)))["P"]
    Pref = G.finish_stats(np.asarray(ref["Y"], np.float64), np.zeros((T, Y.shape[1], 1)))["P"]
    assert np.abs(P / Pref - 1).max() < PROB_RTOL, "per-sample probabilities"
    assert np.abs(Y - ref["Y"]).max() < y_abs * max(1.0, np.abs(ref["Y"]).max()), "per-sample logits"
    pm = res.prob_mean[0].double().cpu().numpy()
    assert np.abs(pm / ref["prob_mean"] - 1).max() < PROB_RTOL, "mean probability"
    pq = res.prob_m2[0].double().cpu().numpy()
    assert np.abs(pq - ref["prob_m2"]).max() < 1e-4 + 2e-2 * np.abs(ref["prob_m2"]).max()
    am = res.attn_mean.double().cpu().numpy()
    aq = res.attn_m2.double().cpu().numpy()
    assert np.abs(am - ref["attn_mean"]).max() < ATTN_ATOL, "attention mean (abs)"
    assert np.abs(aq / max(T - 1, 1) - ref["attn_m2"] / max(T - 1, 1)).max() < ATTN_ATOL, "attention variance (abs)"
    assert np.abs(am / ref["attn_mean"] - 1).max() < rel_mean, "attention mean (rel)"
    scale = np.abs(ref["attn_m2"]).max()
    if T > 1 and scale > 0:
        assert np.abs(aq - ref["attn_m2"]).max() < rel_m2 * scale, "attention M2 (rel to max)"
    if A_ref is not None and res.A is not None:
        A = res.A.double().cpu().numpy()[::A_stride]
        assert np.abs(A - A_ref).max() < ATTN_ATOL
        assert np.abs(A / A_ref - 1).max() < max(rel_mean * 3, 1e-4), "per-sample attention (rel)"


def test_library_loaded_and_exports(mm):
    from mcmil_b200 import _lib
    lib = _lib.load()
    for name in _lib.declared_symbols():
        assert hasattr(lib, name)
    assert lib.mcmil_version() >= 100


def test_exported_masks_match_numpy_philox(mm):
    dev = torch.device("cuda")
    cu = np.array([0, 70, 71, 200], np.int32)
    T, C, seed = 5, 3, 0x1234567890ABCDEF
    fb, ab = mm.export_masks(T, cu, C, seed, 0.1, 0.25, t_offset=3, bag_offset=2, device=dev)
    fb, ab = fb.cpu().numpy().view(np.uint32), ab.cpu().numpy().view(np.uint32)
    for b in range(3):
        n = cu[b + 1] - cu[b]
        kf = PX.feature_keep(seed, 2 + b, 3, T, n, 0.1)
        ka = PX.attn_keep(seed, 2 + b, 3, T, n, C, 0.25)
        got_f = PX.unpack_bits(fb[:, cu[b]:cu[b + 1]], 512)
        assert np.array_equal(got_f, kf)
        got_a = PX.unpack_bits(ab, cu[-1])[:, :, cu[b]:cu[b + 1]]
        assert np.array_equal(got_a, ka)


@pytest.mark.parametrize("impl", IMPLS)
def test_philox7_mode_vs_oracle(mm, impl):
    """philox_rounds=7: exported masks equal the numpy Philox4x32-7 stream and the head matches the
    fp64 oracle driven by those masks."""
    dev = torch.device("cuda")
    N, T, C, seed = 200, 6, 2, 31337
    sd = G.make_weights(61, C, True)
    Hn = G.make_features(700, N)
    kf = PX.feature_keep(seed, 0, 2, T, N, 0.1, rounds=7)
    ka = PX.attn_keep(seed, 0, 2, T, N, C, 0.1, rounds=7)
    fb, ab = mm.export_masks(T, N, C, seed, 0.1, 0.1, t_offset=2, device=dev, philox_rounds=7)
    assert np.array_equal(PX.unpack_bits(fb.cpu().numpy().view(np.uint32), 512), kf)
    assert np.array_equal(PX.unpack_bits(ab.cpu().numpy().view(np.uint32), N), ka)
    assert not np.array_equal(kf, PX.feature_keep(seed, 0, 2, T, N, 0.1, rounds=10))
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    res = mm.mc_head(w, torch.from_numpy(Hn).to(dev), T, seed=seed, t_offset=2, return_attention=True, impl=impl,
                     philox_rounds=7)
    ref = G.mc_head_oracle(sd, Hn, kf, ka, 0.1, 0.1)
    _check(res, ref, impl, T, A_ref=ref["A"])
    with pytest.raises(ValueError):
        mm.mc_head(w, torch.from_numpy(Hn).to(dev), T, philox_rounds=8)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", golden_names())
def test_golden_injected_masks(mm, name, impl):
    """Masks injected (Philox-regenerated, or the reference's own torch draws for *_native):
    CUDA path vs the reference outputs stored in tests/golden/."""
    c = Case(name)
    dev = torch.device("cuda")
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in c.sd.items()}, dev)
    H = torch.from_numpy(c.H).to(dev)
    res = mm.mc_head(w, H, c.T, p_f=c.p_f, p_a=c.p_a, keep_f_bits=_bits(c.keep_f, dev),
                     keep_a_bits=_bits(c.keep_a, dev), return_attention=True, impl=impl)
    _check(res, c.ref, impl, c.T, A_ref=c.ref["A"].astype(np.float64), A_stride=c.A_stride)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", [n for n in golden_names() if not n.endswith("native")])
def test_golden_inkernel_philox(mm, name, impl):
    """Same cases with the masks drawn by the in-kernel Philox (no injection): must land on the
    same reference outputs, because the golden masks ARE that Philox stream."""
    c = Case(name)
    dev = torch.device("cuda")
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in c.sd.items()}, dev)
    H = torch.from_numpy(c.H).to(dev)
    res = mm.mc_head(w, H, c.T, seed=c.mseed, p_f=c.p_f, p_a=c.p_a, t_offset=c.t0, bag_offset=c.bag,
                     return_attention=True, impl=impl)
    assert res.launches == (1 if c.shared else c.C) + 2        # projection launch(es) + rows + columns
    _check(res, c.ref, impl, c.T, A_ref=c.ref["A"].astype(np.float64), A_stride=c.A_stride)


def test_reduction_dispatch_variants_agree(mm):
    """The row kernel has several forms (warp per row in one chunk / in 1024-patch chunks merged on the fly / CTA per
    row with 128-, 256-, 512- or 1024-patch chunks per warp) and the column kernel splits the MC samples over 1..16
    CTAs per tile; which one runs depends on the batch shape.  The same bags packed into batches that select
    different forms give the same outputs up to fp32 summation-order rounding, and every form is run-to-run
    deterministic."""
    dev = torch.device("cuda")
    sd = G.make_weights(72, 2, True)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    T = 40
    lens = [700, 1500, 90, 3000, 5000]       # alone: one chunk per warp of 128 / 256 / 128 / 512 patches, 1024-patch chunks
    g = torch.Generator(device=dev).manual_seed(5)
    Hs = [torch.relu(torch.randn(n, 512, generator=g, device=dev)) for n in lens]
    # (a) each bag alone: few rows -> CTA-per-row kernels, samples split over several CTAs per tile
    alone = [mm.mc_head(w, h, T, seed=3, bag_ids=[i], return_attention=True) for i, h in enumerate(Hs)]
    # (b) packed together with enough filler bags that the warp-per-row kernel (chunked: max_n > 1024) runs
    fill = [torch.relu(torch.randn(64, 512, generator=g, device=dev)) for _ in range(30)]
    allb = Hs + fill
    cu = np.concatenate([[0], np.cumsum([int(h.shape[0]) for h in allb])])
    packed = mm.mc_head(w, torch.cat(allb), T, seed=3, cu_seqlens=cu, bag_ids=list(range(len(allb))), return_attention=True)
    again = mm.mc_head(w, torch.cat(allb), T, seed=3, cu_seqlens=cu, bag_ids=list(range(len(allb))), return_attention=True)
    assert torch.equal(packed.Y, again.Y) and torch.equal(packed.attn_m2, again.attn_m2) and torch.equal(packed.A, again.A)
    for i, a in enumerate(alone):
        sl = slice(int(cu[i]), int(cu[i + 1]))
        assert (a.Y[0] - packed.Y[i]).abs().max().item() <= 2e-6 * max(1.0, packed.Y.abs().max().item())
        assert (a.A / packed.A[:, :, sl] - 1).abs().max().item() < 3e-6
        assert (a.attn_mean / packed.attn_mean[:, sl] - 1).abs().max().item() < 3e-6
        assert (a.attn_m2 - packed.attn_m2[:, sl]).abs().max().item() <= 2e-5 * a.attn_m2.abs().max().item() + 1e-12
        assert (a.prob_mean[0] - packed.prob_mean[i]).abs().max().item() < 1e-6
        assert (a.prob_m2[0] - packed.prob_m2[i]).abs().max().item() <= 1e-5 * max(1.0, a.prob_m2.abs().max().item())


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("shared", [True, False])
def test_ragged_batch_vs_oracle(mm, impl, shared):
    """Packed variable-length batch (incl. a 1-patch bag and tile-boundary sizes) vs the fp64 oracle."""
    dev = torch.device("cuda")
    lens = [200, 77, 333, 1, 128, 129, 64]
    ids = [5, 0, 9, 2, 3, 4, 11]
    T, C, seed = 7, 2, 99
    sd = G.make_weights(21, C, shared)
    Hs = [G.make_features(300 + i, n) for i, n in enumerate(lens)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    res = mm.mc_head(w, torch.from_numpy(np.concatenate(Hs)).to(dev), T, seed=seed, cu_seqlens=cu, bag_ids=ids,
                     t_offset=4, return_attention=True, impl=impl)
    for b, n in enumerate(lens):
        kf = PX.feature_keep(seed, ids[b], 4, T, n, 0.1)
        ka = PX.attn_keep(seed, ids[b], 4, T, n, C, 0.1)
        ref = G.mc_head_oracle(sd, Hs[b], kf, ka, 0.1, 0.1)
        sl = slice(cu[b], cu[b + 1])
        one = mm.MCHeadResult(res.Y[b:b + 1], res.prob_mean[b:b + 1], res.prob_m2[b:b + 1], res.attn_mean[:, sl],
                              res.attn_m2[:, sl], res.A[:, :, sl], T, cu[b:b + 2] - cu[b])
        _check(one, ref, impl, T, A_ref=ref["A"])


@pytest.mark.parametrize("shared", [True, False])
def test_config2_tcgen05_vs_fp32_and_properties(mm, shared):
    """BASELINE config 2 (N=1024, T=100) at full size: tensor-core path vs the fp32 CUDA-core path
    on the same Philox stream, plus size-independent properties."""
    dev = torch.device("cuda")
    N, T = 1024, 100
    sd = G.make_weights(31, 2, shared)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    H = torch.from_numpy(G.make_features(400, N)).to(dev)
    a = mm.mc_head(w, H, T, seed=5, impl="tcgen05", return_attention=True)
    b = mm.mc_head(w, H, T, seed=5, impl="simt_fp32", return_attention=True)
    Pa, Pb = a.probs().double(), b.probs().double()
    assert (Pa / Pb - 1).abs().max().item() < PROB_RTOL
    assert (a.attn_mean - b.attn_mean).abs().max().item() < ATTN_ATOL
    assert (a.attn_mean / b.attn_mean - 1).abs().max().item() < 3e-3
    assert (a.attn_var() - b.attn_var()).abs().max().item() < ATTN_ATOL
    assert (a.attn_m2 - b.attn_m2).abs().max().item() < 3e-2 * b.attn_m2.abs().max().item()
    for r in (a, b):
        assert (r.A.sum(-1) - 1).abs().max().item() < 1e-4          # every softmax row sums to one
        assert (r.attn_mean.sum(-1) - 1).abs().max().item() < 1e-4
        assert (r.prob_mean.sum(-1) - 1).abs().max().item() < 1e-5
        assert (r.attn_m2 >= 0).all() and (r.prob_m2 >= 0).all()
        # statistics are the statistics of the returned samples
        assert (r.A.mean(0) - r.attn_mean).abs().max().item() < 1e-7
        assert (r.A.var(0, unbiased=True) - r.attn_var(1)).abs().max().item() < 1e-9
        assert (r.probs()[0].mean(0) - r.prob_mean[0]).abs().max().item() < 1e-6
    # run-to-run determinism (no float atomics anywhere)
    a2 = mm.mc_head(w, H, T, seed=5, impl="tcgen05", return_attention=True)
    assert torch.equal(a.Y, a2.Y) and torch.equal(a.attn_m2, a2.attn_m2) and torch.equal(a.A, a2.A)
    # a different seed gives different samples
    a3 = mm.mc_head(w, H, T, seed=6, impl="tcgen05")
    assert not torch.equal(a.Y, a3.Y)


def test_config3_ragged_batch_full_size_properties(mm):
    """BASELINE config 3 at full size (256 bags, N ~ U{200..3000}, T=50, ~410 k packed rows): size-independent
    properties of the tensor-core path on the whole batch, tensor-core vs fp32 path on a slice of its bags, and
    independence of a bag's result from the batch it is packed into (masks are keyed by the global bag id)."""
    dev = torch.device("cuda")
    lens = [int(v) for v in np.random.default_rng(0).integers(200, 3001, 256)]
    cu = np.concatenate([[0], np.cumsum(lens)])
    T = 50
    sd = G.make_weights(33, 2, True)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    g = torch.Generator(device=dev).manual_seed(7)
    H = torch.relu(torch.randn(int(cu[-1]), 512, generator=g, device=dev))
    r = mm.mc_head(w, H, T, seed=11, cu_seqlens=cu)
    assert torch.isfinite(r.Y).all() and r.Y.shape == (256, T, 2)
    seg = torch.from_numpy(np.repeat(np.arange(256), lens)).to(dev)
    sums = torch.zeros(2, 256, device=dev).index_add_(1, seg, r.attn_mean)
    assert (sums - 1).abs().max().item() < 2e-4                       # mean attention of every bag sums to one
    assert (r.prob_mean.sum(-1) - 1).abs().max().item() < 1e-5
    assert (r.attn_m2 >= 0).all() and (r.prob_m2 >= 0).all()
    assert (r.probs().mean(1) - r.prob_mean).abs().max().item() < 1e-6
    r2 = mm.mc_head(w, H, T, seed=11, cu_seqlens=cu)
    assert torch.equal(r.Y, r2.Y) and torch.equal(r.attn_m2, r2.attn_m2)   # run-to-run determinism
    # bags 100..103 on their own, with their global ids: same numbers as inside the big batch
    b0, b1 = 100, 104
    sub = mm.mc_head(w, H[cu[b0]:cu[b1]].contiguous(), T, seed=11, cu_seqlens=cu[b0:b1 + 1] - cu[b0],
                     bag_ids=list(range(b0, b1)))
    # (masks and projection are identical bit for bit; the reductions may take another path for another batch shape,
    # so the outputs agree to fp32 summation-order rounding)
    assert (sub.Y - r.Y[b0:b1]).abs().max().item() <= 2e-6 * max(1.0, r.Y.abs().max().item())
    assert (sub.attn_mean / r.attn_mean[:, cu[b0]:cu[b1]] - 1).abs().max().item() < 2e-6
    ref = mm.mc_head(w, H[cu[b0]:cu[b1]].contiguous(), T, seed=11, cu_seqlens=cu[b0:b1 + 1] - cu[b0],
                     bag_ids=list(range(b0, b1)), impl="simt_fp32")
    assert (sub.probs() / ref.probs() - 1).abs().max().item() < PROB_RTOL
    assert (sub.attn_mean - ref.attn_mean).abs().max().item() < ATTN_ATOL
    assert (sub.attn_mean / ref.attn_mean - 1).abs().max().item() < REL["tcgen05"][0]
    # two bags of the full-size batch against the fp64 oracle driven by the exported Philox masks (T = 50, chunked)
    A_big = mm.mc_head(w, H, T, seed=11, cu_seqlens=cu, return_attention=True)
    for b in (100, 255):
        n = lens[b]
        fb, ab = mm.export_masks(T, n, 2, 11, 0.1, 0.1, bag_offset=b, device=dev)
        kf = PX.unpack_bits(fb.cpu().numpy().view(np.uint32), 512)
        ka = PX.unpack_bits(ab.cpu().numpy().view(np.uint32), n)
        o = G.mc_head_oracle(sd, H[cu[b]:cu[b + 1]].cpu().numpy(), kf, ka, 0.1, 0.1)
        sl = slice(int(cu[b]), int(cu[b + 1]))
        one = mm.MCHeadResult(A_big.Y[b:b + 1], A_big.prob_mean[b:b + 1], A_big.prob_m2[b:b + 1], A_big.attn_mean[:, sl],
                              A_big.attn_m2[:, sl], A_big.A[:, :, sl], T, np.array([0, n]))
        _check(one, o, "tcgen05", T, A_ref=o["A"])


def test_config4_large_bag_properties(mm):
    """BASELINE config 4 shape (one bag, N=16384) at a reduced sample count: tensor-core vs fp32 path, and the
    MC-sample split (two shards of the global sample range, merged with the additive Welford form) against the
    single call."""
    from mcmil_b200 import distributed as MD
    dev = torch.device("cuda")
    N, T = 16384, 24
    sd = G.make_weights(35, 2, True)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    g = torch.Generator(device=dev).manual_seed(9)
    H = torch.relu(torch.randn(N, 512, generator=g, device=dev))
    a = mm.mc_head(w, H, T, seed=3)
    b = mm.mc_head(w, H, T, seed=3, impl="simt_fp32")
    assert (a.probs() / b.probs() - 1).abs().max().item() < PROB_RTOL
    assert (a.attn_mean - b.attn_mean).abs().max().item() < ATTN_ATOL
    assert (a.attn_mean / b.attn_mean - 1).abs().max().item() < REL["tcgen05"][0]
    assert (a.attn_mean.sum(-1) - 1).abs().max().item() < 2e-4
    # the first 8 samples at the full N = 16384 against the fp64 oracle driven by the exported Philox masks
    To = 8
    fb, ab = mm.export_masks(To, N, 2, 3, 0.1, 0.1, device=dev)
    kf = PX.unpack_bits(fb.cpu().numpy().view(np.uint32), 512)
    ka = PX.unpack_bits(ab.cpu().numpy().view(np.uint32), N)
    del fb, ab
    o = G.mc_head_oracle(sd, H.cpu().numpy(), kf, ka, 0.1, 0.1, t_chunk=2)
    for impl in IMPLS:
        _check(mm.mc_head(w, H, To, seed=3, return_attention=True, impl=impl), o, impl, To, A_ref=o["A"])
    parts = [mm.mc_head(w, H, 12, seed=3, t_offset=t0) for t0 in (0, 12)]
    assert torch.equal(torch.cat([p.Y for p in parts], 1), a.Y)        # same global samples, bit for bit
    packed = sum(MD.welford_pack(torch.cat([p.attn_mean.reshape(-1), p.prob_mean.reshape(-1)]),
                                 torch.cat([p.attn_m2.reshape(-1), p.prob_m2.reshape(-1)]), 12) for p in parts)
    mean, m2 = MD.welford_unpack(packed, 2 * N + 2)
    assert (mean[:2 * N].view(2, N) - a.attn_mean).abs().max().item() < 1e-7
    assert (m2[:2 * N].view(2, N) - a.attn_m2).abs().max().item() < 1e-6 * max(1.0, a.attn_m2.abs().max().item()) + 1e-9


def test_sample_sharding_equals_single_call(mm):
    """MC-sample split (config 4 mechanics on one GPU): two half-calls with t_offset + the additive
    Welford merge reproduce the single call — masks are keyed by the global sample index."""
    from mcmil_b200 import distributed as D
    dev = torch.device("cuda")
    N, T = 500, 40
    sd = G.make_weights(41, 2, True)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    H = torch.from_numpy(G.make_features(500, N)).to(dev)
    full = mm.mc_head(w, H, T, seed=3)
    packed = None
    Ys = []
    for r in range(4):
        t0, Tl = D.mc_shard(T, r, 4)
        part = mm.mc_head(w, H, Tl, seed=3, t_offset=t0)
        Ys.append(part.Y)
        pk = D.welford_pack(torch.cat([part.attn_mean.reshape(-1), part.prob_mean.reshape(-1)]),
                            torch.cat([part.attn_m2.reshape(-1), part.prob_m2.reshape(-1)]), Tl)
        packed = pk if packed is None else packed + pk
    n = full.attn_mean.numel() + full.prob_mean.numel()
    mean, m2 = D.welford_unpack(packed, n)
    assert int(round(packed[0].item())) == T
    assert torch.equal(torch.cat(Ys, dim=1), full.Y)
    ref_mean = torch.cat([full.attn_mean.reshape(-1), full.prob_mean.reshape(-1)])
    ref_m2 = torch.cat([full.attn_m2.reshape(-1), full.prob_m2.reshape(-1)])
    assert (mean / ref_mean - 1).abs().max().item() < 1e-5
    assert (m2 - ref_m2).abs().max().item() < 1e-3 * ref_m2.abs().max().item() + 1e-12


def test_module_dropin_interface(mm):
    """Same module surface as the reference: state_dict keys load, mc_inference returns the
    (Y (T,1,C), A (T,1,C,N)) 2-tuple (model.py:328), statistics ride along."""
    import torch.nn as nn
    dev = torch.device("cuda")
    for shared in (True, False):
        sd = G.make_weights(51, 2, shared)
        m = mm.MultiHeadGatedAttentionMIL(pretrained=False, shared_attention=shared)
        head_keys = {k for k in m.state_dict() if not k.startswith("feature_extractor")}
        assert head_keys == set(sd.keys())
        m.feature_extractor = nn.Flatten()
        missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not unexpected and not missing
        N, T = 96, 9
        Hn = G.make_features(600, N)
        x = torch.from_numpy(Hn).view(1, N, 512, 1, 1)
        Y, A = m.mc_inference(x, N=T, device="cuda", seed=77)
        assert Y.shape == (T, 1, 2) and A.shape == (T, 1, 2, N) and Y.is_cuda
        assert all(mod.training for mod in m.modules() if isinstance(mod, nn.Dropout))   # model.py:268-271
        assert not m.training
        ref = G.mc_head_oracle(sd, Hn, PX.feature_keep(77, 0, 0, T, N, 0.1), PX.attn_keep(77, 0, 0, T, N, 2, 0.1), 0.1, 0.1)
        assert np.abs(A[:, 0].double().cpu().numpy() - ref["A"]).max() < ATTN_ATOL
        P = torch.softmax(Y[:, 0].double(), -1).cpu().numpy()
        assert np.abs(P / ref["P"] - 1).max() < PROB_RTOL
        assert len(m.mc_inference(x, N=T, device="cuda", legacy_tuple=True)) == 3
        assert m.last_result.count == T
        assert m.auxiliary_loss.scale == 0.5 and m.auxiliary_loss.margin == 1.0 and m.auxiliary_loss.loss_type == "pairwise"
        Ys, As = m.mc_inference_serial(x, N=3, device="cuda")                            # model.py:330-401
        assert Ys.shape == (3, 1, 2) and As.shape == (3, 1, 2, N)
        with pytest.raises(RuntimeError):
            m.mc_inference(torch.zeros(2, 4, 512, 1, 1), N=2, device="cuda")            # bs != 1, model.py:309


@pytest.mark.parametrize("name", forward_golden_names())
def test_forward_eval_and_aux_loss_vs_reference(mm, name):
    """SURVEY §8f-2 through the C-ABI: one all-keep pass of the fused head = the reference's eval-mode forward
    (model.py:211-253); mcmil_aux_pairwise_loss = scale * AuxiliaryLoss (model.py:243-248, 318-326)."""
    c = FwdCase(name)
    dev = torch.device("cuda")
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in c.sd.items()}, dev)
    H = torch.from_numpy(c.H).to(dev)
    Y, A = mm.head_forward_eval(w, H)
    Yn, An = Y[0].double().cpu().numpy(), A.double().cpu().numpy()
    P = np.exp(Yn - Yn.max()); P /= P.sum()
    Pr = np.exp(c.ref["Y"].astype(np.float64) - c.ref["Y"].max()); Pr /= Pr.sum()
    assert np.abs(P / Pr - 1).max() < PROB_RTOL
    assert np.abs(Yn - c.ref["Y"]).max() < 2e-3 * max(1.0, np.abs(c.ref["Y"]).max())
    assert np.abs(An - c.ref["A"]).max() < ATTN_ATOL and np.abs(An / c.ref["A"] - 1).max() < 1e-2
    assert abs(An.sum(-1) - 1).max() < 1e-5
    if c.C < 2:
        return
    A1 = A.view(1, c.C, c.N).contiguous()
    for pos, key in ((True, "aux_pos"), (False, "aux_neg")):
        got = float(mm.aux_pairwise_loss(A1, pos)[0, 0])
        assert abs(got - float(c.ref[key])) < 1e-4 + 1e-3 * abs(float(c.ref[key])), key
        # the kernel itself, fed the reference's attention: fp32 rounding only
        Aref = torch.from_numpy(c.ref["A"]).to(dev).view(1, c.C, c.N).contiguous()
        assert abs(float(mm.aux_pairwise_loss(Aref, pos)[0, 0]) - float(c.ref[key])) < 2e-7
    # per-pass losses of mc_inference (model.py:318-326) with the same Philox masks the golden file used
    res = mm.mc_head(w, H, c.T, seed=c.mseed, return_attention=True)
    for pos, key in ((True, "mc_aux_pos"), (False, "mc_aux_neg")):
        got = mm.aux_pairwise_loss(res.A, pos)[0].double().cpu().numpy()
        assert np.abs(got - c.ref[key]).max() < 1e-4 + 1e-3 * np.abs(c.ref[key]).max(), key


def test_module_forward_eval_fused_matches_torch_path(mm):
    """Module level: eval-mode forward() routes through the fused head and agrees with the plain-torch graph
    (the training path) on Y, A_all and the auxiliary loss; packed batches (bs > 1) included."""
    import torch.nn as nn
    dev = torch.device("cuda")
    for shared, bs, n in ((True, 1, 150), (False, 3, 70)):
        m = mm.MultiHeadGatedAttentionMIL(pretrained=False, shared_attention=shared)
        m.feature_extractor = nn.Flatten()
        m.load_state_dict({k: torch.from_numpy(v) for k, v in G.make_weights(61, 2, shared).items()}, strict=False)
        m = m.to(dev).eval()
        x = torch.from_numpy(np.stack([G.make_features(700 + b, n) for b in range(bs)])).view(bs, n, 512, 1, 1).to(dev)
        with torch.no_grad():
            Yf, Af, _ = m(x)
            m.fused_eval = False
            Yt, At, _ = m(x)
            m.fused_eval = True
        assert Yf.shape == Yt.shape == (bs, 2) and Af.shape == At.shape == (bs, 2, n)
        assert (Yf - Yt).abs().max() < 2e-3 and (Af - At).abs().max() < ATTN_ATOL
        if bs == 1:                                   # targets.item() (model.py:244) needs one target
            for tgt in (1, 0):
                with torch.no_grad():
                    lf = m(x, targets=torch.tensor([tgt], device=dev))[2]
                    m.fused_eval = False
                    lt = m(x, targets=torch.tensor([tgt], device=dev))[2]
                    m.fused_eval = True
                assert abs(float(lf) - float(lt)) < 1e-4
            Y3 = m.mc_inference(x, N=5, device="cuda", targets=torch.tensor([1]), legacy_tuple=True, seed=3)
            assert len(Y3) == 3 and len(Y3[2]) == 5 and all(t.numel() == 1 for t in Y3[2])
        m.train()
        assert m(x)[0].requires_grad                  # training keeps the torch graph


@pytest.mark.parametrize("impl", IMPLS)
def test_fp16_features_entry_is_bit_identical(mm, impl):
    """mcmil_head_forward_f16: half-precision features give exactly the results of the same values passed as fp32
    (the tensor-core path rounds to fp16 anyway; the fp32 path converts exactly)."""
    dev = torch.device("cuda")
    sd = G.make_weights(9, 2, False)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    lens = [130, 77, 1]
    cu = np.concatenate([[0], np.cumsum(lens)])
    H16 = torch.from_numpy(np.concatenate([G.make_features(70 + i, n) for i, n in enumerate(lens)])).to(dev).half()
    a = mm.mc_head(w, H16, 6, seed=4, cu_seqlens=cu, return_attention=True, impl=impl)
    b = mm.mc_head(w, H16.float(), 6, seed=4, cu_seqlens=cu, return_attention=True, impl=impl)
    for x, y in ((a.Y, b.Y), (a.A, b.A), (a.attn_mean, b.attn_mean), (a.attn_m2, b.attn_m2), (a.prob_m2, b.prob_m2)):
        assert torch.equal(x, y)
    with pytest.raises(ValueError):
        mm.mc_head(w, H16.double(), 6, seed=4, cu_seqlens=cu)


def test_feature_range_fp16_bound(mm):
    """The tensor-core path rounds features to fp16 (the reference is fp32, model.py:276-281): large-but-representable
    features (1e3 ... 6e4) still match the oracle; beyond 65504 `validate=True` raises (and impl='simt_fp32' works)."""
    dev = torch.device("cuda")
    N, T = 96, 6
    sd = G.make_weights(81, 2, True)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    kf, ka = PX.feature_keep(5, 0, 0, T, N, 0.1), PX.attn_keep(5, 0, 0, T, N, 2, 0.1)
    for scale in (1e3, 1.2e4):
        Hn = G.make_features(900, N, scale=scale)
        assert 1e3 < Hn.max() < 65504
        ref = G.mc_head_oracle(sd, Hn, kf, ka, 0.1, 0.1)
        res = mm.mc_head(w, torch.from_numpy(Hn).to(dev), T, seed=5, return_attention=True, validate=True)
        assert torch.isfinite(res.Y).all()
        # The fp16 rounding of a feature is RELATIVE (2^-11): with features of this size the pre-activations are in
        # the hundreds, the gates saturate, and the few units near a zero crossing carry an absolute error that grows
        # with the feature scale -> attention within ~1e-3 absolute (a few percent relative) instead of the 1e-4 / 1e-3 of
        # O(1-10) ResNet features (the range north_star's tolerance is stated for; see DESIGN.md "Precision").
        A, Ar = res.A.double().cpu().numpy(), ref["A"]
        assert np.abs(A - Ar).max() < 2e-3 and np.abs(A / Ar - 1).max() < 0.1
        assert np.abs(A.sum(-1) - 1).max() < 1e-5
        assert np.abs(res.Y[0].double().cpu().numpy() - ref["Y"]).max() < 5e-3 * np.abs(ref["Y"]).max()
        exact = mm.mc_head(w, torch.from_numpy(Hn).to(dev), T, seed=5, return_attention=True, impl="simt_fp32")
        assert np.abs(exact.A.double().cpu().numpy() / Ar - 1).max() < 1e-4        # the fp32 path holds the tight bound
    Hbig = torch.from_numpy(G.make_features(900, N, scale=4e4)).to(dev)
    assert Hbig.max() > 65504
    with pytest.raises(ValueError):
        mm.mc_head(w, Hbig, T, seed=5, validate=True)
    with pytest.raises(ValueError):
        mm.mc_head(w, Hbig * float("inf"), T, seed=5, validate=True, impl="simt_fp32")
    ok = mm.mc_head(w, Hbig, T, seed=5, validate=True, impl="simt_fp32")          # fp32 path: no fp16 bound
    assert torch.isfinite(ok.Y).all()
    m = mm.MultiHeadGatedAttentionMIL(pretrained=False)
    m.feature_extractor = torch.nn.Flatten()
    m.validate_features = True
    with pytest.raises(ValueError):
        m.mc_inference(Hbig.view(1, N, 512, 1, 1), N=2, device="cuda")


def test_two_devices_in_one_process(mm):
    """One host process driving two GPUs (the dynamic shared memory opt-in of the kernels is a per-device function
    attribute): same results on both devices."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sd = G.make_weights(91, 2, True)
    Hn = G.make_features(901, 300)
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
        r = mm.mc_head(w, torch.from_numpy(Hn).to(dev), 9, seed=4, return_attention=True)
        torch.cuda.synchronize(dev)
        outs.append((r.Y.cpu(), r.attn_mean.cpu(), r.A.cpu()))
    for x, y in zip(*outs):
        assert torch.equal(x, y)


def test_runner_equals_mc_head(mm):
    """MCHeadRunner (pre-created plan / outputs, one C-ABI call per bag) returns bit-identical results."""
    dev = torch.device("cuda")
    sd = G.make_weights(5, 2, True)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    r = mm.MCHeadRunner(w, 300, 7, return_attention=True)
    for seed in (1, 2):
        H = torch.from_numpy(G.make_features(40 + seed, 300)).to(dev)
        a = r.run(H, seed=seed)
        b = mm.mc_head(w, H, 7, seed=seed, return_attention=True)
        for x, y in ((a.Y, b.Y), (a.A, b.A), (a.attn_mean, b.attn_mean), (a.attn_m2, b.attn_m2), (a.prob_mean, b.prob_mean)):
            assert torch.equal(x, y)
    with pytest.raises(ValueError):
        r.run(torch.zeros(10, 512, device=dev))
    # throughput mode: k private streams, each projection kernel on 1/k of the SMs -> the same bits
    rk = mm.MCHeadRunner(w, 300, 7, return_attention=True, n_streams=3)
    Hs = [torch.from_numpy(G.make_features(50 + i, 300)).to(dev) for i in range(7)]
    outs = []
    for i, Hi in enumerate(Hs):
        res = rk.run(Hi, seed=i)
        if i >= 4:                                     # results stay valid for n_streams further calls
            res.stream.synchronize()
            outs.append((i, res.Y.clone(), res.A.clone(), res.attn_m2.clone()))
    rk.synchronize()
    for i, Y, A, q in outs:
        b = mm.mc_head(w, Hs[i], 7, seed=i, return_attention=True)
        assert torch.equal(Y, b.Y) and torch.equal(A, b.A) and torch.equal(q, b.attn_m2)
    with pytest.raises(ValueError):
        mm.MCHeadRunner(w, 300, 7, n_streams=0)


def test_extractor_modes_agree(mm):
    """channels-last / CUDA-graph execution of the torch extractor (whole-bag batch-stat BN kept)
    gives the eager features; the graph is reused across bags of the same shape."""
    dev = torch.device("cuda")
    torch.manual_seed(0)
    m = mm.MultiHeadGatedAttentionMIL(pretrained=False)
    m.apply(mm.deactivate_batchnorm)
    m.to(dev).eval()
    bags = [torch.rand(1, 12, 3, 224, 224, device=dev) for _ in range(2)]
    with torch.no_grad():
        ref = [m.extract_features(b) for b in bags]
        for mode in ("channels_last", "graph"):
            m.extractor_mode = mode
            for b, r in zip(bags, ref):
                H = m.extract_features(b)
                assert H.shape == r.shape == (12, 512)
                assert (H - r).abs().max() < 2e-3 * max(1.0, float(r.abs().max())), mode
        assert len(m._extractor_runner.graphs) == 1
        Y, A = m.mc_inference(bags[0], N=4, device="cuda", seed=1)
        assert Y.shape == (4, 1, 2) and A.shape == (4, 1, 2, 12)


def test_argument_errors(mm):
    dev = torch.device("cuda")
    sd = G.make_weights(1, 2, True)
    w = mm.HeadWeights({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    H = torch.zeros(10, 512, device=dev)
    with pytest.raises(RuntimeError):
        mm.mc_head(w, H.cpu(), 3)
    with pytest.raises(ValueError):
        mm.mc_head(w, torch.zeros(10, 256, device=dev), 3)
    with pytest.raises(ValueError):
        mm.mc_head(w, H, 0)
    with pytest.raises(ValueError):
        mm.mc_head(w, H, 3, cu_seqlens=[0, 4, 4, 10])       # empty bag
    with pytest.raises(ValueError):
        mm.mc_head(w, H, 3, p_f=1.5)
    with pytest.raises(ValueError):
        mm.mc_head(w, H.double(), 3)
    r = mm.mc_head(w, H, 3, p_f=1.0, p_a=1.0)                # nn.Dropout(p=1): everything dropped
    assert torch.isfinite(r.Y).all() and (r.Y.abs().max().item() == 0.0)
    assert (r.attn_mean - 0.1).abs().max().item() < 1e-6     # all logits 0 -> uniform attention


def test_config5_pipeline_small(mm):
    """infer.py path at a small size with the REAL ResNet-18 extractor (batch-statistics BatchNorm):
    tiling -> features once -> fused head -> attention-map statistics; head checked against the oracle
    on the extractor's own features."""
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = mm.MultiHeadGatedAttentionMIL(pretrained=False, shared_attention=False)
    model.apply(mm.deactivate_batchnorm)
    model.to(dev).eval()
    h, w, T = 700, 520, 8
    from oracle import patcher_oracle as PO
    img = torch.from_numpy(PO.synth_image(3, 3, h, w)).to(dev)
    pt = mm.ImagePatcher(patch_size=224, overlap=0.5, bag_size=-1, empty_thresh=0.5)
    pt.get_tiles(h, w)
    bag, idx, _ = pt.convert_img_to_bag(img)
    assert bag.shape[0] == len(idx) >= 4
    Y, A = model.mc_inference(bag.unsqueeze(0), N=T, device="cuda", seed=5)
    assert Y.shape == (T, 1, 2) and A.shape == (T, 1, 2, len(idx))
    with torch.no_grad():
        H = model.extract_features(bag.unsqueeze(0))
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items() if not k.startswith("feature_extractor")}
    n = len(idx)
    ref = G.mc_head_oracle(sd, H.cpu().numpy(), PX.feature_keep(5, 0, 0, T, n, 0.1), PX.attn_keep(5, 0, 0, T, n, 2, 0.1),
                           0.1, 0.1)
    assert np.abs(A[:, 0].double().cpu().numpy() - ref["A"]).max() < ATTN_ATOL
    P = torch.softmax(Y[:, 0].double(), -1).cpu().numpy()
    assert np.abs(P / ref["P"] - 1).max() < PROB_RTOL
    st = pt.attention_map_stats(A, idx, (h, w))
    mean_ref, std_ref = PO.attention_map_stats(ref["A"], pt.tiles, idx, (1, h, w))
    assert np.abs(st.mean_map().cpu().numpy() - mean_ref[:, 0]).max() < 1e-3
    assert np.abs(st.std_map().cpu().numpy() - std_ref[:, 0]).max() < 1e-3


@pytest.mark.gpu
def test_c_client_runs_the_hot_path_without_python(mm, tmp_path):
    """examples/c_client.c: plain C against the C ABI (weights, plan, forward, Welford outputs); it checks finiteness,
    normalisation, run-to-run bit identity and tcgen05 vs the fp32 CUDA-core path itself and prints "c client ok"."""
    import subprocess
    from tests.test_host_logic import _build_c_client
    exe = str(tmp_path / "c_client")
    r = _build_c_client(exe)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "c client ok" in run.stdout, run.stdout + run.stderr
