"""CPU tests: the oracle against the reference's stored outputs (tests/golden/) and
the Philox restatement against Random123's known-answer vectors."""
import numpy as np
import pytest

from oracle import gamil_oracle as G
from oracle import philox as PX
from tests.cases import Case, FwdCase, forward_golden_names, golden_names


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = PX.philox4x32([np.uint32(c) for c in ctr], key)
        assert tuple(int(g) for g in got) == want
    # Random123 kat_vectors, philox4x32-7 (the optional fast mode)
    kat7 = [
        ((0, 0, 0, 0), (0, 0), (0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x5207DDC2, 0x45165E59, 0x4D8EE751, 0x8C52F662)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0x4DFCCABA, 0x190A87F0, 0xC47362BA, 0xB6B5242A)),
    ]
    for ctr, key, want in kat7:
        got = PX.philox4x32([np.uint32(c) for c in ctr], key, rounds=7)
        assert tuple(int(g) for g in got) == want


def test_threshold_and_rates():
    assert PX.drop_threshold(0.1) == 3277
    assert PX.drop_threshold(0.0) == 0 and PX.drop_threshold(1.0) == 32768
    k = PX.feature_keep(123, 0, 0, 4, 256, 0.1)
    assert abs(k.mean() - (1 - 3277 / 32768)) < 2e-3
    assert PX.feature_keep(1, 0, 0, 2, 8, 0.0).all()
    assert not PX.feature_keep(1, 0, 0, 2, 8, 1.0).any()
    # distinct (bag, t) give distinct streams; t-offset consistency (MC-sample sharding)
    a = PX.feature_keep(5, 0, 0, 6, 16, 0.1)
    b = PX.feature_keep(5, 0, 3, 3, 16, 0.1)
    assert np.array_equal(a[3:], b)
    assert not np.array_equal(a[0], a[1])
    assert not np.array_equal(PX.feature_keep(5, 1, 0, 1, 16, 0.1), a[:1])
    ka = PX.attn_keep(5, 0, 0, 6, 40, 3, 0.1)
    assert np.array_equal(PX.attn_keep(5, 0, 2, 4, 40, 3, 0.1), ka[2:])


def test_feature_mask_stream_statistics():
    """The 16-elements-per-call mask stream behaves like iid Bernoulli(1 - p): drop rate at 4 sigma, and no
    correlation between the pairs that share Philox words (features of a chunk share the refinement byte, rows n
    and n^4 share the primary call, rows n..n+12 share the refinement call), neighbouring samples or K-slices."""
    T, N, p = 8, 512, 0.1
    k = PX.feature_keep(2024, 3, 0, T, N, p).astype(np.float64)            # 2.1 M elements
    q = 3277 / 32768
    n = k.size
    assert abs((1 - k.mean()) - q) < 4 * np.sqrt(q * (1 - q) / n)
    d = 1.0 - k - q                                                        # centred drop indicators

    def corr(a, b):
        return float((a * b).mean() / (q * (1 - q)))

    tol = 5 / np.sqrt(n / 2)                                               # ~5 sigma for 1 M pairs
    assert abs(corr(d[:, :, 0::2], d[:, :, 1::2])) < tol                   # neighbouring features (same chunk)
    assert abs(corr(d[:, :, :-8], d[:, :, 8:])) < tol                      # same lane of neighbouring chunks
    assert abs(corr(d[:, :, :-64], d[:, :, 64:])) < tol                    # neighbouring K-slices
    idx = np.arange(N)
    lo = idx[(idx & 4) == 0]
    assert abs(corr(d[:, lo, :], d[:, lo + 4, :])) < tol                   # rows n, n^4 (same primary call)
    lo8 = idx[(idx & 8) == 0]
    assert abs(corr(d[:, lo8, :], d[:, lo8 + 8, :])) < tol                 # rows n, n+8 (same refinement call)
    assert abs(corr(d[:-1], d[1:])) < tol                                  # consecutive MC samples
    # every feature position has the same rate (no byte of the call is biased)
    per_l = 1 - k.mean(axis=(0, 1))
    assert np.abs(per_l - q).max() < 6 * np.sqrt(q * (1 - q) / (T * N))


def test_philox_masks_are_statistically_equivalent_to_torch_dropout():
    """The counter-based masks cannot reproduce torch's generator stream; what must hold is that the MC
    statistics agree in distribution.  Same head, same bag: T samples with the reference's native torch dropout
    (torch port, bit-identical to the reference for a seed) vs T samples with the Philox masks — class-probability
    mean and per-patch attention mean agree within the Monte-Carlo error of the two estimates."""
    import torch
    from oracle import torch_port as TP
    N, T, C = 48, 1500, 2
    sd = G.make_weights(21, C, True, peaky=3.0)
    H = G.make_features(77, N)
    torch.manual_seed(5)
    Yt, At = TP.mc_head_torch(TP.sd_to_torch(sd), torch.from_numpy(H), T, 0.1, 0.1)
    a = G.finish_stats(Yt.reshape(T, C).double().numpy(), At.reshape(T, C, N).double().numpy())
    kf = PX.feature_keep(99, 0, 0, T, N, 0.1)
    ka = PX.attn_keep(99, 0, 0, T, N, C, 0.1)
    b = G.mc_head_oracle(sd, H, kf, ka, 0.1, 0.1)
    # z-scores of the difference of two independent sample means
    se_p = np.sqrt((a["prob_m2"] + b["prob_m2"]) / (T - 1) / T)
    assert (np.abs(a["prob_mean"] - b["prob_mean"]) / se_p).max() < 4.5
    se_a = np.sqrt((a["attn_m2"] + b["attn_m2"]) / (T - 1) / T)
    z = np.abs(a["attn_mean"] - b["attn_mean"]) / se_a
    assert z.max() < 5.0 and (z ** 2).mean() < 1.6          # ~chi-square with mean 1 over the 96 (head, patch) cells
    # and the spread itself matches: ratio of the variances of the probability samples
    ratio = (b["prob_m2"] / a["prob_m2"])
    assert np.all(ratio > 0.8) and np.all(ratio < 1.25)


def test_pack_bits_roundtrip():
    rng = np.random.default_rng(0)
    k = (rng.random((3, 5, 77)) < 0.9).astype(np.uint8)
    w = PX.pack_bits(k)
    assert w.shape == (3, 5, 3) and w.dtype == np.uint32
    assert np.array_equal(PX.unpack_bits(w, 77), k)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    """fp64 restatement vs the fp32 reference outputs stored by make_golden.py."""
    c = Case(name)
    out = G.mc_head_oracle(c.sd, c.H, c.keep_f, c.keep_a, c.p_f, c.p_a)
    assert np.abs(out["Y"] - c.ref["Y"]).max() < 5e-6
    assert np.abs(out["A"][::c.A_stride] - c.ref["A"]).max() < 2e-7
    assert np.abs(out["prob_mean"] - c.ref["prob_mean"]).max() < 2e-6
    assert np.abs(out["attn_mean"] - c.ref["attn_mean"]).max() < 1e-7
    rel = np.abs(out["attn_m2"] - c.ref["attn_m2"]) / (np.abs(c.ref["attn_m2"]) + 1e-30)
    assert rel.max() < 1e-3


@pytest.mark.parametrize("name", forward_golden_names())
def test_forward_and_aux_oracle_match_reference(name):
    """eval-mode forward (model.py:211-253) and auxiliary loss (model.py:243-248, 318-326,
    405-426) restatements vs the outputs of the live reference stored by make_golden_forward.py."""
    c = FwdCase(name)
    o = G.forward_oracle(c.sd, c.H)
    assert np.abs(o["Y"] - c.ref["Y"]).max() < 5e-6
    assert np.abs(o["A"] - c.ref["A"]).max() < 2e-7
    if c.C >= 2:
        assert abs(G.aux_pairwise_loss(o["A"][1], o["A"][0], True) - c.ref["aux_pos"]) < 1e-6
        assert abs(G.aux_pairwise_loss(o["A"][1], o["A"][0], False) - c.ref["aux_neg"]) < 1e-6
        kf = PX.feature_keep(c.mseed, 0, 0, c.T, c.N, 0.1)
        ka = PX.attn_keep(c.mseed, 0, 0, c.T, c.N, c.C, 0.1)
        A = G.mc_head_oracle(c.sd, c.H, kf, ka, 0.1, 0.1)["A"]
        assert np.abs(G.aux_pairwise_loss(A[:, 1], A[:, 0], True) - c.ref["mc_aux_pos"]).max() < 1e-6
        assert np.abs(G.aux_pairwise_loss(A[:, 1], A[:, 0], False) - c.ref["mc_aux_neg"]).max() < 1e-6


def test_welford_sumform_merge():
    rng = np.random.default_rng(1)
    x = rng.random((40, 7)) * 1e-3
    parts = np.split(x, 8)
    n = s1 = s2 = 0
    for p in parts:
        m = p.mean(0)
        a, b, c = G.welford_sumform(len(p), m, ((p - m) ** 2).sum(0))
        n, s1, s2 = n + a, s1 + b, s2 + c
    cnt, mean, m2 = G.welford_from_sumform(n, s1, s2)
    assert cnt == 40
    assert np.allclose(mean, x.mean(0), rtol=1e-12)
    assert np.allclose(m2, ((x - x.mean(0)) ** 2).sum(0), rtol=1e-9)
