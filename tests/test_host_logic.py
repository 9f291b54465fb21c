"""CPU tests: the C-ABI library loads and exports every symbol include/mcmil_b200.h declares (no
compute calls without a GPU), host-side sharding logic, and the world_size-2 gloo path of the
Welford merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_cabi_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from mcmil_b200 import _lib
    declared = _lib.declared_symbols()
    assert len(declared) >= 12
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mcmil_version() >= 100


def test_no_cpu_fallback():
    import mcmil_b200 as mm
    with pytest.raises(RuntimeError):
        mm.HeadWeights({}, device="cpu")
    m = mm.MultiHeadGatedAttentionMIL(pretrained=False)
    with pytest.raises(RuntimeError):
        m.mc_inference(torch.zeros(1, 2, 3, 32, 32), N=2, device="cpu")
    # every other entry point refuses host tensors as well: nothing routes around the CUDA library
    with pytest.raises(RuntimeError):
        mm.aux_pairwise_loss(torch.zeros(2, 2, 8), True)
    with pytest.raises(RuntimeError):
        m.forward_eval_fused(torch.zeros(1, 2, 3, 32, 32))
    with pytest.raises(RuntimeError):
        mm.mc_head(None, torch.zeros(4, 512), 2)


def test_eval_forward_on_cpu_is_the_plain_torch_graph():
    """forward() only routes through the fused head for CUDA inputs in eval mode under no_grad; on the host it is
    the reference's torch graph (the training path), including the auxiliary loss (model.py:243-248)."""
    import mcmil_b200 as mm
    from oracle import gamil_oracle as G
    sd = G.make_weights(3, 2, True)
    m = mm.MultiHeadGatedAttentionMIL(pretrained=False)
    m.feature_extractor = torch.nn.Flatten()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    m.eval()
    H = G.make_features(11, 40)
    with torch.no_grad():
        Y, A, aux = m(torch.from_numpy(H).view(1, 40, 512, 1, 1), targets=torch.tensor([1]))
    o = G.forward_oracle(sd, H)
    assert np.abs(Y[0].numpy() - o["Y"]).max() < 5e-6 and np.abs(A[0].numpy() - o["A"]).max() < 2e-7
    assert abs(float(aux) - float(G.aux_pairwise_loss(o["A"][1], o["A"][0], True))) < 1e-6


def test_module_state_dict_matches_reference_schema():
    import mcmil_b200 as mm
    from oracle import gamil_oracle as G
    for shared in (True, False):
        m = mm.MultiHeadGatedAttentionMIL(pretrained=False, shared_attention=shared, num_classes=2)
        head = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("feature_extractor")}
        want = {k: v.shape for k, v in G.make_weights(0, 2, shared).items()}
        assert head == want
        # torch forward (training path, not accelerated) keeps the reference's 3-tuple
        m.feature_extractor = torch.nn.Flatten()
        m.eval()
        Y, A, aux = m(torch.rand(1, 6, 512, 1, 1))
        assert Y.shape == (1, 2) and A.shape == (1, 2, 6) and aux is None


def test_lpt_and_mc_shard():
    from mcmil_b200 import distributed as D
    rng = np.random.default_rng(0)
    lens = rng.integers(200, 3001, 256)
    for world in (1, 2, 4, 8):
        parts = D.lpt_assign(lens, world)
        assert sorted(i for p in parts for i in p) == list(range(256))
        loads = [sum(-(-int(lens[i]) // 128) for i in p) for p in parts]
        assert max(loads) - min(loads) <= 24          # one largest bag at most
    for T, world in ((1000, 8), (100, 3), (7, 8), (5, 1)):
        seen = []
        for r in range(world):
            t0, n = D.mc_shard(T, r, world)
            seen += list(range(t0, t0 + n))
        assert seen == list(range(T))


def test_welford_pack_unpack_cpu():
    from mcmil_b200 import distributed as D
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.random((30, 50)) * 1e-3)
    packed = None
    for part in x.split(10):
        mean = part.mean(0)
        pk = D.welford_pack(mean.float(), ((part - mean) ** 2).sum(0).float(), part.shape[0])
        packed = pk if packed is None else packed + pk
    mean, m2 = D.welford_unpack(packed, 50)
    assert torch.allclose(mean.double(), x.mean(0), rtol=1e-6)
    assert torch.allclose(m2.double(), ((x - x.mean(0)) ** 2).sum(0), rtol=1e-4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcmil_b200 import distributed as D
    rng = np.random.default_rng(123)                      # every rank builds the same population
    x = torch.from_numpy(rng.random((T, n)) * 1e-3)
    p = torch.from_numpy(rng.random((T, 2)))
    t0, Tl = D.mc_shard(T, rank, world)
    xs, ps = x[t0:t0 + Tl], p[t0:t0 + Tl]
    (am, pm), (aq, pq), total = D.allreduce_welford(
        [xs.mean(0).float(), ps.mean(0).float()],
        [((xs - xs.mean(0)) ** 2).sum(0).float(), ((ps - ps.mean(0)) ** 2).sum(0).float()], Tl)
    ok = (total == T
          and torch.allclose(am.double(), x.mean(0), rtol=1e-5)
          and torch.allclose(aq.double(), ((x - x.mean(0)) ** 2).sum(0), rtol=1e-3)
          and torch.allclose(pm.double(), p.mean(0), rtol=1e-5)
          and torch.allclose(pq.double(), ((p - p.mean(0)) ** 2).sum(0), rtol=1e-3))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_allreduce_welford_gloo_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, 101, 333, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_auxiliary_loss_surface():
    """The drop-in exposes the reference's `auxiliary_loss` attribute (model.py:149-151) with both loss types
    (model.py:415-438); values against the closed forms."""
    import mcmil_b200 as mm
    from mcmil_b200.model import AuxiliaryLoss
    from oracle import gamil_oracle as G
    m = mm.MultiHeadGatedAttentionMIL(pretrained=False)
    assert (m.auxiliary_loss.loss_type, m.auxiliary_loss.margin, m.auxiliary_loss.scale) == ("pairwise", 1.0, 0.5)
    assert not [k for k in m.state_dict() if "auxiliary" in k]          # no parameters: the state_dict schema is unchanged
    g = torch.Generator().manual_seed(0)
    a, b = torch.softmax(torch.randn(1, 50, generator=g), -1), torch.softmax(torch.randn(1, 50, generator=g), -1)
    for pos in (True, False):
        want = G.aux_pairwise_loss(a.numpy()[0], b.numpy()[0], pos, margin=1.0, scale=1.0)
        assert abs(float(AuxiliaryLoss("pairwise", 1.0, 1.0)(a, b, pos)) - float(want)) < 1e-6
    cs = float((a * b).sum() / (a.norm() * b.norm()))
    assert abs(float(AuxiliaryLoss("cosine")(a, b, True)) - cs) < 1e-6
    assert abs(float(AuxiliaryLoss("cosine")(a, b, False)) - (1 - cs)) < 1e-6
    with pytest.raises(ValueError):
        AuxiliaryLoss("nope")(a, b, True)


def test_bench_reference_arm_uses_the_reference_module_when_present():
    """bench.py --impl reference: the unmodified reference module (baseline/_ref or /root/reference) when it is there,
    else the port; with the same torch seed both give the same (Y, A) — the port stands in faithfully."""
    import bench
    from oracle import torch_port as TP
    head = bench.CpuHead(True, 2)
    assert head.kind in ("reference", "port")
    H = torch.relu(torch.randn(32, 512, generator=torch.Generator().manual_seed(3)))
    torch.manual_seed(11)
    Y, A = head.run(H, 4)
    assert Y.shape == (4, 1, 2) and A.shape == (4, 1, 2, 32)
    if head.kind == "reference":
        torch.manual_seed(11)
        Yp, Ap = TP.mc_head_torch(head.sd, H, 4, 0.1, 0.1)
        assert torch.equal(Y, Yp) and torch.equal(A, Ap)


def test_patcher_grid_must_fit_the_image_and_follows_tiles_assignment():
    import mcmil_b200 as mm
    pt = mm.ImagePatcher(patch_size=32, overlap=0.5)
    pt.get_tiles(100, 80)
    pt._check_image(100, 80)
    with pytest.raises(ValueError):
        pt._check_image(90, 80)              # grid built for a larger image
    with pytest.raises(RuntimeError):
        mm.ImagePatcher()._check_image(10, 10)
    src = pt._tiles_src
    pt.tiles = pt.tiles.copy()               # the reference's dataset assigns patcher.tiles directly
    assert pt._tiles_src is src and pt._tiles_src is not pt.tiles   # stale until the next use refreshes it


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (what the driver runs next to our arm): one JSON line with the contract's keys,
    the CPU arm described in `cpu_baseline`, `e2e` repeating the line's own value with zero copies."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in rec, key
    assert rec["impl"] == "reference" and rec["unit"] == "bags/s" and rec["higher_is_better"] is True
    assert rec["cpu_baseline"]["kind"] in ("reference", "port") and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["value"] == rec["value"] and rec["e2e"]["h2d_bytes_per_step"] == 0


def _build_c_client(out):
    """gcc line of examples/c_client.c (plain C99 against include/mcmil_b200.h, the CUDA runtime and the shipped .so)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "montecarlo-gated-mil_b200", "lib")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O2", "-I" + os.path.join(root, "include"),
           "-I/usr/local/cuda/include", os.path.join(root, "examples", "c_client.c"), "-L" + lib, "-lmcmil_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-lm", "-Wl,-rpath," + lib, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", out]
    return subprocess.run(cmd, capture_output=True, text=True)


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """include/mcmil_b200.h is the drop-in boundary: it must compile as C99 (no C++-isms, no torch types) and a C
    program that uses it must link against the shipped library without any other dependency than the CUDA runtime."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "montecarlo-gated-mil_b200", "lib", "libmcmil_b200.so")):
        import __graft_entry__ as g
        g.build()
    r = _build_c_client(str(tmp_path / "c_client"))
    assert r.returncode == 0, r.stderr
