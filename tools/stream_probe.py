"""Throughput mode of MCHeadRunner: bags/s of single-bag calls round-robin over k private streams, each projection
kernel limited to 1/k of the SMs.  Host loop and CUDA-graph replay (no host in the loop)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import mcmil_b200 as mm                                    # noqa: E402

dev = torch.device("cuda")
w = mm.HeadWeights(bench.make_state_dict(0, True), dev)
H = torch.relu(torch.randn(16 * 1024, 512, device=dev))
calls = 400
import itertools
ks = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (4, 6, 8)
reserves = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0, 8, 16, 28)
for k, reserve in [(1, 0)] + list(itertools.product(ks, reserves)):
    r = mm.MCHeadRunner(w, 1024, 100, n_streams=k, reserve_sms=reserve)
    for i in range(40):
        r.run(H[(i % 16) * 1024:(i % 16 + 1) * 1024], seed=i)
    r.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(calls):
        r.run(H[(i % 16) * 1024:(i % 16 + 1) * 1024], seed=i, sync_input=False)
    t_issue = time.perf_counter() - t0
    r.synchronize(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("streams %d reserve %d: %.2f us per bag (%.0f bags/s), host issue %.2f us per call" % (k, reserve, dt / calls * 1e6, calls / dt, t_issue / calls * 1e6), flush=True)
    if k > 1:
        # the same work without the host: every private stream replays a graph of its own calls
        graphs = []
        for slot in r.slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(slot.stream):
                with torch.cuda.graph(g, stream=slot.stream):
                    for i in range(16):
                        code = r.lib.mcmil_head_forward(r.w._h, slot.plan._h, H[(i % 16) * 1024:].data_ptr(), 0, 0, i, *slot.tail,
                                                        slot.stream.cuda_stream)
                        assert code == 0
            graphs.append((g, slot.stream))
        for g, st in graphs:
            with torch.cuda.stream(st):
                g.replay()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            for g, st in graphs:
                with torch.cuda.stream(st):
                    g.replay()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("streams %d reserve %d: %.2f us per bag replayed from per-stream CUDA graphs" % (k, reserve, dt / (5 * 16 * k) * 1e6), flush=True)
    del r
