"""Single-bag call loop (the reference's bs == 1 serving pattern, infer.py:187-196) for ncu launch lists and
CUDA-event timing: N=1024, T=100, shared attention, MCHeadRunner.  Reports the device time per bag (CUDA events), the
host time spent issuing the calls, and the device time of the same calls replayed from a CUDA graph (no host in the
loop): which of host and device bounds the back-to-back rate."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import mcmil_b200 as mm                                    # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 30
path = sys.argv[2] if len(sys.argv) > 2 else "auto"       # (label only)
graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
dev = torch.device("cuda")
w = mm.HeadWeights(bench.make_state_dict(0, True), dev)
H = torch.relu(torch.randn(16 * 1024, 512, device=dev))
runner = mm.MCHeadRunner(w, 1024, 100)
for i in range(10):
    runner.run(H[(i % 16) * 1024:(i % 16 + 1) * 1024], seed=i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t0 = time.perf_counter()
for i in range(calls):
    runner.run(H[(i % 16) * 1024:(i % 16 + 1) * 1024], seed=i)
t_issue = time.perf_counter() - t0
e1.record()
torch.cuda.synchronize()
print("path %s: %.2f us per bag back to back over %d calls (host issue time %.2f us per call)"
      % (path, e0.elapsed_time(e1) / calls * 1e3, calls, t_issue / calls * 1e6))
if graph:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        K = 32
        with torch.cuda.graph(g, stream=s):
            for i in range(K):
                runner.run(H[(i % 16) * 1024:(i % 16 + 1) * 1024], seed=i)
        for _ in range(3):
            g.replay()
        s.synchronize()
        e0.record(s)
        for _ in range(10):
            g.replay()
        e1.record(s)
        s.synchronize()
    print("path %s: %.2f us per bag replayed from a CUDA graph of %d calls" % (path, e0.elapsed_time(e1) / (10 * K) * 1e3, K))
