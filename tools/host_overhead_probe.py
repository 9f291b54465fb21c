"""Where the host time of a single-bag `mc_head` call goes (cProfile over 2000 calls, device work not waited for)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import mcmil_b200 as mm                                    # noqa: E402

dev = torch.device("cuda")
w = mm.HeadWeights(bench.make_state_dict(0, True), dev)
H = torch.relu(torch.randn(1024, 512, device=dev))
for i in range(50):
    mm.mc_head(w, H, 100, seed=i)
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
for i in range(n):
    mm.mc_head(w, H, 100, seed=i)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("mc_head: %.1f us per call issued (%.1f us per call with the device drained)" % ((t1 - t0) / n * 1e6, (time.perf_counter() - t0) / n * 1e6))
pr = cProfile.Profile()
pr.enable()
for i in range(n):
    mm.mc_head(w, H, 100, seed=i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
