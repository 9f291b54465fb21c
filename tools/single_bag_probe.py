"""Single-bag call loop (the reference's bs == 1 serving pattern, infer.py:187-196) for ncu launch lists and
CUDA-event timing: N=1024, T=100, shared attention, MCHeadRunner."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import mcmil_b200 as mm                                    # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 30
path = sys.argv[2] if len(sys.argv) > 2 else "auto"
dev = torch.device("cuda")
w = mm.HeadWeights(bench.make_state_dict(0, True), dev)
H = torch.relu(torch.randn(16 * 1024, 512, device=dev))
mm.set_reduce_path(path)
runner = mm.MCHeadRunner(w, 1024, 100)
for i in range(10):
    runner.run(H[(i % 16) * 1024:(i % 16 + 1) * 1024], seed=i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(calls):
    runner.run(H[(i % 16) * 1024:(i % 16 + 1) * 1024], seed=i)
e1.record()
torch.cuda.synchronize()
print("path %s: %.2f us per bag back to back over %d calls" % (path, e0.elapsed_time(e1) / calls * 1e3, calls))
