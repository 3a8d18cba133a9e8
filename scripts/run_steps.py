"""Run a few sampled steps of one workload (profiling target for ncu): python scripts/run_steps.py LL 0 [steps]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from distributed_sac_b200.core import Replay, SacCore
wl = sys.argv[1] if len(sys.argv) > 1 else "LL"
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
core = SacCore(bench.core_config(wl, 1, prec), 0, seed=1)
ring = Replay(core, 1 << 16, "device", seed=2)
ring.fill_synthetic(1 << 16, seed=3)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(n):
        core.step_sampled(ring, 1)
torch.cuda.synchronize()
print(core.read_losses(1))
