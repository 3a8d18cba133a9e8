"""One layer-chained LL step (tiny batch) -- a target small enough for compute-sanitizer's synccheck / racecheck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import sac_port as sp
from distributed_sac_b200.core import CoreConfig, SacCore
B = int(os.environ.get("MINI_B", "64"))
spec = sp.SacSpec(batch=B)
core = SacCore(CoreConfig(batch=B), 0, seed=0)
p = sp.init_params(spec, seed=1)
core.set_named(p)
b = sp.synthetic_batch(spec, seed=2)
g = torch.Generator().manual_seed(3)
e1, e2 = torch.randn(B, spec.act_dim, generator=g), torch.randn(B, spec.act_dim, generator=g)
for _ in range(2):
    core.step(*b, e1, e2)
print("losses", core.read_losses(1)[0, 0].tolist(), "launches", core.launches_per_step)
core.close()
