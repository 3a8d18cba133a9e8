import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from distributed_sac_b200.core import Replay, SacCore
wl = sys.argv[1] if len(sys.argv) > 1 else "LL"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for prec in (0, 1):
    core = SacCore(bench.core_config(wl, R, prec), 0, seed=1)
    ring = Replay(core, 1 << 18, "device", seed=2); ring.fill_synthetic(1 << 18, seed=3)
    names = ["sample", "ingest"] + [n for n, _ in core.profile_step(ring, 2)][1:]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        tl = core.graph_timeline(ring, 300)
    print(f"=== {wl} R={R} precision={prec}: step {sum(tl):.1f} us")
    for n, t in zip(names, tl):
        print(f"   {n:28s} {t:7.2f}")
    ring.close(); core.close()
