"""clock64() timeline of CTA (0,0,0) of every chain_kernel launch of one LL step (B200SAC_CHAIN_DBG=1)."""
import os
import sys
os.environ["B200SAC_CHAIN_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import torch
import bench
from distributed_sac_b200 import _lib
from distributed_sac_b200.core import Replay, SacCore
wl = sys.argv[1] if len(sys.argv) > 1 else "LL"
core = SacCore(bench.core_config(wl, 1, 0), 0, seed=1)
ring = Replay(core, 1 << 16, "device", seed=2)
ring.fill_synthetic(1 << 16, seed=3)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    core.step_sampled(ring, 24)
torch.cuda.synchronize()
names = [n for n, _ in core.profile_step(ring, 1)][1:]
out = torch.empty(2 * 64 * 16)
n = C.c_int64(0)
_lib.check(core.lib.b200sac_debug_read(core._h, b"chain_dbg", 0, C.c_void_p(out.data_ptr()), out.numel(), C.byref(n), None))
t = np.frombuffer(out.numpy().tobytes(), dtype=np.int64).reshape(16, 64)
for i, name in enumerate(names):
    k = int(t[i, 63])
    if k <= 0:
        continue
    d = t[i, :k] - t[i, 0]
    print(f"{name:32s} total {d[-1]:6d} cyc | " + " ".join(str(int(x)) for x in np.diff(t[i, :k])))
