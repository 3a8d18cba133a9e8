"""Two-step full-size CARE(M) chain under a given tile-selection env: CUDA vs the oracle port with the CUDA masks forced,
tensor by tensor after each step (which intermediate / gradient separates first)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import sac_port as sp
from _golden import FullCase, care_core_config, core_config, rel_l2, cuda_relu_masks, check_forced
from distributed_sac_b200 import _lib
from distributed_sac_b200.core import SacCore

name = sys.argv[1]
c = FullCase(name)
cfg = care_core_config(c.spec, precision=1) if c.care else core_config(c.spec, precision=1)
core = SacCore(cfg, 0, seed=0)
core.set_named(c.params)
port = c.make_port()
for i in range(c.n_steps):
    core.step(*c.batches[i], c.eps_next[i], c.eps_cur[i])
    forced = cuda_relu_masks(core, c.spec, 0, c.care)
    with sp.ReluTape(forced) as tape:
        o = port.update(*c.batches[i], c.eps_next[i], c.eps_cur[i], want_intermediates=True) if c.care else port.update_SAC(*c.batches[i], c.eps_next[i], c.eps_cur[i])
    flips = check_forced(tape, forced)
    print(f"step {i}: flips {flips}")
    inter = o
    for k in ("y", "q1", "q2", "a_next", "logp_next", "a_cur", "logp_cur", "qmin", "d_action", "d_head"):
        if k in inter:
            ref = torch.as_tensor(inter[k]).float()
            got = core.debug(k).reshape(ref.shape)
            print(f"   {k:12s} {rel_l2(got, ref):.2e}")
    g = core.get_named(_lib.GRADS)
    pg = {k: v.grad for k, v in port.p.items() if getattr(v, "grad", None) is not None} if hasattr(port, "p") else {}
    for k, v in g.items():
        if k in pg:
            e = rel_l2(v.reshape(pg[k].shape), pg[k])
            if e > 5e-5:
                print(f"   grad/{k:24s} {e:.2e}")
    st = port.params() if hasattr(port, "params") else {}
    got = core.get_named()
    bad = {k: float(f"{rel_l2(got[k].reshape(v.shape), v.detach()):.1e}") for k, v in st.items() if k in got and rel_l2(got[k].reshape(v.shape), v.detach()) > 5e-5}
    print("   params beyond 5e-5:", bad)
core.close()
