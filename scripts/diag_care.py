import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
from _golden import CareCase, care_core_config, rel_l2
from distributed_sac_b200.core import SacCore
for name in ("care_small_s4", "care_o_small_s4"):
    c = CareCase(name)
    core = SacCore(care_core_config(c.spec), 0, seed=0)
    core.set_named(c.p_in)
    back = core.get_named()
    bad = [k for k in c.p_in if not torch.equal(back[k].reshape(c.p_in[k].shape), c.p_in[k])]
    print(name, "round-trip mismatches:", bad)
    core.step(*c.step_batch(0), c.eps_next[0], c.eps_cur[0])
    for k, ref in c.i0.items():
        print("   ", k, f"{rel_l2(core.debug(k).reshape(ref.shape), ref):.3e}")
    print("    losses", core.read_losses(1)[0, 0].tolist(), c.losses[0].tolist())
    core.close()
