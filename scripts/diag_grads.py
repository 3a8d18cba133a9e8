"""Diagnostic: per-tensor gradient error of the CUDA step vs the fp64 manual oracle, both precisions."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import sac_manual as smn, sac_port as sp
from _golden import core_config, rel_l2
from distributed_sac_b200 import _lib
from distributed_sac_b200.core import SacCore

which = sys.argv[1] if len(sys.argv) > 1 else "MS"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = {"LL": sp.ll_spec, "VS": sp.vs_spec, "MS": sp.ms_spec}[which]()
p = sp.init_params(spec, seed=seed)
b = sp.synthetic_batch(spec, seed=100 + seed)
gen = torch.Generator().manual_seed(77)
e1 = torch.randn(spec.batch, spec.act_dim, generator=gen); e2 = torch.randn(spec.batch, spec.act_dim, generator=gen)
man = smn.ManualLearner(spec, {k: v.double() for k, v in p.items()}, None, np.float64)
I = man.update_SAC(*b, e1, e2)
ref = {**I["critic_grads"], **I["actor_grads"], "log_alpha": I["alpha_grad"]}
for prec in (0, 1):
    core = SacCore(core_config(spec, precision=prec), 0, seed=0)
    core.set_named(p)
    core.step(*b, e1, e2)
    g = core.get_named(_lib.GRADS)
    worst = max((rel_l2(g[k].reshape(np.asarray(ref[k]).shape), torch.from_numpy(np.asarray(ref[k]))), k) for k in ref)
    bad = [k for k in ref if rel_l2(g[k].reshape(np.asarray(ref[k]).shape), torch.from_numpy(np.asarray(ref[k]))) > 2e-5]
    print(f"--- {which} seed={seed} precision={prec}: worst {worst[0]:.2e} ({worst[1]}); tensors > 2e-5: {bad}")
    print("d_action", rel_l2(core.debug("d_action").reshape(spec.batch, -1), I["d_action"]), "d_head", rel_l2(core.debug("d_head").reshape(spec.batch, -1), I["d_head"]))
    core.close()
