"""Two cores on the same full-size CARE(M) step, one with every tcgen05 launch forced to the unpaired 128-wide tile (the
reference configuration) and one with the cost-model plan: first intermediate / gradient tensor that separates them."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
from _golden import FullCase, care_core_config, core_config, rel_l2
from distributed_sac_b200 import _lib
from distributed_sac_b200.core import SacCore

name = sys.argv[1] if len(sys.argv) > 1 else "full_c10m_s2"
nsteps = int(os.environ.get("DIAG_STEPS", "1"))
envs = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[2:]]
c = FullCase(name)
outs = []
for env in envs:
    for k in ("B200SAC_TC_BN", "B200SAC_TC_PAIR128", "B200SAC_TC_OLDRULE", "B200SAC_TC_NO160"):
        os.environ.pop(k, None)
    os.environ.update(env)
    cfg = care_core_config(c.spec, precision=1) if c.care else core_config(c.spec, precision=1)
    core = SacCore(cfg, 0, seed=0)
    core.set_named(c.params)
    for i in range(nsteps):
        core.step(*c.batches[i], c.eps_next[i], c.eps_cur[i])
    d = {}
    names = ["y", "q1", "q2", "a_next", "logp_next", "a_cur", "logp_cur", "qmin", "d_action", "d_head", "dq_pi"]
    for l in range(3):
        names += [f"hA.{l}", f"hQ.{l}", f"hP.{l}", f"hT.{l}"]
    for n in names:
        try:
            d[n] = core.debug(n).clone()
        except Exception as e:
            pass
    g = core.get_named(_lib.GRADS)
    for k, v in g.items():
        d["grad/" + k] = v.clone()
    outs.append(d)
    core.close()
ref = outs[0]
for i, o in enumerate(outs[1:]):
    print("== vs", envs[i + 1])
    for k in ref:
        if k in o:
            e = rel_l2(o[k], ref[k])
            if e > 2e-5:
                print(f"   {k:28s} {e:.2e}")
    print("   (tensors compared:", len(ref), ")")
    B = c.spec.batch
    ra, rb = ref["dq_pi"].reshape(2, B)[0] != 0, o["dq_pi"].reshape(2, B)[0] != 0
    rows = torch.nonzero(ra != rb).reshape(-1).tolist()
    print("   rows whose min(Q1,Q2) routing differs:", rows)
    da, db = ref["d_action"].reshape(B, -1), o["d_action"].reshape(B, -1)
    per_row = (da - db).norm(dim=1) / da.norm(dim=1).clamp_min(1e-30)
    dh_a, dh_b = ref["d_head"].reshape(B, -1), o["d_head"].reshape(B, -1)
    pr = (dh_a - dh_b).norm(dim=1) / dh_a.norm(dim=1).clamp_min(1e-30)
    print("   d_head rows with relative diff > 1e-3:", int((pr > 1e-3).sum()), " > 1e-2:", int((pr > 1e-2).sum()), " median:", float(pr.median()))
    print("   d_action rows with relative diff > 1e-3:", int((per_row > 1e-3).sum()), " > 1e-2:", int((per_row > 1e-2).sum()), " median:", float(per_row.median()))
    for nm in ("hP.0", "hP.1", "hP.2"):
        ma, mb = ref[nm] > 0, o[nm] > 0
        print("   mask bits differing in", nm, ":", int((ma != mb).sum()))
    top = torch.topk(per_row, 5)
    print("   d_action per-row relative diff, top 5:", [(int(i), float(f"{v:.2e}")) for v, i in zip(top.values, top.indices)],
          " rel-L2 without routing-flipped rows:", float(((da - db)[ra == rb]).norm() / da[ra == rb].norm()))
