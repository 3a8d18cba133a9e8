"""Helpers shared by the CPU (oracle) and GPU (product) parity tests."""
import json
import os

import numpy as np
import torch

import sac_port as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["ll_xavier_s2", "ll_ckpt_s3", "vs_small_s10", "ms_small_s5", "ms_small_unweighted_s3"]

# Parity tolerances: rel 1e-4 on losses / intermediates and
# per-tensor rel-L2 1e-4 on parameters, targets and Adam moments; log_alpha abs 1e-6.
REL = 1e-4


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.name = name
        self.spec = sp.SacSpec(**json.loads(str(z["spec"])))
        self.family = str(z["family"])
        self.n_steps = int(z["n_steps"])
        grab = lambda pre: {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}
        self.p_in, self.m_in, self.v_in = grab("p_in/"), grab("m_in/"), grab("v_in/")
        self.p_out, self.m_out, self.v_out = grab("p_out/"), grab("m_out/"), grab("v_out/")
        self.i0 = grab("i0/")
        self.batch = grab("batch/")
        self.eps_next, self.eps_cur = torch.from_numpy(z["eps_next"]), torch.from_numpy(z["eps_cur"])
        self.step_in, self.step_out = z["step_in"], z["step_out"]
        self.losses = z["losses"]

    def step_batch(self, i):
        return tuple(self.batch[k][i] for k in ("s", "a", "r", "s2", "d"))


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / den if den > 0 else (a - b).norm().item()


def rel_scalar(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def check_state(case, params, m, v, steps, tol=REL, what="state"):
    bad = []
    for k, ref in case.p_out.items():
        if k == "log_alpha":
            err = (params[k].double() - ref.double()).abs().max().item()
            if err > 1e-6:
                bad.append((k, "abs", err))
        else:
            err = rel_l2(params[k], ref)
            if err > tol:
                bad.append((k, "p", err))
    for k, ref in case.m_out.items():
        e1, e2 = rel_l2(m[k], ref), rel_l2(v[k], case.v_out[k])
        if e1 > tol:
            bad.append((k, "m", e1))
        if e2 > tol:
            bad.append((k, "v", e2))
    assert not bad, f"{what} mismatch in {case.name}: {bad[:8]}"
    assert tuple(int(x) for x in steps) == tuple(int(x) for x in case.step_out)


def core_config(spec, replicas=1, **kw):
    from distributed_sac_b200.core import CoreConfig
    d = dict(state_dim=spec.state_dim, act_dim=spec.act_dim, actor_hidden=list(spec.actor_hidden),
             critic_hidden=list(spec.critic_hidden), batch=spec.batch, num_tasks=spec.num_tasks,
             weighted_loss=spec.weighted_loss, replicas=replicas, gamma=spec.gamma, tau=spec.tau,
             reward_scale=spec.reward_scale, lr_actor=spec.lr_actor, lr_critic=spec.lr_critic,
             action_scale=spec.action_scale, beta1=spec.beta1, beta2=spec.beta2, adam_eps=spec.adam_eps)
    d.update(kw)
    return CoreConfig(**d)


class CareCase:
    """Fixture of the CARE(M) learner (oracle/care_port.py names)."""

    def __init__(self, name="care_small_s4"):
        import care_port as cp
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.name = name
        self.spec = cp.CareSpec(**json.loads(str(z["spec"])))
        self.n_steps = int(z["n_steps"])
        grab = lambda pre: {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}
        self.p_in, self.m_in, self.v_in = grab("p_in/"), grab("m_in/"), grab("v_in/")
        self.p_out, self.m_out, self.v_out = grab("p_out/"), grab("m_out/"), grab("v_out/")
        self.i0, self.batch = grab("i0/"), grab("batch/")
        self.eps_next, self.eps_cur = torch.from_numpy(z["eps_next"]), torch.from_numpy(z["eps_cur"])
        self.step_in, self.step_out, self.losses = z["step_in"], z["step_out"], z["losses"]

    step_batch = Case.step_batch


def care_core_config(spec, replicas=1, **kw):
    from distributed_sac_b200.core import CoreConfig
    d = dict(state_dim=spec.state_dim, act_dim=spec.act_dim, actor_hidden=list(spec.actor_hidden),
             critic_hidden=list(spec.critic_hidden), batch=spec.batch, num_tasks=spec.num_tasks,
             weighted_loss=spec.weighted_loss, replicas=replicas, gamma=spec.gamma, tau=spec.tau,
             reward_scale=spec.reward_scale, lr_actor=spec.lr_actor, lr_critic=spec.lr_critic,
             action_scale=spec.action_scale, beta1=spec.beta1, beta2=spec.beta2, adam_eps=spec.adam_eps,
             care=True, num_encoders=spec.num_encoders, mix_hidden=list(spec.mix_hidden), mix_out=spec.mix_out,
             ctx_in=spec.ctx_in, ctx_hidden=list(spec.ctx_hidden), ctx_out=spec.ctx_out, tau_se=spec.tau_se,
             care_original=not spec.modified, emb_dim=spec.emb_dim, lr_ctx=spec.lr_ctx)
    d.update(kw)
    return CoreConfig(**d)
