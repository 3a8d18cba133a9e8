"""Generate tests/golden/*.npz from the UNMODIFIED reference learner (CPU).

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py [case ...]

Every case drives the real `Learner.update()` of the reference with
`memory.sample` replaced by a fixed minibatch and Normal.rsample's noise
injected (oracle/ref_harness.py), and stores inputs, per-step losses, first-step
intermediates (recomputed with the reference's own modules before the step) and
the final parameters / Adam state under canonical names (distributed_sac_b200/names.py).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh                      # noqa: E402
import sac_port as sp                         # noqa: E402
from distributed_sac_b200 import names as nm  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _modules(lrn, family):
    """[(reference module, {ref key: canonical key}, optimizer tag)]"""
    if family in ("LL", "VS"):
        na = len(lrn.actor.layer_intermediate) + 1
        nc = len(lrn.local_critic_1.layer_module) + 1
        return [
            (lrn.actor, nm.actor_key_map(family, na)),
            (lrn.local_critic_1, nm.critic_key_map(family, nc, 1)),
            (lrn.local_critic_2, nm.critic_key_map(family, nc, 2)),
            (lrn.target_critic_1, nm.critic_key_map(family, nc, 1, True)),
            (lrn.target_critic_2, nm.critic_key_map(family, nc, 2, True)),
        ]
    na = len(lrn.actor.mu_log_std_layer) // 2 + 1
    nc = len(lrn.local_critic.Q_function_1) // 2 + 1
    return [
        (lrn.actor, nm.actor_key_map("MS", na)),
        (lrn.local_critic, {**nm.critic_key_map("MS", nc, 1), **nm.critic_key_map("MS", nc, 2)}),
        (lrn.target_critic, {**nm.critic_key_map("MS", nc, 1, True), **nm.critic_key_map("MS", nc, 2, True)}),
    ]


def _named_params(lrn, family):
    """canonical name -> live reference Parameter"""
    out = {}
    for mod, kmap in _modules(lrn, family):
        live = dict(mod.named_parameters())
        for ref_key, canon in kmap.items():
            out[canon] = live[ref_key]
    out["log_alpha"] = lrn.log_alpha
    return out


def _opt_for(lrn, name):
    if name == "log_alpha":
        return lrn.log_alpha_optimizer
    return lrn.actor_optimizer if name.startswith("actor.") else lrn.critic_optimizer


def _adam_snapshot(lrn, named, spec):
    m, v, step = {}, {}, {"critic": 0, "actor": 0, "alpha": 0}
    for name in sp.param_names(spec, sp.TRAINABLE_NETS):
        st = _opt_for(lrn, name).state.get(named[name], {})
        tag = "alpha" if name == "log_alpha" else ("actor" if name.startswith("actor.") else "critic")
        if st:
            m[name], v[name] = st["exp_avg"].clone(), st["exp_avg_sq"].clone()
            step[tag] = int(st["step"])
        else:
            m[name] = torch.zeros_like(named[name].data)
            v[name] = torch.zeros_like(named[name].data)
    return m, v, np.array([step["critic"], step["actor"], step["alpha"]], np.int64)


def _spec_of(lrn, family, weighted):
    if family in ("LL", "VS"):
        ah = [l.out_features for l in lrn.actor.layer_intermediate]
        ch = [lrn.local_critic_1.first_layer.out_features] + [l.out_features for l in lrn.local_critic_1.layer_module[:-1]]
        return sp.SacSpec(state_dim=lrn.state_dim, act_dim=lrn.action_dim, actor_hidden=ah, critic_hidden=ch,
                          batch=lrn.batch_size, gamma=lrn.gamma, tau=lrn.tau, reward_scale=float(lrn.reward_scale),
                          lr_actor=lrn.lr_actor, lr_critic=lrn.lr_critic)
    ah = [m.out_features for m in lrn.actor.mu_log_std_layer if hasattr(m, "out_features")][:-1]
    ch = [m.out_features for m in lrn.local_critic.Q_function_1 if hasattr(m, "out_features")][:-1]
    return sp.SacSpec(state_dim=lrn.actor.state_dim, act_dim=lrn.actor.action_dim, actor_hidden=ah, critic_hidden=ch,
                      batch=lrn.batch_size, num_tasks=lrn.num_tasks, weighted_loss=weighted, gamma=lrn.gamma,
                      tau=lrn.tau, reward_scale=float(lrn.reward_scale), lr_actor=lrn.lr_actor, lr_critic=lrn.lr_critic)


def _load_ll_checkpoint(lrn, path):
    """Do what the reference's broken load_checkpoint intends (LL/learner.py:165-182)."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    lrn.local_critic_1.load_state_dict(ck["local_critic_1"])
    lrn.local_critic_2.load_state_dict(ck["local_critic_2"])
    lrn.target_critic_1.load_state_dict(ck["target_critic_1"])
    lrn.target_critic_2.load_state_dict(ck["target_critic_2"])
    lrn.actor.load_state_dict(ck["actor"])
    lrn.critic_optimizer.load_state_dict(ck["critic_optimizer"])
    lrn.actor_optimizer.load_state_dict(ck["actor_optimizer"])
    lrn.log_alpha.data = ck["log_alpha"].data.clone()
    lrn.log_alpha_optimizer.load_state_dict(ck["log_alpha_optimizer"])


def make_case(name, family, n_steps, cfg_overrides=None, seed=0, checkpoint=None, data_seed=1234):
    weighted = bool((cfg_overrides or {}).get("use_weighted_loss", family == "MS"))
    lrn, _ = rh.make_learner(family, cfg_overrides, seed=seed)
    if checkpoint:
        _load_ll_checkpoint(lrn, checkpoint)
    spec = _spec_of(lrn, family, weighted)
    named = _named_params(lrn, family)
    d = {"spec": json.dumps(spec.to_json()), "family": family, "n_steps": n_steps}
    for k, p in named.items():
        d["p_in/" + k] = p.detach().clone().numpy()
    m, v, step = _adam_snapshot(lrn, named, spec)
    for k in m:
        d["m_in/" + k], d["v_in/" + k] = m[k].numpy(), v[k].numpy()
    d["step_in"] = step

    g = torch.Generator().manual_seed(data_seed + 17)
    batches, eps_n, eps_c, losses = [], [], [], []
    for i in range(n_steps):
        batches.append(sp.synthetic_batch(spec, seed=data_seed + i))
        eps_n.append(torch.randn(spec.batch, spec.act_dim, generator=g))
        eps_c.append(torch.randn(spec.batch, spec.act_dim, generator=g))

    # first-step intermediates from the reference's own modules (pre-update)
    s, a, r, s2, dn = batches[0]
    with torch.no_grad(), rh.injected_eps([eps_n[0], eps_c[0]]):
        if family in ("LL", "VS"):
            alpha = lrn.log_alpha.exp()
            a2, lp2 = lrn.actor.get_action_log_prob(s2)
            qt = torch.min(lrn.target_critic_1(s2, a2), lrn.target_critic_2(s2, a2))
            q1, q2 = lrn.local_critic_1(s, a), lrn.local_critic_2(s, a)
            ac, lpc = lrn.actor.get_action_log_prob(s)
        else:
            alpha = lrn.get_log_alpha(s).exp()
            a2, lp2, _ = lrn.actor.get_action_log_prob_log_std(mtobss=s2)
            qt = torch.min(*lrn.target_critic(mtobss=s2, action=a2))
            q1, q2 = lrn.local_critic(mtobss=s, action=a)
            ac, lpc, _ = lrn.actor.get_action_log_prob_log_std(mtobss=s)
        y = lrn.reward_scale * r + lrn.gamma * (1 - dn) * (qt - alpha * lp2)
    for k, t in dict(y=y, q1=q1, q2=q2, a_next=a2, logp_next=lp2, a_cur=ac, logp_cur=lpc).items():
        d["i0/" + k] = t.numpy()

    for i in range(n_steps):
        lrn.memory.sample = (lambda b: (lambda: tuple(t.clone() for t in b)))(batches[i])
        with rh.injected_eps([eps_n[i], eps_c[i]]) as q:
            res = lrn.update()
            assert not q
        losses.append(list(res) + [float("nan")] * (3 - len(res)))
    d["losses"] = np.array(losses, np.float64)
    for j, key in enumerate(("s", "a", "r", "s2", "d")):
        d["batch/" + key] = np.stack([b[j].numpy() for b in batches])
    d["eps_next"] = np.stack([e.numpy() for e in eps_n])
    d["eps_cur"] = np.stack([e.numpy() for e in eps_c])
    for k, p in named.items():
        d["p_out/" + k] = p.detach().clone().numpy()
    m, v, step = _adam_snapshot(lrn, named, spec)
    for k in m:
        d["m_out/" + k], d["v_out/" + k] = m[k].numpy(), v[k].numpy()
    d["step_out"] = step
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  losses[0]={losses[0]}  losses[-1]={losses[-1]}")


CASES = {
    # full-size LunarLander learner, seeded Xavier init, distinct target nets, fresh Adam
    "ll_xavier_s2": dict(family="LL", n_steps=2, seed=0),
    # shipped trained checkpoint incl. its Adam state (step 55001)
    "ll_ckpt_s3": dict(family="LL", n_steps=3, seed=1,
                       checkpoint=os.path.join(rh.REF_ROOT, "saved_models/LunarLander_Distributed_SAC/checkpoint_165000.tar")),
    # VSAC learner with ragged small dims, 10 chained steps
    "vs_small_s10": dict(family="VS", n_steps=10, seed=2,
                         cfg_overrides=dict(batch_size=96, actor_hidden_dim=[64, 48, 32], critic_hidden_dim=[40, 72, 56])),
    # MTSAC one-hot head, weighted loss (== mean/B), per-task alpha
    "ms_small_s5": dict(family="MS", n_steps=5, seed=3,
                        cfg_overrides=dict(batch_size=120, actor=dict(actor_hidden_dim=[64, 64, 64]),
                                           critic=dict(critic_hidden_dim=[64, 64, 64]))),
    "ms_small_unweighted_s3": dict(family="MS", n_steps=3, seed=4,
                                   cfg_overrides=dict(batch_size=120, use_weighted_loss=False,
                                                      actor=dict(actor_hidden_dim=[48, 64, 32]),
                                                      critic=dict(critic_hidden_dim=[32, 64, 48]))),
}

if __name__ == "__main__":
    todo = sys.argv[1:] or list(CASES)
    for c in todo:
        make_case(c, **CASES[c])
