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


def _care_named(lrn):
    """canonical name -> live reference Parameter for the CARE(M) learner (oracle/care_port.py names)."""
    out = {}
    def mlp(seq, net):
        lin = [m for m in seq if hasattr(m, "out_features")]
        for i, m in enumerate(lin):
            out[f"{net}.{i}.weight"], out[f"{net}.{i}.bias"] = m.weight, m.bias
    mlp(lrn.actor.mu_log_std_layer, "actor")
    mlp(lrn.local_critic.Q_function_1, "q1"); mlp(lrn.local_critic.Q_function_2, "q2")
    mlp(lrn.target_critic.Q_function_1, "q1_target"); mlp(lrn.target_critic.Q_function_2, "q2_target")
    for pre, se in (("cse", lrn.local_critic.state_encoder), ("tse", lrn.target_critic.state_encoder)):
        mixl = [m for m in se.mixture_encoders.mixtureEncoders if hasattr(m, "W")]
        for l, m in enumerate(mixl):
            out[f"{pre}.mix.{l}.W"], out[f"{pre}.mix.{l}.b"] = m.W, m.b
        for name, seq in (("trunk", se.trunk), ("ctx", getattr(se, "mlp_context", []))):
            for j, m in enumerate([m for m in seq if hasattr(m, "out_features")]):
                out[f"{pre}.{name}.{j}.weight"], out[f"{pre}.{name}.{j}.bias"] = m.weight, m.bias
    out["embedding"] = lrn.context_encoder.embedding[0].weight
    if not lrn.use_modified_care:        # CARE(O): embedding = Sequential(Embedding, ReLU, header), then mlp
        ce = lrn.context_encoder
        lin = [m for m in ce.embedding[2] if hasattr(m, "out_features")] + [m for m in ce.mlp if hasattr(m, "out_features")]
        for j, m in enumerate(lin):
            out[f"cenc.{j}.weight"], out[f"cenc.{j}.bias"] = m.weight, m.bias
    out["log_alpha"] = lrn.log_alpha
    return out


def make_care_case(name, n_steps, cfg_overrides, seed=0, data_seed=4321, variant="C10"):
    import care_port as cp
    lrn, _ = rh.make_learner(variant, cfg_overrides, seed=seed)
    enc = dict(lrn.encoder_cfg)
    if variant == "C1":        # MT1_Distributed_CARE is CARE(O) with these two constants hard-coded
        lrn.use_modified_care = False                                      # (no such switch in MT1/src/learner.py)
        enc.setdefault("RoBERTa_embedding_dim", lrn.context_encoder.embedding[0].weight.shape[1])   # C1/context_encoder.py:30-36
        enc.setdefault("state_encoder_tau", 0.05)                          # C1/learner.py:311
    lin = lambda seq: [m.out_features for m in seq if hasattr(m, "out_features")]
    spec = cp.CareSpec(state_dim=lrn.actor.state_dim, act_dim=lrn.actor.action_dim, num_tasks=lrn.num_tasks,
                       actor_hidden=lin(lrn.actor.mu_log_std_layer)[:-1], critic_hidden=lin(lrn.local_critic.Q_function_1)[:-1],
                       batch=lrn.batch_size, num_encoders=int(enc["num_encoders"]), mix_hidden=list(enc["hidden_dims_mixtureEnc"]),
                       mix_out=int(enc["output_dim_mixtureEnc"]), ctx_in=int(enc["RoBERTa_embedding_dim"]),
                       ctx_hidden=list(enc["hidden_dims_contextEnc"]), ctx_out=int(enc["output_dim_contextEnc"]),
                       tau_se=float(enc["state_encoder_tau"]), weighted_loss=bool(lrn.use_modified_care), gamma=lrn.gamma,
                       tau=lrn.tau, reward_scale=float(lrn.reward_scale), lr_actor=lrn.lr_actor, lr_critic=lrn.lr_critic,
                       modified=bool(lrn.use_modified_care), emb_dim=int(enc["embedding_dim_contextEnc"]),
                       lr_ctx=float(enc["lr_contextEnc"]))
    named = _care_named(lrn)
    d = {"spec": json.dumps(spec.to_json()), "family": variant, "n_steps": n_steps}
    for k, p in named.items():
        d["p_in/" + k] = p.detach().clone().numpy()
    trainable = [k for k in named if not (k.startswith("tse.") or "_target" in k or k == "embedding")]
    for k in trainable:
        d["m_in/" + k] = np.zeros_like(d["p_in/" + k]); d["v_in/" + k] = np.zeros_like(d["p_in/" + k])
    d["step_in"] = np.zeros(3 if lrn.use_modified_care else 4, np.int64)
    g = torch.Generator().manual_seed(data_seed + 17)
    batches, eps_n, eps_c, losses = [], [], [], []
    for i in range(n_steps):
        batches.append(cp.synthetic_batch(spec, seed=data_seed + i))
        eps_n.append(torch.randn(spec.batch, spec.act_dim, generator=g))
        eps_c.append(torch.randn(spec.batch, spec.act_dim, generator=g))
    s, a, r, s2, dn = batches[0]
    with torch.no_grad(), rh.injected_eps([eps_n[0], eps_c[0]]):
        alpha = lrn.get_log_alpha(s).exp()
        z = lrn.context_encoder.forward(s)
        a2, lp2, _ = lrn.actor.get_action_log_prob_log_std(mtobss=s2, z_context=z)
        qt = torch.min(*lrn.target_critic.forward(mtobss=s2, z_context=z, action=a2))
        q1, q2 = lrn.local_critic.forward(mtobss=s, z_context=z, action=a)
        ac, lpc, _ = lrn.actor.get_action_log_prob_log_std(mtobss=s, z_context=z)
        y = lrn.reward_scale * r + lrn.gamma * (1 - dn) * (qt - alpha * lp2)
    for k, t in dict(y=y, q1=q1, q2=q2, a_next=a2, logp_next=lp2, a_cur=ac, logp_cur=lpc).items():
        d["i0/" + k] = t.numpy()
    for i in range(n_steps):
        lrn.memory.sample = (lambda b: (lambda: tuple(t.clone() for t in b)))(batches[i])
        with rh.injected_eps([eps_n[i], eps_c[i]]) as q:
            res = lrn.update()
            assert not q
        losses.append(list(res))
    d["losses"] = np.array(losses, np.float64)
    for j, key in enumerate(("s", "a", "r", "s2", "d")):
        d["batch/" + key] = np.stack([b[j].numpy() for b in batches])
    d["eps_next"] = np.stack([e.numpy() for e in eps_n]); d["eps_cur"] = np.stack([e.numpy() for e in eps_c])
    for k, p in named.items():
        d["p_out/" + k] = p.detach().clone().numpy()
    opts = {"critic": lrn.critic_optimizer, "actor": lrn.actor_optimizer, "alpha": lrn.log_alpha_optimizer,
            "ctx": lrn.context_encoder_optimizer}
    steps = {"critic": 0, "actor": 0, "alpha": 0, "ctx": 0}
    for k in trainable:
        tag = "alpha" if k == "log_alpha" else ("actor" if k.startswith("actor.") else ("ctx" if k.startswith("cenc.") else "critic"))
        st = opts[tag].state[named[k]]
        d["m_out/" + k], d["v_out/" + k] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
        steps[tag] = int(st["step"])
    d["step_out"] = np.array([steps["critic"], steps["actor"], steps["alpha"]] + ([] if lrn.use_modified_care else [steps["ctx"]]), np.int64)
    # the actor's tied encoder must equal the critic's after update() (learner.py:402)
    ase = dict(lrn.actor.state_encoder.named_parameters()); cse = dict(lrn.local_critic.state_encoder.named_parameters())
    assert all(torch.equal(ase[k], cse[k]) for k in ase)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  losses[0]={losses[0]}  losses[-1]={losses[-1]}")


CARE_CASES = {
    "care_small_s4": dict(n_steps=4, seed=7, cfg_overrides=dict(
        batch_size=120, actor=dict(actor_hidden_dim=[64, 48, 32]), critic=dict(critic_hidden_dim=[40, 72, 56]),
        encoder=dict(hidden_dims_contextEnc=[24, 20], output_dim_contextEnc=16, embedding_dim_contextEnc=16,
                     hidden_dims_mixtureEnc=[28], output_dim_mixtureEnc=12, num_encoders=4))),
}

CARE_CASES["care_o_small_s4"] = dict(n_steps=4, seed=8, cfg_overrides=dict(
    use_modified_care=False, batch_size=120, actor=dict(actor_hidden_dim=[64, 48, 32]), critic=dict(critic_hidden_dim=[40, 72, 56]),
    encoder=dict(hidden_dims_contextEnc=[24, 20], output_dim_contextEnc=12, embedding_dim_contextEnc=12,
                 hidden_dims_mixtureEnc=[28], output_dim_mixtureEnc=12, num_encoders=4)))

# MT1_Distributed_CARE (one task, plain ReplayBuffer, CARE(O) arithmetic) driven through ITS OWN learner.py
CARE_CASES["care_mt1_small_s3"] = dict(n_steps=3, seed=9, variant="C1", cfg_overrides=dict(
    batch_size=96, actor=dict(actor_hidden_dim=[48, 40, 32]), critic=dict(critic_hidden_dim=[56, 44, 36]),
    encoder=dict(hidden_dims_contextEnc=[20, 24], output_dim_contextEnc=12, embedding_dim_contextEnc=12,   # C1 asserts ctx_out == mix_out; trunk eats z
                 hidden_dims_mixtureEnc=[24], output_dim_mixtureEnc=12, num_encoders=3)))

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

# ---------------------------------------------------------------------------------------------------------------
# Full-size cases at the CONFIGURED shapes (BASELINE.json configs 3-5, long LL chain).  A full dump of a 1.7 M-parameter
# learner (parameters, targets, two Adam moments, before and after) would be ~50 MB per case, so these fixtures are
# "summaries": the INPUTS are regenerated from seeds on the test side (parameters from the port's seeded init -- loaded
# into the unmodified reference learner here --, minibatches and noise from seeded generators; torch's CPU generators
# are deterministic for a given torch build, and the GPU box runs this image), and of the reference's OUTPUTS the file
# keeps per-step losses, the first step's forward intermediates in full, and per tensor its sum, its L2 norm and 1024
# seeded sample elements.  tests/test_oracle_fullsize.py pins the port to them on CPU (same ATen ops: ~1e-7); the GPU
# tests compare the CUDA step with them on every quantity that no ReLU kink can move (forward values, losses) and with
# the port -- masks forced, kinks proven -- on the rest (tests/_golden.py::kink_checked_step).
# ---------------------------------------------------------------------------------------------------------------
N_SAMPLE = 1024


def summarize(d, prefix, tensors, seed=99):
    g = torch.Generator().manual_seed(seed)
    for k, t in tensors.items():
        t = t.detach().reshape(-1).double()
        n = t.numel()
        idx = torch.randperm(n, generator=g)[:min(n, N_SAMPLE)].sort().values
        d[f"{prefix}/{k}/sum"] = np.float64(t.sum().item())
        d[f"{prefix}/{k}/l2"] = np.float64(t.norm().item())
        d[f"{prefix}/{k}/idx"] = idx.numpy().astype(np.int64)
        d[f"{prefix}/{k}/val"] = t[idx].float().numpy()


def _set_params(named, params):
    with torch.no_grad():
        for k, prm in named.items():
            prm.data.copy_(params[k].reshape(prm.shape))


def make_full_case(name, family, n_steps, param_seed, data_seed, cfg_overrides=None):
    weighted = bool((cfg_overrides or {}).get("use_weighted_loss", family == "MS"))
    lrn, _ = rh.make_learner(family, cfg_overrides, seed=0)
    spec = _spec_of(lrn, family, weighted)
    params = sp.init_params(spec, seed=param_seed)
    named = _named_params(lrn, family)
    _set_params(named, params)
    d = {"spec": json.dumps(spec.to_json()), "family": family, "n_steps": n_steps, "param_seed": param_seed,
         "data_seed": data_seed, "kind": "summary"}
    g = torch.Generator().manual_seed(data_seed + 17)
    batches, eps_n, eps_c, losses = [], [], [], []
    for i in range(n_steps):
        batches.append(sp.synthetic_batch(spec, seed=data_seed + i))
        eps_n.append(torch.randn(spec.batch, spec.act_dim, generator=g))
        eps_c.append(torch.randn(spec.batch, spec.act_dim, generator=g))
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
    summarize(d, "p_out", {k: p_.detach() for k, p_ in named.items()})
    m, v, step = _adam_snapshot(lrn, named, spec)
    summarize(d, "m_out", m); summarize(d, "v_out", v)
    d["step_out"] = step
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  losses[0]={losses[0]}  losses[-1]={losses[-1]}")


def make_full_care_case(name, n_steps, param_seed, data_seed, modified=True):
    import care_port as cp
    lrn, _ = rh.make_learner("C10", dict(use_modified_care=modified), seed=0)
    spec = cp.CareSpec(modified=modified, weighted_loss=modified)
    params = cp.init_params(spec, seed=param_seed)
    named = _care_named(lrn)
    _set_params(named, params)
    with torch.no_grad():       # the actor's own state encoder starts tied to the critic's (learner.py:130-134, :402)
        ase = dict(lrn.actor.state_encoder.named_parameters()); cse = dict(lrn.local_critic.state_encoder.named_parameters())
        for k in ase:
            ase[k].data.copy_(cse[k].data)
    d = {"spec": json.dumps(spec.to_json()), "family": "C10", "n_steps": n_steps, "param_seed": param_seed,
         "data_seed": data_seed, "kind": "summary"}
    g = torch.Generator().manual_seed(data_seed + 17)
    batches, eps_n, eps_c, losses = [], [], [], []
    for i in range(n_steps):
        batches.append(cp.synthetic_batch(spec, seed=data_seed + i))
        eps_n.append(torch.randn(spec.batch, spec.act_dim, generator=g))
        eps_c.append(torch.randn(spec.batch, spec.act_dim, generator=g))
    s, a, r, s2, dn = batches[0]
    with torch.no_grad(), rh.injected_eps([eps_n[0], eps_c[0]]):
        alpha = lrn.get_log_alpha(s).exp()
        z = lrn.context_encoder.forward(s)
        a2, lp2, _ = lrn.actor.get_action_log_prob_log_std(mtobss=s2, z_context=z)
        qt = torch.min(*lrn.target_critic.forward(mtobss=s2, z_context=z, action=a2))
        q1, q2 = lrn.local_critic.forward(mtobss=s, z_context=z, action=a)
        ac, lpc, _ = lrn.actor.get_action_log_prob_log_std(mtobss=s, z_context=z)
        y = lrn.reward_scale * r + lrn.gamma * (1 - dn) * (qt - alpha * lp2)
    for k, t in dict(y=y, q1=q1, q2=q2, a_next=a2, logp_next=lp2, a_cur=ac, logp_cur=lpc).items():
        d["i0/" + k] = t.numpy()
    for i in range(n_steps):
        lrn.memory.sample = (lambda b: (lambda: tuple(t.clone() for t in b)))(batches[i])
        with rh.injected_eps([eps_n[i], eps_c[i]]) as q:
            res = lrn.update()
            assert not q
        losses.append(list(res))
    d["losses"] = np.array(losses, np.float64)
    summarize(d, "p_out", {k: p_.detach() for k, p_ in named.items()})
    trainable = [k for k in named if not (k.startswith("tse.") or "_target" in k or k == "embedding")]
    opts = {"critic": lrn.critic_optimizer, "actor": lrn.actor_optimizer, "alpha": lrn.log_alpha_optimizer,
            "ctx": lrn.context_encoder_optimizer}
    m, v, steps = {}, {}, {"critic": 0, "actor": 0, "alpha": 0, "ctx": 0}
    for k in trainable:
        tag = "alpha" if k == "log_alpha" else ("actor" if k.startswith("actor.") else ("ctx" if k.startswith("cenc.") else "critic"))
        st = opts[tag].state[named[k]]
        m[k], v[k] = st["exp_avg"].clone(), st["exp_avg_sq"].clone()
        steps[tag] = int(st["step"])
    summarize(d, "m_out", m); summarize(d, "v_out", v)
    d["step_out"] = np.array([steps["critic"], steps["actor"], steps["alpha"]] + ([] if modified else [steps["ctx"]]), np.int64)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  losses[0]={losses[0]}  losses[-1]={losses[-1]}")


FULL_CASES = {
    "full_vs_s3": dict(family="VS", n_steps=3, param_seed=21, data_seed=700),                 # config 3: 39/4/400^3, B 1024
    "full_ms_s3": dict(family="MS", n_steps=3, param_seed=22, data_seed=710),                 # config 4: B 1280, weighted loss
    "full_ll_s100": dict(family="LL", n_steps=100, param_seed=23, data_seed=720),             # 100 chained steps
}
FULL_CARE_CASES = {
    "full_c10m_s2": dict(n_steps=2, param_seed=24, data_seed=730, modified=True),             # config 5: CARE(M) B 1280 K 6
    "full_c10o_s2": dict(n_steps=2, param_seed=25, data_seed=740, modified=False),
}

if __name__ == "__main__":
    todo = sys.argv[1:] or (list(CASES) + list(CARE_CASES) + list(FULL_CASES) + list(FULL_CARE_CASES))
    for c in todo:
        if c in FULL_CASES:
            make_full_case(c, **FULL_CASES[c])
        elif c in FULL_CARE_CASES:
            make_full_care_case(c, **FULL_CARE_CASES[c])
        elif c in CARE_CASES:
            make_care_case(c, **CARE_CASES[c])
        else:
            make_case(c, **CASES[c])
