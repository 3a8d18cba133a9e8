"""CPU oracle for the CARE gradient step (both variants) -- TEST INFRASTRUCTURE, NOT PRODUCT.

Clean-room PyTorch fp32 autograd restatement of MT10_Distributed_CARE:
  CARE(M) `use_modified_care: true`  (BASELINE.json config 5): frozen embedding -> per-encoder mlp_context, weighted loss
  CARE(O) `use_modified_care: false` (also the algorithm of MT1_Distributed_CARE): TRAINABLE context encoder
          z = mlp(header(relu(E[t]))) with its own Adam (context_encoder.py:59-89, learner.py:137-141,399),
          z feeds the attention trunk (detached) and is concatenated as-is; plain (unweighted) losses.  Same import rules as
oracle/sac_port.py.  Pinned to the unmodified reference by the fixture
tests/golden/care_small_s4.npz (oracle/gen_golden.py drives the real
`Learner.update()`), checked in tests/test_oracle_golden.py.

Reference sites restated (paths relative to /root/reference/MT10_Distributed_CARE/src):
  step ordering, detach rules, ties .. learner.py:281-404 (update_SAC + update)
  4 optimizers ...................... learner.py:136-158  (actor opt = mu_log_std_layer ONLY)
  context encoder (M) ............... context_encoder.py:49-58,110-127 (frozen RoBERTa embedding lookup)
  mixture of encoders ............... state_encoder.py:132-221 (einsum 'kio,bi->kbo', W ~ randn)
  attention + context MLP ........... state_encoder.py:75-96   (softmax(trunk(z.detach())), mlp_context)
  actor / critic over encoded state . model.py:50-88,208-222
  weighted losses (== mean/B) ....... model.py:148-157,256-267

Canonical parameter names (on top of oracle/sac_port.py's actor/q1/q2/*_target/log_alpha):
  cse.mix.{l}.W (K,in,out) | cse.mix.{l}.b (K,1,out)      critic state encoder, mixture layer l
  cse.trunk.{j}.weight|bias                              attention trunk (nn.Linear layout)
  cse.ctx.{j}.weight|bias                                mlp_context
  tse.*                                                  target critic's state encoder
  embedding                                              (T, 768) frozen
  cenc.{j}.weight|bias                                   CARE(O) only: context encoder = header (2 Linear) + mlp, trainable
The actor's state encoder is always a hard copy of `cse` taken at the end of update()
(learner.py:402), so it is not a separate set of names: `ase.*` is derived on export.
"""
import math
from dataclasses import dataclass, field, asdict
from typing import List

import torch
import torch.nn.functional as F

import sac_port as sp


@dataclass
class CareSpec:
    state_dim: int = 39
    act_dim: int = 4
    num_tasks: int = 10
    actor_hidden: List[int] = field(default_factory=lambda: [400, 400, 400])
    critic_hidden: List[int] = field(default_factory=lambda: [400, 400, 400])
    batch: int = 1280
    num_encoders: int = 6
    mix_hidden: List[int] = field(default_factory=lambda: [50])     # hidden_dims_mixtureEnc (also the trunk's)
    mix_out: int = 50                                               # output_dim_mixtureEnc
    ctx_in: int = 768                                               # RoBERTa_embedding_dim
    ctx_hidden: List[int] = field(default_factory=lambda: [50, 50]) # hidden_dims_contextEnc
    ctx_out: int = 50                                               # output_dim_contextEnc
    tau_se: float = 0.05                                            # state_encoder_tau
    weighted_loss: bool = True                                      # use_modified_care => weighted (== /B)
    modified: bool = True                                           # use_modified_care
    emb_dim: int = 50                                               # embedding_dim_contextEnc (CARE(O) header width)
    lr_ctx: float = 3e-4                                            # lr_contextEnc
    gamma: float = 0.99
    tau: float = 0.005
    reward_scale: float = 1.0
    lr_actor: float = 3e-4
    lr_critic: float = 3e-4
    action_scale: float = 1.0
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8

    @property
    def obs_dim(self):
        return self.state_dim + self.num_tasks

    @property
    def enc_dim(self):          # policy_input_dim
        return self.ctx_out + self.mix_out

    def to_json(self):
        return asdict(self)

    def mlp_spec(self) -> sp.SacSpec:
        """The actor / Q MLPs seen as a plain SAC problem over the encoded state."""
        return sp.SacSpec(state_dim=self.enc_dim, act_dim=self.act_dim, actor_hidden=list(self.actor_hidden),
                          critic_hidden=list(self.critic_hidden), batch=self.batch, num_tasks=0)


def mix_dims(spec):
    d = [spec.state_dim] + list(spec.mix_hidden) + [spec.mix_out]
    return list(zip(d[:-1], d[1:]))


def trunk_dims(spec):
    d = [spec.ctx_in if spec.modified else spec.emb_dim] + list(spec.mix_hidden) + [spec.num_encoders]
    return list(zip(d[:-1], d[1:]))


def ctx_dims(spec):
    """CARE(M): mlp_context inside every state encoder.  CARE(O): none there."""
    if not spec.modified:
        return []
    d = [spec.ctx_in] + list(spec.ctx_hidden) + [spec.ctx_out]
    return list(zip(d[:-1], d[1:]))


def cenc_dims(spec):
    """CARE(O) context encoder: header Linear(768, 2e), Linear(2e, e) then mlp e -> ctx_hidden -> ctx_out."""
    if spec.modified:
        return []
    d = [spec.ctx_in, 2 * spec.emb_dim, spec.emb_dim] + list(spec.ctx_hidden) + [spec.ctx_out]
    return list(zip(d[:-1], d[1:]))


def cenc_names(spec):
    n = []
    for j, _ in enumerate(cenc_dims(spec)):
        n += [f"cenc.{j}.weight", f"cenc.{j}.bias"]
    return n


def encoder_names(spec, prefix):
    n = []
    for l, _ in enumerate(mix_dims(spec)):
        n += [f"{prefix}.mix.{l}.W", f"{prefix}.mix.{l}.b"]
    for j, _ in enumerate(trunk_dims(spec)):
        n += [f"{prefix}.trunk.{j}.weight", f"{prefix}.trunk.{j}.bias"]
    for j, _ in enumerate(ctx_dims(spec)):
        n += [f"{prefix}.ctx.{j}.weight", f"{prefix}.ctx.{j}.bias"]
    return n


def init_params(spec: CareSpec, seed=0, embedding=None):
    g = torch.Generator().manual_seed(seed)
    p = {k: v for k, v in sp.init_params(spec.mlp_spec(), seed=seed).items() if k != "log_alpha"}
    p["log_alpha"] = torch.zeros(spec.num_tasks)
    for pre in ("cse", "tse"):
        for l, (i, o) in enumerate(mix_dims(spec)):          # randn init (state_encoder.py:146-153)
            p[f"{pre}.mix.{l}.W"] = torch.randn(spec.num_encoders, i, o, generator=g)
            p[f"{pre}.mix.{l}.b"] = torch.randn(spec.num_encoders, 1, o, generator=g)
        for name, dims in (("trunk", trunk_dims(spec)), ("ctx", ctx_dims(spec))):
            for j, (i, o) in enumerate(dims):
                bound = math.sqrt(6.0 / (i + o))
                p[f"{pre}.{name}.{j}.weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
                p[f"{pre}.{name}.{j}.bias"] = torch.zeros(o)
    for j, (i, o) in enumerate(cenc_dims(spec)):
        bound = math.sqrt(6.0 / (i + o))
        p[f"cenc.{j}.weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        p[f"cenc.{j}.bias"] = torch.zeros(o)
    p["embedding"] = embedding.clone().float() if embedding is not None else torch.randn(spec.num_tasks, spec.ctx_in, generator=g) * 0.3
    return p


def synthetic_batch(spec: CareSpec, seed=1234, batch=None):
    s = sp.SacSpec(state_dim=spec.state_dim, act_dim=spec.act_dim, num_tasks=spec.num_tasks, batch=batch or spec.batch)
    return sp.synthetic_batch(s, seed=seed)


def _seq(p, prefix, x, n, last_relu=False):
    for j in range(n):
        x = F.linear(x, p[f"{prefix}.{j}.weight"], p[f"{prefix}.{j}.bias"])
        if j < n - 1 or last_relu:
            x = torch.relu(x)
    return x


def state_encode(spec, p, pre, z_context, mtobs, detach_z_encs=False, tag=None):
    """stateEncoder.forward (state_encoder.py:98-129).  `tag` names the pass for sac_port.ReluTape (test hook)."""
    x = mtobs[:, :spec.state_dim]
    nl = len(mix_dims(spec))
    for l in range(nl):
        W, b = p[f"{pre}.mix.{l}.W"], p[f"{pre}.mix.{l}.b"]
        x = (torch.einsum("kio,bi->kbo", W, x) if x.dim() == 2 else torch.einsum("kio,kbi->kbo", W, x)) + b
        if l < nl - 1:
            x = sp.relu_tagged(x, None if tag is None else f"{pre}.mix:{tag}:{l}")
    z_encs = x.transpose(1, 0)                                   # (B, K, out)
    if detach_z_encs:
        z_encs = z_encs.detach()
    alpha = torch.softmax(_seq(p, f"{pre}.trunk", z_context.detach(), len(trunk_dims(spec))), dim=-1).unsqueeze(-1)
    z_enc = (z_encs * alpha).sum(dim=1)
    z_enc = z_enc / alpha.sum(dim=1)
    zc = _seq(p, f"{pre}.ctx", z_context, len(ctx_dims(spec))) if spec.modified else z_context
    return torch.cat([zc, z_enc], dim=1)


def context_encode(spec, p, tid):
    """contextEncoder.forward (context_encoder.py:110-127)."""
    if spec.modified:
        return p["embedding"][tid]
    return _seq(p, "cenc", torch.relu(p["embedding"][tid]), len(cenc_dims(spec)))


class CarePortLearner:
    def __init__(self, spec: CareSpec, params, adam_state=None):
        self.spec = spec
        frozen = lambda k: ("_target" in k) or k.startswith("tse.") or k == "embedding"
        self.p = {k: torch.nn.Parameter(v.detach().clone().float(), requires_grad=not frozen(k)) for k, v in params.items()}
        # the actor's own copy of the state encoder (tied to the critic's at the end of every update)
        for k in encoder_names(spec, "cse"):
            self.p["ase" + k[3:]] = torch.nn.Parameter(self.p[k].detach().clone(), requires_grad=True)
        ms = spec.mlp_spec()
        self.actor_names = sp.param_names(ms, ("actor",), False)
        self.critic_names = encoder_names(spec, "cse") + sp.param_names(ms, ("q1", "q2"), False)   # learner.py:150-153 order
        self.opt_actor = torch.optim.Adam([self.p[n] for n in self.actor_names], lr=spec.lr_actor)
        self.opt_critic = torch.optim.Adam([self.p[n] for n in self.critic_names], lr=spec.lr_critic)
        self.opt_alpha = torch.optim.Adam([self.p["log_alpha"]], lr=spec.lr_actor)
        self.ctx_names = cenc_names(spec)
        self.opt_ctx = torch.optim.Adam([self.p[n] for n in self.ctx_names], lr=spec.lr_ctx) if self.ctx_names else None
        if adam_state is not None:
            self.load_adam(adam_state)

    def trainable_names(self):
        return self.actor_names + self.critic_names + ["log_alpha"] + self.ctx_names

    def _opt_of(self, name):
        if name == "log_alpha":
            return self.opt_alpha, 2
        if name.startswith("cenc."):
            return self.opt_ctx, 3
        return (self.opt_actor, 1) if name.startswith("actor.") else (self.opt_critic, 0)

    def load_adam(self, st):
        for name in self.trainable_names():
            opt, slot = self._opt_of(name)
            opt.state[self.p[name]] = {"step": torch.tensor(float(st["step"][slot])),
                                       "exp_avg": torch.as_tensor(st["m"][name]).clone().float(),
                                       "exp_avg_sq": torch.as_tensor(st["v"][name]).clone().float()}

    def adam_state(self):
        m, v, step = {}, {}, [0, 0, 0] + ([0] if self.ctx_names else [])
        for name in self.trainable_names():
            opt, slot = self._opt_of(name)
            st = opt.state.get(self.p[name], None)
            if not st:
                m[name] = torch.zeros_like(self.p[name].data)
                v[name] = torch.zeros_like(self.p[name].data)
            else:
                m[name], v[name] = st["exp_avg"].clone(), st["exp_avg_sq"].clone()
                step[slot] = int(st["step"])
        return {"m": m, "v": v, "step": tuple(step)}

    def params(self):
        return {k: v.detach().clone() for k, v in self.p.items() if not k.startswith("ase.")}

    def _policy(self, se_prefix, z, obs, eps, detach, tag=None):
        spec, p = self.spec, self.p
        A = spec.act_dim
        enc = state_encode(spec, p, se_prefix, z, obs, detach_z_encs=detach, tag=tag)
        out = sp.mlp(p, "actor", enc, tag)
        mu, log_std = out[:, :A], torch.clamp(out[:, A:], -20, 2)
        std = torch.exp(log_std)
        u = mu + std * eps
        k = spec.action_scale
        act = k * sp.tanh_tagged(u, None if tag is None else f"tanh:{tag}")
        gauss = -((u - mu) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
        logp = (gauss - torch.log(k * (1 - (act / k) ** 2 + 1e-6))).sum(-1, keepdim=True)
        return act, logp, torch.log(std)

    def _q(self, se_prefix, qa, qb, z, obs, act, detach=False, tag=None):
        enc = state_encode(self.spec, self.p, se_prefix, z, obs, detach_z_encs=detach, tag=tag)
        x = torch.cat([enc, act], -1)
        return sp.mlp(self.p, qa, x, tag), sp.mlp(self.p, qb, x, tag)

    def update(self, s, a, r, s2, d, eps_next=None, eps_cur=None, want_intermediates=False):
        """Learner.update() minus sampling (learner.py:377-404 + 281-369)."""
        spec, p = self.spec, self.p
        B = s.shape[0]
        eps_next = torch.randn(B, spec.act_dim) if eps_next is None else eps_next
        eps_cur = torch.randn(B, spec.act_dim) if eps_cur is None else eps_cur
        tid = torch.argmax(s[:, -spec.num_tasks:], dim=1)
        alpha = p["log_alpha"].detach()[tid].exp().unsqueeze(1)
        for opt in (self.opt_critic, self.opt_actor, self.opt_alpha) + ((self.opt_ctx,) if self.opt_ctx else ()):
            opt.zero_grad()
        div = float(B) if spec.weighted_loss else 1.0
        z = context_encode(spec, p, tid)                          # contextEncoder.forward

        with torch.no_grad():
            a2, logp2, _ = self._policy("ase", z, s2, eps_next, False, "next")
            qt1, qt2 = self._q("tse", "q1_target", "q2_target", z, s2, a2, tag="next")
            y = spec.reward_scale * r + spec.gamma * (1 - d) * (torch.min(qt1, qt2) - alpha * logp2)

        q1, q2 = self._q("cse", "q1", "q2", z, s, a, tag="cur")
        q_loss = torch.mean((y - q1) ** 2) / div + torch.mean((y - q2) ** 2) / div
        q_loss.backward()                                         # also deposits d/d(context encoder) in CARE(O)
        self.opt_critic.step()

        a_cur, logp, log_std = self._policy("ase", z.detach(), s, eps_cur, True, "cur")
        q1n, q2n = self._q("cse", "q1", "q2", z.detach(), s, a_cur, detach=True, tag="pi")
        qmin = sp.min_tagged(q1n, q2n, "route:pi")
        pi_loss = torch.mean(-(qmin - alpha * logp)) / div
        pi_loss.backward()
        self.opt_actor.step()

        la = p["log_alpha"][tid].unsqueeze(1)
        alpha_loss = -(la * (logp.detach() + (-float(spec.act_dim)))).mean()
        alpha_loss.backward()
        self.opt_alpha.step()

        with torch.no_grad():
            ms = spec.mlp_spec()
            for q in ("q1", "q2"):
                for i, _ in enumerate(sp.layer_dims(ms, q)):
                    for kind in ("weight", "bias"):
                        t, l = p[f"{q}_target.{i}.{kind}"], p[f"{q}.{i}.{kind}"]
                        t.copy_(spec.tau * l + (1.0 - spec.tau) * t)
            for k in encoder_names(spec, "cse"):
                t, l = p["tse" + k[3:]], p[k]
                t.copy_(spec.tau_se * l + (1.0 - spec.tau_se) * t)
        if self.opt_ctx is not None:
            self.opt_ctx.step()                                   # update(): context_encoder_optimizer.step() (learner.py:399)
        with torch.no_grad():
            # the hard tie (learner.py:402)
            for k in encoder_names(spec, "cse"):
                p["ase" + k[3:]].copy_(p[k])

        entropy = (0.5 * spec.act_dim * (1.0 + math.log(2 * math.pi)) + log_std.detach().sum(-1)).mean()
        out = {"critic_loss": q_loss.item(), "actor_loss": pi_loss.item(), "alpha_loss": alpha_loss.item(),
               "entropy": entropy.item()}
        if want_intermediates:
            out.update(y=y, q1=q1.detach(), q2=q2.detach(), a_next=a2, logp_next=logp2, a_cur=a_cur.detach(),
                       logp_cur=logp.detach(), qmin=qmin.detach())
        return out

    update_SAC = update
