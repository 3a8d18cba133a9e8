"""Second CPU oracle: the SAC step with HAND-DERIVED gradients in numpy.

TEST INFRASTRUCTURE, NOT PRODUCT (same import rules as oracle/sac_port.py).

The CUDA path does not use autograd; it implements the closed-form backward
written out here.  This file restates that math stage by stage in numpy (fp32
or fp64) so the derivation can be checked on CPU against the autograd port
(tests/test_oracle_manual.py) and so every intermediate the kernels produce
(dQ, layer gradients, d-action, d-mu / d-log_std ...) has a reference value.

Follows /root/reference/LunarLander_Distributed_SAC/src/learner.py:203-239 and
model.py:38-65,117-142 (MT deltas: MT10_Distributed_MTSAC/src/learner.py:253-325).

Derivation notes (B = batch, A = act_dim, k = action scale, c = 1/B or 1/B^2):
  critic:  L = c * sum_i (y_i-Q1_i)^2 + (y_i-Q2_i)^2      dL/dQk_i = 2c (Qk_i - y_i)
  actor :  L = c * sum_i (alpha_i logpi_i - min(Q1,Q2)_i)
           dL/dQmin_i = -c (to the smaller head; Q1 on ties), dL/dlogpi_i = glp = c alpha_i
           u = mu + std*eps, t = tanh(u), a = k t, diff = u - mu (as rounded), var = std^2
           logpi = sum_j [ -diff^2/(2 var) - log std - log sqrt(2pi) - log(k (1 - (a/k)^2 + 1e-6)) ]
           d_act = dL/da + glp * 2 (a/k) / (k (1 - (a/k)^2 + 1e-6))
           g_u   = d_act * k (1-t^2) - glp * diff / var
           dL/dmu  = g_u + glp * diff / var
           dL/dstd = g_u * eps + glp * (diff^2 / std^3 - 1/std)
           dL/dlog_std = dL/dstd * std * [ -20 <= raw <= 2 ]
           (analytically dL/dmu = d_act k(1-t^2), dL/dlog_std = that*eps*std - glp; the long
            form is kept because the reference evaluates it in fp32, where diff != std*eps
            and 1 - t^2 is quantised for saturated samples -- SURVEY.md §7 "quirks")
  alpha :  dL/dlog_alpha[t] = -(1/B) sum_{i in t} (logpi_i + Hbar),  Hbar = -A
  Adam  :  torch _single_tensor_adam
"""
import math

import numpy as np

import sac_port as sp


def _np(d, dt):
    return {k: np.asarray(v, dtype=dt).copy() for k, v in d.items()}


def mlp_fwd(p, net, x, n):
    """Returns pre-head activations list hs (hs[0] = input) and the head output."""
    hs = [x]
    for i in range(n):
        z = hs[-1] @ p[f"{net}.{i}.weight"].T + p[f"{net}.{i}.bias"]
        if i < n - 1:
            hs.append(np.maximum(z, 0))
        else:
            return hs, z


def mlp_bwd(p, net, hs, dout, n, need_wgrad=True):
    """dout: gradient at the head output. Returns ({name: grad}, d_input)."""
    g = {}
    d = dout
    for i in reversed(range(n)):
        if need_wgrad:
            g[f"{net}.{i}.weight"] = d.T @ hs[i]
            g[f"{net}.{i}.bias"] = d.sum(0)
        d = d @ p[f"{net}.{i}.weight"]
        if i > 0:
            d = d * (hs[i] > 0)
    return g, d


def _cr(fn, x, dt):
    """Correctly-rounded transcendental: evaluate in fp64, round once to dt.
    torch's CPU fp32 tanh/exp/log (Sleef u10) agree with this on 98-99.9 % of
    inputs, numpy's fp32 tanh only on ~70 %; the CUDA kernels do the same."""
    return fn(np.asarray(x, np.float64)).astype(dt)


def policy_fwd(spec, p, obs, eps, dt):
    A, k = spec.act_dim, dt(spec.action_scale)
    n = len(spec.actor_hidden) + 1
    hs, out = mlp_fwd(p, "actor", obs, n)
    mu, raw = out[:, :A], out[:, A:]
    ls = np.clip(raw, -20, 2)
    std = _cr(np.exp, ls, dt)
    u = mu + std * eps
    t = _cr(np.tanh, u, dt)
    act = k * t
    diff = u - mu                       # NOT std*eps: keep the reference's fp32 quantisation
    var = std * std
    gauss = -(diff * diff) / (2 * var) - _cr(np.log, std, dt) - dt(math.log(math.sqrt(2 * math.pi)))
    sq = (act / k) ** 2
    jac = k * (1 - sq + dt(1e-6))
    logp = (gauss - _cr(np.log, jac, dt)).sum(-1, keepdims=True)
    return dict(hs=hs, mu=mu, raw=raw, std=std, t=t, act=act, logp=logp, log_std=_cr(np.log, std, dt),
                diff=diff, var=var, jac=jac)


def adam(p, g, m, v, step, lr, b1, b2, eps, dt):
    """torch.optim.adam._single_tensor_adam; `step` is the post-increment count."""
    m[...] = m + (g - m) * dt(1 - b1)
    v[...] = v * dt(b2) + dt(1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = np.sqrt(v) / dt(math.sqrt(bc2)) + dt(eps)
    p[...] = p - dt(lr / bc1) * (m / denom)


class ManualLearner:
    def __init__(self, spec, params, adam_state=None, dtype=np.float32):
        self.spec, self.dt = spec, dtype
        self.p = _np(params, dtype)
        tr = sp.param_names(spec, sp.TRAINABLE_NETS)
        if adam_state is None:
            self.m = {k: np.zeros_like(self.p[k]) for k in tr}
            self.v = {k: np.zeros_like(self.p[k]) for k in tr}
            self.step = [0, 0, 0]
        else:
            self.m, self.v = _np(adam_state["m"], dtype), _np(adam_state["v"], dtype)
            self.step = [int(x) for x in adam_state["step"]]

    def update_SAC(self, s, a, r, s2, d, eps_next, eps_cur):
        spec, p, dt = self.spec, self.p, self.dt
        s, a, r, s2, d, eps_next, eps_cur = (np.asarray(x, dtype=dt) for x in (s, a, r, s2, d, eps_next, eps_cur))
        B, A = s.shape[0], spec.act_dim
        nA, nC = len(spec.actor_hidden) + 1, len(spec.critic_hidden) + 1
        T = spec.num_tasks
        tid = np.argmax(s[:, -T:], 1) if T > 0 else np.zeros(B, np.int64)
        alpha = np.exp(p["log_alpha"])[tid][:, None]
        c = dt(1.0 / B / (B if spec.weighted_loss else 1))
        I = {}

        # ---- target ----
        pn = policy_fwd(spec, p, s2, eps_next, dt)
        x2 = np.concatenate([s2, pn["act"]], 1)
        qt1 = mlp_fwd(p, "q1_target", x2, nC)[1]
        qt2 = mlp_fwd(p, "q2_target", x2, nC)[1]
        y = dt(spec.reward_scale) * r + dt(spec.gamma) * (1 - d) * (np.minimum(qt1, qt2) - alpha * pn["logp"])
        I.update(y=y, a_next=pn["act"], logp_next=pn["logp"])

        # ---- critic update ----
        x = np.concatenate([s, a], 1)
        grads = {}
        closs = dt(0)
        for q in ("q1", "q2"):
            hs, qv = mlp_fwd(p, q, x, nC)
            I[q] = qv
            closs = closs + ((y - qv) ** 2).sum() * c
            g, _ = mlp_bwd(p, q, hs, 2 * c * (qv - y), nC)
            grads.update(g)
        self.step[0] += 1
        for k_, g in grads.items():
            adam(p[k_], g, self.m[k_], self.v[k_], self.step[0], spec.lr_critic, spec.beta1, spec.beta2, spec.adam_eps, dt)
        I["critic_grads"] = grads

        # ---- actor update (uses the UPDATED critics) ----
        pc = policy_fwd(spec, p, s, eps_cur, dt)
        xa = np.concatenate([s, pc["act"]], 1)
        h1, q1n = mlp_fwd(p, "q1", xa, nC)
        h2, q2n = mlp_fwd(p, "q2", xa, nC)
        pick1 = (q1n <= q2n)
        qmin = np.where(pick1, q1n, q2n)
        aloss = ((alpha * pc["logp"] - qmin) * c).sum()
        dq1 = np.where(pick1, -c, dt(0)).astype(dt)
        dq2 = np.where(pick1, dt(0), -c).astype(dt)
        _, dx1 = mlp_bwd(p, "q1", h1, dq1, nC, need_wgrad=False)
        _, dx2 = mlp_bwd(p, "q2", h2, dq2, nC, need_wgrad=False)
        da = (dx1 + dx2)[:, -A:]
        # Same evaluation order as autograd on the reference's expression graph, so the
        # fp32 quantisation of saturated samples (|tanh| -> 1, tiny std) is reproduced:
        t, std, diff, var = pc["t"], pc["std"], pc["diff"], pc["var"]
        k = dt(spec.action_scale)
        glp = c * alpha                                  # dL/dlogpi, per sample
        d_act = da + glp * (2 * (pc["act"] / k) / k) * k / pc["jac"]   # -log(k(1-(a/k)^2+1e-6)) branch
        g_u_t = d_act * k * (1 - t * t)                  # through a = k tanh(u)
        g_u = g_u_t + glp * (-(diff) / var)              # + Gaussian d/du
        dmu = g_u + glp * (diff / var)                   # Gaussian d/dmu (cancels the line above)
        dstd = g_u * eps_cur + glp * ((diff * diff) / (var * std) - 1 / std)
        dls = dstd * std * ((pc["raw"] >= -20) & (pc["raw"] <= 2))
        dout = np.concatenate([dmu, dls], 1).astype(dt)
        ag, _ = mlp_bwd(p, "actor", pc["hs"], dout, nA)
        self.step[1] += 1
        for k_, g in ag.items():
            adam(p[k_], g, self.m[k_], self.v[k_], self.step[1], spec.lr_actor, spec.beta1, spec.beta2, spec.adam_eps, dt)
        I.update(a_cur=pc["act"], logp_cur=pc["logp"], qmin=qmin, d_action=da, d_head=dout, actor_grads=ag)

        # ---- temperature ----
        hbar = dt(-A)
        gal = np.zeros_like(p["log_alpha"])
        np.add.at(gal, tid, (-(pc["logp"] + hbar) / dt(B))[:, 0])
        self.step[2] += 1
        adam(p["log_alpha"], gal, self.m["log_alpha"], self.v["log_alpha"], self.step[2], spec.lr_actor,
             spec.beta1, spec.beta2, spec.adam_eps, dt)
        I["alpha_grad"] = gal

        # ---- Polyak ----
        for q in ("q1", "q2"):
            for i in range(nC):
                for kind in ("weight", "bias"):
                    tkey, lkey = f"{q}_target.{i}.{kind}", f"{q}.{i}.{kind}"
                    p[tkey] = dt(spec.tau) * p[lkey] + dt(1.0 - spec.tau) * p[tkey]

        ent = (dt(0.5 * A * (1.0 + math.log(2 * math.pi))) + pc["log_std"].sum(-1)).mean()
        I.update(critic_loss=float(closs), actor_loss=float(aloss), entropy=float(ent))
        return I
