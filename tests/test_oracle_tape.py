"""CPU: the oracle's test hooks (oracle/sac_port.py::ReluTape, min_tagged, tanh_tagged) are transparent -- forcing the port's OWN
ReLU masks, min(Q1,Q2) routing and tanh(u) values back into it reproduces the unforced step (bit for bit for the masks and the routing), and a forced tanh that
differs in the last bits moves the actor gradient of a saturated row by far more than the parity tolerance (the ill-conditioning
the GPU tests neutralise by forcing the CUDA step's own tanh)."""
import torch

import sac_port as sp
from _golden import rel_l2


def _run(spec, params, batch, e1, e2, forced=None):
    port = sp.PortLearner(spec, params)
    with sp.ReluTape(forced) as tape:
        out = port.update_SAC(*batch, e1, e2)
    return port, out, tape


def test_forcing_the_ports_own_decisions_is_the_identity():
    spec = sp.SacSpec(state_dim=8, act_dim=2, actor_hidden=[32, 24], critic_hidden=[24, 32], batch=48)
    params = sp.init_params(spec, seed=3)
    batch = sp.synthetic_batch(spec, seed=5)
    g = torch.Generator().manual_seed(7)
    e1, e2 = torch.randn(spec.batch, spec.act_dim, generator=g), torch.randn(spec.batch, spec.act_dim, generator=g)
    p0, o0, t0 = _run(spec, params, batch, e1, e2)
    masks = {tag: (z > 0) for tag, z in t0.z.items() if not tag.startswith("tanh:")}
    assert "route:pi" in masks and any(t.startswith("actor:cur:") for t in masks) and {"tanh:cur", "tanh:next"} <= set(t0.z)
    # ReLU masks + min routing: exactly the identity
    p1, o1, _ = _run(spec, params, batch, e1, e2, masks)
    assert o0["critic_loss"] == o1["critic_loss"] and o0["actor_loss"] == o1["actor_loss"]
    a, b = p0.params(), p1.params()
    assert all(torch.equal(a[k], b[k]) for k in a)
    sa, sb = p0.adam_state(), p1.adam_state()
    assert all(torch.equal(sa["m"][k], sb["m"][k]) and torch.equal(sa["v"][k], sb["v"][k]) for k in sa["m"])
    # + the port's own tanh values: the derivative 1 - t*t is evaluated outside torch's fused tanh backward, which already
    # shows in the 5th digit of the actor moments (the conditioning this hook exists for); well inside the 1e-4 parity bar
    forced = dict(masks, **{tag: z.clone() for tag, z in t0.z.items() if tag.startswith("tanh:")})
    p2, o2, _ = _run(spec, params, batch, e1, e2, forced)
    assert abs(o0["actor_loss"] - o2["actor_loss"]) <= 1e-6 * abs(o0["actor_loss"])
    c, sc = p2.params(), p2.adam_state()
    assert all(rel_l2(c[k], a[k]) <= 1e-6 for k in a if k != "log_alpha")
    assert all(rel_l2(sc["m"][k], sa["m"][k]) <= 1e-4 for k in sa["m"] if k != "log_alpha")


def test_a_last_bit_of_a_saturated_tanh_moves_the_actor_gradient():
    spec = sp.SacSpec(state_dim=8, act_dim=2, actor_hidden=[32, 24], critic_hidden=[24, 32], batch=48)
    params = sp.init_params(spec, seed=3)
    params = dict(params)
    head_b = [k for k in params if k.startswith("actor.") and k.endswith(".bias")][-1]
    params[head_b] = params[head_b].clone()
    params[head_b][0] = 7.5                      # mean of action 0 deep in the saturated region: tanh(u) = 1 - O(1e-6)
    batch = sp.synthetic_batch(spec, seed=5)
    g = torch.Generator().manual_seed(7)
    e1, e2 = torch.randn(spec.batch, spec.act_dim, generator=g), torch.randn(spec.batch, spec.act_dim, generator=g)
    p0, _o0, t0 = _run(spec, params, batch, e1, e2)
    t = t0.z["tanh:cur"].clone()
    sat = (1 - t[:, 0].abs()) < 5e-6
    assert int(sat.sum()) > 0, "the constructed case has no saturated row"
    nudged = t.clone()
    nudged[:, 0] = torch.where(sat, torch.nextafter(t[:, 0], torch.zeros_like(t[:, 0])), t[:, 0])     # one ulp towards zero
    assert float((nudged - t).abs().max()) <= 1.2e-7
    p1, _o1, _t1 = _run(spec, params, batch, e1, e2, {"tanh:cur": nudged})
    m0, m1 = p0.adam_state()["m"], p1.adam_state()["m"]
    worst = max(rel_l2(m1[k], m0[k]) for k in m0 if k.startswith("actor."))
    assert worst > 1e-3, f"one ulp of tanh moved the actor moments by only {worst:.2e}"
