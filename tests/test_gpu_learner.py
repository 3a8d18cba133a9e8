"""GPU: the reference-shaped Learner / ReplayBuffer surface (drop-in boundary)."""
import json
import os
import pickle

import pytest
import torch

import redis_stub
from _golden import Case, REL, rel_l2, rel_scalar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


LL_CFG = {"num_tasks": 10, "device": "cuda", "buffer_size": 1e6, "reward_scale": 1, "batch_size": 256, "gamma": 0.99,
          "lr_actor": 3e-4, "lr_critic": 3e-4, "log_alpha": 0, "tau": 0.005, "num_learn": 1, "num_time_step": 1,
          "random_step": 5000, "start_memory_len": 5000}


def _ll_learner(tmp_path, **kw):
    from distributed_sac_b200.learner import Learner
    p = tmp_path / "cfg.json"
    p.write_text(json.dumps(LL_CFG))
    os.chdir(tmp_path)
    return Learner(str(p), write_mode=False, server=redis_stub.StrictRedis(host=str(tmp_path)), **kw)


def test_ll_learner_matches_fixture_through_reference_api(cuda, tmp_path):
    """Load the fixture through load_state_dict-style handles with the REFERENCE key names, run
    update_SAC() on the fixture minibatches, compare losses and the actor blob players would pull."""
    from distributed_sac_b200 import _lib, names
    c = Case("ll_ckpt_s3")
    lrn = _ll_learner(tmp_path)
    inv = lambda km, src: {ref: src[canon] for ref, canon in km.items()}
    lrn.actor.load_state_dict(inv(names.actor_key_map("LL", 3), c.p_in))
    lrn.local_critic_1.load_state_dict(inv(names.critic_key_map("LL", 3, 1), c.p_in))
    lrn.local_critic_2.load_state_dict(inv(names.critic_key_map("LL", 3, 2), c.p_in))
    lrn.target_critic_1.load_state_dict(inv(names.critic_key_map("LL", 3, 1, True), c.p_in))
    lrn.target_critic_2.load_state_dict(inv(names.critic_key_map("LL", 3, 2, True), c.p_in))
    lrn.core.set_named({"log_alpha": c.p_in["log_alpha"]}, strict=False)
    lrn.core.set_named(c.m_in, _lib.ADAM_M)
    lrn.core.set_named(c.v_in, _lib.ADAM_V)
    lrn.core.set_steps(c.step_in)
    for i in range(c.n_steps):
        s, a, r, s2, d = c.step_batch(i)
        dev = (i % 2 == 0)       # alternate device / host minibatches: both entry points
        mv = (lambda t: t.cuda()) if dev else (lambda t: t)
        cl, al = lrn.update_SAC(mv(s), mv(a), mv(r), mv(s2), mv(d), None, eps_next=mv(c.eps_next[i]), eps_cur=mv(c.eps_cur[i]))
        assert rel_scalar(cl, c.losses[i, 0]) <= REL and rel_scalar(al, c.losses[i, 1]) <= REL
    blob = pickle.loads(pickle.dumps(lrn.get_parameters()))       # what run() publishes to players
    assert set(blob) == {"actor"}
    km = names.actor_key_map("LL", 3)
    assert set(blob["actor"]) == set(km)
    for ref, canon in km.items():
        assert blob["actor"][ref].shape == c.p_out[canon].shape
        assert rel_l2(blob["actor"][ref], c.p_out[canon]) <= REL
    lrn.memory.stop()


def test_checkpoint_round_trip_in_reference_format(cuda, tmp_path):
    lrn = _ll_learner(tmp_path, seed=3)
    lrn.memory.ring.fill_synthetic(20000, seed=1)
    for _ in range(5):
        cl, al = lrn.update()
        assert cl == cl and al == al
    path = lrn.save_checkpoint(15)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert {"episode_idx", "total_step", "local_critic_1", "local_critic_2", "critic_optimizer", "target_critic_1",
            "target_critic_2", "actor", "actor_optimizer", "log_alpha", "log_alpha_optimizer", "alpha"} <= set(ck)
    assert set(ck["local_critic_1"]) == {"first_layer.weight", "first_layer.bias", "layer_module.0.weight",
                                         "layer_module.0.bias", "layer_module.1.weight", "layer_module.1.bias"}
    assert len(ck["critic_optimizer"]["state"]) == 12 and int(ck["critic_optimizer"]["state"][0]["step"]) == 5
    # a plain torch Adam accepts the saved optimizer state (the format players / tools rely on)
    ps = [torch.nn.Parameter(torch.zeros_like(ck["critic_optimizer"]["state"][i]["exp_avg"])) for i in range(12)]
    torch.optim.Adam(ps, lr=3e-4).load_state_dict(ck["critic_optimizer"])
    before = lrn.core.export_arena()
    lrn2 = _ll_learner(tmp_path, seed=9, checkpoint_path=path)
    assert torch.equal(lrn2.core.export_arena(), before)
    assert lrn2.core.get_steps() == (5, 5, 5)
    s, a, r, s2, d = lrn.memory.sample()
    assert s.is_cuda and s.shape == (256, 8) and a.shape == (256, 2) and r.shape == (256, 1) and d.shape == (256, 1)
    g = torch.Generator().manual_seed(0)
    e1, e2 = torch.randn(256, 2, generator=g), torch.randn(256, 2, generator=g)
    assert lrn.update_SAC(s, a, r, s2, d, None, eps_next=e1.cuda(), eps_cur=e2.cuda()) == \
        lrn2.update_SAC(s, a, r, s2, d, None, eps_next=e1.cuda(), eps_cur=e2.cuda())
    lrn.memory.stop()
    lrn2.memory.stop()


def test_run_loop_with_redis_stub(cuda, tmp_path):
    """Learner.run(): players push pickled tuples to the 'sample' list, the ReplayBuffer thread drains
    them, the learner publishes 'parameters' / 'update_iteration' and loss lists (LL/learner.py:278-316)."""
    import numpy as np
    lrn = _ll_learner(tmp_path, seed=1)
    lrn.write_mode = True
    lrn.start_memory_len = 1200
    rng = np.random.default_rng(0)
    for _ in range(1500):      # tuple layout of LL/player.py:115-122
        tup = (rng.standard_normal(8), rng.uniform(-1, 1, 2).astype(np.float32), float(rng.standard_normal()),
               rng.standard_normal(8), bool(rng.random() < 0.01))
        lrn.server.rpush("sample", pickle.dumps(tup))
    n = lrn.run(max_updates=7)
    assert n == 7 and len(lrn.memory) == 1500
    assert pickle.loads(lrn.server.get("update_iteration")) == 18          # update_delay = 3
    params = pickle.loads(lrn.server.get("parameters"))
    assert params["actor"]["mu_log_std_layer.weight"].shape == (4, 256)
    assert lrn.server.llen("critic_loss") == 7 and lrn.server.llen("alpha") == 7
    lrn.memory.stop()
