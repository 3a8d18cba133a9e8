"""CPU, build container only (skipped where no reference checkout is reachable -- e.g. on the GPU box): the oracle ports
and the host-side wire formats against the UNMODIFIED reference, live.

* oracle/sac_port.py / care_port.py reproduce the reference learner's update() step for step (the committed fixtures under
  tests/golden/ are frozen outputs of the same comparison; this test re-derives them from the code in front of us);
* the bytes Learner.run() publishes (Learner.parameters_blob: a pickle stream built once whose float payloads the library gathers from the arena) load into
  the reference's own Actor via load_state_dict -- what Player.pull_parameters does (LL/player.py:75-85);
* a checkpoint written by the drop-in learner's save path (reference-written fixtures round-tripped on the GPU side, see
  tests/test_gpu_checkpoint.py) has the key set the reference's load_checkpoint() reads."""
import pickle

import numpy as np
import pytest
import torch

import ref_harness as rh
import sac_port as sp
from _golden import rel_l2, rel_scalar

pytestmark = pytest.mark.skipif(not rh.available(), reason="no reference checkout (B200SAC_REFERENCE, /root/reference, baseline/_ref)")


def _set(named, params):
    with torch.no_grad():
        for k, p in named.items():
            p.data.copy_(params[k].reshape(p.shape))


@pytest.mark.parametrize("family,overrides", [
    ("LL", dict(batch_size=64)),
    ("MS", dict(batch_size=60, actor=dict(actor_hidden_dim=[32, 48]), critic=dict(critic_hidden_dim=[40, 24]))),
])
def test_port_follows_the_live_reference(family, overrides):
    import gen_golden as gg
    lrn, _ = rh.make_learner(family, overrides, seed=3)
    spec = gg._spec_of(lrn, family, family == "MS")
    params = sp.init_params(spec, seed=5)
    named = gg._named_params(lrn, family)
    _set(named, params)
    port = sp.PortLearner(spec, params)
    g = torch.Generator().manual_seed(1)
    for i in range(4):
        b = sp.synthetic_batch(spec, seed=300 + i)
        e1, e2 = torch.randn(spec.batch, spec.act_dim, generator=g), torch.randn(spec.batch, spec.act_dim, generator=g)
        lrn.memory.sample = (lambda bb: (lambda: tuple(t.clone() for t in bb)))(b)
        with rh.injected_eps([e1, e2]) as q:
            res = lrn.update()
            assert not q
        o = port.update_SAC(*b, e1, e2)
        assert rel_scalar(o["critic_loss"], res[0]) <= 1e-5 and rel_scalar(o["actor_loss"], res[1]) <= 1e-5, (i, o, res)
    got = port.params()
    for k, p in named.items():
        if k == "log_alpha":
            assert (got[k] - p.detach()).abs().max().item() <= 1e-6
        else:
            assert rel_l2(got[k], p.detach()) <= 1e-5, (k, rel_l2(got[k], p.detach()))


def test_published_blob_loads_into_the_reference_actor():
    """parameters_blob() is host logic: exercised here with a stand-in core whose published views are the state of a live
    reference Actor; the unpickled blob must load into a second reference Actor and make it identical."""
    from distributed_sac_b200 import names
    from distributed_sac_b200.learner import _BaseLearner
    lrn, mod = rh.make_learner("LL", dict(batch_size=64), seed=7)
    km = names.actor_key_map("LL", 3)
    sd = lrn.actor.state_dict()

    # stand-in core: the arena of a LunarLander learner filled with the live reference Actor's state; the device gather of
    # b200sac_blob_* is emulated on the CPU (tests/test_host_logic.py::BlobStubLib)
    import numpy as np
    import distributed_sac_b200.core as core_mod
    from distributed_sac_b200.core import CoreConfig, SacCore, layout
    from test_host_logic import BlobStubLib
    cfg = CoreConfig(batch=64)
    table, arena, _tr = layout(cfg)
    flat = np.zeros(arena, np.float32)
    core = object.__new__(SacCore)
    core.lib, core._h, core.cfg, core.table = BlobStubLib(flat), None, cfg, table
    for ref, canon in km.items():
        off, rows, cols, _t, _o, pitch = table[canon]
        flat[off:off + rows * pitch].reshape(rows, pitch)[:, :cols] = sd[ref].detach().numpy().reshape(rows, cols)
    shim = _BaseLearner.__new__(_BaseLearner)
    shim.core = core
    shim._key_map = lambda net: km
    real_stream = core_mod._stream
    core_mod._stream = lambda: None
    try:
        blob = shim.parameters_blob()
    finally:
        core_mod._stream = real_stream
    params = pickle.loads(blob)                                   # Player.pull_parameters: _pickle.loads(server.get('parameters'))
    other, _ = rh.make_learner("LL", dict(batch_size=64), seed=99)
    assert not torch.equal(other.actor.state_dict()["mu_log_std_layer.weight"], sd["mu_log_std_layer.weight"])
    other.actor.load_state_dict(params["actor"])                  # strict: every key, every shape
    for k, v in sd.items():
        assert torch.equal(other.actor.state_dict()[k], v), k
    x = torch.randn(5, 8)
    assert torch.equal(other.actor(x)[0], lrn.actor(x)[0])


def test_reference_load_checkpoint_reads_the_keys_we_write():
    """Key sets of the reference-written fixtures == what each reference load_checkpoint() indexes (LL/learner.py:165-182,
    MS:176-190, C10:200-217); tests/test_gpu_checkpoint.py proves the drop-in learner writes exactly these trees."""
    import os
    from _golden import GOLDEN
    want = {
        "ref_ckpt_ll_small": {"episode_idx", "total_step", "local_critic_1", "local_critic_2", "critic_optimizer", "target_critic_1",
                              "target_critic_2", "actor", "actor_optimizer", "log_alpha", "log_alpha_optimizer", "alpha"},
        "ref_ckpt_ms_small": {"update_iteration", "total_step", "local_critic", "critic_optimizer", "target_critic", "actor",
                              "actor_optimizer", "log_alpha", "log_alpha_optimizer", "alpha"},
        "ref_ckpt_c10m_small": {"update_iteration", "total_step", "context_encoder", "context_encoder_optimizer", "local_critic",
                                "critic_optimizer", "target_critic", "actor", "actor_optimizer", "log_alpha", "log_alpha_optimizer", "alpha"},
    }
    for name, keys in want.items():
        ck = torch.load(os.path.join(GOLDEN, name + ".tar"), map_location="cpu", weights_only=False)
        assert set(ck) == keys, (name, set(ck) ^ keys)
    # and the reference's own modules accept the fixture's state_dicts (the format is theirs)
    lrn, _ = rh.make_learner("LL", dict(batch_size=64), seed=1)
    ck = torch.load(os.path.join(GOLDEN, "ref_ckpt_ll_small.tar"), map_location="cpu", weights_only=False)
    lrn.actor.load_state_dict(ck["actor"])
    lrn.local_critic_1.load_state_dict(ck["local_critic_1"])
    lrn.critic_optimizer.load_state_dict(ck["critic_optimizer"])
    assert int(lrn.critic_optimizer.state_dict()["state"][0]["step"]) == 2
