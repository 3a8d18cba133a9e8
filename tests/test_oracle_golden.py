"""CPU: the oracle port reproduces the fixtures generated from the real reference."""
import math

import pytest
import torch

import sac_port as sp
from _golden import CASES, Case, REL, check_state, rel_l2, rel_scalar


@pytest.mark.parametrize("name", CASES)
def test_port_matches_reference_fixture(name):
    c = Case(name)
    torch.set_num_threads(4)
    lrn = sp.PortLearner(c.spec, c.p_in, {"m": c.m_in, "v": c.v_in, "step": tuple(c.step_in)})
    for i in range(c.n_steps):
        out = lrn.update_SAC(*c.step_batch(i), eps_next=c.eps_next[i], eps_cur=c.eps_cur[i],
                             want_intermediates=(i == 0))
        if i == 0:
            for k, ref in c.i0.items():
                assert rel_l2(out[k], ref) <= REL, (k, rel_l2(out[k], ref))
        assert rel_scalar(out["critic_loss"], c.losses[i, 0]) <= REL
        assert rel_scalar(out["actor_loss"], c.losses[i, 1]) <= REL
        if not math.isnan(c.losses[i, 2]):
            assert rel_scalar(out["entropy"], c.losses[i, 2]) <= REL
    st = lrn.adam_state()
    check_state(c, lrn.params(), st["m"], st["v"], st["step"])
