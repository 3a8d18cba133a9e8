"""CPU: the CARE(M) oracle port reproduces the fixture generated from the real reference."""
import torch

import care_port as cp
from _golden import CareCase, REL, check_state, rel_l2, rel_scalar


import pytest


@pytest.mark.parametrize("name", ["care_small_s4", "care_o_small_s4", "care_mt1_small_s3"])
def test_care_port_matches_reference_fixture(name):
    c = CareCase(name)
    torch.set_num_threads(4)
    lrn = cp.CarePortLearner(c.spec, c.p_in)
    for i in range(c.n_steps):
        out = lrn.update(*c.step_batch(i), eps_next=c.eps_next[i], eps_cur=c.eps_cur[i], want_intermediates=(i == 0))
        if i == 0:
            for k, ref in c.i0.items():
                assert rel_l2(out[k], ref) <= REL, (k, rel_l2(out[k], ref))
        assert rel_scalar(out["critic_loss"], c.losses[i, 0]) <= REL
        assert rel_scalar(out["actor_loss"], c.losses[i, 1]) <= REL
        assert rel_scalar(out["entropy"], c.losses[i, 2]) <= REL
    st = lrn.adam_state()
    check_state(c, lrn.params(), st["m"], st["v"], st["step"])
