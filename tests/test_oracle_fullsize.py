"""CPU: the oracle ports against the full-size fixtures of the UNMODIFIED reference (BASELINE.json configs 3-5 at their
configured shapes -- VSAC 39/4/400^3/B1024, MTSAC B1280 weighted, CARE(M)/CARE(O) B1280 K6 -- and a 100-step LunarLander
chain; oracle/gen_golden.py FULL_CASES).  Same ATen ops in the same order: the port reproduces the reference to ~1e-6,
which pins the oracle the GPU tests compare with (tests/test_gpu_fullsize.py) to the real code at the real sizes."""
import math

import pytest
import torch

from _golden import FULL_CARE_CASES, FULL_CASES, FullCase, rel_l2, rel_scalar

TIGHT = 2e-5      # port vs reference: identical ATen kernels; losses and sampled state far below the 1e-4 parity bar


@pytest.mark.parametrize("name", FULL_CASES + FULL_CARE_CASES)
def test_port_reproduces_reference_at_full_size(name):
    c = FullCase(name)
    torch.set_num_threads(min(8, torch.get_num_threads()))
    port = c.make_port()
    for i in range(c.n_steps):
        o = port.update_SAC(*c.batches[i], c.eps_next[i], c.eps_cur[i], want_intermediates=(i == 0))
        if i == 0:
            for k, ref in c.i0.items():
                assert rel_l2(o[k].detach().reshape(ref.shape), ref) <= TIGHT, (k, rel_l2(o[k].detach().reshape(ref.shape), ref))
        assert rel_scalar(o["critic_loss"], c.losses[i, 0]) <= TIGHT, ("critic_loss", i, o["critic_loss"], c.losses[i, 0])
        assert rel_scalar(o["actor_loss"], c.losses[i, 1]) <= TIGHT, ("actor_loss", i)
        if not math.isnan(c.losses[i, 2]):
            assert rel_scalar(o["entropy"], c.losses[i, 2]) <= TIGHT, ("entropy", i)
    st = port.adam_state()
    c.check_summary("p_out", port.params(), TIGHT, "port parameters")
    c.check_summary("m_out", st["m"], TIGHT, "port Adam m")
    c.check_summary("v_out", st["v"], TIGHT, "port Adam v")
    assert tuple(int(x) for x in st["step"]) == tuple(int(x) for x in c.step_out)
