"""Generate the reference-WRITTEN checkpoint fixtures tests/golden/ref_ckpt_*.tar (+ .npz) (CPU, build container only).

    python oracle/gen_checkpoints.py

For each family a toy-sized UNMODIFIED reference learner takes two update() steps (so that every Adam state exists),
calls ITS OWN save_checkpoint() -- the file that lands on disk is committed as the fixture --, then takes one more
step on a stored minibatch; losses and the resulting parameters go into the .npz next to it.  The GPU tests
(tests/test_gpu_checkpoint.py) feed the .tar to the drop-in learner's load_checkpoint(), check every tensor / Adam
moment / step counter against the file, replay the stored step, and compare the structure of a checkpoint saved by the
drop-in learner with the reference's (keys, nesting, shapes, dtypes) -- SURVEY.md 8(f) rank 3.
"""
import glob
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh                      # noqa: E402
import sac_port as sp                         # noqa: E402
import care_port as cp                        # noqa: E402
import gen_golden as gg                       # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

ENC = dict(hidden_dims_contextEnc=[24, 20], output_dim_contextEnc=12, embedding_dim_contextEnc=12,
           hidden_dims_mixtureEnc=[28], output_dim_mixtureEnc=12, num_encoders=4)
CASES = {
    "ref_ckpt_ll_small": dict(variant="LL", overrides=dict(batch_size=64)),
    "ref_ckpt_vs_small": dict(variant="VS", overrides=dict(batch_size=96, actor_hidden_dim=[64, 48, 32], critic_hidden_dim=[40, 72, 56])),
    "ref_ckpt_ms_small": dict(variant="MS", overrides=dict(batch_size=120, actor=dict(actor_hidden_dim=[48, 64, 32]),
                                                           critic=dict(critic_hidden_dim=[32, 64, 48]))),
    "ref_ckpt_c10m_small": dict(variant="C10", overrides=dict(batch_size=120, actor=dict(actor_hidden_dim=[64, 48, 32]),
                                                              critic=dict(critic_hidden_dim=[40, 72, 56]),
                                                              encoder=dict(ENC, output_dim_contextEnc=16, embedding_dim_contextEnc=16))),
    "ref_ckpt_c10o_small": dict(variant="C10", overrides=dict(use_modified_care=False, batch_size=120, actor=dict(actor_hidden_dim=[64, 48, 32]),
                                                              critic=dict(critic_hidden_dim=[40, 72, 56]), encoder=dict(ENC))),
}


def make(name, variant, overrides):
    work = tempfile.mkdtemp(prefix="b200sac_ckpt_")
    lrn, _ = rh.make_learner(variant, overrides, seed=11, workdir=work)
    care = variant == "C10"
    if care:
        enc = dict(lrn.encoder_cfg)
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
        batch_of = lambda seed: cp.synthetic_batch(spec, seed=seed)
        named_of = gg._care_named
    else:
        weighted = bool((overrides or {}).get("use_weighted_loss", variant == "MS"))
        spec = gg._spec_of(lrn, variant, weighted)
        batch_of = lambda seed: sp.synthetic_batch(spec, seed=seed)
        named_of = lambda l: gg._named_params(l, variant)
    g = torch.Generator().manual_seed(5)
    old = os.getcwd()
    os.chdir(work)
    try:
        for i in range(2):
            b = batch_of(900 + i)
            lrn.memory.sample = (lambda bb: (lambda: tuple(t.clone() for t in bb)))(b)
            e = [torch.randn(spec.batch, spec.act_dim, generator=g) for _ in range(2)]
            with rh.injected_eps(e):
                lrn.update()
        lrn.total_step = 4321
        before = set(glob.glob(os.path.join(work, "**", "*.tar"), recursive=True))
        os.makedirs(os.path.dirname(lrn.save_model_path + "x"), exist_ok=True)     # (the reference concatenates its path, LL/learner.py:163)
        lrn.save_checkpoint(12)
        new = set(glob.glob(os.path.join(work, "**", "*.tar"), recursive=True)) - before
        assert len(new) == 1, new
        shutil.copy(new.pop(), os.path.join(OUT, name + ".tar"))
        # one more step on a stored minibatch
        b = batch_of(950)
        e1, e2 = torch.randn(spec.batch, spec.act_dim, generator=g), torch.randn(spec.batch, spec.act_dim, generator=g)
        lrn.memory.sample = lambda: tuple(t.clone() for t in b)
        with rh.injected_eps([e1, e2]):
            res = lrn.update()
    finally:
        os.chdir(old)
    with open(os.path.join(work, "cfg.json")) as f:
        cfg_used = json.load(f)
    d = {"spec": json.dumps(spec.to_json()), "variant": variant, "cfg": json.dumps(cfg_used),
         "losses": np.array(list(res) + [float("nan")] * (3 - len(res)), np.float64), "eps_next": e1.numpy(), "eps_cur": e2.numpy()}
    for j, key in enumerate(("s", "a", "r", "s2", "d")):
        d["batch/" + key] = b[j].numpy()
    for k, p in named_of(lrn).items():
        d["p_out/" + k] = p.detach().clone().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, os.path.getsize(os.path.join(OUT, name + ".tar")) // 1024, "KB tar;", "losses", list(res))


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        make(n, **CASES[n])
