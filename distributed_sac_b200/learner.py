"""Drop-in `Learner` classes: the reference's Python surface over the CUDA hot path.

Same constructor signatures, cfg JSONs, method names, return values, state_dict
keys and checkpoint format as the reference learners; `update()` /
`update_SAC()` run in libb200sac instead of ~2000 eager ATen ops.

  LunarLanderLearner  <- LunarLander_Distributed_SAC/src/learner.py:21-316
  VSACLearner         <- MT1_Distributed_VSAC/src/learner.py:22-322
  MTSACLearner        <- MT10_Distributed_MTSAC/src/learner.py:22-414

What is deliberately NOT copied: `to_device` forgetting the second
critics, `load_checkpoint` touching a non-existent `actor.optimizer` (ours loads),
the checkpoint filename concatenated onto the directory name without a slash (ours
joins paths), `buffer_size` hard-coded to 1e5 in the LL learner (kept as the LL default
but overridable).  Redis / Ray / TensorBoard are out of scope: the learner takes a
`server` object (anything with redis-py's set/get/rpush/delete/scan_iter/pipeline API);
if none is given and `redis` is importable it connects to localhost like the reference.
"""
import itertools
import json
import os
import pickle as _pickle
import time
from datetime import datetime

import numpy as np
import torch

from . import _lib, names
from .core import CoreConfig, Replay, SacCore
from .replay_buffer import ReplayBuffer, MTReplayBuffer


class Decoder(json.JSONDecoder):
    """Numeric *strings* become ints (LL/utils.py:4-20)."""

    def decode(self, s):
        return self._coerce(super().decode(s))

    def _coerce(self, o):
        if isinstance(o, str):
            try:
                return int(o)
            except ValueError:
                return o
        if isinstance(o, dict):
            return {k: self._coerce(v) for k, v in o.items()}
        if isinstance(o, list):
            return [self._coerce(v) for v in o]
        return o


def cfg_read(path):
    with open(path, "r") as f:
        return json.loads(f.read(), cls=Decoder)


def _default_server():
    try:
        import redis
    except ImportError as e:       # pragma: no cover - environment dependent
        raise RuntimeError("no `server` given and the `redis` package is not installed") from e
    return redis.StrictRedis(host="localhost")


class _Handle:
    """Stand-in for the nn.Module arguments of Learner.soft_update(local, target, tau)."""

    def __init__(self, learner, net):
        self._learner, self.net = learner, net

    def state_dict(self):
        return self._learner._module_state_dict(self.net)

    def load_state_dict(self, sd):
        self._learner._load_module_state_dict(self.net, sd)


class _BaseLearner:
    family = "LL"
    log_name = "LunarLander_Distributed_SAC"
    _iteration_key = "episode_idx"

    # ---- construction ---------------------------------------------------------------------
    def _init_common(self, cfg_path, write_mode, save_period, checkpoint_path, server, device_index, seed,
                     replay_where, precision=1):
        self.cfg = cfg_read(cfg_path)
        self.precision = precision
        self.write_mode = write_mode
        self.save_period = save_period
        self.total_step = 0
        self.episode_idx = 0
        self.datetime = str(datetime.now())[:-7]
        self._set_dims()
        self.device = torch.device(self.cfg.get("device", "cuda"))
        if self.device.type != "cuda":
            raise RuntimeError("the B200 learner only runs on CUDA devices (cfg 'device' must be 'cuda')")
        self.device_index = device_index if self.device.index is None else self.device.index
        self.server = server if server is not None else _default_server()
        for key in list(self.server.scan_iter()):      # LL/learner.py:34-36
            self.server.delete(key)
        self.core = SacCore(self._core_config(), self.device_index, seed=seed)
        self.memory = self._make_memory(replay_where)
        self.memory.start()
        self.save_model_path = os.path.join("saved_models", self.log_name, self.datetime)
        self.log_file = os.path.join("log", "log_" + self.log_name, f"{self.log_name}_log_{self.datetime}.txt")
        self.alpha = torch.ones(max(self.num_tasks, 1))     # stale build-time value the reference also saves
        if checkpoint_path is not None:
            self.load_checkpoint(checkpoint_path)
            print("######## load checkpoint completely ########")

    def _core_config(self):
        c = self.cfg
        return CoreConfig(state_dim=self.state_dim, act_dim=self.action_dim, actor_hidden=list(self.actor_hidden_dim),
                          critic_hidden=list(self.critic_hidden_dim), batch=self.batch_size, num_tasks=self.num_tasks,
                          weighted_loss=self.use_weighted_loss, replicas=1, precision=self.precision, gamma=float(self.gamma), tau=float(self.tau),
                          reward_scale=float(self.reward_scale), lr_actor=float(self.lr_actor),
                          lr_critic=float(self.lr_critic),
                          action_scale=(self.action_bound[1] - self.action_bound[0]) / 2,
                          log_alpha_init=float(c["log_alpha"]))

    # ---- reference method surface -------------------------------------------------------------
    def update(self):
        """Learner.update(): sample a minibatch, one SAC gradient step, return python floats."""
        l = self.memory.step_core(self.core)          # (critic, actor, alpha loss, entropy) python floats
        return (l[0], l[1]) if self.num_tasks == 0 else (l[0], l[1], l[3])

    def update_SAC(self, states, actions, rewards, next_states, dones, alpha=None, retain_graph=False,
                   eps_next=None, eps_cur=None):
        """Learner.update_SAC() on an explicit minibatch (device or host tensors).  `alpha` is accepted
        for signature compatibility; the step uses exp(log_alpha) snapshotted in-kernel exactly like update()."""
        if states.is_cuda:
            self.core.step(states, actions, rewards, next_states, dones, eps_next, eps_cur)
            losses = self.core.read_losses(1)[0]
        else:
            losses = self.core.step_host(states, actions, rewards, next_states, dones, eps_next, eps_cur)
        return self._loss_tuple(losses)

    def update_many(self, n):
        """n pipelined update() calls; returns a [n][k] tensor of the per-step losses."""
        self.core.step_sampled(self.memory.ring, n)
        L = self.core.read_losses(min(n, 1024))[:, 0]
        return L[:, [0, 1]] if self.num_tasks == 0 else L[:, [0, 1, 3]]

    def _loss_tuple(self, losses):
        l = losses.reshape(-1, 4)[0]
        if self.num_tasks == 0:
            return float(l[0]), float(l[1])
        return float(l[0]), float(l[1]), float(l[3])

    def optimizer_zero_grad(self):
        """No-op: gradients never outlive a step in the fused path."""

    def soft_update(self, local_model=None, target_model=None, tau=None):
        """theta_target = tau*theta_local + (1-tau)*theta_target for both critics (LL/learner.py:126-137).
        The step itself already applies the per-step Polyak update; this entry point exists for the
        hard copy in run() (tau=1.0) and for callers that drive it by hand.  Called with the same
        (local, target) pair twice in a row -- as run() does for critic 1 and 2 -- the second call of a
        tau=1.0 copy is idempotent."""
        tau = self.tau if tau is None else tau
        self.core.soft_update(float(tau))

    def wait_until_memoryReady(self):
        while len(self.memory) <= self.start_memory_len:
            time.sleep(0.1)

    def act_batch(self, states, stochastic=True):
        """Actor.get_action (LL/model.py:67-82) for a batch of environments at once with the learner's current actor:
        states [n][obs_dim] (n <= 2*batch_size) -> actions [n][action_dim] on the CPU."""
        # the LunarLander actor's deterministic action is k*mu, every other family's k*tanh(mu) (LL/model.py:78-80, VS/model.py:78-80)
        return self.core.act(torch.as_tensor(states), stochastic=stochastic, no_tanh=(self.family == "LL"))

    # ---- parameter publication (LL/learner.py:272-276; consumed by Player.pull_parameters, player.py:75-85) ----
    _published = ("actor",)

    def publish_begin(self):
        """Start an asynchronous snapshot of the published modules (b200sac_publish_begin): a device-side copy in
        stream order after the steps enqueued so far, then D2H into pinned memory on a private stream."""
        self._pub_maps = [(net, self._key_map(net)) for net in self._published]
        # (+ the temperature: the logger's alpha then comes from the same consistent snapshot, not from a second,
        #  stream-ordered read that would wait for whatever step is running by then)
        self.core.publish_begin({canon for _, m in self._pub_maps for canon in m.values()} | {"log_alpha"})

    def publish_wait(self):
        named = self.core.publish_wait()
        return {net: {ref: named[canon] for ref, canon in m.items()} for net, m in self._pub_maps}

    def get_parameters(self):
        """{'actor': state_dict on the CPU} with the reference's key names -- only the published slices cross PCIe."""
        self.publish_begin()
        return self.publish_wait()

    def parameters_blob(self, blocking=True):
        """The bytes Learner.run() stores under the Redis key 'parameters' (LL/learner.py:298-299): _pickle.dumps of
        {'actor': state_dict_on_cpu[, ...]} with the reference's key names -- what Player.pull_parameters unpickles
        (LL/player.py:75-85).  The shapes never change, so the pickle stream is built ONCE (from tensors of the same shapes)
        and handed to the library as a byte image with the positions of the float payloads (b200sac_blob_template); per
        publication a kernel gathers the current floats into a device copy of the image and ONE D2H copy brings the finished
        byte string to pinned memory -- the host never touches a tensor.
        blocking=True: publish now (get_parameters + dumps semantics); False: collect the blob started by blob_begin()."""
        self._ensure_blob()
        if blocking:
            self.core.blob_begin()
        return self.core.blob_wait()[0]

    def _ensure_blob(self):
        if getattr(self, "_blob_ready", False):
            return
        self._pub_maps = [(net, self._key_map(net)) for net in self._published]
        rng = np.random.default_rng(12345)
        fake = {net: {ref: torch.from_numpy(rng.random(self.core.tensor_shape(canon), dtype=np.float32) + 1.0) for ref, canon in m.items()}
                for net, m in self._pub_maps}
        blob = _pickle.dumps(fake)
        slots = []
        for net, m in self._pub_maps:
            for ref, canon in m.items():
                payload = fake[net][ref].numpy().tobytes()
                at = blob.find(payload)
                if at < 0 or blob.find(payload, at + 1) >= 0:
                    raise RuntimeError(f"cannot locate the payload of {net}.{ref} in the pickle stream")
                slots.append((canon, at))
        # (+ the temperature behind the image: the logger's alpha then comes from the same consistent snapshot)
        self.core.blob_template(blob, slots, extra=("log_alpha",))
        self._blob_ready = True

    def _blob_from_views(self, views):
        """Host-side variant (snapshot views -> bytes patched into the pickle template); kept for callers of
        publish_begin()/core.publish_views()."""
        tpl = getattr(self, "_blob_tpl", None)
        if tpl is None:
            rng = np.random.default_rng(12345)
            fake = {net: {ref: torch.from_numpy(rng.random(views[canon].shape, dtype=np.float32) + 1.0) for ref, canon in m.items()}
                    for net, m in self._pub_maps}
            blob = _pickle.dumps(fake)
            where = []
            for net, m in self._pub_maps:
                for ref, canon in m.items():
                    payload = fake[net][ref].numpy().tobytes()
                    at = blob.find(payload)
                    if at < 0 or blob.find(payload, at + 1) >= 0:
                        raise RuntimeError(f"cannot locate the payload of {net}.{ref} in the pickle stream")
                    where.append((canon, at, views[canon].shape))
            tpl = self._blob_tpl = (blob, where)
        blob, where = tpl
        buf = bytearray(blob)
        for canon, at, shape in where:
            np.frombuffer(buf, dtype=np.float32, count=int(np.prod(shape)), offset=at).reshape(shape)[...] = views[canon]
        return bytes(buf)

    def my_print(self, content):
        os.makedirs(os.path.dirname(self.log_file), exist_ok=True)
        with open(self.log_file, "a") as writer:
            print(content)
            writer.write(content + "\n")

    # ---- state_dict plumbing -----------------------------------------------------------------
    def _key_map(self, net):
        na, nc = len(self.actor_hidden_dim) + 1, len(self.critic_hidden_dim) + 1
        if net == "actor":
            return names.actor_key_map(self.family, na)
        which = 1 if "1" in net else 2
        return names.critic_key_map(self.family, nc, which, target="target" in net)

    def _module_state_dict(self, net, which=_lib.PARAMS, named=None):
        named = named if named is not None else self.core.get_named(which)
        return {ref: named[canon].clone() for ref, canon in self._key_map(net).items()}

    def _load_module_state_dict(self, net, sd, which=_lib.PARAMS):
        km = self._key_map(net)
        missing = set(km) - set(sd)
        if missing:
            raise KeyError(f"state_dict for {net} lacks {sorted(missing)}")
        self.core.set_named({km[k]: v for k, v in sd.items() if k in km}, which, strict=False)

    @property
    def log_alpha(self):
        return self.core.read_named("log_alpha")

    # ---- checkpointing (reference format, LL/learner.py:144-182) ------------------------------
    def _adam_state_dict(self, canon_names, lr, step, n_frozen_first=0):
        """torch.optim.Adam.state_dict() layout.  n_frozen_first: parameters that sit in the optimizer's param group before
        the trainable ones but never get state (the frozen nn.Embedding of the CARE context encoder is parameter 0 of
        `Adam(context_encoder.parameters())`, MT10_Distributed_CARE/src/learner.py:137-141)."""
        m, v = self.core.get_named(_lib.ADAM_M), self.core.get_named(_lib.ADAM_V)
        state = {n_frozen_first + i: {"step": torch.tensor(float(step)), "exp_avg": m[n], "exp_avg_sq": v[n]}
                 for i, n in enumerate(canon_names)} if step > 0 else {}
        # the param-group keys are whatever this torch's Adam writes (they changed across releases); values = Adam defaults
        group = dict(torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=lr).state_dict()["param_groups"][0])
        group["params"] = list(range(n_frozen_first + len(canon_names)))
        return {"state": state, "param_groups": [group]}

    def _canon(self, nets):
        out = []
        for net in nets:
            out += list(self._key_map(net).values())
        return out

    def _load_adam(self, sd, canon_names, slot, n_frozen_first=0):
        m, v, step = {}, {}, 0
        for i, n in enumerate(canon_names):
            st = sd["state"].get(n_frozen_first + i)
            if st is None:
                continue
            m[n], v[n] = st["exp_avg"], st["exp_avg_sq"]
            step = int(st["step"])
        if m:
            self.core.set_named(m, _lib.ADAM_M, strict=False)
            self.core.set_named(v, _lib.ADAM_V, strict=False)
        steps = list(self.core.get_steps())
        steps[slot] = step
        self.core.set_steps(steps)

    def _checkpoint_path(self, idx):
        return os.path.join(self.save_model_path, f"checkpoint_{idx}.tar")

    def save_checkpoint(self, episode_idx):
        named = self.core.get_named()
        sc, sa, sl = self.core.get_steps()[:3]
        # LL / VSAC write 'episode_idx' (LL/learner.py:146, VS:148); MTSAC / CARE / MT1-CARE write and READ 'update_iteration'
        # (MS/learner.py:159,178; C10:180,202; C1:147,169)
        state = {self._iteration_key: episode_idx, "total_step": self.total_step}
        state.update(self._critic_checkpoint_entries(named))
        state["critic_optimizer"] = self._adam_state_dict(self._canon(("q1", "q2")), self.lr_critic, sc)
        state["actor"] = self._module_state_dict("actor", named=named)
        state["actor_optimizer"] = self._adam_state_dict(self._canon(("actor",)), self.lr_actor, sa)
        state["log_alpha"] = named["log_alpha"].clone()
        state["log_alpha_optimizer"] = self._adam_state_dict(["log_alpha"], self.lr_actor, sl)
        state["alpha"] = self.alpha.clone()
        os.makedirs(self.save_model_path, exist_ok=True)
        path = self._checkpoint_path(episode_idx)
        torch.save(state, path)
        return path

    def load_checkpoint(self, path):
        ck = torch.load(path, map_location="cpu", weights_only=False)
        self.episode_idx = ck.get("episode_idx", ck.get("update_iteration", 0))
        self.total_step = ck.get("total_step", 0)
        self._load_critic_checkpoint_entries(ck)
        self._load_module_state_dict("actor", ck["actor"])
        self.core.set_named({"log_alpha": ck["log_alpha"].detach().reshape(-1)}, strict=False)
        self._load_adam(ck["critic_optimizer"], self._canon(("q1", "q2")), 0)
        self._load_adam(ck["actor_optimizer"], self._canon(("actor",)), 1)
        self._load_adam(ck["log_alpha_optimizer"], ["log_alpha"], 2)
        if "alpha" in ck:
            self.alpha = ck["alpha"].detach().reshape(-1).clone()

    # ---- Redis-facing loop (LL/learner.py:191-201,278-316) -------------------------------------
    def write(self, update_iteration, critic_loss, actor_loss, entropy=None, log_alpha=None):
        self.server.rpush("critic_loss", _pickle.dumps((update_iteration, critic_loss)))
        self.server.rpush("actor_loss", _pickle.dumps((update_iteration, actor_loss)))
        if entropy is not None:
            self.server.rpush("entropy", _pickle.dumps((update_iteration, entropy)))
        alphas = np.exp(np.array(log_alpha, dtype=np.float32)) if log_alpha is not None else self.log_alpha.exp().numpy()
        self.server.rpush("alpha", _pickle.dumps((update_iteration, alphas)))

    def run(self, max_updates=None):
        self.server.set("update_iteration", _pickle.dumps(-1))
        self.server.set("parameters", self.parameters_blob(blocking=True))
        self.wait_until_memoryReady()
        self.my_print("######################### Start train #########################")
        self.soft_update(None, None, 1.0)           # copy parameters to target
        # Pipelined like this: while the GPU runs step k+1 (enqueued together with the assembly of its blob), the host hands
        # blob k to Redis -- the reference does the same things strictly one after the other (learner.py:296-316).  Two
        # blob images in the library make that safe.
        done = 0
        iters = (i for i in itertools.count() if i % self.update_delay == 0)
        update_iteration = next(iters)
        core = self.core
        self.memory.enqueue_step(core)                 # update(), split so that everything below overlaps the GPU
        core.blob_begin()
        while True:
            cur = update_iteration
            res = self._loss_tuple(core.read_losses(1)[0])             # step `cur` has finished
            last = max_updates is not None and done + 1 >= max_updates
            if cur % self.save_period == 0:
                self.save_checkpoint(cur)              # before step cur+1 is enqueued: the arena is the state after `cur`
            if not last:                               # the GPU goes on with step cur+1 while blob `cur` is still crossing PCIe
                update_iteration = next(iters)
                self.memory.enqueue_step(core)
                core.blob_begin()
            blob, extra = core.blob_wait()             # the OLDEST uncollected blob = the one of step `cur`
            self.server.set("update_iteration", _pickle.dumps(cur))
            self.server.set("parameters", blob)
            if self.write_mode:
                self.write(cur, *res, log_alpha=extra["log_alpha"])
                if cur % self.print_period == 0:
                    self.my_print("[Learner] Update_iteration: {0:<6} \t | actor_loss : {1:5.3f} \t | critic_loss : {2:5.3f}".format(
                        cur, res[1], res[0]))
            done += 1
            if last:
                return done


class _TwoCriticMixin:
    """LL / VS: four separate Critic modules in the checkpoint."""

    def _critic_checkpoint_entries(self, named):
        return {k: self._module_state_dict(net, named=named) for k, net in
                (("local_critic_1", "q1"), ("local_critic_2", "q2"), ("target_critic_1", "q1_target"),
                 ("target_critic_2", "q2_target"))}

    def _load_critic_checkpoint_entries(self, ck):
        for k, net in (("local_critic_1", "q1"), ("local_critic_2", "q2"), ("target_critic_1", "q1_target"),
                       ("target_critic_2", "q2_target")):
            self._load_module_state_dict(net, ck[k])

    @property
    def local_critic_1(self): return _Handle(self, "q1")
    @property
    def local_critic_2(self): return _Handle(self, "q2")
    @property
    def target_critic_1(self): return _Handle(self, "q1_target")
    @property
    def target_critic_2(self): return _Handle(self, "q2_target")
    @property
    def actor(self): return _Handle(self, "actor")


class LunarLanderLearner(_TwoCriticMixin, _BaseLearner):
    """LunarLander_Distributed_SAC/src/learner.py: Learner(cfg_path, update_delay=3, print_period=10,
    write_mode=True, save_period=1000, checkpoint_path=None)."""
    family = "LL"
    log_name = "LunarLander_Distributed_SAC"

    def __init__(self, cfg_path, update_delay=3, print_period=10, write_mode=True, save_period=1000,
                 checkpoint_path=None, *, server=None, device_index=0, seed=0, replay_where="host", buffer_size=None,
                 precision=1):
        self.update_delay, self.print_period = update_delay, print_period
        self._buffer_size_override = buffer_size
        self._init_common(cfg_path, write_mode, save_period, checkpoint_path, server, device_index, seed, replay_where,
                          precision)

    def _set_dims(self):
        c = self.cfg
        self.gamma, self.lr_actor, self.lr_critic = c["gamma"], c["lr_actor"], c["lr_critic"]
        self.batch_size, self.tau, self.reward_scale = int(c["batch_size"]), c["tau"], c["reward_scale"]
        self.start_memory_len = c["start_memory_len"]
        self.buffer_size = int(self._buffer_size_override or 1e5)      # LL/learner.py:72
        self.action_dim, self.state_dim, self.action_bound = 2, 8, [-1.0, 1.0]     # LL/learner.py:79-81
        self.actor_hidden_dim, self.critic_hidden_dim = [256, 256], [256, 256]     # LL/learner.py:102-107
        self.num_tasks, self.use_weighted_loss = 0, False

    def _make_memory(self, where):
        return ReplayBuffer(self.buffer_size, self.batch_size, 0, self.device, server=self.server, core=self.core,
                            where=where)


class VSACLearner(_TwoCriticMixin, _BaseLearner):
    """MT1_Distributed_VSAC/src/learner.py: Learner(cfg_path, write_mode, save_period, checkpoint_path)."""
    family = "VS"
    log_name = "MT1_Distributed_VSAC"

    def __init__(self, cfg_path, write_mode=True, save_period=1000, checkpoint_path=None, *, server=None,
                 device_index=0, seed=0, replay_where="host", precision=1):
        self._init_common(cfg_path, write_mode, save_period, checkpoint_path, server, device_index, seed, replay_where,
                          precision)

    def _set_dims(self):
        c = self.cfg
        self.gamma, self.lr_actor, self.lr_critic = c["gamma"], c["lr_actor"], c["lr_critic"]
        self.batch_size, self.tau, self.reward_scale = int(c["batch_size"]), c["tau"], c["reward_scale"]
        self.start_memory_len, self.buffer_size = c["start_memory_len"], int(c["buffer_size"])
        self.print_period, self.update_delay = int(c["print_period_learner"]), c["update_delay"]
        self.actor_hidden_dim, self.critic_hidden_dim = c["actor_hidden_dim"], c["critic_hidden_dim"]
        self.action_dim, self.state_dim, self.action_bound = 4, 39, [-1.0, 1.0]    # VS/learner.py:80-82
        self.num_tasks, self.use_weighted_loss = 0, False

    def _make_memory(self, where):
        return ReplayBuffer(self.buffer_size, self.batch_size, 0, self.device, server=self.server, core=self.core,
                            where=where)


class MTSACLearner(_BaseLearner):
    """MT10_Distributed_MTSAC/src/learner.py: Learner(train_classes, train_tasks, cfg_path, write_mode,
    save_period, checkpoint_path).  One learner over T tasks, one-hot appended to the state."""
    family = "MS"
    log_name = "MT10_Distributed_MTSAC"
    _iteration_key = "update_iteration"

    def __init__(self, train_classes, train_tasks, cfg_path, write_mode=True, save_period=1000, checkpoint_path=None,
                 *, server=None, device_index=0, seed=0, replay_where="host", precision=1):
        self.train_classes, self.train_tasks = train_classes, train_tasks
        self._init_common(cfg_path, write_mode, save_period, checkpoint_path, server, device_index, seed, replay_where,
                          precision)

    def _set_dims(self):
        c = self.cfg
        self.actor_cfg, self.critic_cfg = c["actor"], c["critic"]
        self.gamma = c["gamma"]
        self.lr_actor, self.lr_critic = self.actor_cfg["lr_actor"], self.critic_cfg["lr_critic"]
        self.batch_size, self.tau, self.reward_scale = int(c["batch_size"]), c["tau"], c["reward_scale"]
        self.start_memory_len, self.buffer_size = c["start_memory_len"], int(c["buffer_size"])
        self.print_period, self.update_delay = int(c["print_period_learner"]), c["update_delay"]
        self.num_tasks = int(c["num_tasks"])
        self.use_weighted_loss = bool(c.get("use_weighted_loss", False))
        self.state_dim, self.action_dim = int(self.actor_cfg["state_dim"]), int(self.actor_cfg["action_dim"])
        self.action_bound = self.actor_cfg["action_bound"]
        self.actor_hidden_dim = self.actor_cfg["actor_hidden_dim"]
        self.critic_hidden_dim = self.critic_cfg["critic_hidden_dim"]

    def _make_memory(self, where):
        return MTReplayBuffer(self.buffer_size, self.batch_size, 0, self.device, self.num_tasks, server=self.server,
                              core=self.core, where=where)

    def _critic_checkpoint_entries(self, named):
        loc = {**self._module_state_dict("q1", named=named), **self._module_state_dict("q2", named=named)}
        tgt = {**self._module_state_dict("q1_target", named=named), **self._module_state_dict("q2_target", named=named)}
        return {"local_critic": loc, "target_critic": tgt}

    def _load_critic_checkpoint_entries(self, ck):
        for net in ("q1", "q2"):
            self._load_module_state_dict(net, ck["local_critic"])
            self._load_module_state_dict(net + "_target", ck["target_critic"])

    def get_log_alpha(self, mtobss):
        """(batch, 1) log_alpha of each row's task (MS/learner.py:213-233)."""
        one_hots = torch.as_tensor(mtobss)[:, -self.num_tasks:].float().cpu()
        return one_hots @ self.log_alpha.reshape(-1, 1)


class CARELearner(MTSACLearner):
    """MT10_Distributed_CARE/src/learner.py with `use_modified_care: true` (CARE(M)):
    Learner(train_classes, train_tasks, cfg_path, write_mode, save_period, checkpoint_path).

    One actor / twin-Q critic over the encoded state [mlp_context(z) | attention-mixed K encoders(state)],
    z = frozen RoBERTa embedding of the task (cfg 'encoder' block; metadata JSONs are read relative to the
    working directory exactly like context_encoder.py:31-36).  The actor's state encoder is a hard copy of the
    critic's after every update (learner.py:402), so it is exported from the same tensors."""
    family = "C10"
    log_name = "MT10_Distributed_CARE"

    def _set_dims(self):
        super()._set_dims()
        c = self.cfg
        self.use_modified_care = bool(c.get("use_modified_care", False))
        self.use_weighted_loss = self.use_modified_care    # learner.py:313,349: use_weighted_loss = use_modified_care
        self.encoder_cfg = c["encoder"]

    def _core_config(self):
        cc = super()._core_config()
        e = self.encoder_cfg
        cc.care = True
        cc.num_encoders = int(e["num_encoders"])
        cc.mix_hidden = [int(x) for x in e["hidden_dims_mixtureEnc"]]
        cc.mix_out = int(e["output_dim_mixtureEnc"])
        cc.ctx_in = int(e.get("RoBERTa_embedding_dim", 768))      # MT1_Distributed_CARE's cfg has no such key (width of the JSON rows)
        cc.ctx_hidden = [int(x) for x in e["hidden_dims_contextEnc"]]
        cc.ctx_out = int(e["output_dim_contextEnc"])
        cc.tau_se = float(e.get("state_encoder_tau", 0.05))        # hard-coded 0.05 in MT1_Distributed_CARE/src/learner.py:311
        cc.care_original = not self.use_modified_care     # CARE(O): trainable context encoder with its own Adam
        cc.emb_dim = int(e["embedding_dim_contextEnc"])
        cc.lr_ctx = float(e["lr_contextEnc"])
        return cc

    def _init_common(self, *a, **kw):
        super()._init_common(*a, **kw)
        e = self.encoder_cfg
        emb_path, names_path = e.get("pretrained_embedding_json_path"), e.get("task_name_json_path")
        if emb_path and names_path and not (os.path.exists(emb_path) and os.path.exists(names_path)):
            # the reference fails fast here (context_encoder.py:31-36 opens both files); training on the random
            # stand-in table would publish meaningless context vectors without any error
            if not os.environ.get("B200SAC_ALLOW_RANDOM_EMBEDDING"):
                raise FileNotFoundError(f"CARE task-embedding files not found (cwd {os.getcwd()}): {emb_path}, {names_path}; "
                                        "set B200SAC_ALLOW_RANDOM_EMBEDDING=1 to train on a random embedding table")
        if emb_path and names_path and os.path.exists(emb_path) and os.path.exists(names_path):
            table, order = cfg_read(emb_path), cfg_read(names_path)
            E = torch.tensor([table[n] for n in order], dtype=torch.float32)       # context_encoder.py:43-47
            if E.shape != (self.num_tasks, self.core.cfg.ctx_in):
                raise ValueError(f"pretrained embedding has shape {tuple(E.shape)}, cfg says ({self.num_tasks}, {self.core.cfg.ctx_in})")
            self.core.set_named({"embedding": E}, strict=False)

    def _enc_map(self, prefix):
        cc = self.core.cfg
        n_ctx = len(cc.ctx_hidden) + 1 if self.use_modified_care else 0      # CARE(O) has no mlp_context in the state encoder
        return names.care_encoder_key_map(len(cc.mix_hidden) + 1, len(cc.mix_hidden) + 1, n_ctx, prefix)

    def _cenc_map(self):
        """CARE(O) context encoder keys: embedding = Sequential(Embedding, ReLU, header[Linear, ReLU, Linear, ReLU]),
        mlp = build_mlp(...) (context_encoder.py:59-89)."""
        m = {"embedding.2.0.weight": "cenc.0.weight", "embedding.2.0.bias": "cenc.0.bias",
             "embedding.2.2.weight": "cenc.1.weight", "embedding.2.2.bias": "cenc.1.bias"}
        for j in range(len(self.core.cfg.ctx_hidden) + 1):
            for kind in ("weight", "bias"):
                m[f"mlp.{2 * j}.{kind}"] = f"cenc.{2 + j}.{kind}"
        return m

    def _key_map(self, net):
        if net == "context_encoder":
            km = {"embedding.0.weight": "embedding"}
            if not self.use_modified_care:
                km.update(self._cenc_map())
            return km
        if net == "actor":            # policy MLP + the tied copy of the critic's state encoder
            return {**names.actor_key_map("MS", len(self.actor_hidden_dim) + 1), **self._enc_map("cse")}
        if net in ("critic_se", "target_se"):
            return self._enc_map("cse" if net == "critic_se" else "tse")
        which = 1 if "1" in net else 2
        return names.critic_key_map("MS", len(self.critic_hidden_dim) + 1, which, target="target" in net)

    _published = ("context_encoder", "actor")      # C10/learner.py:412-417

    def _critic_checkpoint_entries(self, named):
        d = super()._critic_checkpoint_entries(named)
        d["local_critic"].update(self._module_state_dict("critic_se", named=named))
        d["target_critic"].update(self._module_state_dict("target_se", named=named))
        d["context_encoder"] = self._module_state_dict("context_encoder", named=named)
        # Adam(context_encoder.parameters()): parameter 0 is the frozen embedding (no state); CARE(M) has nothing else, the
        # reference still saves -- and its load_checkpoint indexes -- the (empty) optimizer state (C10/learner.py:184,206)
        cenc = [] if self.use_modified_care else list(self._cenc_map().values())
        step = 0 if self.use_modified_care else self.core.get_steps()[3]
        d["context_encoder_optimizer"] = self._adam_state_dict(cenc, self.core.cfg.lr_ctx, step, n_frozen_first=1)
        return d

    def _load_critic_checkpoint_entries(self, ck):
        super()._load_critic_checkpoint_entries(ck)
        self._load_module_state_dict("critic_se", ck["local_critic"])
        self._load_module_state_dict("target_se", ck["target_critic"])
        if "context_encoder" in ck:
            self._load_module_state_dict("context_encoder", ck["context_encoder"])
        if not self.use_modified_care and "context_encoder_optimizer" in ck:
            self._load_adam(ck["context_encoder_optimizer"], list(self._cenc_map().values()), 3, n_frozen_first=1)

    def _canon(self, nets):
        out = []
        for net in nets:
            if net == "actor":        # actor_optimizer holds mu_log_std_layer only (learner.py:146-149)
                out += list(names.actor_key_map("MS", len(self.actor_hidden_dim) + 1).values())
            else:
                out += list(self._key_map(net).values())
        if tuple(nets) == ("q1", "q2"):   # critic_optimizer = local_critic.parameters(): state encoder first (learner.py:150-153)
            out = list(self._enc_map("cse").values()) + out
        return out


# the reference modules all call their class `Learner`
Learner = LunarLanderLearner
