"""SacCore: thin Python object over one libb200sac handle (R co-scheduled learners).

Host-side plumbing only -- torch is used for device buffers / streams; all math
runs in the CUDA library.  The reference-shaped `Learner` classes
(distributed_sac_b200/learner.py) are built on this.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib


@dataclass
class CoreConfig:
    state_dim: int = 8
    act_dim: int = 2
    actor_hidden: List[int] = field(default_factory=lambda: [256, 256])
    critic_hidden: List[int] = field(default_factory=lambda: [256, 256])
    batch: int = 256
    num_tasks: int = 0
    weighted_loss: bool = False
    replicas: int = 1
    precision: int = 0
    gamma: float = 0.99
    tau: float = 0.005
    reward_scale: float = 1.0
    lr_actor: float = 3e-4
    lr_critic: float = 3e-4
    lr_alpha: Optional[float] = None      # the reference uses lr_actor (LL/learner.py:124)
    action_scale: float = 1.0
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    log_alpha_init: float = 0.0
    # CARE(M) state encoders (cfg "encoder" block of MT10_Distributed_CARE_cfg.json)
    care: bool = False
    num_encoders: int = 6
    mix_hidden: List[int] = field(default_factory=lambda: [50])
    mix_out: int = 50
    ctx_in: int = 768
    ctx_hidden: List[int] = field(default_factory=lambda: [50, 50])
    ctx_out: int = 50
    tau_se: float = 0.05
    care_original: bool = False          # CARE(O): trainable context encoder (use_modified_care: false)
    emb_dim: int = 50
    lr_ctx: float = 3e-4

    @property
    def obs_dim(self):
        return self.state_dim + self.num_tasks

    def to_c(self):
        c = _lib.Cfg()
        c.state_dim, c.act_dim, c.num_tasks = self.state_dim, self.act_dim, self.num_tasks
        c.n_actor_hidden, c.n_critic_hidden = len(self.actor_hidden), len(self.critic_hidden)
        if max(c.n_actor_hidden, c.n_critic_hidden) > _lib.MAX_HIDDEN:
            raise ValueError(f"at most {_lib.MAX_HIDDEN} hidden layers")
        for i, v in enumerate(self.actor_hidden):
            c.actor_hidden[i] = int(v)
        for i, v in enumerate(self.critic_hidden):
            c.critic_hidden[i] = int(v)
        c.batch, c.weighted_loss, c.replicas, c.precision = self.batch, int(self.weighted_loss), self.replicas, self.precision
        c.gamma, c.tau, c.reward_scale = self.gamma, self.tau, self.reward_scale
        c.lr_actor, c.lr_critic = self.lr_actor, self.lr_critic
        c.lr_alpha = self.lr_actor if self.lr_alpha is None else self.lr_alpha
        c.action_scale = self.action_scale
        c.beta1, c.beta2, c.adam_eps = self.beta1, self.beta2, self.adam_eps
        c.log_alpha_init = self.log_alpha_init
        c.care = (2 if self.care_original else 1) if self.care else 0
        if self.care:
            c.emb_dim, c.lr_ctx = self.emb_dim, self.lr_ctx
            c.num_encoders, c.mix_out, c.ctx_in, c.ctx_out = self.num_encoders, self.mix_out, self.ctx_in, self.ctx_out
            c.n_mix_hidden, c.n_ctx_hidden = len(self.mix_hidden), len(self.ctx_hidden)
            for i, v in enumerate(self.mix_hidden):
                c.mix_hidden[i] = int(v)
            for i, v in enumerate(self.ctx_hidden):
                c.ctx_hidden[i] = int(v)
            c.tau_se = self.tau_se
        return c


def layout(cfg: CoreConfig):
    """Parameter layout table from the C library (pure host call, works without a GPU).
    Returns ({name: (offset, rows, cols, trainable, opt)}, arena_floats, trainable_floats)."""
    lib = _lib.load()
    c = cfg.to_c()
    n, arena, train = C.c_int32(0), C.c_int64(0), C.c_int64(0)
    _lib.check(lib.b200sac_layout(C.byref(c), None, 0, C.byref(n), C.byref(arena), C.byref(train)))
    descs = (_lib.TensorDesc * n.value)()
    _lib.check(lib.b200sac_layout(C.byref(c), descs, n.value, C.byref(n), C.byref(arena), C.byref(train)))
    table = {d.name.decode(): (d.offset, d.rows, d.cols, d.trainable, d.opt, d.pitch) for d in descs}
    return table, arena.value, train.value


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """torch's current stream as a cudaStream_t (the raw-stream query skips building a torch.cuda.Stream object: ~2 us per call
    saved on the per-update path)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SacCore:
    def __init__(self, cfg: CoreConfig, device: int = 0, seed: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("SacCore needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self.table, self.arena_floats, self.trainable_floats = layout(cfg)
        self._h = C.c_void_p(0)
        c = cfg.to_c()
        torch.cuda.set_device(self.device)
        _lib.check(self.lib.b200sac_create(C.byref(c), device, C.c_uint64(seed), C.byref(self._h)))
        n = C.c_int32(0)
        _lib.check(self.lib.b200sac_launches_per_step(self._h, C.byref(n)))
        self.launches_per_step = n.value
        self.steps_done = 0
        self._loss_buf = (C.c_float * (4 * max(1, cfg.replicas)))()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b200sac_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ------------------------------------------------------------------------
    def _arena_n(self, which):
        return self.arena_floats if which == _lib.PARAMS else self.trainable_floats

    def export_arena(self, which=_lib.PARAMS, replica=0) -> torch.Tensor:
        out = torch.empty(self._arena_n(which), dtype=torch.float32)
        _lib.check(self.lib.b200sac_export(self._h, which, replica, _ptr(out), out.numel(), _stream()))
        return out

    def import_arena(self, flat: torch.Tensor, which=_lib.PARAMS, replica=0):
        flat = flat.detach().to(torch.float32).contiguous()
        assert flat.numel() == self._arena_n(which)
        _lib.check(self.lib.b200sac_import(self._h, which, replica, _ptr(flat), flat.numel(), _stream()))

    def arena_view(self, which=_lib.PARAMS):
        """(device pointer, floats per replica) of an arena (replicas contiguous)."""
        p, n = C.c_void_p(0), C.c_int64(0)
        _lib.check(self.lib.b200sac_arena_ptr(self._h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def _unpack(self, name, flat, off):
        """One layout-table entry read out of `flat` (float32 numpy array) at float offset `off`, as a fresh CPU tensor in
        the reference's shape.  numpy copies on purpose: a 256x256 torch .clone() fans out to the intra-op thread pool,
        which costs milliseconds on a busy host -- this runs after every update (publication path)."""
        return torch.from_numpy(np.array(self._view(name, flat, off), dtype=np.float32, order="C", copy=True))

    def _view(self, name, flat, off):
        """numpy view (no copy) of one layout-table entry inside `flat`, in the reference's shape."""
        _, rows, cols, _t, _o, pitch = self.table[name]
        a = flat[off:off + rows * pitch].reshape(rows, pitch)[:, :cols]
        if ".mix." in name:            # stored as [K][out][in] | [K][out]; the reference holds (K,in,out) | (K,1,out)
            K = self.cfg.num_encoders
            a = (a.reshape(K, rows // K, cols).transpose(0, 2, 1) if name.endswith(".W") else a.reshape(K, 1, rows // K))
        elif not (name.endswith(".weight") or name == "embedding"):
            a = a.reshape(-1)
        return a

    def act(self, obs, stochastic=True, eps=None, replica=0, no_tanh=False) -> torch.Tensor:
        """Batched Actor.get_action for n <= 2*batch (CARE: batch) observation rows (CPU or CUDA tensor [n][obs_dim]); returns a
        CPU tensor [n][act_dim].  eps ([n][act_dim]) injects the noise; stochastic=False gives k*tanh(mu), or k*mu with
        no_tanh=True (the LunarLander actor's deterministic action)."""
        obs = obs.to(torch.float32).contiguous()
        n = obs.shape[0]
        e = None if eps is None else eps.to(torch.float32).contiguous()
        out = torch.empty(n, self.cfg.act_dim)
        _lib.check(self.lib.b200sac_act(self._h, replica, n, _ptr(obs), _ptr(e), 1 if stochastic else (-1 if no_tanh else 0), _ptr(out), _stream()))
        return out

    # ---- publication path (Learner.get_parameters, LL/learner.py:272-276) ---------------------------
    def publish_begin(self, tensor_names, replica=0):
        """Enqueue a consistent snapshot of the named parameter tensors and start its async copy to pinned host
        memory (b200sac_publish_begin); steps enqueued afterwards overlap it.  Collect with publish_wait()."""
        ents = sorted((self.table[n][0], self.table[n][1] * self.table[n][5], n) for n in tensor_names)
        ranges, where = [], {}                   # merge neighbouring tensors (16-B alignment gaps of <= 3 floats are copied along)
        for off, cnt, n in ents:
            if ranges and 0 <= off - (ranges[-1][0] + ranges[-1][1]) <= 3:
                ranges[-1][1] = off + cnt - ranges[-1][0]
            else:
                ranges.append([off, cnt])
            where[n] = sum(r[1] for r in ranges[:-1]) + (off - ranges[-1][0])
        offs = (C.c_int64 * len(ranges))(*[r[0] for r in ranges])
        cnts = (C.c_int64 * len(ranges))(*[r[1] for r in ranges])
        _lib.check(self.lib.b200sac_publish_begin(self._h, replica, len(ranges), offs, cnts, _stream()))
        self._pub_where = where

    def read_named(self, name, which=_lib.PARAMS, replica=0) -> torch.Tensor:
        """One small tensor (e.g. log_alpha) straight from the device: a few floats cross PCIe, not the arena."""
        off, rows, cols, _t, _o, pitch = self.table[name]
        out = torch.empty(rows * pitch)
        _lib.check(self.lib.b200sac_read_range(self._h, which, replica, off, rows * pitch, _ptr(out), _stream()))
        return self._unpack(name, out.numpy(), 0)

    def publish_views(self) -> Dict[str, np.ndarray]:
        """Like publish_wait(), but returns numpy VIEWS of the pinned snapshot in the reference's shapes (no copy); valid
        until the next publish_begin()."""
        ptr, n = C.POINTER(C.c_float)(), C.c_int64()
        _lib.check(self.lib.b200sac_publish_wait(self._h, C.byref(ptr), C.byref(n)))
        flat = np.ctypeslib.as_array(ptr, shape=(n.value,))
        return {name: self._view(name, flat, at) for name, at in self._pub_where.items()}

    def publish_wait(self) -> Dict[str, torch.Tensor]:
        ptr, n = C.POINTER(C.c_float)(), C.c_int64()
        _lib.check(self.lib.b200sac_publish_wait(self._h, C.byref(ptr), C.byref(n)))
        flat = np.ctypeslib.as_array(ptr, shape=(n.value,))       # view of the pinned buffer; _unpack copies out of it
        return {name: self._unpack(name, flat, at) for name, at in self._pub_where.items()}

    # ---- blob publication (b200sac_blob_*): the published byte string is assembled on the device -------------------
    def tensor_shape(self, name):
        """Shape of tensor `name` as the reference's state_dict holds it."""
        off, rows, _cols, _t, _o, pitch = self.table[name]
        return tuple(self._view(name, np.broadcast_to(np.float32(0), (off + rows * pitch,)), off).shape)

    def blob_index_map(self, name) -> np.ndarray:
        """Arena index of every element of tensor `name`, in the reference's shape and C order (pitch padding and the
        CARE mixture transposition are resolved here, once)."""
        off, rows, _cols, _t, _o, pitch = self.table[name]
        idx = np.arange(off + rows * pitch, dtype=np.int64)
        return np.ascontiguousarray(self._view(name, idx, off)).reshape(-1)

    def blob_template(self, image: bytes, slots, extra=(), replica=0):
        """Register the byte image.  slots: [(tensor name, byte offset of its payload inside `image`)]; extra: tensor names
        whose floats are appended BEHIND the image bytes (4-byte aligned; e.g. log_alpha for the logger) and returned by
        blob_wait() as numpy arrays."""
        src, dst = [], []
        for name, at in slots:
            m = self.blob_index_map(name)
            src.append(m)
            dst.append(at + 4 * np.arange(m.size, dtype=np.int64))
        total = (len(image) + 3) & ~3
        self._blob_extra = []
        for name in extra:
            m = self.blob_index_map(name)
            src.append(m)
            dst.append(total + 4 * np.arange(m.size, dtype=np.int64))
            self._blob_extra.append((name, total, m.size))
            total += 4 * m.size
        src = np.ascontiguousarray(np.concatenate(src), dtype=np.int32)
        dst = np.ascontiguousarray(np.concatenate(dst), dtype=np.int32)
        img = np.zeros(total, dtype=np.uint8)
        img[:len(image)] = np.frombuffer(image, dtype=np.uint8)
        _lib.check(self.lib.b200sac_blob_template(self._h, replica, C.c_void_p(img.ctypes.data), total, src.size,
                                                  C.c_void_p(src.ctypes.data), C.c_void_p(dst.ctypes.data)))
        self._blob_len = len(image)

    def blob_begin(self):
        _lib.check(self.lib.b200sac_blob_begin(self._h, _stream()))

    def blob_wait(self):
        """(published bytes, {extra tensor name: float32 array}) of the oldest blob begun and not yet collected."""
        ptr, n = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.b200sac_blob_wait(self._h, C.byref(ptr), C.byref(n)))
        blob = C.string_at(ptr.value, self._blob_len)
        extra = {name: np.ctypeslib.as_array(C.cast(ptr.value + at, C.POINTER(C.c_float)), shape=(cnt,)).copy()
                 for name, at, cnt in self._blob_extra}
        return blob, extra

    def get_named(self, which=_lib.PARAMS, replica=0) -> Dict[str, torch.Tensor]:
        flat = self.export_arena(which, replica).numpy()
        return {name: self._unpack(name, flat, d[0]) for name, d in self.table.items() if which == _lib.PARAMS or d[3]}

    def set_named(self, tensors: Dict[str, torch.Tensor], which=_lib.PARAMS, replica=0, strict=True):
        flat = self.export_arena(which, replica)
        seen = set()
        for name, t in tensors.items():
            if name not in self.table:
                if strict:
                    raise KeyError(name)
                continue
            off, rows, cols, trainable, _, pitch = self.table[name]
            if which != _lib.PARAMS and not trainable:
                continue
            t = torch.as_tensor(t, dtype=torch.float32)
            if ".mix." in name and name.endswith(".W"):
                t = t.permute(0, 2, 1).contiguous()          # (K,in,out) -> [K][out][in]
            t = t.reshape(-1)
            if t.numel() != rows * cols:
                raise ValueError(f"{name}: expected {rows * cols} elements, got {t.numel()}")
            flat[off:off + rows * pitch].reshape(rows, pitch)[:, :cols] = t.reshape(rows, cols)
            seen.add(name)
        if strict:
            need = {n for n, d in self.table.items() if which == _lib.PARAMS or d[3]}
            if need - seen:
                raise KeyError(f"missing tensors: {sorted(need - seen)[:5]}")
        self.import_arena(flat, which, replica)

    def get_steps(self, replica=0):
        s = (C.c_int64 * 4)()
        _lib.check(self.lib.b200sac_get_steps(self._h, replica, s))
        n = 4 if (self.cfg.care and self.cfg.care_original) else 3
        return tuple(int(x) for x in s)[:n]     # (critic, actor, alpha[, context encoder])

    def set_steps(self, steps, replica=0):
        cur = list(self.get_steps(replica)) + [0]
        vals = [int(x) for x in steps] + cur[len(steps):4]
        s = (C.c_int64 * 4)(*vals[:4])
        _lib.check(self.lib.b200sac_set_steps(self._h, replica, s))

    def soft_update(self, tau: float):
        _lib.check(self.lib.b200sac_soft_update(self._h, float(tau), _stream()))

    # ---- stepping -----------------------------------------------------------------------
    def _shape_check(self, s, a, r, s2, d, eps_next, eps_cur):
        c = self.cfg
        R, B = c.replicas, c.batch
        want = dict(s=c.obs_dim, a=c.act_dim, r=1, s2=c.obs_dim, d=1)
        for k, t in dict(s=s, a=a, r=r, s2=s2, d=d).items():
            if t.numel() != R * B * want[k]:
                raise ValueError(f"{k}: expected {R}x{B}x{want[k]} elements, got {tuple(t.shape)}")
        for k, t in dict(eps_next=eps_next, eps_cur=eps_cur).items():
            if t is not None and t.numel() != R * B * c.act_dim:
                raise ValueError(f"{k}: expected {R}x{B}x{c.act_dim} elements, got {tuple(t.shape)}")

    def step(self, s, a, r, s2, d, eps_next=None, eps_cur=None):
        """One gradient step from DEVICE tensors [R][B][w] (asynchronous)."""
        ts = [None if t is None else t.to(self.device, torch.float32).contiguous() for t in (s, a, r, s2, d, eps_next, eps_cur)]
        self._shape_check(*ts)
        self._keep = ts          # keep inputs alive until the stream has consumed them
        _lib.check(self.lib.b200sac_step(self._h, *[_ptr(t) for t in ts], _stream()))
        self.steps_done += 1

    def step_host(self, s, a, r, s2, d, eps_next=None, eps_cur=None, want_losses=True):
        """One gradient step from HOST tensors through the pinned staging path; returns
        losses [R][4] = (critic, actor, alpha, entropy) if want_losses (synchronises)."""
        ts = [None if t is None else t.detach().to("cpu", torch.float32).contiguous() for t in (s, a, r, s2, d, eps_next, eps_cur)]
        self._shape_check(*ts)
        out = torch.empty(self.cfg.replicas, 4) if want_losses else None
        _lib.check(self.lib.b200sac_step_host(self._h, *[_ptr(t) for t in ts], _ptr(out), _stream()))
        self.steps_done += 1
        return out

    def step_sampled(self, replay: "Replay", n_steps: int = 1):
        _lib.check(self.lib.b200sac_step_sampled(self._h, replay._h, int(n_steps), _stream()))
        self.steps_done += int(n_steps)

    def prepare(self, replay: "Replay"):
        """Instantiate every CUDA graph the sampled path can launch for this ring (no step is run)."""
        _lib.check(self.lib.b200sac_prepare(self._h, replay._h, _stream()))

    def update_sampled(self, replay: "Replay"):
        """sample + one gradient step + that step's losses of replica 0 as python floats (critic, actor, alpha, entropy):
        one library call, no tensor round trips -- the body of Learner.update()."""
        buf = self._loss_buf
        _lib.check(self.lib.b200sac_update(self._h, replay._h, buf, _stream()))
        self.steps_done += 1
        return buf[0], buf[1], buf[2], buf[3]

    def read_losses(self, n_last: int = 1) -> torch.Tensor:
        out = torch.empty(n_last, self.cfg.replicas, 4)
        _lib.check(self.lib.b200sac_read_losses(self._h, n_last, _ptr(out), _stream()))
        return out

    def profile_step(self, replay: "Replay", iters: int = 20):
        """[(kernel label, mean ms)] per launch of one step (eager run with CUDA events)."""
        cap = 256
        out = (C.c_float * cap)()
        n = C.c_int32(0)
        names = C.create_string_buffer(8192)
        _lib.check(self.lib.b200sac_profile_step(self._h, replay._h, int(iters), out, cap, C.byref(n), names, 8192, _stream()))
        self.steps_done += int(iters)
        labels = names.value.decode().split(";")
        return [(labels[i], float(out[i])) for i in range(n.value)]

    def graph_timeline(self, replay: "Replay", iters: int = 200):
        """Mean in-graph start-to-start time (us) of every launch of one step: [sample, ingest, plan...]."""
        cap = 256
        out = (C.c_float * cap)()
        n = C.c_int32(0)
        _lib.check(self.lib.b200sac_graph_timeline(self._h, replay._h, int(iters), out, cap, C.byref(n), _stream()))
        self.steps_done += int(iters) + 6
        return [float(out[i]) for i in range(n.value)]

    def debug(self, name: str, replica=0) -> torch.Tensor:
        """Per-step intermediates (b200sac_debug_read): y, q1, ..., and the hidden activations "hA.<l>", "hQ.<l>", "hP.<l>",
        "hT.<l>", "mixH.<inst>.<l>" whose sign patterns are the ReLU masks the step used."""
        cap = self.cfg.batch * max(2 * self.cfg.act_dim, 1)
        if name == "psave":
            cap = 2 * self.cfg.batch * self.cfg.act_dim * 8
        elif name[0] == "h" or name.startswith("mixH"):
            widest = max(list(self.cfg.actor_hidden) + list(self.cfg.critic_hidden) + [4 * ((w + 3) // 4) * self.cfg.num_encoders
                                                                                        for w in self.cfg.mix_hidden])
            cap = 2 * self.cfg.batch * widest
        out = torch.empty(cap)
        n = C.c_int64(0)
        _lib.check(self.lib.b200sac_debug_read(self._h, name.encode(), replica, _ptr(out), cap, C.byref(n), _stream()))
        return out[:n.value].clone()


class Replay:
    """Replay ring in HBM (where='device') or pinned host DRAM (where='host')."""

    def __init__(self, core: SacCore, capacity: int, where: str = "device", seed: int = 0):
        self.core = core
        self.lib = core.lib
        self._h = C.c_void_p(0)
        w = {"device": 0, "host": 1}[where]
        _lib.check(self.lib.b200sac_replay_create(core._h, int(capacity), w, C.c_uint64(seed), C.byref(self._h)))
        self.where = where

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value and self.core._h.value:
            self.lib.b200sac_replay_destroy(self._h)
        self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, s, a, r, s2, d, replica=0):
        ts = [np.ascontiguousarray(np.asarray(t, dtype=np.float32)) for t in (s, a, r, s2, d)]
        n = ts[2].size
        _lib.check(self.lib.b200sac_replay_push(self._h, replica, n, *[C.c_void_p(t.ctypes.data) for t in ts]))

    def fill_synthetic(self, n: int, seed: int = 1234):
        _lib.check(self.lib.b200sac_replay_fill_synthetic(self._h, int(n), C.c_uint64(seed), _stream()))

    def size(self, replica=0) -> int:
        n = C.c_int64(0)
        _lib.check(self.lib.b200sac_replay_size(self._h, replica, C.byref(n)))
        return n.value

    def sample(self, replica=0, with_indices=False):
        c = self.core.cfg
        B = c.batch
        s, a, r = torch.empty(B, c.obs_dim), torch.empty(B, c.act_dim), torch.empty(B, 1)
        s2, d = torch.empty(B, c.obs_dim), torch.empty(B, 1)
        idx = torch.empty(B, dtype=torch.int64)
        _lib.check(self.lib.b200sac_replay_sample(self._h, replica, _ptr(s), _ptr(a), _ptr(r), _ptr(s2), _ptr(d), _ptr(idx)))
        return (s, a, r, s2, d, idx) if with_indices else (s, a, r, s2, d)
