"""Drive the UNMODIFIED reference learner on CPU (TEST INFRASTRUCTURE ONLY).

Works only where /root/reference exists (the build container).  Nothing under
tests -m gpu, smoke() or bench.py may import this file: the GPU box has no
/root/reference.  It is used by oracle/gen_golden.py to produce the committed
fixtures under tests/golden/, by tests/test_oracle_vs_reference.py (skipped when the
reference is absent) to pin oracle/sac_port.py to the real code, and by bench.py's CPU
arm (`cpu_baseline.kind == "reference"`) when a reference checkout is reachable
(B200SAC_REFERENCE, /root/reference or baseline/_ref) -- never by the product path.

How the reference is made importable:
  * `redis` -> oracle/redis_stub.py, installed in sys.modules first;
  * cwd -> a temp dir, because Learner.__init__ creates ./log and saved_models
    (LunarLander_Distributed_SAC/src/learner.py:83-98);
  * cfg copied with "device": "cpu";
  * eps injection: torch.distributions.normal._standard_normal is what
    Normal.rsample() calls (LL/model.py:55); we replace it with a FIFO.
"""
import contextlib
import importlib
import json
import os
import sys
import tempfile

import torch

def _find_reference():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("B200SAC_REFERENCE"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "LunarLander_Distributed_SAC", "src")):
            return cand
    return os.environ.get("B200SAC_REFERENCE", "/root/reference")


REF_ROOT = _find_reference()

VARIANTS = {
    # name: (src dir, cfg file)
    "LL": ("LunarLander_Distributed_SAC/src", "cfg/LunarLanderContinuous-v2_Distributed_SAC_cfg.json"),
    "VS": ("MT1_Distributed_VSAC/src", "cfg/MT1_Distributed_VSAC_cfg.json"),
    "MS": ("MT10_Distributed_MTSAC/src", "cfg/MT10_Distributed_MTSAC_cfg.json"),
    "C1": ("MT1_Distributed_CARE/src", "cfg/MT1_Distributed_CARE_cfg.json"),
    "C10": ("MT10_Distributed_CARE/src", "cfg/MT10_Distributed_CARE_cfg.json"),
}

_REF_MODULE_NAMES = ("learner", "model", "utils", "replay_buffer", "replay_buffers", "logger",
                     "context_encoder", "state_encoder", "player")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, VARIANTS["LL"][0]))


def _purge():
    for name in _REF_MODULE_NAMES:
        sys.modules.pop(name, None)
    for v in VARIANTS.values():
        p = os.path.join(REF_ROOT, v[0])
        while p in sys.path:
            sys.path.remove(p)


def import_variant(variant):
    """Import the reference's `learner` module for one variant (fresh)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import redis_stub
    redis_stub.install()
    _purge()
    sys.path.insert(0, os.path.join(REF_ROOT, VARIANTS[variant][0]))
    return importlib.import_module("learner")


def make_learner(variant, cfg_overrides=None, seed=0, workdir=None):
    """Construct the reference Learner on CPU.  Returns (learner, module)."""
    mod = import_variant(variant)
    workdir = workdir or tempfile.mkdtemp(prefix="b200sac_ref_")
    with open(os.path.join(REF_ROOT, VARIANTS[variant][1])) as f:
        cfg = json.load(f)
    cfg["device"] = "cpu"
    if "encoder" in cfg:
        cfg["encoder"]["device"] = "cpu"
    for k, v in (cfg_overrides or {}).items():
        if isinstance(v, dict) and isinstance(cfg.get(k), dict):
            cfg[k].update(v)
        else:
            cfg[k] = v
    cfg_path = os.path.join(workdir, "cfg.json")
    with open(cfg_path, "w") as f:
        json.dump(cfg, f)
    # CARE reads cfg/metadata/*.json relative to cwd (C10/context_encoder.py:31-36)
    if not os.path.exists(os.path.join(workdir, "cfg")):
        os.symlink(os.path.join(REF_ROOT, "cfg"), os.path.join(workdir, "cfg"))
    old = os.getcwd()
    os.chdir(workdir)
    try:
        torch.manual_seed(seed)
        if variant == "LL":
            lrn = mod.Learner(cfg_path, write_mode=True)
        elif variant == "VS":
            lrn = mod.Learner(cfg_path, write_mode=True, save_period=10 ** 9, checkpoint_path=None)
        else:
            with open(os.path.join(REF_ROOT, "cfg/metadata/mt10_ordered_task_name.json")) as f:
                names = json.load(f)
            names = names if isinstance(names, list) else list(names)
            if variant == "C1":
                names = names[:1]
            lrn = mod.Learner(None, names, cfg_path, write_mode=True, save_period=10 ** 9, checkpoint_path=None)
    finally:
        os.chdir(old)
    return lrn, mod


@contextlib.contextmanager
def injected_eps(eps_list):
    """Make Normal.rsample() consume the given (B, act) tensors in order.

    The reference draws twice per step: next-state actor, then current-state
    actor (LL/learner.py:207,221)."""
    import torch.distributions.normal as tdn
    queue = [torch.as_tensor(e, dtype=torch.float32) for e in eps_list]
    orig = tdn._standard_normal

    def fake(shape, dtype, device):
        e = queue.pop(0)
        assert tuple(e.shape) == tuple(shape), (e.shape, shape)
        return e.to(dtype=dtype, device=device)

    tdn._standard_normal = fake
    try:
        yield queue
    finally:
        tdn._standard_normal = orig


def fill_memory(lrn, variant, batch):
    """Append transitions (tensors s, a, r, s2, d; MT: s carries the one-hot task id in its last columns) to the reference
    ReplayBuffer's own deque(s), as its Redis-drain thread would (LL/replay_buffer.py:52-59, MS/replay_buffers.py:57-63)."""
    import numpy as np
    s, a, r, s2, d = [t.numpy() for t in batch]
    mem = lrn.memory
    E = mem.experience
    if hasattr(mem, "memories"):
        T = len(mem.memories)
        task = np.argmax(s[:, -T:], axis=1)
        for i in range(s.shape[0]):
            mem.memories[int(task[i])].append(E(s[i].astype(np.float64), a[i], float(r[i, 0]), s2[i].astype(np.float64), bool(d[i, 0])))
    else:
        for i in range(s.shape[0]):
            mem.memory.append(E(s[i].astype(np.float64), a[i], float(r[i, 0]), s2[i].astype(np.float64), bool(d[i, 0])))
