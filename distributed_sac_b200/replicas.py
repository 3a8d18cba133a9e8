"""Multi-GPU = independent learner replicas: no gradient all-reduce, one
collective in the whole job -- the broadcast of the initial parameter arena so that replicas can
start from identical weights.  Host-side plumbing over torch.distributed (NCCL on GPUs, gloo in the
CPU tests)."""
from typing import List

import torch


def shard_replicas(n_replicas: int, world_size: int) -> List[List[int]]:
    """Round-robin placement of replica ids on ranks: 10 replicas on 8 ranks -> 2,2,1,1,1,1,1,1
    (BASELINE.json config 4: 10 task-replica learners over 8 GPUs)."""
    if n_replicas < 1 or world_size < 1:
        raise ValueError("n_replicas and world_size must be >= 1")
    out = [[] for _ in range(world_size)]
    for r in range(n_replicas):
        out[r % world_size].append(r)
    return out


def broadcast_flat(flat: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """Broadcast one flat fp32 tensor from `src` (in place) and return it."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def broadcast_initial_params(core, src: int = 0, group=None, same_across_local_replicas: bool = False):
    """Make every rank's learner(s) start from rank `src`'s parameters.

    core: SacCore.  Exports each local replica's parameter arena to a device tensor, broadcasts it
    over NCCL (NVLink/NVSwitch: a few MB, one shot) and imports it back.  Adam state starts at zero
    everywhere and is not sent.  Returns the number of floats broadcast."""
    from . import _lib
    _ptr, n = core.arena_view(_lib.PARAMS)
    R = core.cfg.replicas
    flat = torch.empty(n * R, dtype=torch.float32, device=core.device)
    for r in range(R):
        _lib.check(core.lib.b200sac_export(core._h, _lib.PARAMS, r, flat[r * n:(r + 1) * n].data_ptr(), n, None))
    if same_across_local_replicas:
        flat.view(R, n)[1:] = flat.view(R, n)[0]
    broadcast_flat(flat, src, group)
    for r in range(R):
        _lib.check(core.lib.b200sac_import(core._h, _lib.PARAMS, r, flat[r * n:(r + 1) * n].data_ptr(), n, None))
    torch.cuda.synchronize(core.device)
    return int(flat.numel())
