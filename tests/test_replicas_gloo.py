"""CPU, world_size 2 over gloo: the N>1 host logic (replica placement + the one broadcast)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from distributed_sac_b200.replicas import broadcast_flat, shard_replicas
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)          # each rank starts with different weights
        flat = torch.randn(344600, generator=g)
        before = flat.clone()
        broadcast_flat(flat, src=0)
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        mine = shard_replicas(10, world)[rank]
        # weak scaling bookkeeping like bench.py: whole-job steps = sum over ranks, time = max over ranks
        t = torch.tensor([1.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, same, torch.equal(flat, before), mine, float(t)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, same0, unchanged0, mine0, tmax0), (r1, same1, unchanged1, mine1, tmax1) = res
    assert same0 and same1                       # identical arenas everywhere after the broadcast
    assert unchanged0 and not unchanged1         # rank 0 was the source
    assert mine0 == [0, 2, 4, 6, 8] and mine1 == [1, 3, 5, 7, 9]
    assert tmax0 == tmax1 == 2.0


def test_shard_replicas_config4_placement():
    from distributed_sac_b200.replicas import shard_replicas
    assert [len(x) for x in shard_replicas(10, 8)] == [2, 2, 1, 1, 1, 1, 1, 1]
    assert sorted(sum(shard_replicas(10, 8), [])) == list(range(10))
    with pytest.raises(ValueError):
        shard_replicas(0, 8)
