"""CPU test of the N>1 path's host logic: world-size-2 gloo, frame pairs sharded round-robin, no data collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fldr_vfi_b200 import sharding
    items = [float(i) for i in range(n_items)]
    local = sharding.run_sharded(items, lambda v: v * v, rank, world)
    gathered = sharding.gather_host(local)
    tmax = sharding.max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((gathered, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_pairs_are_partitioned_exactly_once_world2():
    world, n_items = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert tmax == 2.0                                   # max over ranks of (1 + rank)
    merged = {}
    for part in gathered:
        assert not (set(part) & set(merged)), "a frame pair was processed by two ranks"
        merged.update(part)
    assert sorted(merged) == list(range(n_items))
    assert all(merged[i] == float(i * i) for i in merged)
    assert [len(p) for p in gathered] == [4, 3]          # round-robin keeps ranks within one item


def test_shard_helpers():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fldr_vfi_b200 import sharding
    assert sharding.shard_indices(64, 3, 8) == list(range(3, 64, 8))
    assert sharding.shard_counts(64, 8) == [8] * 8
    assert sharding.shard_counts(5, 4) == [2, 1, 1, 1]
    assert sharding.gather_host("x") == ["x"]            # no process group: single rank
    assert sharding.max_over_ranks(3.5) == 3.5
