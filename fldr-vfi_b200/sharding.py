"""Frame-pair sharding across ranks.

Triplets / frame pairs are independent units (the reference's test batch is 1, utils.py:150; no state crosses
pairs), so the multi-GPU story is one process per GPU, pair i -> rank i mod world, weights replicated, and NO
collective on the data path.  ``torch.distributed`` is used only to gather small per-rank results (timings, PSNR)
on the host side and for barriers around timed regions.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Indices of the items rank ``rank`` owns: i with i % world == rank (round-robin keeps ranks within one item)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def shard_counts(n_items, world):
    return [len(range(r, n_items, world)) for r in range(world)]


def gather_host(obj):
    """Gather one small picklable object per rank on every rank (host side; no device collective involved)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def max_over_ranks(value, device=None):
    """Max of a python float over ranks - the timing reduction bench.py uses (device-timed, max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def run_sharded(items, fn, rank, world):
    """Apply ``fn`` to this rank's items; returns {global_index: result}.  Results stay rank-local."""
    return {i: fn(items[i]) for i in shard_indices(len(items), rank, world)}
