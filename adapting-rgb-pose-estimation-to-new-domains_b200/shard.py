"""Multi-GPU layout of the path: images / frames are independent, so rank r of N takes the
items with index % N == r and no data-path collective exists.  The only
exchanges are the optional gather of decoded persons (a few KB per frame) and the max-over-ranks
of a timing -- both through torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_indices(n_items, rank, world):
    return np.arange(rank, n_items, world, dtype=np.int64)


def gather_results(local_results, n_items):
    """local_results: list of dicts carrying their global 'index'.  Returns the full list in
    input order on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = sorted(local_results, key=lambda r: r["index"])
    else:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, local_results)
        out = sorted((r for p in parts for r in p), key=lambda r: r["index"])
    assert [r["index"] for r in out] == list(range(n_items))
    return out


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
