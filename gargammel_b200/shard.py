"""Multi-GPU sharding of the fragment id space: rank r of P takes the r-th contiguous slice, so the
rank outputs concatenated in rank order reproduce the single-GPU byte stream; no collective on the data path, only a
final reduction of the per-rank counts."""


def rank_slice(total, rank, world):
    """[lo, hi) of `total` ids for `rank` (same arithmetic as bin/gargammel-b200's per-GPU workers)."""
    return total * rank // world, total * (rank + 1) // world


def reduce_counts(values, device=None):
    """Sum per-rank counters across the process group (torch.distributed, nccl or gloo); identity without a group."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(round(x)) for x in t.tolist()]


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
