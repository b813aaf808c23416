"""Partition of independent stereo sequences across ranks (one process per GPU, no collectives
on the data path).  Pure host logic, exercised on CPU with gloo."""


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` of `world` (first ranks get the remainder)."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Device-timed numbers are reported as the max over ranks (bench contract)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
