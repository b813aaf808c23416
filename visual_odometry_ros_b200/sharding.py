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


def split_lba_problem(p: dict, world: int, rank: int) -> dict:
    """This rank's part of a local-BA window for the landmark-sharded solve (vo_lba_solve_dist): ALL frames / poses, the
    landmarks shard_range(n_points, world, rank) and their observations (CSR re-based).  Landmark order is preserved, so
    concatenating the ranks' point outputs in rank order restores the original landmark order."""
    import numpy as np
    lo, hi = shard_range(int(p["n_points"]), world, rank)
    ptr = np.asarray(p["obs_ptr"], np.int64)
    o0, o1 = int(ptr[lo]), int(ptr[hi])
    q = dict(p)
    q["n_points"] = hi - lo
    q["n_obs"] = o1 - o0
    q["points"] = np.ascontiguousarray(np.asarray(p["points"], np.float64).reshape(-1, 3)[lo:hi])
    q["obs_ptr"] = (ptr[lo:hi + 1] - o0).astype(np.int32)
    q["obs_frame"] = np.ascontiguousarray(np.asarray(p["obs_frame"], np.int32)[o0:o1])
    q["obs_right"] = np.ascontiguousarray(np.asarray(p["obs_right"], np.uint8)[o0:o1])
    q["obs_px"] = np.ascontiguousarray(np.asarray(p["obs_px"], np.float64).reshape(-1, 2)[o0:o1])
    q["landmark_range"] = (lo, hi)
    return q
