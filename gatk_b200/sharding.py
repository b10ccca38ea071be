"""Region sharding for one-process-per-GPU runs (bench.py under torchrun, and any multi-process caller).

The path shards by active region with no data-path collective: every rank owns a
disjoint slice of regions and computes it alone; only the timing/accounting scalars are reduced.
"""
import torch
import torch.distributed as dist


def region_slice(rank, world, regions_per_rank):
    """Weak scaling: rank r owns regions [r*n, (r+1)*n).  Returns (first_region, n_regions)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * regions_per_rank, regions_per_rank


def split_units(n_units, rank, world):
    """Strong scaling over an existing unit list: contiguous, balanced to within one unit."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_timing(seconds, totals, device="cpu"):
    """max over ranks of every entry of `seconds`, sum over ranks of every entry of `totals`."""
    t = torch.tensor(list(seconds), dtype=torch.float64, device=device)
    s = torch.tensor(list(totals), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [float(x) for x in s.tolist()]
