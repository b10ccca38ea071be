"""Region sharding for one-process-per-GPU runs (bench.py under torchrun, and any multi-process caller).

The path shards by active region with no data-path collective: every rank owns a
disjoint slice of regions and computes it alone; only the timing/accounting scalars are reduced.
"""
import os

import numpy as np
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


_FIELDS = ("read_bases", "base_q", "ins_q", "del_q", "gcp", "read_off", "hap_bases", "hap_off", "units")


def _scratch_dir(need_bytes, preferred="/dev/shm"):
    import shutil
    for d in (preferred, "/tmp"):
        try:
            if os.path.isdir(d) and shutil.disk_usage(d).free > need_bytes + (64 << 20):
                return d
        except OSError:
            pass
    return "/tmp"


def collect_shards(batch, rank, world, tag="gphmm", scratch="/dev/shm", pinned=False, group=None, extra=None):
    """One node: every rank's shard on rank 0, back to back in rank order, as ONE batch (None on the other ranks).
    No collective carries data: the shards travel as files under `scratch`, the two barriers are
    the only communication.  This is how bench.py builds the batch it gives to one handle over all GPUs of the box.
    `group`: process group of the barriers (a gloo group keeps waiting ranks off their GPUs); `extra`: one more array per
    rank (e.g. its results), returned as a list in rank order next to the batch."""
    from .native import Batch
    if world == 1:
        return batch if extra is None else (batch, [extra])
    ready = dist.is_available() and dist.is_initialized()
    if not ready:
        raise RuntimeError("collect_shards needs an initialised process group")
    need = world * (batch.input_bytes() + (0 if extra is None else extra.nbytes) + 16 * len(batch.read_off))
    path = lambda r: os.path.join(_scratch_dir(need, scratch), "%s_%s_shard%d.npz" % (tag, os.environ.get("MASTER_PORT", "0"), r))
    if rank != 0:
        arrays = {k: getattr(batch, k) for k in _FIELDS}
        if extra is not None:
            arrays["extra"] = extra
        np.savez(path(rank), **arrays)
    dist.barrier(group=group)
    whole, extras = None, None
    if rank == 0:
        parts, extras = [batch], [extra]
        for r in range(1, world):
            with np.load(path(r)) as z:
                parts.append(Batch(*[z[k] for k in _FIELDS]))
                extras.append(z["extra"] if "extra" in z.files else None)
        whole = Batch.concat(parts, pinned=pinned)
    dist.barrier(group=group)
    if rank != 0:
        try:
            os.remove(path(rank))
        except OSError:
            pass
    return whole if extra is None else (whole, extras)
