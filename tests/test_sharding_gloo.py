"""world_size-2 gloo test of the N>1 host logic (no GPU): disjoint region shards, max/sum reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gatk_b200 import sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = sharding.region_slice(rank, world, 3)
    b = synth.config2(n, first_region=first)
    secs, tot = sharding.reduce_timing([1.0 + rank, 5.0 - rank], [b.cells(), b.pairs()])
    lo, hi = sharding.split_units(7, rank, world)
    whole = sharding.collect_shards(b, rank, world, tag="gphmm_test")
    if rank == 0:
        # the collected batch is the unsharded one, field by field
        ref = synth.config2(n * world)
        for name in ("read_bases", "base_q", "ins_q", "del_q", "gcp", "read_off", "hap_bases", "hap_off", "units"):
            assert np.array_equal(getattr(whole, name), getattr(ref, name)), name
    else:
        assert whole is None
    q.put((rank, first, n, b.cells(), b.pairs(), secs, tot, lo, hi, bytes(b.read_bases[:16])))
    dist.destroy_process_group()


def test_two_rank_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, n0, c0, p0, s0, t0, lo0, hi0, head0), (r1, f1, n1, c1, p1, s1, t1, lo1, hi1, head1) = res
    assert (f0, n0, f1, n1) == (0, 3, 3, 3)          # disjoint weak-scaling slices
    assert head0 != head1                             # different regions => different reads
    assert s0 == s1 == [2.0, 5.0]                     # max over ranks
    assert t0 == t1 == [float(c0 + c1), float(p0 + p1)]  # sum over ranks
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 7)
    # the two shards together are exactly the unsharded batch
    whole = synth.config2(6)
    assert whole.cells() == c0 + c1 and whole.pairs() == p0 + p1


def test_split_units_covers_everything():
    for n in (0, 1, 5, 8, 13):
        for world in (1, 2, 4, 8):
            spans = [sharding.split_units(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
