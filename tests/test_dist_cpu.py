"""CPU (gloo, world_size 2 and 3) tests of the multi-GPU host plumbing: shard ranges and the padded range all-gather."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ranges(n, block, world):
    from onbody_b200.api import load_library
    L = load_library()
    L.onb_shard_range_for.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    out = []
    for r in range(world):
        lo, hi = C.c_uint64(), C.c_uint64()
        assert L.onb_shard_range_for(n, block, r, world, C.byref(lo), C.byref(hi)) == 0
        out.append((int(lo.value), int(hi.value)))
    return out


@pytest.mark.parametrize("n,block,world", [(100000, 128, 2), (100000, 128, 8), (129, 128, 4), (100, 128, 3), (10 ** 7, 128, 8), (999, 50, 7)])
def test_shard_ranges_partition_the_leaves(n, block, world):
    rs = _ranges(n, block, world)
    assert rs[0][0] == 0 and rs[-1][1] == n
    for (a, b), (c, d) in zip(rs, rs[1:]):
        assert b == c and a <= b
    for lo, hi in rs:
        assert lo % block == 0 and (hi % block == 0 or hi == n)
    sizes = [hi - lo for lo, hi in rs]
    assert max(sizes) - min(sizes) <= block                      # balanced to one leaf


def _worker(rank, world, port, n, block, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from onbody_b200.multigpu import allgather_ranges
        ranges = _ranges(n, block, world)
        truth = torch.arange(n + 256, dtype=torch.float32) * 0.5
        plane = torch.full((n + 256,), -1.0)
        lo, hi = ranges[rank]
        plane[lo:hi] = truth[lo:hi]                               # each rank owns only its range
        scratch = allgather_ranges(plane, ranges, rank, world)
        ok = bool(torch.equal(plane[:n], truth[:n]))
        # second plane reuses the scratch buffer
        plane2 = torch.zeros(n + 256); plane2[lo:hi] = 3.0 * truth[lo:hi]
        allgather_ranges(plane2, ranges, rank, world, scratch)
        ok = ok and bool(torch.equal(plane2[:n], 3.0 * truth[:n]))
        # the bench's reductions: max of the times, sum of the work
        red = torch.tensor([10.0 + rank, 100.0 * (rank + 1)], dtype=torch.float64)
        mx = red.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ok = ok and mx[0].item() == 10.0 + world - 1 and sm[1].item() == 100.0 * world * (world + 1) / 2
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5000), (3, 1000)])
def test_padded_range_allgather_gloo(world, n):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 128, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
