"""CPU tests of the multi-GPU host logic (no GPU needed): the partition arithmetic of csrc/plan.cu against a brute-force
restatement, and - with gloo at world size 2 and 3 - the data movement the C++ communicator performs with NCCL (in-place
all-gather of equal leaf-aligned chunks; packed exchange of the per-level node intervals), driven by the plan the library
computes, on host tensors."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAXL = 40


def _lib():
    from onbody_b200.api import load_library
    L = load_library()
    L.onb_shard_range_for.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.onb_shard_chunk_for.restype = C.c_uint64
    L.onb_shard_chunk_for.argtypes = [C.c_uint64, C.c_int, C.c_int]
    u32p = C.POINTER(C.c_uint32)
    L.onb_plan_query.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, u32p, u32p, u32p, u32p, u32p, u32p]
    return L


def _ranges(n, block, world):
    L = _lib()
    out = []
    for r in range(world):
        lo, hi = C.c_uint64(), C.c_uint64()
        assert L.onb_shard_range_for(n, block, r, world, C.byref(lo), C.byref(hi)) == 0
        out.append((int(lo.value), int(hi.value)))
    return out


def _plan(n, block, world, rank):
    L = _lib()
    arr = lambda k: np.zeros(k, np.uint32)
    own_lo, own_hi, need_lo, need_hi, nsh, sh = arr(MAXL), arr(MAXL), arr(MAXL), arr(MAXL), arr(MAXL), arr(MAXL * world)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    lv = L.onb_plan_query(n, block, world, rank, MAXL, p(own_lo), p(own_hi), p(need_lo), p(need_hi), p(nsh), p(sh))
    assert lv > 0, lv
    shared = [list(sh[l * world: l * world + nsh[l]]) for l in range(lv)]
    return lv, own_lo[:lv], own_hi[:lv], need_lo[:lv], need_hi[:lv], shared


def _tree_shape(n, block):
    """the reference's VAM-split shape (Tree.hpp:83-87, barneshut.hpp:663): {node id: (first particle, count)}"""
    nleaf = 1 + (n - 1) // block
    levels = 1 + int(np.floor(np.log2(2 * nleaf - 1)))
    nodes = {1: (0, n)}
    stack = [1]
    while stack:
        i = stack.pop()
        io, num = nodes[i]
        if num > block:
            pm = io + block * (1 << int(np.floor(np.log2((num - 1) // block))))
            nodes[2 * i] = (io, pm - io); nodes[2 * i + 1] = (pm, io + num - pm)
            stack += [2 * i, 2 * i + 1]
    return levels, nodes


@pytest.mark.parametrize("n,block,world", [(100000, 128, 2), (100000, 128, 8), (129, 128, 4), (100, 128, 3), (10 ** 7, 128, 8), (999, 50, 7)])
def test_shard_ranges_partition_the_leaves(n, block, world):
    rs = _ranges(n, block, world)
    chunk = int(_lib().onb_shard_chunk_for(n, block, world))
    assert rs[0][0] == 0 and rs[-1][1] == n and chunk % block == 0
    for (a, b), (c, d) in zip(rs, rs[1:]):
        assert b == c and a <= b
    for r, (lo, hi) in enumerate(rs):
        assert lo == min(n, r * chunk) and hi == min(n, (r + 1) * chunk)      # equal chunks: one in-place all-gather per plane
        assert (lo % block == 0 or lo == n) and (hi % block == 0 or hi == n)
    assert chunk * world >= n and chunk * world < n + (world + 1) * block       # the slack every plane is allocated with


@pytest.mark.parametrize("n,block,world", [(70000, 128, 3), (100000, 128, 8), (1000, 128, 4), (129, 128, 2), (5000, 64, 5), (300000, 128, 2), (128 * 64, 128, 8)])
def test_plan_matches_brute_force(n, block, world):
    levels, nodes = _tree_shape(n, block)
    rs = _ranges(n, block, world)
    seen_shared = None
    for rank in range(world):
        lv, own_lo, own_hi, need_lo, need_hi, shared = _plan(n, block, world, rank)
        assert lv == levels
        lo, hi = rs[rank]
        for l in range(levels):
            ids = [i for i in nodes if (1 << l) <= i < (2 << l)]
            own = sorted(i for i in ids if hi > lo and nodes[i][0] >= lo and nodes[i][0] + nodes[i][1] <= hi)
            need = sorted(i for i in ids if nodes[i][0] < hi and nodes[i][0] + nodes[i][1] > lo)
            assert list(range(own_lo[l], own_hi[l])) == own, (rank, l)
            assert list(range(need_lo[l], need_hi[l])) == need, (rank, l)
            strad = sorted(i for i in ids if nodes[i][1] > block and not any(nodes[i][0] >= a and nodes[i][0] + nodes[i][1] <= b for a, b in rs))
            assert sorted(shared[l]) == strad, (rank, l)
        seen_shared = seen_shared or shared
        assert shared == seen_shared                                              # every rank recomputes the same straddling nodes
    # every non-leaf node is owned by exactly one rank or straddles
    for l in range(levels):
        ids = [i for i in nodes if (1 << l) <= i < (2 << l) and nodes[i][1] > block]
        cover = []
        for rank in range(world):
            _, own_lo, own_hi, _, _, _ = _plan(n, block, world, rank)
            cover += [i for i in range(own_lo[l], own_hi[l]) if nodes[i][1] > block]
        assert sorted(cover + seen_shared[l]) == sorted(ids)


def _worker(rank, world, port, n, block, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = _lib()
        chunk = int(L.onb_shard_chunk_for(n, block, world))
        lo, hi = _ranges(n, block, world)[rank]
        # (1) particle planes: what comm.cu does with one in-place ncclAllGather - rank r owns [r*chunk, (r+1)*chunk)
        cap = chunk * world
        truth = torch.arange(cap, dtype=torch.float32) * 0.5
        plane = torch.full((cap,), -1.0)
        plane[lo:hi] = truth[lo:hi]
        parts = [plane[r * chunk:(r + 1) * chunk] for r in range(world)]
        dist.all_gather(parts, plane[rank * chunk:(rank + 1) * chunk].clone())
        ok = bool(torch.equal(plane[:n], truth[:n]))
        # (2) equivalent strengths (dist.cu k_eqx_pack / k_eqx_unpack): every rank packs the blocks of the nodes it owns, level by
        #     level, into ITS chunk of a staging buffer, one in-place all-gather of equal chunks, every rank scatters the others'
        #     chunks back by walking the same per-level intervals; then the straddling nodes locally
        ebs = 4
        levels, nodes = _tree_shape(n, block)
        numnodes = 1 << levels
        want = torch.zeros(numnodes * ebs)
        for i, (io, num) in nodes.items():
            if num > block:
                want[i * ebs:(i + 1) * ebs] = float(i)
        eq = torch.zeros(numnodes * ebs)
        plans = [_plan(n, block, world, r) for r in range(world)]
        owned = [[i for l in range(levels - 1) for i in range(int(plans[r][1][l]), int(plans[r][2][l]))] for r in range(world)]
        _, own_lo, own_hi, _, _, shared = plans[rank]
        for i in owned[rank]:
            if nodes[i][1] > block:
                eq[i * ebs:(i + 1) * ebs] = float(i)
        chunk_blocks = max(len(o) for o in owned)
        stage = torch.full((world * chunk_blocks * ebs,), -7.0)
        for k, i in enumerate(owned[rank]):
            stage[(rank * chunk_blocks + k) * ebs:(rank * chunk_blocks + k + 1) * ebs] = eq[i * ebs:(i + 1) * ebs]
        if chunk_blocks:
            parts = [stage[r * chunk_blocks * ebs:(r + 1) * chunk_blocks * ebs] for r in range(world)]
            dist.all_gather(parts, stage[rank * chunk_blocks * ebs:(rank + 1) * chunk_blocks * ebs].clone())
        for r in range(world):
            if r != rank:
                for k, i in enumerate(owned[r]):
                    eq[i * ebs:(i + 1) * ebs] = stage[(r * chunk_blocks + k) * ebs:(r * chunk_blocks + k + 1) * ebs]
        for l in range(levels):
            for i in shared[l]:
                eq[i * ebs:(i + 1) * ebs] = float(i)
        ok = ok and bool(torch.equal(eq, want))
        # (3) the bench's reductions: max of the times, sum of the work
        red = torch.tensor([10.0 + rank, 100.0 * (rank + 1)], dtype=torch.float64)
        mx = red.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ok = ok and mx[0].item() == 10.0 + world - 1 and sm[1].item() == 100.0 * world * (world + 1) / 2
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5000), (3, 1000)])
def test_exchange_pattern_gloo(world, n):
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
