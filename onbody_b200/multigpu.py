"""Host-side plumbing of the multi-GPU source build: each rank builds its share of the source tree
(onb_make_tree_range), the ranks exchange their tree-order ranges of the particle planes with one NCCL all-gather per
plane over NVLink (torch.distributed on tensors that alias the library's device memory), then every rank completes the
node arrays locally (onb_finish_tree). Only plumbing lives here; the kernels are in csrc/."""
import torch
import torch.distributed as dist


import os
# Two ways to replicate unequal, leaf-aligned ranges (both bit-identical, tools/check_multi.py):
#   uneven : every rank's range straight into place (torch issues one grouped ncclBroadcast per rank, no staging buffer)
#   padded : one ncclAllGather of equal padded chunks + unpack copies
# Measured on B200s at N=1e7 (5 source planes): 2 GPUs uneven 1.34 ms / padded 1.59 ms; 8 GPUs uneven 2.8 ms / padded 2.2 ms.
# ONB_ALLGATHER=auto (default: uneven up to 2 ranks, padded above) | uneven | padded
_UNEVEN = {"ok": True, "mode": os.environ.get("ONB_ALLGATHER", "auto")}


def shard_ranges(session, n, world):
    return [session.shard_particle_range(n, r, world) for r in range(world)]


def allgather_ranges(plane, ranges, rank, world, scratch=None, group=None):
    """In-place all-gather of a 1-D tensor whose slice [lo_r, hi_r) is valid on rank r. Ranges are contiguous and leaf
    aligned but not equal, so the collective is padded to the longest (the tensor must extend >= chunk past every lo).
    Works on CUDA tensors over NCCL and on CPU tensors over gloo (tests)."""
    if world == 1:
        return scratch
    want_uneven = _UNEVEN["mode"] == "uneven" or (_UNEVEN["mode"] == "auto" and world <= 2)
    if dist.get_backend(group) != "gloo" and _UNEVEN["ok"] and want_uneven:
        # NCCL: every rank's range goes straight to its final place on every other rank (torch issues one grouped
        # ncclBroadcast per rank for unequal sizes) - no padding, no staging buffer, no unpack copies
        lo, hi = ranges[rank]
        try:
            dist.all_gather([plane[a:b] for a, b in ranges], plane[lo:hi], group=group)
            return scratch
        except Exception:          # older torch: fall back to the padded collective below, once and for all
            _UNEVEN["ok"] = False
    chunk = max(hi - lo for lo, hi in ranges)
    lo, hi = ranges[rank]
    if scratch is None or scratch.numel() < world * chunk or scratch.device != plane.device:
        scratch = torch.empty(world * chunk, dtype=plane.dtype, device=plane.device)
    out = scratch[: world * chunk]
    mine = plane[lo:lo + chunk]
    if mine.numel() < chunk:
        raise ValueError("plane is not padded enough for a padded all-gather")
    if dist.get_backend(group) == "gloo":
        dist.all_gather([out[r * chunk:(r + 1) * chunk] for r in range(world)], mine.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(out, mine, group=group)
    for r, (a, b) in enumerate(ranges):
        if r != rank:
            plane[a:b].copy_(out[r * chunk: r * chunk + (b - a)])
    return scratch


def exchange_planes(session, which, n, rank, world, ranges=None, group=None, scratch=None):
    """all-gather of every particle plane of set `which` (0 sources: x, r, s; 1 targets: x, r)"""
    if world == 1:
        return scratch
    ranges = ranges or shard_ranges(session, n, world)
    fields = session.source_fields() if which == 0 else list(range(session.PD)) + [3]
    for f in fields:
        plane = session.plane_tensor(which, f, n + 256)      # the library pads every plane by >= one leaf
        scratch = allgather_ranges(plane, ranges, rank, world, scratch, group)
    torch.cuda.synchronize()
    return scratch


def _tick(times, key, t0):
    import time
    if times is not None:
        times[key] = times.get(key, 0.0) + (time.perf_counter() - t0) * 1e3
    return time.perf_counter()


def build_sources_distributed(session, n, rank, world, scratch=None, times=None):
    """source tree + equivalent particles, replicated on every rank with 1/world of the sorting work each"""
    import time
    t = time.perf_counter()
    lo, hi = session.shard_particle_range(n, rank, world)
    session.make_tree_range(0, lo, hi); t = _tick(times, "src_tree_range", t)
    scratch = exchange_planes(session, 0, n, rank, world, scratch=scratch); t = _tick(times, "src_allgather", t)
    if world > 1:
        session.finish_tree(0); t = _tick(times, "src_finish", t)
    session.upward(0); t = _tick(times, "upward", t)
    return scratch


def build_targets_sharded(session, n, rank, world, scratch=None, times=None):
    """target tree: each rank sorts (and later refines and evaluates) only its own leaves, but the centres of the
    ancestor nodes enter the dual-tree MAC (ongrav3d.cpp:338) and depend on every particle below them, so the coordinate
    planes are exchanged as well and the node arrays completed bottom-up - before the in-leaf refinement, as in the
    reference (makeTree's finishTree runs before refineTree)."""
    import time
    t = time.perf_counter()
    lo, hi = session.shard_particle_range(n, rank, world)
    session.make_tree_range(1, lo, hi); t = _tick(times, "tgt_tree_range", t)
    scratch = exchange_planes(session, 1, n, rank, world, scratch=scratch); t = _tick(times, "tgt_allgather", t)
    if world > 1:
        session.finish_tree(1); t = _tick(times, "tgt_finish", t)
        session.set_build_range(1, lo, hi)
    session.refine(1); t = _tick(times, "refine", t)
    session.upward(1); t = _tick(times, "tgt_equiv", t)
    return scratch


def build_both_distributed(session, nsrc, ntarg, rank, world, scratch=None, times=None):
    """both trees of a multi-GPU step: the two range-restricted builds run concurrently on the device
    (onb_make_trees_range, two streams), then the source side is completed (exchange, node arrays, upward pass),
    then the target side (exchange, node arrays, in-leaf refinement of the rank's own leaves, equivalent points)"""
    import time
    t = time.perf_counter()
    slo, shi = session.shard_particle_range(nsrc, rank, world)
    tlo, thi = session.shard_particle_range(ntarg, rank, world)
    session.make_trees_range(slo, shi, tlo, thi); t = _tick(times, "both_trees_range", t)
    scratch = exchange_planes(session, 0, nsrc, rank, world, scratch=scratch); t = _tick(times, "src_allgather", t)
    scratch = exchange_planes(session, 1, ntarg, rank, world, scratch=scratch); t = _tick(times, "tgt_allgather", t)
    # node arrays of both trees, upward pass + packing | in-leaf refinement of the rank's own leaves + equivalent points:
    # the two sides on two streams in one call (onb_prepare_eval)
    session.prepare_eval(world > 1, tlo, thi); t = _tick(times, "finish_upward_refine", t)
    return scratch
