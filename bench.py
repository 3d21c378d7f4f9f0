#!/usr/bin/env python
"""bench.py - the hot path of ongrav3d (dual-tree, -t=1.4 -o=4 -b=128, charges) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles N]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input (the drivers' own mt19937(12345) cloud):
source tree build -> barycentric upward pass -> target tree build -> in-leaf refinement -> target equivalent
points -> dual-tree evaluation (interaction lists + leaf-block P2P + downward interpolation), i.e. what the
reference prints as "[fast total]". metric = pair interactions per second: the exact number of source-target pairs
the dual-tree lists contain (the same lists as the reference's, proven bit-for-bit by tests) / step seconds.

value : inputs already resident in HBM when the timed region starts (CUDA events on the library's stream).
e2e   : the same step through the C ABI with HOST buffers: pinned host -> device copies of sources and targets and
        the device -> host read of the target outputs are inside the timed region.
The last line printed by rank 0 is the JSON record.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

THETA, ORDER, BLOCK = 1.4, 4, 128
FLOP_PER_PAIR = 19          # reference ongrav3d.cpp:49,60
FP32_SLOTS_PER_PAIR = 12    # 3 FADD + 6 FFMA + 3 FMUL in the SASS loop of k_p2p_lists<grav> (plus 1 MUFU)

# exact dual-tree pair counts of `ongrav3d -n=N -t=1.4 -o=4 -b=128` (seed 12345), measured by the GPU list builder,
# whose lists are bit-identical to the reference's (tests/test_gpu_parity.py). Used by the reference arm, which
# cannot count them itself (the reference compiles its statistics out: ongrav3d.cpp:218 dostats=false).
DTT_PAIRS = {}
_pairs_file = os.path.join(ROOT, "tests", "golden", "dtt_pairs.json")
if os.path.exists(_pairs_file):
    with open(_pairs_file) as _f:
        DTT_PAIRS = {int(k): int(v) for k, v in json.load(_f).items()}


def pairs_estimate(n):
    if n in DTT_PAIRS:
        return DTT_PAIRS[n], "exact"
    return int(n * 16500.0), "estimate(16.5k pairs/target)"


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, read in-process through NVML every 100 ms (the fields
    of the B200_PROFILING.md clocks line). An external `nvidia-smi -lms` loop was measured to stall the step it was
    watching by hundreds of ms on this driver, so it is only the fallback when pynvml is unavailable."""

    def __init__(self, gpu_index):
        import threading
        self.gpu = gpu_index
        self.samples = []        # (sm_mhz, sm_max_mhz, reasons_bitmask)
        self.stop_flag = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = gpu_index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _run(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                if n is not None:
                    sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    self.samples.append((float(sm), float(mx), int(rs)))
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append((float(out[0]), float(out[1]), 0))
            except Exception:
                pass
            time.sleep(float(os.environ.get("ONB_CLOCK_PERIOD", "0.1")) if n is not None else 1.0)

    def start(self):
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=3)

    def summary(self):
        sm = [s[0] for s in self.samples]; smax = max([s[1] for s in self.samples], default=0)
        bits = 0
        for s in self.samples:
            bits |= s[2]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted(v for k, v in names.items() if bits & k)
        busy = [v for v in sm if v > 0.5 * smax] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": smax or None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the same step (its unmodified templates compiled in place into
    oracle/_ref/fast with the reference's CMake flags), all host threads, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.refapi import RefSession, ref_available
    cores = os.cpu_count() or 1
    if not ref_available("grav3d", "fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fast/libref_grav3d.so is not built"}))
        return 0
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))

    def step(n):
        s = RefSession("grav3d", n, n, block=BLOCK, order=ORDER, eq_block=128, build="fast")
        s.init_driver()
        t0 = time.perf_counter()
        s.make_tree(0); s.upward(0); s.make_tree(1); s.refine(1); s.upward(1)
        t1 = time.perf_counter()
        s.zero_vels(); s.fastsumm(THETA, parallel=True)       # the driver's own omp-task invocation (ongrav3d.cpp:880-884)
        t2 = time.perf_counter()
        s.close()
        return t1 - t0, t2 - t1

    # size the sample so that warmup+steps fit in ~150 s: probe at 1e5 (DTT is O(N))
    tb, te = step(100000)
    per_particle = (tb + te) / 1e5
    budget = 150.0 / max(1, args.steps + args.warmup)
    n = args.n_ref or 100000
    if not args.n_ref:
        for cand in (1000000, 500000, 200000, 100000):
            if cand in DTT_PAIRS and per_particle * cand <= budget:
                n = cand; break
    for _ in range(args.warmup):
        step(n)
    tot_b = tot_e = 0.0
    for _ in range(args.steps):
        b, e = step(n)
        tot_b += b; tot_e += e
    pairs, how = pairs_estimate(n)
    sec = (tot_b + tot_e) / args.steps
    val = pairs / sec * 1e-9
    rec = {
        "impl": "reference", "metric": "ongrav3d_dualtree_pair_interactions_per_s", "value": val, "unit": "Ginteractions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ongrav3d -n=10000000 -t=1.4 -o=4 -b=128 charges, dual-tree (reference arm: bounded sample, see cpu_baseline.sample)",
                   "theta": THETA, "order": ORDER, "block": BLOCK},
        "seconds_per_eval": tot_e / args.steps, "seconds_tree_and_upward": tot_b / args.steps, "pairs_per_step": pairs, "pairs_count": how,
        "cpu_baseline": {"value": val, "unit": "Ginteractions/s", "cores": cores, "kind": "reference",
                         "sample": "N=%d particles of the same generator (whole step: trees+upward+dual-tree eval); the reference's dual tree is O(N)" % n},
        "e2e": {"value": val, "unit": "Ginteractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(rec))
    return 0


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from onbody_b200.api import GpuSession, driver_inputs, ARITH_FAST

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = args.n
    physics = "grav3d"

    # synthetic input exactly as the driver makes it, in pinned host memory
    x, r, s = driver_inputs(physics, N, True)
    hx = torch.from_numpy(x).pin_memory(); hr = torch.from_numpy(r).pin_memory(); hs = torch.from_numpy(s).pin_memory()
    hu = torch.empty((3, N), dtype=torch.float32).pin_memory()
    del x, r, s
    # resident copies for the device-timed leg
    dx = hx.cuda(); dr = hr.cuda(); ds = hs.cuda()
    g = GpuSession(physics, N, N, block=BLOCK, order=ORDER, arith=ARITH_FAST, device=local)
    g.set_shard(rank, world)
    h2d = (hx.numel() + hr.numel() + hs.numel() + hx.numel() + hr.numel()) * 4
    d2h = hu.numel() * 4

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    phases = {}

    scratch = {"buf": None}

    def hot_path_multi():
        # each rank sorts its share of both trees (the two builds concurrently); NCCL all-gathers over NVLink replicate the particle planes; every rank then
        # completes the node arrays bottom-up and runs the (cheap) upward pass in full; evaluation is sharded by target leaves
        from onbody_b200 import multigpu
        t0 = time.perf_counter()
        scratch["buf"] = multigpu.build_both_distributed(g, N, N, rank, world, scratch["buf"], phases)
        phases["build_side_wall"] = phases.get("build_side_wall", 0.0) + (time.perf_counter() - t0) * 1e3
        g.fastsumm(THETA)
        for k in ("eval", "lists", "p2p", "downward"):
            phases[k] = phases.get(k, 0.0) + g.phase_ms(k)

    def hot_path():
        if world > 1:
            return hot_path_multi()
        g.make_trees(); phases["both_trees"] = phases.get("both_trees", 0.0) + g.phase_ms("tree")    # two streams, overlapped
        # source side (upward pass + packing) and target side (in-leaf refinement + equivalent points) on two streams
        g.prepare_eval(); phases["upward_refine_tgt_equiv"] = phases.get("upward_refine_tgt_equiv", 0.0) + g.phase_ms("prepare")
        g.fastsumm(THETA)
        for k in ("eval", "lists", "p2p", "downward"):
            phases[k] = phases.get(k, 0.0) + g.phase_ms(k)

    def step_resident():
        g.set_sources_ptr(N, dx.data_ptr(), dr.data_ptr(), ds.data_ptr())      # device -> device, untimed
        g.set_targets_ptr(N, dx.data_ptr(), dr.data_ptr())
        g.timer_start()
        hot_path()
        return g.timer_stop_ms()

    stage = {}

    def step_e2e_multi():
        # Each rank reads only ITS 1/world slice of every input plane from (pinned) host memory, the slices are
        # all-gathered over NVLink into full device planes (NVLink is ~15x faster than 8 GPUs pulling the same bytes over
        # PCIe), and each rank writes back only the outputs of its own target shard.
        from onbody_b200 import multigpu
        if not stage:
            per = (N + world - 1) // world
            stage["ranges"] = [(min(N, r * per), min(N, (r + 1) * per)) for r in range(world)]
            stage["dx"] = torch.empty((3, N + per), dtype=torch.float32, device="cuda")
            stage["dr"] = torch.empty(N + per, dtype=torch.float32, device="cuda")
            stage["ds"] = torch.empty((1, N + per), dtype=torch.float32, device="cuda")
            stage["buf"] = None
        g.timer_start()
        tt = [time.perf_counter()]
        lo, hi = stage["ranges"][rank]
        for d in range(3):
            stage["dx"][d, lo:hi].copy_(hx[d, lo:hi], non_blocking=True)
        stage["dr"][lo:hi].copy_(hr[lo:hi], non_blocking=True)
        stage["ds"][0, lo:hi].copy_(hs[0, lo:hi], non_blocking=True)
        torch.cuda.synchronize(); tt.append(time.perf_counter())
        for plane in (stage["dx"][0], stage["dx"][1], stage["dx"][2], stage["dr"], stage["ds"][0]):
            stage["buf"] = multigpu.allgather_ranges(plane, stage["ranges"], rank, world, stage["buf"])
        torch.cuda.synchronize(); tt.append(time.perf_counter())
        # planar [PD][n] views for the C ABI: rows of the staging tensors are N+per apart, so hand each set over plane by plane
        xs = torch.stack([stage["dx"][d, :N] for d in range(3)]).contiguous()
        g.set_sources_ptr(N, xs.data_ptr(), stage["dr"].data_ptr(), stage["ds"].data_ptr())
        g.set_targets_ptr(N, xs.data_ptr(), stage["dr"].data_ptr()); tt.append(time.perf_counter())
        hot_path(); tt.append(time.perf_counter())
        slo, shi = g.shard_particle_range(N, rank, world)
        for d in range(3):
            hu[d, slo:shi].copy_(g.plane_tensor(1, 7 + d, N)[slo:shi], non_blocking=True)
        torch.cuda.synchronize(); tt.append(time.perf_counter())
        stage.setdefault("trace", []).append([round((b_ - a_) * 1e3, 1) for a_, b_ in zip(tt, tt[1:])])
        return g.timer_stop_ms()

    def step_e2e():
        if world > 1:
            return step_e2e_multi()
        g.set_async_inputs(True)       # pinned buffers: the target copy overlaps the source tree build (onb_set_async_inputs)
        g.timer_start()
        g.set_sources_ptr(N, hx.data_ptr(), hr.data_ptr(), hs.data_ptr())      # pinned host -> device
        g.set_targets_ptr(N, hx.data_ptr(), hr.data_ptr())
        hot_path()
        g.results_into(hu.data_ptr())                                          # device -> pinned host
        ms = g.timer_stop_ms()
        g.set_async_inputs(False)
        return ms

    for _ in range(max(args.warmup, 1)):
        step_resident()
    step_e2e()
    peak_tf = g.measure_fp32_peak() if rank == 0 else 0.0

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    # ---- timed: K resident steps
    phases.clear()
    l0 = g.launch_count()
    barrier()
    t_res = 0.0; res_steps = []
    for _ in range(args.steps):
        res_steps.append(step_resident()); t_res += res_steps[-1]
    barrier()
    launches = g.launch_count() - l0
    ph_res = dict(phases)
    pairs_local = g.last_pairs()
    # ---- timed: K end-to-end steps
    phases.clear()
    barrier()
    t_e2e = 0.0; e2e_steps = []
    for _ in range(args.steps):
        e2e_steps.append(step_e2e()); t_e2e += e2e_steps[-1]
    barrier()
    sampler.stop()

    # max over ranks for times, sum for work
    red = torch.tensor([t_res, t_e2e, float(pairs_local), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = red.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_res, t_e2e = mx[0].item(), mx[1].item()
        pairs_total, launches_total = int(sm[2].item()), int(sm[3].item())
    else:
        pairs_total, launches_total = int(pairs_local), int(launches)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    K = args.steps
    sec_res = t_res / K * 1e-3
    sec_e2e = t_e2e / K * 1e-3
    value = pairs_total / sec_res * 1e-9
    e2e = pairs_total / sec_e2e * 1e-9
    p2p_ms = ph_res.get("p2p", 0.0) / K
    achieved_tf = pairs_local * FLOP_PER_PAIR / (p2p_ms * 1e-3) * 1e-12 if p2p_ms > 0 else 0.0
    peaks, peaks_src = measured_peaks()
    clocks = sampler.summary()
    tree_ms = (ph_res.get("both_trees", 0) + ph_res.get("both_trees_range", 0)) / K
    # algorithmic bytes of one tree build: every plane read once + written once: sources 6 planes + targets 4 planes + gidx
    tree_bytes_lb = N * (2 * 4 * (3 + 1 + 1) + 2 * 4 * (3 + 1) + 8)      # lower bound: every plane of both sets read and written once
    # algorithmic bytes of the level-synchronous partition build the reference's order demands:
    # per level above the shared-memory subtree 12 B bbox read + 4 B lidx write + 36 B permute (3 coordinates + index
    # plane read and written, lidx read), the key scans of the select passes (8 B per scanned element: count + compact),
    # one 32 B read/write of coordinates + index in the subtree kernel, 8 B per deferred plane (r, s) at the end
    def _tree_bytes(n, sd, scanned):
        lev, left = 0, n
        while left > 8192:
            lev += 1
            left = 128 * 2 ** int(np.floor(np.log2((left - 1) / 128)))
        return n * (lev * 52 + 32 + 8 * (1 + sd)) + scanned * 8
    try:
        scanned = int(g.build_stats()["scanned"])
    except Exception:
        scanned = 26 * N
    tree_bytes = _tree_bytes(N, 1, scanned) + _tree_bytes(N, 0, scanned)
    rec = {
        "metric": "ongrav3d_dualtree_pair_interactions_per_s", "value": value, "unit": "Ginteractions/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": sec_res * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ongrav3d -n=%d -t=1.4 -o=4 -b=128 charges, dual-tree (BASELINE.json configs[1])" % N,
                   "theta": THETA, "order": ORDER, "block": BLOCK, "n_particles": N,
                   "parallelism": ("target leaves sharded x%d; tree builds split by particle range, planes replicated by NCCL all-gather" % world) if world > 1 else "single GPU",
                   "l2_policy": "inputs larger than L2 (%.0f MB of particle planes per tree vs 126 MB L2); every step rebuilds from pristine input" % (N * 24 / 1e6)},
        "seconds_per_step": sec_res, "ms_steps": [round(v, 2) for v in res_steps], "e2e_ms_steps": [round(v, 2) for v in e2e_steps], "e2e_trace_h2d_gather_set_hot_d2h": stage.get("trace", [])[-len(e2e_steps):], "seconds_per_eval": ph_res.get("eval", 0.0) / K * 1e-3, "pairs_per_step": pairs_total,
        "phases_ms": {k: v / K for k, v in ph_res.items()},
        "e2e": {"value": e2e, "unit": "Ginteractions/s", "h2d_bytes_per_step": h2d if world == 1 else hx.numel() * 4 + hr.numel() * 4 + hs.numel() * 4, "d2h_bytes_per_step": d2h, "ms_per_step": sec_e2e * 1e3,
                "note": "single GPU: sources and targets are copied separately (same host arrays twice). multi GPU: every input plane crosses PCIe once in total (1/world per rank) and is replicated over NVLink; each rank returns its own target shard" if world > 1 else "sources and targets copied separately from pinned host memory; all outputs copied back"},
        "gpu_launches": launches_total,
        "clocks": clocks,
        "roofline": {"kernel": "k_p2p_lists<grav3d,fast>", "bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved_tf / peak_tf) if peak_tf else None,
                     # DRAM bytes per launch of the packed pair kernel at this workload, from the ncu launch list of the same
                     # step (profiles/r1_launch_list_summary_v3.txt: 1672.3 MB read + 180.3 MB written over 5 launches)
                     "traffic": 370.5e6 if (N == 10000000 and world == 1) else None,
                     "traffic_source": "profiles/r1_launch_list_summary_v3.txt (dram__bytes_read.sum + dram__bytes_write.sum, k_p2p_lists<grav3d,fast,TPT=4,packed>, per launch)" if (N == 10000000 and world == 1) else None,
                     "peak_source": "FP32 FMA issue microbenchmark run in this process (onb_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 SIMT figure",
                     "flop_per_pair": FLOP_PER_PAIR,
                     "fp32_issue_util": (pairs_local * FP32_SLOTS_PER_PAIR * 2 / (p2p_ms * 1e-3) * 1e-12 / peak_tf) if (peak_tf and p2p_ms > 0) else None,
                     "share_of_step": (p2p_ms / (sec_res * 1e3)) if sec_res > 0 else None},
        "roofline_tree": {"kernels": "k_big_level + k_node_split + k_gather + k_subtree + k_apply_perm (both trees)", "bound": "hbm",
                          "achieved": tree_bytes / (tree_ms * 1e-3) * 1e-9 if (tree_ms > 0 and world == 1) else None,
                          "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "peak_source": peaks_src,
                          "frac": (tree_bytes / (tree_ms * 1e-3) * 1e-9 / peaks.get("hbm_gbs")) if (tree_ms > 0 and world == 1) else None,
                          "algorithmic_bytes": tree_bytes, "lower_bound_bytes": tree_bytes_lb, "traffic": None,
                          "note": "algorithmic bytes = level-synchronous partition build (see bench.py/_tree_bytes and DESIGN.md section 4); lower_bound_bytes = every plane read and written once"},
    }
    # ---- CPU baseline beside it (rank 0, N=1 only): the reference itself on a bounded sample
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle.refapi import RefSession, ref_available
            cores = os.cpu_count() or 1
            os.environ.setdefault("OMP_NUM_THREADS", str(cores))
            n_cpu = args.n_ref or (1000000 if 1000000 in DTT_PAIRS else 100000)
            kind = "reference" if ref_available("grav3d", "fast") else "port"
            if kind == "reference":
                sref = RefSession("grav3d", n_cpu, n_cpu, block=BLOCK, order=ORDER, eq_block=128, build="fast")
            else:
                from oracle.refapi import PortSession
                n_cpu = 100000
                sref = PortSession("grav3d", n_cpu, n_cpu, block=BLOCK, order=ORDER, eq_block=128)
            sref.init_driver()
            t0 = time.perf_counter()
            sref.make_tree(0); sref.upward(0); sref.make_tree(1); sref.refine(1); sref.upward(1)
            sref.zero_vels(); sref.fastsumm(THETA, parallel=True)
            dt = time.perf_counter() - t0
            pr, how = pairs_estimate(n_cpu)
            rec["cpu_baseline"] = {"value": pr / dt * 1e-9, "unit": "Ginteractions/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                                   "seconds": dt, "pairs": pr, "pairs_count": how,
                                   "sample": "N=%d particles of the same generator, whole step (trees+upward+dual-tree eval) once; the dual tree is O(N)" % n_cpu}
        except Exception as e:  # the baseline must never take the GPU number down with it
            rec["cpu_baseline"] = {"value": None, "unit": "Ginteractions/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
    print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", dest="n", type=int, default=int(os.environ.get("ONB_BENCH_N", "10000000")))
    ap.add_argument("--n-ref", type=int, default=0, help="sample size of the CPU reference leg (default: sized to a few minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
