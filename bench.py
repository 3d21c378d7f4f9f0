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
N > 1 : one process per GPU; the communicator is the C++ library's own (csrc/comm.cu, NCCL); torch.distributed only ships the
        128-byte NCCL id, reduces the timings (max over ranks) and sums the work. The hot path is the same three C-ABI calls.
Also in the record: cold_ms_per_step (a step on never-seen particles), accuracy (rms error against the direct sum on a target
sample, the drivers' own metric), exchange (device times of the NVLink collectives), device_memory_peak_bytes_per_gpu.
--particles 1000000000 --gpus 8 is BASELINE configs[4] (lean memory mode is selected automatically).
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
def _force_omp_threads(cores):
    """torchrun presets OMP_NUM_THREADS=1 for its workers; the reference arm must use every host core. Set the variable before
    libgomp is loaded AND tell an already-loaded libgomp directly. Returns the thread count OpenMP will actually use."""
    os.environ["OMP_NUM_THREADS"] = str(cores)
    import ctypes
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(cores))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return cores


def run_reference(args):
    """The reference's own CPU implementation of the same step (its unmodified templates compiled in place into
    oracle/_ref/fast with the reference's CMake flags), all host threads. The full configuration (N = 1e7) when
    warmup+steps of it fit in about ten minutes on this host, otherwise the largest bounded sample that does."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.refapi import RefSession, ref_available
    cores = os.cpu_count() or 1
    if not ref_available("grav3d", "fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fast/libref_grav3d.so is not built"}))
        return 0
    threads = _force_omp_threads(cores)

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

    # probe at 1e5 (the dual tree is O(N)), then take the largest size whose warmup+steps fit the time budget
    tb, te = step(100000)
    per_particle = (tb + te) / 1e5
    budget = float(os.environ.get("ONB_REF_BUDGET_S", "600")) / max(1, args.steps + args.warmup)
    n = args.n_ref or 100000
    if not args.n_ref:
        for cand in (args.n, 1000000, 500000, 200000, 100000):
            if cand in DTT_PAIRS and per_particle * cand * 1.15 <= budget:
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
    same = n == args.n
    rec = {
        "impl": "reference", "metric": "ongrav3d_dualtree_pair_interactions_per_s", "value": val, "unit": "Ginteractions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ongrav3d -n=%d -t=1.4 -o=4 -b=128 charges, dual-tree (BASELINE.json configs[1])" % args.n,
                   "theta": THETA, "order": ORDER, "block": BLOCK, "n_particles": args.n, "n_particles_run": n, "same_config": same,
                   "note": "full configuration" if same else "knowingly NOT the same configuration: warmup+steps of N=%d would take %.0f s on this host (probe: %.2e s per particle), "
                           "so the arm runs the bounded sample N=%d; the metric is a rate and the dual tree is O(N)" % (args.n, per_particle * args.n * (args.steps + args.warmup), per_particle, n)},
        "seconds_per_eval": tot_e / args.steps, "seconds_tree_and_upward": tot_b / args.steps, "pairs_per_step": pairs, "pairs_count": how,
        "cpu_baseline": {"value": val, "unit": "Ginteractions/s", "cores": threads, "kind": "reference", "host_cores": cores,
                         "sample": "N=%d particles of the same generator (whole step: trees+upward+dual-tree eval), %d OpenMP threads; the reference's dual tree is O(N)" % (n, threads)},
        "e2e": {"value": val, "unit": "Ginteractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(rec))
    return 0


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from onbody_b200.api import GpuSession, driver_inputs, ARITH_FAST, MEM_LEAN, comm_unique_id

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))      # plumbing only: barriers, reductions, the NCCL id
    N = args.n
    physics = "grav3d"
    lean = args.lean or N > 450000000                 # BASELINE configs[4]: N = 1e9 on 8 GPUs needs the lean memory mode
    big = N > 200000000

    g = GpuSession(physics, N, N, block=BLOCK, order=ORDER, arith=ARITH_FAST, device=local)
    if lean:
        g.set_memory_mode(MEM_LEAN)
    if world > 1:
        # the communicator lives in the C++ library (csrc/comm.cu): rank 0 makes the NCCL id, torch only ships its 128 bytes
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        g.comm_init_rank(rank, world, box[0])
    comm = g.comm_info()

    # synthetic input exactly as the driver makes it (mt19937(12345) on the host). Up to 2e8 particles every rank generates
    # and pins its own copy; above, rank 0 generates and the resident device copies are replicated over NVLink, each rank
    # keeping only its own slice on the host (what the sliced end-to-end leg reads).
    chunk_in = ((N + world - 1) // world + 31) // 32 * 32
    s_lo, s_hi = min(N, rank * chunk_in), min(N, (rank + 1) * chunk_in)
    if not big or world == 1:
        x, r, s = driver_inputs(physics, N, True)
        hx = torch.from_numpy(x); hr = torch.from_numpy(r); hs = torch.from_numpy(s)
        if not big:
            hx, hr, hs = hx.pin_memory(), hr.pin_memory(), hs.pin_memory()
        del x, r, s
        dx = hx.cuda(); dr = hr.cuda(); ds = hs.cuda()
        hx_ptr, hr_ptr, hs_ptr = hx.data_ptr(), hr.data_ptr(), hs.data_ptr()
    else:
        dx = torch.empty((3, N), dtype=torch.float32, device="cuda"); dr = torch.empty(N, dtype=torch.float32, device="cuda")
        ds = torch.empty((1, N), dtype=torch.float32, device="cuda")
        if rank == 0:
            x, r, s = driver_inputs(physics, N, True)
            dx.copy_(torch.from_numpy(x)); dr.copy_(torch.from_numpy(r)); ds.copy_(torch.from_numpy(s))
            del x, r, s
        for t in (dx, dr, ds):
            dist.broadcast(t, src=0)
        # host slices; the C ABI takes the base pointer of the full plane and, with sliced inputs on, reads only [s_lo, s_hi)
        hx = torch.empty((3, s_hi - s_lo), dtype=torch.float32).pin_memory(); hx.copy_(dx[:, s_lo:s_hi])
        hr = torch.empty(s_hi - s_lo, dtype=torch.float32).pin_memory(); hr.copy_(dr[s_lo:s_hi])
        hs = torch.empty((1, s_hi - s_lo), dtype=torch.float32).pin_memory(); hs.copy_(ds[:, s_lo:s_hi])
        hx_ptr = hr_ptr = hs_ptr = None
    t_lo, t_hi = g.shard_particle_range(N, rank, world)
    if big and world > 1:
        hu = torch.empty((3, t_hi - t_lo), dtype=torch.float32).pin_memory()
    else:
        hu = torch.empty((3, N), dtype=torch.float32).pin_memory() if not big else torch.empty((3, N), dtype=torch.float32)
    h2d = (3 + 1 + 1 + 3 + 1) * N * 4           # sources x,r,s + targets x,r
    d2h = 3 * N * 4

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    phases = {}

    def hot_path():
        # the same three calls on 1 and on N GPUs: with a communicator attached the library sorts only this rank's leaf range of
        # both trees, exchanges leaf records / source planes / equivalent strengths over NVLink (overlapped with the local
        # work) and evaluates the rank's target shard
        g.make_trees(); phases["both_trees"] = phases.get("both_trees", 0.0) + g.phase_ms("tree")
        g.prepare_eval(); phases["upward_refine_tgt_equiv"] = phases.get("upward_refine_tgt_equiv", 0.0) + g.phase_ms("prepare")
        if world > 1:
            for k in ("ag_src_planes", "bcast_eq_strengths"):
                phases[k] = phases.get(k, 0.0) + max(0.0, g.phase_ms(k))
        g.fastsumm(THETA)
        for k in ("eval", "lists", "p2p", "downward"):
            phases[k] = phases.get(k, 0.0) + g.phase_ms(k)

    def step_resident(xsrc=None):
        xs = dx if xsrc is None else xsrc
        g.set_sliced_inputs(False)
        g.set_sources_ptr(N, xs.data_ptr(), dr.data_ptr(), ds.data_ptr())      # device -> device, untimed
        g.set_targets_ptr(N, xs.data_ptr(), dr.data_ptr())
        g.timer_start()
        hot_path()
        return g.timer_stop_ms()

    def step_e2e():
        g.set_async_inputs(True)       # pinned buffers: the target copy overlaps the source tree build (onb_set_async_inputs)
        if world > 1:
            g.set_sliced_inputs(True)  # each rank pulls 1/world of every plane over its own PCIe link; NVLink replicates
        g.timer_start()
        if hx_ptr is not None:
            g.set_sources_ptr(N, hx_ptr, hr_ptr, hs_ptr)                       # (pinned) host -> device
            g.set_targets_ptr(N, hx_ptr, hr_ptr)
        else:
            # only this rank's slice exists on the host: hand over per-plane base pointers shifted so that index s_lo is the slice start
            w = s_hi - s_lo
            xs = [hx.data_ptr() + 4 * (d * w - s_lo) for d in range(3)]
            g.set_planes_ptr(N, xs, hr.data_ptr() - 4 * s_lo, [hs.data_ptr() - 4 * s_lo])
        hot_path()
        if world > 1:
            if big:
                g.shard_results_into(hu.data_ptr() - 4 * t_lo, stride=t_hi - t_lo)    # each rank returns its own target shard
            else:
                g.shard_results_into(hu.data_ptr())
        else:
            g.results_into(hu.data_ptr())                                      # device -> pinned host
        ms = g.timer_stop_ms()
        g.set_async_inputs(False); g.set_sliced_inputs(False)
        return ms

    mem_peak = 0
    for _ in range(max(args.warmup, 1)):
        step_resident()
        mem_peak = max(mem_peak, g.device_memory()[0])
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
    phases_last = {k: v / args.steps for k, v in ph_res.items()}
    pairs_local = g.last_pairs()
    pool_used, pool_cap = g.phase_ms("dtt_pool_used"), g.phase_ms("dtt_pool_cap")
    # ---- timed: K end-to-end steps
    phases.clear()
    barrier()
    t_e2e = 0.0; e2e_steps = []
    for _ in range(args.steps):
        e2e_steps.append(step_e2e()); t_e2e += e2e_steps[-1]
    barrier()
    sampler.stop()
    mem_peak = max(mem_peak, g.device_memory()[0])
    # ---- a COLD step: particles the context has never seen (a time-stepping caller's normal case). Nothing is cached between
    # evaluations - the dual-tree lists are bump-allocated on the device - so this must cost what a repeated step costs.
    dx2 = dx * 0.98 if not big else None
    cold_ms = cold_attempts = None
    if dx2 is not None:
        barrier()
        cold_ms = step_resident(dx2); cold_attempts = g.phase_ms("dtt_attempts")
        barrier()
        del dx2
    # ---- accuracy against the direct sum on a sample of the targets (ongrav3d.cpp:556-568,782-789), distributed like the run
    err = None
    if args.check_error or big:
        tskip = max(1, N // 2048)
        step_resident()
        uf = torch.empty((3, N), dtype=torch.float32) if not big else None
        def shard_u():
            buf = torch.empty((3, t_hi - t_lo), dtype=torch.float32)
            g.shard_results_into(buf.data_ptr() - 4 * t_lo, stride=t_hi - t_lo)
            return buf
        fast_u = shard_u()
        g.zero_vels(); g.naive(tskip)
        naive_u = shard_u()
        k0 = (t_lo + tskip - 1) // tskip
        idx = torch.arange(k0 * tskip, t_hi, tskip) - t_lo
        e = (fast_u[0, idx].double() - naive_u[0, idx].double())
        acc = torch.tensor([float((e * e).sum()), float((naive_u[0, idx].double() ** 2).sum()), float(idx.numel())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        err = {"rms_vs_direct": float((acc[0] / acc[1]).sqrt()), "samples": int(acc[2].item()), "tskip": tskip,
               "note": "x-component, every tskip-th target in tree order, the drivers' own error metric (ongrav3d.cpp:782-789); reference: about 1e-4 at theta 1.4"}

    # max over ranks for times, sum for work
    red = torch.tensor([t_res, t_e2e, float(pairs_local), float(launches), float(mem_peak), float(cold_ms or 0.0)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = red.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_res, t_e2e, mem_peak = mx[0].item(), mx[1].item(), mx[4].item()
        cold_ms = mx[5].item() if cold_ms is not None else None
        pairs_total, launches_total = int(sm[2].item()), int(sm[3].item())
    else:
        pairs_total, launches_total = int(pairs_local), int(launches)
    exch = None
    if world > 1:
        exch = {k: {"ms": phases_last.get(k), "bytes_per_rank": g.phase_ms(k + "_bytes"),
                    "GBps_algorithmic": (g.phase_ms(k + "_bytes") / (phases_last[k] * 1e-3) * 1e-9) if phases_last.get(k) else None}
                for k in ("ag_src_planes", "bcast_eq_strengths")}
    # tear down in the same order on every rank: the library's communicator first, then torch's
    if rank != 0:
        g.close()
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
                   "parallelism": ("target leaves sharded x%d; both tree builds split by leaf range; leaf records, source planes and equivalent strengths exchanged by the library's own NCCL communicator (%s, NCCL %s)" % (world, comm["transport"], comm["nccl_version"])) if world > 1 else "single GPU",
                   "memory_mode": "lean" if lean else "normal",
                   "l2_policy": "inputs larger than L2 (%.0f MB of particle planes per tree vs 126 MB L2); every step rebuilds from pristine input" % (N * 24 / 1e6)},
        "seconds_per_step": sec_res, "ms_steps": [round(v, 2) for v in res_steps], "e2e_ms_steps": [round(v, 2) for v in e2e_steps], "seconds_per_eval": ph_res.get("eval", 0.0) / K * 1e-3, "pairs_per_step": pairs_total,
        "phases_ms": {k: v / K for k, v in ph_res.items()},
        # a step on particles the context has never seen (positions scaled by 0.98): nothing is cached between evaluations
        "cold_ms_per_step": cold_ms, "cold_dtt_attempts": cold_attempts,
        "dtt_list_pool": {"entries_used": pool_used, "entries_allocated": pool_cap},
        "device_memory_peak_bytes_per_gpu": mem_peak,
        "exchange": exch,
        "accuracy": err,
        "e2e": {"value": e2e, "unit": "Ginteractions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": sec_e2e * 1e3,
                "note": "multi GPU: every input plane crosses PCIe once in total (onb_set_sliced_inputs: 1/world per rank, replicated by the library over NVLink); each rank returns its own target shard" if world > 1 else "sources and targets copied separately from pinned host memory; all outputs copied back"},
        "gpu_launches": launches_total,
        "clocks": clocks,
        "roofline": {"kernel": "k_p2p_lists<grav3d,fast>", "bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved_tf / peak_tf) if peak_tf else None,
                     # DRAM bytes per launch of the packed pair kernel at this workload, from the ncu launch list of the same
                     # step (profiles/r2_launch_list_summary.txt: 1671.7 MB read + 181.1 MB written over 5 launches)
                     "traffic": 370.6e6 if (N == 10000000 and world == 1) else None,
                     "traffic_source": "profiles/r2_launch_list_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, k_p2p_lists<grav3d,fast,TPT=4,packed>, per launch)" if (N == 10000000 and world == 1) else None,
                     "peak_source": "FP32 FMA issue microbenchmark run in this process (onb_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 SIMT figure",
                     "flop_per_pair": FLOP_PER_PAIR,
                     "fp32_issue_util": (pairs_local * FP32_SLOTS_PER_PAIR * 2 / (p2p_ms * 1e-3) * 1e-12 / peak_tf) if (peak_tf and p2p_ms > 0) else None,
                     "share_of_step": (p2p_ms / (sec_res * 1e3)) if sec_res > 0 else None},
        "roofline_tree": {"kernels": "k_big_level + k_node_split + k_gather + k_subtree + k_apply_perm (both trees)", "bound": "hbm",
                          "achieved": tree_bytes / (tree_ms * 1e-3) * 1e-9 if (tree_ms > 0 and world == 1) else None,
                          "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "peak_source": peaks_src,
                          "frac": (tree_bytes / (tree_ms * 1e-3) * 1e-9 / peaks.get("hbm_gbs")) if (tree_ms > 0 and world == 1) else None,
                          "algorithmic_bytes": tree_bytes, "lower_bound_bytes": tree_bytes_lb,
                          # measured DRAM bytes of all tree kernels of one step (both trees), ncu launch list of the same workload
                          "traffic": 16.25e9 if (N == 10000000 and world == 1) else None,
                          "traffic_source": "profiles/r2_launch_list_summary.txt (k_big_level + k_gather + k_subtree + k_node_split + k_apply_perm + k_copy_back, dram read + write)" if (N == 10000000 and world == 1) else None,
                          "note": "algorithmic bytes = level-synchronous partition build (see bench.py/_tree_bytes and DESIGN.md section 4); lower_bound_bytes = every plane read and written once"},
        # the barycentric upward pass (source side of onb_prepare_eval): SURVEY 8d models it as 52 B x N of HBM traffic, ncu shows it is
        # bound by instruction issue and shared memory instead (profiles/r2_ncu_full_upward.txt: SM throughput 80 %, DRAM 4 %)
        "roofline_upward": {"kernel": "k_upward<3,1,5> (17 one-level launches per tree)", "bound": "hbm", "unit": "GB/s", "peak": peaks.get("hbm_gbs"), "peak_source": peaks_src,
                            "algorithmic_bytes": 52 * N, "ms": 1.79 if (N == 10000000 and world == 1) else None,
                            "achieved": (52 * N / 1.79e-3 * 1e-9) if (N == 10000000 and world == 1) else None,
                            "frac": (52 * N / 1.79e-3 * 1e-9 / peaks.get("hbm_gbs")) if (N == 10000000 and world == 1) else None,
                            "traffic": 423.9e6 if (N == 10000000 and world == 1) else None,
                            "source": "profiles/r2_launch_list_summary.txt (duration and DRAM bytes of the 34 k_upward launches of one step, both trees), profiles/r2_ncu_full_upward.txt"},
    }
    # ---- CPU baseline beside it (rank 0, N=1 only): the reference itself on a bounded sample
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle.refapi import RefSession, ref_available
            cores = _force_omp_threads(os.cpu_count() or 1)
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
    print(json.dumps(rec), flush=True)
    g.close()
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
    ap.add_argument("--lean", action="store_true", help="lean memory mode (automatic above 4.5e8 particles)")
    ap.add_argument("--no-check-error", dest="check_error", action="store_false",
                    help="skip the accuracy leg (rms error against the direct sum on ~2048 sample targets, after the timed regions)")
    ap.add_argument("--check-error", dest="check_error", action="store_true", default=True, help=argparse.SUPPRESS)
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
