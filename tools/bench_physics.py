#!/usr/bin/env python
"""Step times of the other BASELINE.json configurations (onvort3d dual-tree / boxwise, onvortgrad3d boxwise, ...) on 1..8
B200s. Same phase sequence and multi-GPU plumbing as bench.py (which stays on the headline ongrav3d workload); inputs
resident in HBM; prints one JSON line on rank 0.
    python tools/bench_physics.py vort3d dualtree 10000000 1.4 [steps]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_physics.py vortgrad3d boxwise 10000000 1.4"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from onbody_b200.api import GpuSession, driver_inputs, comm_unique_id

FLOPS = {"grav3d": 19, "vort3d": 28, "vortgrad3d": 64, "vort2d": 13, "vort2dtr": 15}
physics = sys.argv[1]; method = sys.argv[2]; N = int(float(sys.argv[3])); theta = float(sys.argv[4]); steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, r, s = driver_inputs(physics, N, True)
dx = torch.from_numpy(x).cuda(); dr = torch.from_numpy(r).cuda(); ds = torch.from_numpy(s).cuda()
g = GpuSession(physics, N, N, device=local)
if world > 1:      # the library's own NCCL communicator; torch only ships the id
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    g.comm_init_rank(rank, world, box[0])


def step():
    g.set_sources_ptr(N, dx.data_ptr(), dr.data_ptr(), ds.data_ptr()); g.set_targets_ptr(N, dx.data_ptr(), dr.data_ptr())
    g.timer_start()
    g.make_trees()                      # with a communicator attached: range builds + exchanges inside the library
    if method == "dualtree":
        g.prepare_eval()
    else:
        g.upward(0)
    g.zero_vels()
    if method == "dualtree":
        g.fastsumm(theta)
    else:
        g.treecode3(theta)
    return g.timer_stop_ms(), g.phase_ms("eval"), g.phase_ms("p2p"), g.last_pairs()


for _ in range(2):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
tot = ev = pp = 0.0
for _ in range(steps):
    a, b, c, pairs = step(); tot += a; ev += b; pp += c
red = torch.tensor([tot, ev, pp, float(pairs)], dtype=torch.float64, device="cuda")
if world > 1:
    mx = red.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); sm = red.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    tot, ev, pp, pairs = mx[0].item(), mx[1].item(), mx[2].item(), sm[3].item()
if rank == 0:
    peak = g.measure_fp32_peak()
    cpu = None
    if world == 1 and os.environ.get("ONB_CPU_BASELINE"):
        # the reference's own OpenMP path on the host cores, bounded sample of the same workload (checker code, timing only)
        import ctypes
        from oracle.refapi import RefSession, ref_available
        if ref_available(physics, "fast"):
            cores = os.cpu_count() or 1
            os.environ["OMP_NUM_THREADS"] = str(cores)
            try:
                ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
            except OSError:
                pass
            n_cpu = int(float(os.environ["ONB_CPU_BASELINE"]))
            o = RefSession(physics, n_cpu, n_cpu, build="fast"); o.init_driver()
            t0 = time.perf_counter()
            o.make_tree(0); o.upward(0); o.make_tree(1)
            if method == "dualtree":
                o.refine(1); o.upward(1); o.zero_vels(); o.fastsumm(theta, parallel=True)
            else:
                o.zero_vels(); o.treecode3(theta)
            cpu = {"kind": "reference", "cores": cores, "n": n_cpu, "seconds": time.perf_counter() - t0,
                   "sample": "N=%d of the same generator, whole step, %d OpenMP threads" % (n_cpu, cores)}
    print(json.dumps({"cpu_baseline": cpu, "physics": physics, "method": method, "n": N, "theta": theta, "n_gpus": world, "ms_per_step": tot / steps, "ms_eval": ev / steps,
                      "ms_p2p_max_rank": pp / steps, "pairs": int(pairs), "Ginteractions_per_s": pairs / (tot / steps) * 1e-6,
                      "p2p_TFLOPs_per_gpu": (pairs / world) * FLOPS[physics] / (pp / steps) * 1e-9, "fp32_peak_TFLOPs": peak}))
if world > 1:
    dist.destroy_process_group()
