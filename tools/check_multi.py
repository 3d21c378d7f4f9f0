"""N-rank run against the single-GPU run of the same step, bit for bit (launch with torchrun, 2+ GPUs):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from onbody_b200.api import GpuSession, driver_inputs
from onbody_b200 import multigpu
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300000
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, r, s = driver_inputs("grav3d", N, True)
g = GpuSession("grav3d", N, N, device=local); g.set_shard(rank, world)
g.set_sources(x, r, s); g.set_targets(x, r)
multigpu.build_both_distributed(g, N, N, rank, world)
g.zero_vels(); g.fastsumm(1.4)
lo, hi = g.shard_particle_range(N, rank, world)
mine = g.parts(1, ("u", "gidx", "x"))
ok = True
if True:
    h = GpuSession("grav3d", N, N, device=local)
    h.set_sources(x, r, s); h.set_targets(x, r)
    h.make_trees(); h.upward(0); h.refine(1); h.upward(1); h.zero_vels(); h.fastsumm(1.4)
    ref = h.parts(1, ("u", "gidx", "x"))
    for k in ("u", "gidx", "x"):
        a = np.ascontiguousarray(mine[k][..., lo:hi]); b = np.ascontiguousarray(ref[k][..., lo:hi])
        same = np.array_equal(a.view(np.uint8), b.view(np.uint8))
        ok = ok and same
        print("rank %d %s[%d:%d] bit-identical to the single-GPU run: %s" % (rank, k, lo, hi, same), flush=True)
    es_m = g.parts(2, ("s",))["s"]; es_r = h.parts(2, ("s",))["s"]
    same = np.array_equal(es_m.view(np.uint8), es_r.view(np.uint8)); ok = ok and same
    print("rank %d equivalent strengths identical: %s" % (rank, same), flush=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("CHECK_MULTI", "PASS" if t.item() == 1.0 else "FAIL", "world", world, "all-gather mode:", multigpu._UNEVEN["mode"], flush=True)
dist.destroy_process_group()
