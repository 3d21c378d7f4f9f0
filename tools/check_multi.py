"""N-rank run (the library's own NCCL communicator, csrc/comm.cu) against the single-GPU run of the same step, bit for bit.
Launch with torchrun, 2+ GPUs; torch.distributed only ships the 128-byte NCCL id and collects the verdict:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi.py [N] [physics] [lean]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from onbody_b200.api import GpuSession, driver_inputs, comm_unique_id, shard_range_for, MEM_LEAN

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300000
physics = sys.argv[2] if len(sys.argv) > 2 else "grav3d"
lean = len(sys.argv) > 3 and sys.argv[3] == "lean"
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, r, s = driver_inputs(physics, N, True)
g = GpuSession(physics, N, N, device=local)
if lean:
    g.set_memory_mode(MEM_LEAN)
box = [comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
g.comm_init_rank(rank, world, box[0])
info = g.comm_info()
g.set_sliced_inputs(True)                       # each rank pulls 1/world of the inputs over PCIe, NVLink replicates them
g.set_sources(x, r, s); g.set_targets(x, r)
g.make_trees()
t_m = g.tree(1); s_m = g.tree(0)
g.prepare_eval()
es_m = None if lean else g.parts(2, ("s",))["s"]
g.zero_vels(); g.fastsumm(1.4)
lo, hi = shard_range_for(N, 128, rank, world)
mine = g.parts(1, ("u", "gidx", "x"))
g.zero_vels(); g.treecode3(1.4); tc3 = g.parts(1, ("u",))["u"]
ok = True
h = GpuSession(physics, N, N, device=local)
h.set_sources(x, r, s); h.set_targets(x, r)
h.make_trees(); s_r = h.tree(0); t_r = h.tree(1)
h.upward(0); h.refine(1); h.upward(1); h.zero_vels(); h.fastsumm(1.4)
ref = h.parts(1, ("u", "gidx", "x"))
h.zero_vels(); h.treecode3(1.4); tc3_r = h.parts(1, ("u",))["u"]
def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))
for k in ("u", "gidx", "x"):
    v = same(mine[k][..., lo:hi], ref[k][..., lo:hi]); ok = ok and v
    print("rank %d %s[%d:%d] bit-identical to the single-GPU run: %s" % (rank, k, lo, hi, v), flush=True)
v = same(tc3[:, lo:hi], tc3_r[:, lo:hi]); ok = ok and v
print("rank %d boxwise u identical: %s" % (rank, v), flush=True)
for name, a, b in (("source", s_m, s_r), ("target", t_m, t_r)):
    v = all(same(a[k], b[k]) for k in ("num", "ioffset", "nc", "ns", "nr", "x", "pr") + (("s",) if name == "source" else ()))
    ok = ok and v
    print("rank %d %s node arrays identical: %s" % (rank, name, v), flush=True)
if es_m is not None:
    v = same(es_m, h.parts(2, ("s",))["s"]); ok = ok and v
    print("rank %d equivalent strengths identical: %s" % (rank, v), flush=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("CHECK_MULTI", "PASS" if t.item() == 1.0 else "FAIL", "world", world, "N", N, physics, "lean" if lean else "normal", info, flush=True)
g.close(); h.close()
dist.destroy_process_group()
