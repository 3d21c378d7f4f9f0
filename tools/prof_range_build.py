"""the range-restricted build of both trees as rank `r` of `R` runs it (explicit *_range calls, no communicator): the kernels
whose latency bounds the build side of a multi-GPU step.   python tools/prof_range_build.py [N] [rank] [nranks] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onbody_b200.api import GpuSession, driver_inputs
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
x, r, s = driver_inputs("grav3d", N, True)
g = GpuSession("grav3d", N, N); g.set_shard(rank, world)
lo, hi = g.shard_particle_range(N, rank, world)
for rep in range(reps):
    g.set_sources(x, r, s); g.set_targets(x, r)
    g.make_trees_range(lo, hi, lo, hi)
    both = g.phase_ms("tree")
    g.set_sources(x, r, s); g.make_tree_range(0, lo, hi); one = g.phase_ms("tree")
    print("rep %d: both range builds (two streams) %.3f ms, source tree alone %.3f ms, range [%d,%d)" % (rep, both, one, lo, hi), flush=True)
