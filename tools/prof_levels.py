"""per-level table of one dual-tree evaluation (ONB_DTT_PROF=1): list sizes, pairs, list/downward/pair-kernel times.
    ONB_DTT_PROF=1 python tools/prof_levels.py [N] [rank nranks]     (rank/nranks: one shard of a multi-GPU run, emulated)"""
import os, sys
os.environ.setdefault("ONB_DTT_PROF", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onbody_b200.api import GpuSession
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
g = GpuSession("grav3d", N, N)
if len(sys.argv) > 3:
    g.set_shard(int(sys.argv[2]), int(sys.argv[3]))
g.init_driver(); g.make_trees(); g.prepare_eval()
for it in range(2):
    print("---- evaluation %d" % it, file=sys.stderr, flush=True)
    g.zero_vels(); g.fastsumm(1.4)
print("eval %.3f ms lists %.3f p2p %.3f downward %.3f" % (g.phase_ms("eval"), g.phase_ms("lists"), g.phase_ms("p2p"), g.phase_ms("downward")))
