"""one full hot-path step at N (default 1e7) after one warm-up step: the command profiled for profiles/*launch_list*"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onbody_b200.api import GpuSession, driver_inputs
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x, r, s = driver_inputs("grav3d", N, True)
g = GpuSession("grav3d", N, N)
for rep in range(reps):
    g.set_sources(x, r, s); g.set_targets(x, r)
    g.make_trees(); g.prepare_eval(); g.zero_vels(); g.fastsumm(1.4)       # the three calls of bench.py's hot path
    print("step %d: trees %.2f upward+refine+eq.targets %.2f eval %.2f (lists %.2f p2p %.2f down %.2f) pairs %d" % (
        rep, g.phase_ms("tree"), g.phase_ms("prepare"), g.phase_ms("eval"), g.phase_ms("lists"), g.phase_ms("p2p"), g.phase_ms("downward"), g.last_pairs()), flush=True)
